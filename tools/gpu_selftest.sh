#!/bin/bash
# First-contact GPU run: every stage in its own process with its own timeout, logs under gpurun_out/.
# A trap / hang in one stage (e.g. a tcgen05 protocol bug) must not take the others down.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
stage() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  echo "=== $name ===" | tee -a $OUT/summary.txt
  timeout "$t" "$@" > $OUT/$name.log 2>&1
  local rc=$?
  echo "$name rc=$rc" | tee -a $OUT/summary.txt
  tail -n 6 $OUT/$name.log | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
for s in "$@"; do
  case $s in
    ops_safe) stage ops_safe 600 python -m pytest tests/test_ops_gpu.py -q -x -k "not tc" ;;
    ops_tc_bf16) stage ops_tc_bf16 300 python -m pytest tests/test_ops_gpu.py -q -k "tc_bf16" ;;
    ops_tc_tf32) stage ops_tc_tf32 300 python -m pytest tests/test_ops_gpu.py -q -k "tc_3xtf32" ;;
    vit) stage vit 900 python -m pytest tests/test_vit_gpu.py -q -s ;;
    refsuite) stage refsuite 600 python -m pytest tests/test_reference_suite_gpu.py -q ;;
    sae) stage sae 900 python -m pytest tests/test_sae_gpu.py -q -s ;;
    smoke) stage smoke 300 python __graft_entry__.py smoke ;;
    bench_fp32) stage bench_fp32 900 python bench.py --steps 5 --warmup 3 --dtype fp32 ;;
    bench_bf16) stage bench_bf16 900 python bench.py --steps 5 --warmup 3 --dtype bf16 ;;
    bench_sae) stage bench_sae 900 python bench.py --workload sae --steps 10 --warmup 3 ;;
    alltests) stage alltests 1500 python -m pytest tests -m gpu -q -x ;;
    *) stage "custom" 900 bash -c "$s" ;;
  esac
done
echo "=== done ===" >> $OUT/summary.txt
