#!/bin/bash
# round-2 run 2: bring-up of the fused encoder -> TopK path
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2b_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 12 $OUT/$name.log | cut -c1-1500 >> $S; }
: > $S
stage r2b_fused 600 python -m pytest tests/test_sae_gpu.py -q -x -k "fused"
stage r2b_sae 900 python -m pytest tests/test_sae_gpu.py tests/test_sae_dense_gpu.py tests/test_sae_splice.py -q -x
stage r2b_smoke 400 python __graft_entry__.py smoke
for c in 8 6 4; do
  PRISMA_SAE_C_KEEP=$c stage r2b_bench_c$c 600 python bench.py --workload sae --steps 20 --warmup 5
done
PRISMA_SAE_M_CAND=64 stage r2b_bench_m64 600 python bench.py --workload sae --steps 20 --warmup 5
PRISMA_SAE_M_CAND=40 stage r2b_bench_m40 600 python bench.py --workload sae --steps 20 --warmup 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2b_launches_sae.csv python bench.py --workload sae --steps 3 --warmup 2 > $OUT/r2b_ncu_list.log 2>&1
echo "ncu list rc=$?" >> $S
cat $S
