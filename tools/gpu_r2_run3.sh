#!/bin/bash
# round-2 run 3: select-kernel fix, backward without shared atomics, bulk-copy Adam, full-size parity tests, ncu captures
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2c_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-600 >> $S; }
: > $S
stage r2c_sae 900 python -m pytest tests/test_sae_gpu.py tests/test_sae_dense_gpu.py tests/test_sae_gated_gpu.py -q -x
stage r2c_parity 900 python -m pytest tests/test_parity_full_gpu.py -q -x -s
PRISMA_SAE_C_KEEP=8 stage r2c_bench_c8 600 python bench.py --workload sae --steps 20 --warmup 5
PRISMA_SAE_C_KEEP=6 stage r2c_bench_c6 600 python bench.py --workload sae --steps 20 --warmup 5
PB_SAE_ADAM=rows stage r2c_bench_adamrows 600 python bench.py --workload sae --steps 20 --warmup 5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cand_select|k_enc_cand|k_sae_adam_bulk|k_sae_grads|k_sae_decode' --launch-skip 30 -c 7 -o $OUT/r2c_sae_kernels python bench.py --workload sae --steps 3 --warmup 3 > $OUT/r2c_ncu_full.log 2>&1
echo "ncu full rc=$?" >> $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/r2c_launches_sae.csv python bench.py --workload sae --steps 3 --warmup 2 > $OUT/r2c_ncu_list.log 2>&1
echo "ncu list rc=$?" >> $S
cat $S
