#!/bin/bash
# round-2 run 9 (1 GPU): full suite after PDL / scan / gradient-kernel changes, default bench, PDL A/B, ncu evidence
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2h_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-600 >> $S; }
: > $S
stage r2h_tests 1800 python -m pytest tests -m gpu -q
stage r2h_bench 900 python bench.py --steps 20 --warmup 5
PB_PDL=0 stage r2h_bench_nopdl 600 python bench.py --workload sae --steps 20 --warmup 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/r2h_launches_sae.csv python bench.py --workload sae --steps 3 --warmup 2 > $OUT/r2h_ncu_list.log 2>&1
echo "ncu list rc=$?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cand_select|k_enc_cand|k_sae_adam_bulk|k_sae_grads|k_sae_decode' --launch-skip 40 -c 6 -o $OUT/r2h_sae_kernels python bench.py --workload sae --steps 3 --warmup 3 > $OUT/r2h_ncu_full.log 2>&1
echo "ncu full rc=$?" >> $S
cat $S
