#!/bin/bash
# round-2 run 6 (1 GPU): full GPU suite with the pair GEMM as default, transcoder / substitution-loss tests, default bench
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2e_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-600 >> $S; }
: > $S
stage r2e_tests 1800 python -m pytest tests -m gpu -q
stage r2e_smoke 400 python __graft_entry__.py smoke
stage r2e_bench 900 python bench.py --steps 20 --warmup 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/r2e_launches_sae.csv python bench.py --workload sae --steps 3 --warmup 2 > $OUT/r2e_ncu_list.log 2>&1
echo "ncu list rc=$?" >> $S
cat $S
