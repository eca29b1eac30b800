#!/bin/bash
# round-2 run 16 (1 GPU): full GPU suite, smoke, default bench, launch list of the forward-only record
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2o_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 6 $OUT/$name.log | cut -c1-1200 >> $S; }
: > $S
stage r2o_suite 1200 python -m pytest tests -q -m gpu
stage r2o_smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
stage r2o_bench 600 python bench.py
stage r2o_fwd_list 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r2o_launches_fwd.csv python bench.py --workload sae_fwd --steps 10 --warmup 3
cat $S
