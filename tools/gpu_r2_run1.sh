#!/bin/bash
# round-2 first GPU contact: smoke, full GPU suite, default bench (+ reference arm)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $OUT/r2_summary.txt; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $OUT/r2_summary.txt; tail -n 5 $OUT/$name.log >> $OUT/r2_summary.txt; }
: > $OUT/r2_summary.txt
stage r2_smoke 400 python __graft_entry__.py smoke
stage r2_tests 1500 python -m pytest tests -m gpu -q -x
stage r2_bench 900 python bench.py --steps 20 --warmup 5
stage r2_bench_ref 600 python bench.py --impl reference --steps 20 --warmup 5
cat $OUT/r2_summary.txt
