#!/bin/bash
# round-2 run 11 (1 GPU): bf16 SAE (cfg #5 class) tests + bench, criterion (f) on every loose bf16 ViT key, full GPU suite
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2j_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-1500 >> $S; }
: > $S
stage r2j_bf16 600 python -m pytest tests/test_sae_bf16_gpu.py tests/test_vit_gpu.py -q -x -s -k "bf16"
stage r2j_cfg5 600 python bench.py --workload cfg5 --steps 20 --warmup 5
stage r2j_suite 1500 python -m pytest tests -q -m gpu -x
stage r2j_bench 900 python bench.py
cat $S
