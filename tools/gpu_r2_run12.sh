#!/bin/bash
# round-2 run 12 (1 GPU): bf16 SAE test, CTA-pair candidate GEMM (tests + A/B bench), loose-key bf16 ViT criterion, rest of the suite
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2k_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-1500 >> $S; }
: > $S
stage r2k_bf16 600 python -m pytest tests/test_sae_bf16_gpu.py -q -x -s
stage r2k_vitbf16 600 python -m pytest tests/test_vit_gpu.py -q -x -s -k "bf16_matches_oracle"
stage r2k_pair 600 python -m pytest tests/test_sae_gpu.py -q -x -k "fused"
PB_ENC_PAIR=0 stage r2k_bench_single 600 python bench.py --workload sae --steps 40 --warmup 5
PB_ENC_PAIR=1 stage r2k_bench_pair 600 python bench.py --workload sae --steps 40 --warmup 5
PB_ENC_PAIR=1 stage r2k_cfg5_pair 600 python bench.py --workload cfg5 --steps 20 --warmup 5
stage r2k_suite 1500 python -m pytest tests -q -m gpu
cat $S
