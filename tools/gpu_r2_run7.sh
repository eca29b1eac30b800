#!/bin/bash
# round-2 run 7 (2 GPUs): DP parity + timing, multicast vs peer path
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2f_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 6 $OUT/$name.log | cut -c1-600 >> $S; }
: > $S
stage r2f_sae 900 python -m pytest tests/test_sae_gpu.py tests/test_sae_splice.py tests/test_parity_full_gpu.py -q
stage r2f_bench1 600 python bench.py --workload sae --steps 20 --warmup 5
stage r2f_dp 900 python -m pytest tests/test_sae_dp_gpu.py -q -x -s
stage r2f_bench2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
PRISMA_P2P_MULTICAST=0 stage r2f_bench2_peer 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
cat $S
