#!/bin/bash
# round-2 run 8 (8 GPUs): SAE data-parallel step, multicast vs peer exchange, cfg #4 pipeline
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2g_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 4 $OUT/$name.log | cut -c1-400 >> $S; }
: > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
PRISMA_P2P_MULTICAST=0 stage r2g_bench8_peer 600 $TR --master-port 29621 bench.py --gpus 8 --workload sae --steps 20 --warmup 5
stage r2g_bench8_mc 600 $TR --master-port 29622 bench.py --gpus 8 --workload sae --steps 20 --warmup 5
PRISMA_P2P_MULTICAST=0 stage r2g_cfg4 900 $TR --master-port 29623 bench.py --gpus 8 --workload cfg4 --dtype bf16 --steps 8 --warmup 2
cat $S
