#!/bin/bash
# round-2 run 17 (2 GPUs): event trace of the data-parallel step with the deferred decoder push on / off
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2p_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 3 $OUT/$name.log | cut -c1-300 >> $S; }
: > $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
PRISMA_P2P_TRACE=1 PRISMA_P2P_OVERLAP=1 stage r2p_trace_ov 300 $TR --master-port 29661 bench.py --gpus 2 --workload sae --steps 40 --warmup 5
PRISMA_P2P_TRACE=1 PRISMA_P2P_OVERLAP=0 stage r2p_trace_noov 300 $TR --master-port 29662 bench.py --gpus 2 --workload sae --steps 40 --warmup 5
cat $S
