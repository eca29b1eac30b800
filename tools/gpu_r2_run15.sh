#!/bin/bash
# round-2 run 15 (8 GPUs, short): SAE scaling after the deferral-order fix, rotated peer order, templated reduce-scatter, merged launches
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2n_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 3 $OUT/$name.log | cut -c1-500 >> $S; }
: > $S
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
PRISMA_P2P_OVERLAP=0 stage r2n_bench8_noov 300 $TR8 --master-port 29652 bench.py --gpus 8 --workload sae --steps 40 --warmup 5
PRISMA_P2P_OVERLAP=1 stage r2n_bench8_ov 300 $TR8 --master-port 29651 bench.py --gpus 8 --workload sae --steps 40 --warmup 5
cat $S
