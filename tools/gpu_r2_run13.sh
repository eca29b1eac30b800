#!/bin/bash
# round-2 run 13 (8 GPUs): DP parity at 8 ranks (fp32 engine + bf16 trainer), SAE scaling with the deferred decoder all-gather on / off,
# cfg #5 at 8 ranks, 4-rank point of the curve
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2l_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 4 $OUT/$name.log | cut -c1-600 >> $S; }
: > $S
stage r2l_dp8 600 python -m pytest tests/test_sae_dp_gpu.py -q -x -s -k "8"
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
stage r2l_bench8 600 $TR8 --master-port 29631 bench.py --gpus 8 --workload sae --steps 40 --warmup 5
PRISMA_P2P_OVERLAP=0 stage r2l_bench8_noov 600 $TR8 --master-port 29632 bench.py --gpus 8 --workload sae --steps 40 --warmup 5
stage r2l_cfg5_8 600 $TR8 --master-port 29633 bench.py --gpus 8 --workload cfg5 --steps 20 --warmup 5
stage r2l_bench4 600 $TR4 --master-port 29634 bench.py --gpus 4 --workload sae --steps 40 --warmup 5
cat $S
