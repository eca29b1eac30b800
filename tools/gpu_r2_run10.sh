#!/bin/bash
# round-2 run 10 (2 GPUs): DP tests after the deferred W_dec all-gather / new reduce-scatter, overlap on vs off, default bench at N=2
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2i_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 6 $OUT/$name.log | cut -c1-900 >> $S; }
: > $S
stage r2i_dp 900 python -m pytest tests/test_sae_dp_gpu.py -q -x -s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
stage r2i_bench2 600 $TR --master-port 29621 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
PRISMA_P2P_OVERLAP=0 stage r2i_bench2_noov 600 $TR --master-port 29622 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
stage r2i_bench2_all 900 $TR --master-port 29623 bench.py --gpus 2
stage r2i_bench2_long 600 $TR --master-port 29624 bench.py --gpus 2 --workload sae --steps 100 --warmup 10
cat $S
