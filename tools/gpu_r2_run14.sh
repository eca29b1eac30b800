#!/bin/bash
# round-2 run 14 (2 GPUs): DP after the deferral-order fix and the merged small launches -- parity (fp32 engine, bf16 trainer), bench N=2 on / off
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2m_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 5 $OUT/$name.log | cut -c1-700 >> $S; }
: > $S
stage r2m_dp 600 python -m pytest tests/test_sae_dp_gpu.py -q -x -s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
stage r2m_bench2 600 $TR --master-port 29641 bench.py --gpus 2 --workload sae --steps 100 --warmup 10
PRISMA_P2P_OVERLAP=0 stage r2m_bench2_noov 600 $TR --master-port 29642 bench.py --gpus 2 --workload sae --steps 100 --warmup 10
cat $S
