#!/bin/bash
# round-2 run 4 (2 GPUs): DP parity gate through the trainer, multicast probe, new select / Adam / backward kernels
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
S=$OUT/r2d_summary.txt
stage() { local name=$1 t=$2; shift 2; echo "=== $name ===" >> $S; timeout "$t" "$@" > $OUT/$name.log 2>&1; echo "$name rc=$?" >> $S; tail -n 8 $OUT/$name.log | cut -c1-700 >> $S; }
: > $S
nvidia-smi topo -m > $OUT/r2d_topo.txt 2>&1
stage r2d_sae 900 python -m pytest tests/test_sae_gpu.py tests/test_parity_full_gpu.py -q -x
stage r2d_refsuite 900 python -m pytest tests/test_reference_suite_verbatim_gpu.py -q
PB_GEMM_TC_VARIANT=3 PROBE_CHECK=1 stage r2d_gemm_pair 600 python tools/gemm_probe.py
stage r2d_gemm_base 600 python tools/gemm_probe.py
stage r2d_dp 900 python -m pytest tests/test_sae_dp_gpu.py -q -x -s
PRISMA_P2P_MULTICAST=0 stage r2d_dp_peer 900 python -m pytest tests/test_sae_dp_gpu.py -q -x -s
stage r2d_bench1 600 python bench.py --workload sae --steps 20 --warmup 5
PB_SAE_ADAM=rows stage r2d_bench1_rows 600 python bench.py --workload sae --steps 20 --warmup 5
stage r2d_bench2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
PRISMA_P2P_MULTICAST=0 stage r2d_bench2_peer 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload sae --steps 20 --warmup 5
stage r2d_bench2_all 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5
cat $S
