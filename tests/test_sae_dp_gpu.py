"""Multi-GPU: the NVLink P2P data-parallel SAE step equals single-process training (needs >= 2 GPUs; gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dp_training_matches_single_process_reference(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert f"DP_RESULT world={world} ok=True" in out.stdout, tail
    print([ln for ln in out.stdout.splitlines() if ln.startswith("DP_RESULT")])


@pytest.mark.parametrize("world", [2, 8])
def test_dp_bf16_trainer_follows_the_fp32_trajectory(world):
    """cfg #5 class: VisionSAETrainer(p2p_group=...) with bf16 storage -- fp32 masters in peer memory, bf16 export after the deferred
    decoder all-gather -- against the reference's fp32 run of tests/golden/sae_bf16_v.pt."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29720 + world), os.path.join(ROOT, "tests", "dp_worker.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, DP_MODE="bf16_trainer"))
    tail = (out.stdout + out.stderr)[-3000:]
    assert out.returncode == 0, tail
    assert f"DP_RESULT world={world} ok=True" in out.stdout, tail
    print([ln for ln in out.stdout.splitlines() if ln.startswith("DP_RESULT")])
