"""bench.py contract checks that need no GPU: the reference arm prints one well-formed JSON line from the CPU oracle port, and
the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "cpu_baseline", "e2e")


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=900, env=e)


@pytest.mark.parametrize("workload", ["vit", "sae", "cfg5", "sae_fwd"])
def test_reference_arm_prints_one_json_line(workload):
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", workload)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["vs_baseline"] is None and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    assert (d["unit"], d["config"]["workload"][:3]) == (("images/s", "vit") if workload == "vit" else ("tokens/s", "sae"))


def test_reference_arm_nonzero_ranks_exit_quietly():
    out = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2", env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal")
def test_product_arm_fails_loudly_without_a_gpu():
    out = _run("--steps", "1", "--warmup", "1")
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
