"""bench.py contract on CPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the driver
reads, rank != 0 prints nothing, and the GPU arm refuses to run without a B200 (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_json_line():
    r = _run(["--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "0", "--ref-prompt-len", "8", "--ref-new-tokens", "3"])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_tokens_per_sec" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_falcon_family():
    """BASELINE config 4 is a Falcon model: the CPU arm must build that family too (FalconForCausalLM, eager)."""
    r = _run(["--impl", "reference", "--workload", "tiny-falcon", "--steps", "1", "--warmup", "0", "--ref-prompt-len", "8", "--ref-new-tokens", "3"])
    assert r.returncode == 0, r.stderr[-800:]
    d = json.loads([l for l in r.stdout.strip().splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["value"] > 0 and "tiny-falcon" in d["config"]["workload"]


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--workload", "tiny", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_gpu():
    if torch.cuda.is_available():
        import pytest

        pytest.skip("GPU present")
    r = _run(["--workload", "tiny", "--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
