"""Continuous-batching policy of the serve host (host/scheduler.h) against a fake engine on CPU:
compiled with g++ and run; see host/test_scheduler.cpp for the assertions (ids independent of batch mates, sharing
happens, max_batch respected, no slot leaks, bad requests rejected)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_batch_scheduler_with_fake_engine(tmp_path):
    exe = str(tmp_path / "test_scheduler")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wall", "-o", exe, os.path.join(ROOT, "host", "test_scheduler.cpp")],
                   check=True, cwd=os.path.join(ROOT, "host"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "SCHEDULER TEST OK" in r.stdout, r.stdout + r.stderr


def test_host_sampler_draws_from_the_filtered_softmax(tmp_path):
    """host/sampler.h (opt-in sampling of the serve host): chi-square against the exact temperature/top_k/top_p distribution."""
    exe = str(tmp_path / "test_sampler")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "host", "test_sampler.cpp")], check=True,
                   cwd=os.path.join(ROOT, "host"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "SAMPLER TEST OK" in r.stdout, r.stdout + r.stderr
