"""Stream-K partition arithmetic (csrc/sk_partition.h, shared by tc_gemm_sk_kernel and its launcher): exhaustive host check
that every tile has one owner, contributor parts are first segments of the CTAs right after the owner, and the owner's own
count of them is right — for all BASELINE projection shapes (TP 1-8), odd shapes and SM counts.  CPU only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_k_partition_exhaustive(tmp_path):
    exe = str(tmp_path / "sk_partition_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "sk_partition_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK "), r.stdout + r.stderr
    assert int(r.stdout.split()[1]) > 5000
