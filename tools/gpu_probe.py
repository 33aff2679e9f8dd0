"""Stage-by-stage probe of the engine on a GPU box; prints (flushed) before/after every call so a hang is
attributable.  Usage: python tools/gpu_probe.py '{"use_pdl":0,"use_graph":0,"gemm_path":"gemv"}' [model]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def say(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, flush=True)


T0 = time.time()
mode = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
model = sys.argv[2] if len(sys.argv) > 2 else "tiny"
from oracle import synth  # noqa: E402
from substratus_b200 import Engine  # noqa: E402

cfg = {"tiny": synth.TINY_GQA, "7b": synth.LLAMA2_7B}[model]
d = tempfile.mkdtemp()
json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
say("creating engine", mode)
e = Engine(d, dict(mode, weights="synthetic", seed=1, max_batch=4, max_seq_len=256))
say("created; hbm GB", e.info.hbm_bytes_allocated / 1e9)
rng = np.random.default_rng(0)
for plen in (1, 2, 5, 33):
    s = e.seq_create()
    say("prefill len", plen)
    nxt, lg = e.prefill([s], [rng.integers(0, cfg["vocab_size"], plen).tolist()], want_logits=True)
    say("  -> next", nxt, "logit absmax", float(np.abs(lg).max()), "finite", bool(np.isfinite(lg).all()))
    say("decode 3 steps")
    out, _ = e.decode([s], nxt, 3)
    say("  ->", out.tolist())
    e.seq_free(s)
say("batch of 3 ragged")
toks, _ = e.generate([rng.integers(0, cfg["vocab_size"], n).tolist() for n in (3, 17, 40)], 5)
say("  ->", toks.tolist())
for k in ("qkv", "o", "gate_up", "down", "lm_head", "attn"):
    ms, by = e.bench_kernel(k, rows=1, ctx=200, iters=8)
    say(f"bench {k}: {ms*1e3:.1f} us, {by/ms/1e6:.1f} GB/s")
e.close()
say("PROBE OK")
