"""Per-kernel-class device time of prefill (512 tokens) and of a decode step at batch B (Llama-2-7B shapes)."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth
from substratus_b200 import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
extra = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
d = tempfile.mkdtemp()
json.dump(synth.LLAMA2_7B, open(os.path.join(d, "config.json"), "w"))
e = Engine(d, dict(extra, weights="synthetic", max_batch=32, max_seq_len=700, profile_forward=1))
rng = np.random.default_rng(0)
prompts = [rng.integers(0, 32000, 512).tolist() for _ in range(B)]
for it in range(2):
    sids = [e.seq_create() for _ in range(B)]
    nxt, _ = e.prefill(sids, prompts)
    p = e.profile()
    if it == 1:
        print(f"prefill B={B}: total {sum(p.values()):.2f} ms ", {k: round(v, 2) for k, v in sorted(p.items(), key=lambda x: -x[1])})
    e.decode(sids, nxt, 8)
    p = e.profile()
    if it == 1:
        print(f"decode  B={B}: per step {sum(p.values())/8:.3f} ms ", {k: round(v / 8, 3) for k, v in sorted(p.items(), key=lambda x: -x[1])})
    for s in sids:
        e.seq_free(s)
