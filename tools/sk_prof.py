"""Where a batched-decode stream-K projection spends its time: per-CTA globaltimer stamps of the LAST launch of
`bench_kernel(<class>, rows=32)` on Llama-2-7B shapes.  Needs the stamped library: SSB_LIB_VARIANT=skprof python tools/sk_prof.py
Columns (us from the earliest CTA's entry, median [min..max] over CTAs): entry, setup done, first weight tile landed,
last MMA committed, epilogue saw the last accumulator, epilogue done, "exit" (thread 0 after the final barrier — in practice
the time the producer lane ARRIVED there: the timer read issues before the warp blocks; do not read it as the exit time)."""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from substratus_b200 import Engine
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
extra = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
d = tempfile.mkdtemp()
json.dump(synth.LLAMA2_7B, open(os.path.join(d, "config.json"), "w"))
e = Engine(d, dict(extra, weights="synthetic", max_batch=rows, max_seq_len=700, sk_prof=1))
names = ["entry", "setup", "tile0", "mma_end", "acc_seen", "epi_end", "exit"]
for k in ("o", "qkv", "down", "gate_up"):
    for iters in (1, 16):  # 1 = an isolated launch, 16 = the last of a back-to-back run (PDL overlap with its predecessor)
        ms, by = e.bench_kernel(k, rows=rows, ctx=576, iters=iters)
        t = e.debug_read("sk_prof")
        t = t[t[:, 0] >= 0]
        seg = t[:, 7]
        print(f"{k:8s} iters {iters:2d}: {ms*1e3:7.2f} us/launch  {by/ms/1e6:7.1f} GB/s  ctas {len(t)}", flush=True)
        for i, nme in enumerate(names):
            c = t[:, i]
            c = c[c >= 0]
            print(f"    {nme:9s} median {np.median(c):7.2f}  [{c.min():7.2f} .. {c.max():7.2f}]")
        for a, b in ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 6)):
            dlt = t[:, b] - t[:, a]
            print(f"    {names[a]:>9s} -> {names[b]:9s} median {np.median(dlt):6.2f}  max {dlt.max():6.2f}")
