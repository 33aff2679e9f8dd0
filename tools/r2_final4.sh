#!/bin/bash
# Final 4-GPU call: the driver's command at N = 4 on the final library (profile refresh)
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611"
( time timeout -k 20 900 $TR bench.py --gpus 4 --steps 5 --warmup 3 ) > $O/rf4_bench_full_n4.log 2>&1
tail -4 $O/rf4_bench_full_n4.log | cut -c1-300
python - <<'PY'
import json
for ln in open("gpurun_out/rf4_bench_full_n4.log"):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    b32 = d.get("batch32", {})
    print(f'{d["value"]:8.1f} tok/s {d["decode_ms_per_token"]:.3f} ms/tok frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.1f} | b32 {b32.get("value", 0):.0f} tok/s TTFT {b32.get("ttft_ms_p50", 0):.0f} ms')
    for k in ("tp_parity", "llama2_70b", "falcon_40b"):
        if k in d: print("  ", k, json.dumps(d[k])[:600])
PY
