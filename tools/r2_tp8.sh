#!/bin/bash
# Round-2 validation at 8 GPUs (charged 8 x): the driver's command, then the world-4 TP parity cases that failed their
# single-draw oracle criterion in the first 4-GPU run (criterion fixed: tests/test_tp_gpu.py).
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
( time timeout -k 20 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 ) > $O/r2_bench_full_n8.log 2>&1
tail -4 $O/r2_bench_full_n8.log | cut -c1-300
timeout -k 20 600 python -m pytest tests/test_tp_gpu.py -q -k "4 and (default or tp_mega3 or two_shot or gemv_multikernel_ll or falcon)" 2>&1 | tail -5 | tee $O/r2_tp_parity_w4_on_n8.log
tail -12 $O/parity_tp.txt
python - <<'PY'
import json
for ln in open("gpurun_out/r2_bench_full_n8.log"):
    if not ln.startswith("{"): continue
    d = json.loads(ln)
    b32 = d.get("batch32", {})
    print(f'{d["value"]:8.1f} tok/s {d["decode_ms_per_token"]:.3f} ms/tok frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.1f} | b32 {b32.get("value", 0):.0f} tok/s TTFT {b32.get("ttft_ms_p50", 0):.0f} ms')
    for k in ("tp_parity", "llama2_70b", "tp"):
        if k in d: print("  ", k, json.dumps(d[k])[:1200])
PY
