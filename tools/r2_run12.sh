#!/bin/bash
# Round-2, twelfth GPU call (ONE GPU): final-default A/Bs — calibrated row shares at gain 0.5, batch-32 attention context splits.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r12_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
for P in '{}' '{"sm_balance": 1, "sm_balance_gain": 0.5}' '{}' '{"sm_balance": 1, "sm_balance_gain": 0.5}' '{"sm_balance": 1, "sm_balance_gain": 0.35}'; do
  timeout -k 20 300 $B --no-batch32 --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r12_bench.jsonl | cut -c1-60
done
for S in 0 2 3 4; do
  timeout -k 20 400 $B --batch 32 --engine-params "{\"attn_splits\": $S}" 2>&1 | tail -1 | tee -a $O/r12_bench.jsonl | cut -c1-60
done
python - <<'PY'
import json
for ln in open("gpurun_out/r12_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]:<2d} {json.dumps(d["config"].get("engine_params", {})):44s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f} attn {d["roofline"]["per_kernel_class_gbs"]["attn"]} GB/s')
PY
