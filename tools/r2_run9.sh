#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r9_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
for V in fma "" notune fma "" notune; do
  for P in '{}' '{"sm_tune": 0}'; do
    SSB_LIB_VARIANT=$V timeout -k 20 300 $B --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r9_bench.jsonl | cut -c1-60
  done
done
python - <<'PY'
import json
for ln in open("gpurun_out/r9_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d.get("engine", "?")[28:]:16s} {json.dumps(d["config"].get("engine_params", {})):16s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
