#!/bin/bash
# Round-2, eighth GPU call (ONE GPU): self-tuned row shares (sm_tune) on top of the tensor-pipe consumers.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r8_bench.jsonl
B="python bench.py --steps 4 --warmup 4 --no-cpu-baseline --no-extras --no-batch32"
timeout -k 20 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
for P in '{}' '{"sm_tune": 0}' '{"sm_tune": 8}'; do
  timeout -k 20 300 $B --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r8_bench.jsonl | cut -c1-100
done
timeout -k 20 300 $B --batch 4 2>&1 | tail -1 | tee -a $O/r8_bench.jsonl | cut -c1-100
timeout -k 20 300 $B --batch 4 --engine-params '{"sm_tune": 0}' 2>&1 | tail -1 | tee -a $O/r8_bench.jsonl | cut -c1-100
timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 3 2>&1 | tail -1 | tee -a $O/r8_bench.jsonl | cut -c1-100
timeout -k 20 400 $B --workload llama2-13b --steps 2 --warmup 3 2>&1 | tail -1 | tee -a $O/r8_bench.jsonl | cut -c1-100
timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 2>&1 | tail -26 | tee $O/r8_mega_skew.log
timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 '{"sm_tune": 0}' 2>&1 | tail -26 | tee $O/r8_mega_skew_notune.log
python - <<'PY'
import json
for ln in open("gpurun_out/r8_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {json.dumps(d["config"].get("engine_params", {})):20s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
