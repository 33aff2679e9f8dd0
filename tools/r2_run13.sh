#!/bin/bash
# Round-2, thirteenth GPU call (ONE GPU): proj_rows_kernel on the tensor pipe with up to 6 ring stages: gate + A/B lines
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r13_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
timeout -k 20 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r13_suite.log
timeout -k 20 300 $B 2>&1 | tail -1 | tee -a $O/r13_bench.jsonl | cut -c1-60
timeout -k 20 300 $B --engine-params '{"use_mega": 0}' 2>&1 | tail -1 | tee -a $O/r13_bench.jsonl | cut -c1-60
timeout -k 20 400 $B --workload falcon-40b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r13_bench.jsonl | cut -c1-60
timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 1 --engine-params '{"use_mega": 0}' 2>&1 | tail -1 | tee -a $O/r13_bench.jsonl | cut -c1-60
timeout -k 20 300 $B --batch 4 --engine-params '{"use_mega": 0}' 2>&1 | tail -1 | tee -a $O/r13_bench.jsonl | cut -c1-60
python - <<'PY'
import json
for ln in open("gpurun_out/r13_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]:<2d} {json.dumps(d["config"].get("engine_params", {})):18s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.2f} classes {d["roofline"]["per_kernel_class_gbs"]}')
PY
