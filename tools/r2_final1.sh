#!/bin/bash
# Final single-GPU call: batch-6 path A/B (CUDA-core/tensor-pipe row kernel in two passes vs tcgen05 stream-K), then the
# driver's own commands for the profile JSONs.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/rf_bench.jsonl
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras --no-batch32"
for P in '{}' '{"tc_min_rows": 5}'; do
  timeout -k 20 300 $B --batch 6 --engine-params "$P" 2>&1 | tail -1 | tee -a $O/rf_bench.jsonl | cut -c1-60
done
timeout -k 20 300 $B --batch 3 2>&1 | tail -1 | tee -a $O/rf_bench.jsonl | cut -c1-60
( time timeout -k 20 900 python bench.py --steps 20 --warmup 5 ) > $O/rf_bench_full.log 2>&1; tail -4 $O/rf_bench_full.log | cut -c1-400
( time timeout -k 20 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/rf_bench_ref.log 2>&1; tail -4 $O/rf_bench_ref.log | cut -c1-300
python smoke_run.py 2>&1 | tail -2
python - <<'PY'
import json
for ln in open("gpurun_out/rf_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]:<2d} {json.dumps(d["config"].get("engine_params", {})):22s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
