#!/bin/bash
# Round-2, tenth GPU call (ONE GPU): gate of the final default library (full GPU suite), stamps compiled out A/B, then the
# driver's own commands (default line with every sub-object; reference arm).
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r10_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
timeout -k 20 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r10_suite.log
for V in "" noprof "" noprof; do
  SSB_LIB_VARIANT=$V timeout -k 20 300 $B 2>&1 | tail -1 | tee -a $O/r10_bench.jsonl | cut -c1-60
done
( time timeout -k 20 900 python bench.py --steps 20 --warmup 5 ) > $O/r10_bench_full.log 2>&1; tail -4 $O/r10_bench_full.log | cut -c1-4500
( time timeout -k 20 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/r10_bench_ref.log 2>&1; tail -4 $O/r10_bench_ref.log | cut -c1-1200
python smoke_run.py 2>&1 | tail -3
python - <<'PY'
import json
for ln in open("gpurun_out/r10_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d.get("engine", "?")[28:]:16s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
