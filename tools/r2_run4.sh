#!/bin/bash
# Round-2, fourth GPU call (ONE GPU): persistent kernel v3 — epilogue input prefetch, CTA-cooperative attention (G <= 4),
# per-SM weighted row shares — parity gate, then A/B lines and the timeline / skew of the new default.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r4_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
echo "== 1. GPU suite incl. true widths"
timeout -k 20 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/r4_suite.log
echo "== 2. 7B batch 1: default / equal shares / per-warp attention / both off"
for P in '{}' '{"sm_balance": 0}' '{"mega_attn_tile": 0}' '{"sm_balance": 0, "mega_attn_tile": 0}' '{"sm_balance_gain": 1.5}' '{"sm_balance_gain": 0.6}'; do
  timeout -k 20 300 $B --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r4_bench.jsonl | cut -c1-160
done
echo "== 3. phase timeline and skew (7B, new default)"
timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r4_mega_prof_7b.log
timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 2>&1 | tee $O/r4_mega_skew_7b.log
echo "== 4. 70B TP1 and 13B batch 1"
timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r4_bench.jsonl | cut -c1-160
timeout -k 20 400 $B --workload llama2-13b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r4_bench.jsonl | cut -c1-160
timeout -k 20 300 python tools/mega_prof.py 1 llama2-70b 2>&1 | tee $O/r4_mega_prof_70b.log
echo "== 5. batch 2 and 4 (persistent kernel, BT = 2 / 4)"
for BB in 2 4; do timeout -k 20 300 $B --batch $BB 2>&1 | tail -1 | tee -a $O/r4_bench.jsonl | cut -c1-160; done
python - <<'PY'
import json
for ln in open("gpurun_out/r4_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {json.dumps(d["config"].get("engine_params", {})):44s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.2f} ms')
PY
