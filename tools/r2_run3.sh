#!/bin/bash
# Round-2, third GPU call (ONE GPU): the new default persistent kernel (stall-driven L2 lookahead, light barrier, CTA-tile
# GQA attention) and the 256-token prefill tiles: parity gate, bit-identity vs the round-1 kernel, A/B bench lines, phase
# timeline + all-CTA skew.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r3_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
echo "== 1. GPU suite incl. true widths"
timeout -k 20 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/r3_suite.log
echo "== 2. bit-identity: default vs r1mega"
timeout -k 20 200 python tools/dump_logits.py $O/r3_logits_default.npz 2>&1 | tail -1
SSB_LIB_VARIANT=r1mega timeout -k 20 200 python tools/dump_logits.py $O/r3_logits_r1mega.npz 2>&1 | tail -1
python tools/ab_bitexact.py $O/r3_logits_r1mega.npz $O/r3_logits_default.npz | tee $O/r3_bitexact.log
echo "== 3. 7B batch 1: default / r1mega / ahead8 / ahead20"
for V in "" r1mega ahead8 ahead20; do
  SSB_LIB_VARIANT=$V timeout -k 20 300 $B --no-batch32 2>&1 | tail -1 | tee -a $O/r3_bench.jsonl | cut -c1-200
done
echo "== 4. phase timeline and skew (7B)"
timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r3_mega_prof_7b.log
timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 2>&1 | tee $O/r3_mega_skew_7b.log
echo "== 5. 70B TP1: default / per-warp attention / multi-kernel"
for P in '{}' '{"mega_attn_tile": 0}' '{"use_mega": 0}'; do
  timeout -k 20 400 $B --no-batch32 --workload llama2-70b --steps 2 --warmup 1 --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r3_bench.jsonl | cut -c1-200
done
timeout -k 20 300 python tools/mega_prof.py 1 llama2-70b 2>&1 | tee $O/r3_mega_prof_70b.log
echo "== 6. prefill tiles: 128 forced vs heuristic (TTFT batch 1 and batch 32)"
timeout -k 20 400 $B --engine-params '{"tc_tn_prefill": 128}' 2>&1 | tail -1 | tee -a $O/r3_bench.jsonl | cut -c1-200
timeout -k 20 400 $B 2>&1 | tail -1 | tee -a $O/r3_bench.jsonl | cut -c1-200
python - <<'PY'
import json
for ln in open("gpurun_out/r3_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed", ln[:200]); continue
    b32 = d.get("batch32", {})
    print(f'{d.get("engine", "?")[28:]:18s} {d["config"]["workload"][:14]:14s} {json.dumps(d["config"].get("engine_params", {})):26s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f} '
          f'TTFT {d["ttft_ms_p50"]:.2f} ms | b32 {b32.get("value", 0):7.1f} tok/s frac {b32.get("hbm_roofline_frac", 0):.3f} TTFT {b32.get("ttft_ms_p50", 0):.1f}')
PY
