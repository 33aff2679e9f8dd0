#!/bin/bash
# Round-2, sixth GPU call (ONE GPU): the L2 lookahead actually compiled in (runs 3-5 had lost its #define), FHFMA for >= 2
# rows, calibration gain sweep.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r6_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
echo "== 1. bit-identity vs the no-lookahead build, then the parity files that cover the persistent kernel"
timeout -k 20 200 python tools/dump_logits.py $O/r6_logits_default.npz 2>&1 | tail -1
SSB_LIB_VARIANT=noahead timeout -k 20 200 python tools/dump_logits.py $O/r6_logits_noahead.npz 2>&1 | tail -1
python tools/ab_bitexact.py $O/r6_logits_noahead.npz $O/r6_logits_default.npz | tail -6
timeout -k 20 900 python -m pytest tests/test_parity_gpu.py tests/test_fullwidth_gpu.py -m gpu -q 2>&1 | tail -4
echo "== 2. 7B batch 1: lookahead off / stall-driven 14 / 8 / floor 4"
for V in noahead "" ahead8 min4; do
  SSB_LIB_VARIANT=$V timeout -k 20 300 $B 2>&1 | tail -1 | tee -a $O/r6_bench.jsonl | cut -c1-120
done
echo "== 3. calibration gain (default lib)"
for G in 0.25 0.5; do
  timeout -k 20 300 $B --engine-params "{\"sm_balance\": 1, \"sm_balance_gain\": $G}" 2>&1 | tail -1 | tee -a $O/r6_bench.jsonl | cut -c1-120
done
echo "== 4. batch 2 / 4 (FHFMA), 13B, 70B"
for BB in 2 4; do timeout -k 20 300 $B --batch $BB 2>&1 | tail -1 | tee -a $O/r6_bench.jsonl | cut -c1-120; done
timeout -k 20 400 $B --workload llama2-13b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r6_bench.jsonl | cut -c1-120
timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r6_bench.jsonl | cut -c1-120
echo "== 5. timeline of the new default"
timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r6_mega_prof_7b.log
python - <<'PY'
import json
for ln in open("gpurun_out/r6_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed", ln[:200]); continue
    print(f'{d.get("engine", "?")[28:]:18s} {d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {json.dumps(d["config"].get("engine_params", {})):44s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
