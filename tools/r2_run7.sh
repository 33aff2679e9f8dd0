#!/bin/bash
# Round-2, seventh GPU call (ONE GPU): the tensor-pipe consumer (MG_MMA) of the persistent kernel: oracle parity gate (the
# summation order changes, so bit-identity does not apply), then A/B lines.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r7_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
echo "== 1. parity gate: mma variant through the oracle suites"
SSB_LIB_VARIANT=mma timeout -k 20 1200 python -m pytest tests/test_parity_gpu.py tests/test_fullwidth_gpu.py tests/test_gguf_gpu.py -m gpu -q 2>&1 | tail -6 | tee $O/r7_parity_mma.log
echo "== 2. 7B batch 1 / 2 / 4"
for V in "" mma mmana mmamin4; do
  for BB in 1 2 4; do
    SSB_LIB_VARIANT=$V timeout -k 20 300 $B --batch $BB 2>&1 | tail -1 | tee -a $O/r7_bench.jsonl | cut -c1-120
  done
done
echo "== 3. 13B / 70B"
for V in mma mmana; do
  SSB_LIB_VARIANT=$V timeout -k 20 400 $B --workload llama2-13b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r7_bench.jsonl | cut -c1-120
  SSB_LIB_VARIANT=$V timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r7_bench.jsonl | cut -c1-120
done
echo "== 4. timeline (mmamin4)"
SSB_LIB_VARIANT=mmamin4 timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r7_mega_prof_mmamin4.log
python - <<'PY'
import json
for ln in open("gpurun_out/r7_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed", ln[:200]); continue
    print(f'{d.get("engine", "?")[28:]:18s} {d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
