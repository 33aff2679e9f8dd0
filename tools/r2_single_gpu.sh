#!/bin/bash
# Round-2 runbook, ONE GPU: first hardware run of everything DESIGN.md section 9 lists for a single device.
#   (here, free)   make -C substratus_b200/csrc variants        # builds lib/libsubstratus_b200.{fhfma,fhfma12,synclight,cw12,skprefetch}.so
#   gpurun --timeout 1500 -- 'bash tools/r2_single_gpu.sh'
# Every step has its own timeout; results land in gpurun_out/r2_single_*.{log,jsonl}.
set -u
mkdir -p gpurun_out
echo "== 1. full GPU suite on the default library (includes the tests written after the round-1 GPU budget ran out)"
timeout -k 20 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_single_suite.log
timeout -k 20 200 python tools/dump_logits.py gpurun_out/r2_logits_default.npz 2>&1 | tail -1
echo "== 2. default bench line"
timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r2_single_bench.jsonl
for V in fhfma fhfma12 synclight cw12 skprefetch; do
  [ -f substratus_b200/lib/libsubstratus_b200.$V.so ] || { echo "variant $V not built (make -C substratus_b200/csrc variants)"; continue; }
  echo "== 3. variant $V: parity subset, then the same bench line"
  SSB_LIB_VARIANT=$V timeout -k 20 400 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_single_parity_$V.log
  SSB_LIB_VARIANT=$V timeout -k 20 200 python tools/dump_logits.py gpurun_out/r2_logits_$V.npz 2>&1 | tail -1
  python tools/ab_bitexact.py gpurun_out/r2_logits_default.npz gpurun_out/r2_logits_$V.npz 2>&1 | tee gpurun_out/r2_single_bitexact_$V.log
  if grep -q "passed" gpurun_out/r2_single_parity_$V.log && ! grep -q "failed" gpurun_out/r2_single_parity_$V.log; then
    SSB_LIB_VARIANT=$V timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r2_single_bench.jsonl
  fi
done
python - <<'PY'
import json
for ln in open("gpurun_out/r2_single_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        continue
    b32 = d.get("batch32", {})
    print(f'{d.get("engine", "?"):55s} B=1 {d["value"]:8.1f} tok/s (frac {d["roofline"]["frac"]:.3f})  B=32 {b32.get("value", 0):8.1f}  TTFT {d["ttft_ms_p50"]:.1f} ms')
PY
