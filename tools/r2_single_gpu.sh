#!/bin/bash
# Round-2 runbook, ONE GPU: first hardware run of everything DESIGN.md section 9 lists for a single device.
#   (here, free)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_fma tools/ubench_fma.cu
#   (here, free)   make -C substratus_b200/csrc variants   # lib/libsubstratus_b200.{fhfma,fhfma12,synclight,synctree,l2ahead,combo,combo2,cw12,skprefetch,sk2cta,attnlite}.so
#   gpurun --timeout 1500 -- 'bash tools/r2_single_gpu.sh'
# Every step has its own timeout; results land in gpurun_out/r2_single_*.{log,jsonl,npz}.
# fhfma / synclight / skprefetch change neither arithmetic nor summation order, so they must reproduce the default
# library's logits BIT FOR BIT (tools/dump_logits.py at batch 1 / 3 / 12 -> tools/ab_bitexact.py); the 12-warp variants
# may differ in the last ulp of the RMSNorm statistics and fall back to the oracle parity suite as their gate.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r2_single_bench.jsonl
echo "== 1. full GPU suite on the default library (includes the tests written after the round-1 GPU budget ran out)"
timeout -k 20 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_single_suite.log
echo "== 1b. FFMA vs FHFMA.BF16 issue rate (reads the fhfma experiment)"
[ -x tools/bin/ubench_fma ] && timeout -k 5 60 tools/bin/ubench_fma 2>&1 | tee gpurun_out/r2_ubench_fma.log
echo "== 2. default library: reference logits, bench lines at batch 1 (+32)"
timeout -k 20 200 python tools/dump_logits.py gpurun_out/r2_logits_default.npz 2>&1 | tail -1
timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee -a gpurun_out/r2_single_bench.jsonl
echo "== 2b. per-kernel-class bandwidth, HBM-streamed vs L2-resident (@l2): default vs the mixed-FMA loop.  Round 1 read the"
echo "       equal '@l2' and HBM rates (~46 GB/s/SM) as an L2->SM fabric cap; the TMA service rate is 46 B/clk/SM = 88 GB/s/SM,"
echo "       so if the consumers were the limit the @l2 numbers of the fhfma build rise above the default's"
timeout -k 20 200 python tools/kbench.py '{"use_mega": 0}' 1 2>&1 | tee gpurun_out/r2_kbench_default.log
[ -f substratus_b200/lib/libsubstratus_b200.fhfma.so ] && SSB_LIB_VARIANT=fhfma timeout -k 20 200 python tools/kbench.py '{"use_mega": 0}' 1 2>&1 | tee gpurun_out/r2_kbench_fhfma.log
# variant : extra bench flags (mega-kernel variants only matter at batch <= 4; skprefetch only at batch >= 8)
for SPEC in "fhfma:--no-batch32" "fhfma12:--no-batch32" "cw12:--no-batch32" "synclight:--no-batch32" "synctree:--no-batch32" "l2ahead:--no-batch32" "combo:--no-batch32" "combo2:--no-batch32" "skprefetch:--batch 32" "sk2cta:--batch 32" "attnlite:--batch 32"; do
  V=${SPEC%%:*}; FLAGS=${SPEC#*:}
  [ -f substratus_b200/lib/libsubstratus_b200.$V.so ] || { echo "variant $V not built (make -C substratus_b200/csrc variants)"; continue; }
  echo "== 3. variant $V"
  SSB_LIB_VARIANT=$V timeout -k 20 200 python tools/dump_logits.py gpurun_out/r2_logits_$V.npz 2>&1 | tail -1
  python tools/ab_bitexact.py gpurun_out/r2_logits_default.npz gpurun_out/r2_logits_$V.npz > gpurun_out/r2_single_bitexact_$V.log 2>&1
  GATE=$?
  tail -6 gpurun_out/r2_single_bitexact_$V.log
  if [ $GATE -ne 0 ]; then
    # expected for the 12-consumer-warp variants only: the RMSNorm sum of squares is accumulated over a different thread
    # count, so the last ulp of rstd may move; they must then pass the oracle parity suite instead
    echo "variant $V is not bit-identical to the default library: running the parity suite as the gate"
    SSB_LIB_VARIANT=$V timeout -k 20 400 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2_single_parity_$V.log
    grep -q "passed" gpurun_out/r2_single_parity_$V.log && ! grep -q "failed\|error" gpurun_out/r2_single_parity_$V.log && GATE=0
  fi
  if [ $GATE -eq 0 ]; then
    SSB_LIB_VARIANT=$V timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $FLAGS 2>&1 | tail -1 | tee -a gpurun_out/r2_single_bench.jsonl
  else
    echo "variant $V failed its gate: not benchmarked"
  fi
done
echo "== 4. batch-32 decode attention: context-split sweep on the default library (heuristic gives 1 split at 32 x 32 heads)"
for S in 2 4; do
  timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 32 --engine-params "{\"attn_splits\": $S}" 2>&1 | tail -1 | tee -a gpurun_out/r2_single_bench.jsonl
done
python - <<'PY'
import json
for ln in open("gpurun_out/r2_single_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        continue
    b32 = d.get("batch32", {})
    print(f'{d.get("engine", "?") + " " + json.dumps(d["config"].get("engine_params", {})):70s} B={d["config"]["batch"]:<3d} {d["value"]:9.1f} tok/s  frac {d["roofline"]["decode_step"]["frac"]:.3f}  '
          f'batch32 {b32.get("value", 0):8.1f}  TTFT {d["ttft_ms_p50"]:.1f} ms')
PY
