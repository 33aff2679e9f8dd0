#!/bin/bash
# Round-2, first GPU call (ONE GPU): box facts, barrier / FMA microbenchmarks, the GPU suite incl. the true-width parity
# tests, then the compile-time variant sweep of DESIGN.md section 9 (bit-identity gate -> bench line).
#   gpurun --timeout 1800 -- 'bash tools/r2_run1.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r2_single_bench.jsonl $O/parity_fullwidth.txt
echo "== 0. box"; (nproc; free -g | head -2; df -h /tmp /dev/shm | tail -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader) 2>&1 | tee $O/r2_box.txt
echo "== 1. microbenchmarks"
timeout -k 5 120 tools/bin/ubench_gridbar 2>&1 | tee $O/r2_ubench_gridbar.log
[ -x tools/bin/ubench_fma ] && timeout -k 5 60 tools/bin/ubench_fma 2>&1 | tee $O/r2_ubench_fma.log
echo "== 2. GPU suite (everything but the true-width file)"
timeout -k 20 900 python -m pytest tests -m gpu -q --ignore=tests/test_fullwidth_gpu.py 2>&1 | tail -8 | tee $O/r2_suite.log
echo "== 3. true-width parity"
timeout -k 20 1200 python -m pytest tests/test_fullwidth_gpu.py -m gpu -q --durations=8 2>&1 | tail -25 | tee $O/r2_fullwidth.log
echo "== 4. default library: logits dump, phase timeline, bench (batch 1 + 32)"
timeout -k 20 200 python tools/dump_logits.py $O/r2_logits_default.npz 2>&1 | tail -1
timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r2_mega_prof_default.log
timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>&1 | tail -1 | tee -a $O/r2_single_bench.jsonl
for SPEC in "combo2:--no-batch32" "combo:--no-batch32" "fhfma:--no-batch32" "l2ahead:--no-batch32" "synclight:--no-batch32" "synctree:--no-batch32" "cw12:--no-batch32" "fhfma12:--no-batch32" "skprefetch:--batch 32" "sk2cta:--batch 32" "attnlite:--batch 32"; do
  V=${SPEC%%:*}; FLAGS=${SPEC#*:}
  [ -f substratus_b200/lib/libsubstratus_b200.$V.so ] || { echo "variant $V not built"; continue; }
  echo "== 5. variant $V"
  SSB_LIB_VARIANT=$V timeout -k 20 200 python tools/dump_logits.py $O/r2_logits_$V.npz 2>&1 | tail -1
  python tools/ab_bitexact.py $O/r2_logits_default.npz $O/r2_logits_$V.npz > $O/r2_bitexact_$V.log 2>&1
  GATE=$?
  tail -6 $O/r2_bitexact_$V.log
  if [ $GATE -ne 0 ]; then
    echo "variant $V is not bit-identical to the default library: oracle parity suite as the gate"
    SSB_LIB_VARIANT=$V timeout -k 20 400 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3 | tee $O/r2_parity_$V.log
    grep -q "passed" $O/r2_parity_$V.log && ! grep -q "failed\|error" $O/r2_parity_$V.log && GATE=0
  fi
  if [ $GATE -eq 0 ]; then
    SSB_LIB_VARIANT=$V timeout -k 20 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras $FLAGS 2>&1 | tail -1 | tee -a $O/r2_single_bench.jsonl
    case $V in combo2|combo|l2ahead) SSB_LIB_VARIANT=$V timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r2_mega_prof_$V.log;; esac
  else
    echo "variant $V failed its gate: not benchmarked"
  fi
done
echo "== 6. the full default line (all BASELINE configs as sub-objects, CPU baseline) and the reference arm"
( time timeout -k 20 900 python bench.py --steps 5 --warmup 3 ) > $O/r2_bench_full.log 2>&1; tail -4 $O/r2_bench_full.log | cut -c1-3000
( time timeout -k 20 600 python bench.py --impl reference --steps 4 --warmup 1 ) > $O/r2_bench_ref.log 2>&1; tail -4 $O/r2_bench_ref.log | cut -c1-1500
python - <<'PY'
import json
for ln in open("gpurun_out/r2_single_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        continue
    b32 = d.get("batch32", {})
    print(f'{d.get("engine", "?"):40s} B={d["config"]["batch"]:<3d} {d["value"]:9.1f} tok/s  frac {d["roofline"]["decode_step"]["frac"]:.3f}  '
          f'batch32 {b32.get("value", 0):8.1f}  TTFT {d["ttft_ms_p50"]:.1f} ms')
PY
