#!/bin/bash
# Round-2, fifth GPU call (ONE GPU): consumer-side FHFMA and the prefetch floor (latency-bound far SMs), ring-style SM
# calibration; batch 4 with FHFMA.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r5_bench.jsonl
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-batch32"
echo "== 1. parity gate for the variants that change instruction selection only: bit-identity vs default"
timeout -k 20 200 python tools/dump_logits.py $O/r5_logits_default.npz 2>&1 | tail -1
for V in fh fhmin4; do
  SSB_LIB_VARIANT=$V timeout -k 20 200 python tools/dump_logits.py $O/r5_logits_$V.npz 2>&1 | tail -1
  python tools/ab_bitexact.py $O/r5_logits_default.npz $O/r5_logits_$V.npz | tail -6
done
echo "== 2. 7B batch 1"
for V in "" fh min4 fhmin4 fhmin8; do
  for P in '{}' '{"sm_balance": 0}'; do
    SSB_LIB_VARIANT=$V timeout -k 20 300 $B --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r5_bench.jsonl | cut -c1-120
  done
done
echo "== 3. skew with the ring-style calibration (default lib, then fhmin4)"
timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 2>&1 | tail -12 | tee $O/r5_mega_skew_default.log
SSB_LIB_VARIANT=fhmin4 timeout -k 20 200 python tools/mega_skew.py llama2-7b 1 2>&1 | tail -24 | tee $O/r5_mega_skew_fhmin4.log
SSB_LIB_VARIANT=fhmin4 timeout -k 20 200 python tools/mega_prof.py 1 2>&1 | tee $O/r5_mega_prof_fhmin4.log
echo "== 4. batch 4 and 70B with FHFMA"
for V in "" fhmin4; do
  SSB_LIB_VARIANT=$V timeout -k 20 300 $B --batch 4 2>&1 | tail -1 | tee -a $O/r5_bench.jsonl | cut -c1-120
  SSB_LIB_VARIANT=$V timeout -k 20 400 $B --workload llama2-70b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r5_bench.jsonl | cut -c1-120
done
python - <<'PY'
import json
for ln in open("gpurun_out/r5_bench.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed", ln[:200]); continue
    print(f'{d.get("engine", "?")[28:]:18s} {d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {json.dumps(d["config"].get("engine_params", {})):22s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f}')
PY
