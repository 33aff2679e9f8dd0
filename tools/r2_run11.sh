#!/bin/bash
# Round-2, eleventh GPU call (ONE GPU): per-head QKV -> attention flags, staged prefill QKV epilogue; gate + A/B + launch lists.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/r11_bench.jsonl $O/*.ncu-rep
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
timeout -k 20 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $O/r11_suite.log
for P in '{}' '{"mega_head_flags": 0}' '{}' '{"mega_head_flags": 0}'; do
  timeout -k 20 300 $B --no-batch32 --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r11_bench.jsonl | cut -c1-60
done
timeout -k 20 400 $B 2>&1 | tail -1 | tee -a $O/r11_bench.jsonl | cut -c1-60
timeout -k 20 400 $B --no-batch32 --workload llama2-70b --steps 2 --warmup 1 2>&1 | tail -1 | tee -a $O/r11_bench.jsonl | cut -c1-60
timeout -k 20 400 $B --no-batch32 --workload llama2-70b --steps 2 --warmup 1 --engine-params '{"mega_head_flags": 0}' 2>&1 | tail -1 | tee -a $O/r11_bench.jsonl | cut -c1-60
timeout -k 20 300 $B --no-batch32 --batch 4 2>&1 | tail -1 | tee -a $O/r11_bench.jsonl | cut -c1-60
NB="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch32 --no-extras --graph 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm_sk_kernel|attn_decode_kernel|rmsnorm_kernel|argmax_kernel|embed_kernel' \
    -s 4000 -c 460 --csv --log-file $O/ncu_r02_launches_decode_b32.csv $NB --batch 32 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm_kernel|attn_prefill_kernel|rmsnorm_kernel|proj_rows_kernel|argmax_kernel|embed_kernel' \
    -s 0 -c 240 --csv --log-file $O/ncu_r02_launches_prefill_b1.csv $NB > /dev/null 2>&1
ls -la $O/ncu_r02_launches_* | cut -c30-
python - <<'PY'
import json, csv, collections
for ln in open("gpurun_out/r11_bench.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    b32 = d.get("batch32", {})
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]} {json.dumps(d["config"].get("engine_params", {})):26s} {d["value"]:8.1f} tok/s frac {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.2f} | b32 {b32.get("value",0):.0f} TTFT {b32.get("ttft_ms_p50",0):.0f}')
for f in ("gpurun_out/ncu_r02_launches_decode_b32.csv", "gpurun_out/ncu_r02_launches_prefill_b1.csv"):
    try:
        rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    except OSError:
        continue
    if len(rows) < 2: print(f, "empty"); continue
    h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        k = r[ki].split("(")[0][:40]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f, "total", round(tot / 1e3, 1), "us over", sum(v[0] for v in agg.values()), "launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"   {k:42s} n={v[0]:4d}  {v[1]/1e3:9.1f} us  {100*v[1]/tot:5.1f} %  avg {v[1]/v[0]/1e3:7.2f} us")
PY
