#!/bin/bash
# N GPUs: persistent kernel vs multi-kernel path after proj_rows_kernel moved to the tensor pipe (batch 1..4; TP exchange modes)
set -u
N=${1:-1}
O=gpurun_out
mkdir -p $O
rm -f $O/r14_bench_n$N.jsonl
if [ $N -eq 1 ]; then RUN="python"; else RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"; fi
B="bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --no-extras --no-batch32"
if [ $N -eq 1 ]; then
  for BB in 2 3 4; do for P in '{}' '{"use_mega": 0}'; do
    timeout -k 20 300 $RUN $B --batch $BB --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r14_bench_n$N.jsonl | cut -c1-60
  done; done
else
  for WL in llama2-7b llama2-70b; do for P in '{}' '{"tp_mega": 0}'; do
    timeout -k 20 600 $RUN $B --workload $WL --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r14_bench_n$N.jsonl | cut -c1-60
  done; done
  timeout -k 20 600 $RUN $B --workload falcon-40b 2>&1 | tail -1 | tee -a $O/r14_bench_n$N.jsonl | cut -c1-60
fi
python - <<PY
import json
for ln in open("$O/r14_bench_n$N.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:200]); continue
    print(f'{d["config"]["workload"][:14]:14s} B={d["config"]["batch"]:<2d} tp{d["n_gpus"]} {json.dumps(d["config"].get("engine_params", {})):18s} {d["value"]:8.1f} tok/s {d["decode_ms_per_token"]:.3f} ms frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.2f}')
PY
