#!/bin/bash
# quick tensor-parallel check after a kernel change (N GPUs): parity tests, LL / pull / NCCL allreduce latency, bench lines
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout -k 20 900 python -m pytest tests/test_tp_gpu.py -q 2>&1 | tail -6 | tee $O/r2_tp_parity_n$N.log
cat $O/parity_tp.txt | tail -24
timeout -k 20 200 $TR tools/nccl_ar_bench.py 2>&1 | tail -1 | tee $O/r2_nccl_ar_n$N.json
timeout -k 20 200 $TR tools/ar_bench.py 2>&1 | tail -12 | tee $O/r2_ar_bench_n$N.log
for P in '{}' '{"tp_mega": 0}' '{"tp_mega": 0, "tp_ll": 0}'; do
  timeout -k 20 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-batch32 --no-extras --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r2_tp_quick_n$N.jsonl | cut -c1-100
done
timeout -k 20 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-batch32 --no-extras --workload falcon-40b 2>&1 | tail -1 | tee -a $O/r2_tp_quick_n$N.jsonl | cut -c1-100
timeout -k 20 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-batch32 --no-extras --workload falcon-40b --engine-params '{"tp_ll": 0}' 2>&1 | tail -1 | tee -a $O/r2_tp_quick_n$N.jsonl | cut -c1-100
python - <<PY
import json
for ln in open("$O/r2_tp_quick_n$N.jsonl"):
    try: d = json.loads(ln)
    except ValueError: print("unparsed", ln[:300]); continue
    print(f'{d["config"]["workload"][:16]:16s} {json.dumps(d["config"].get("engine_params", {})):30s} {d["value"]:8.1f} tok/s {d["decode_ms_per_token"]:.3f} ms/tok frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.1f}')
PY
