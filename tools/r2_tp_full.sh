#!/bin/bash
# Round-2 tensor-parallel validation at N GPUs (charged N x): TP parity tests, NCCL baseline, the driver's own command
# (full line with tp_parity / llama2_70b / falcon_40b sub-objects), then the two-shot prefill allreduce A/B.
#   gpurun --gpus 4 --timeout 1500 -- 'bash tools/r2_tp_full.sh 4'
set -u
N=${1:-4}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "== 1. TP parity tests (world 2 and 4 as available)"
timeout -k 20 900 python -m pytest tests/test_tp_gpu.py -q 2>&1 | tail -8 | tee $O/r2_tp_parity_n$N.log
cat $O/parity_tp.txt 2>/dev/null | tail -20
echo "== 2. NCCL baseline latency"
timeout -k 20 200 $TR tools/nccl_ar_bench.py 2>&1 | tail -1 | tee $O/r2_nccl_ar_n$N.json
echo "== 3. the driver's command at N=$N (all sub-objects)"
( time timeout -k 20 1200 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > $O/r2_bench_full_n$N.log 2>&1
tail -5 $O/r2_bench_full_n$N.log | cut -c1-6000
echo "== 4. prefill allreduce A/B: one-shot pull vs two-shot (TTFT at batch 1 and 32)"
for P in '{"tp_two_shot": 1}'; do
  timeout -k 20 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --no-extras --engine-params "$P" 2>&1 | tail -1 | tee -a $O/r2_tp_twoshot_n$N.jsonl | cut -c1-300
done
python - <<PY
import json
for f in ("$O/r2_bench_full_n$N.log", "$O/r2_tp_twoshot_n$N.jsonl"):
    for ln in open(f):
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        b32 = d.get("batch32", {})
        print(f'{json.dumps(d["config"].get("engine_params", {})):22s} {d["value"]:8.1f} tok/s {d["decode_ms_per_token"]:.3f} ms/tok frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f} TTFT {d["ttft_ms_p50"]:.1f} | b32 {b32.get("value", 0):.0f} tok/s TTFT {b32.get("ttft_ms_p50", 0):.0f} ms')
        for k in ("tp_parity", "llama2_70b", "falcon_40b", "tp"):
            if k in d:
                print("  ", k, json.dumps(d[k])[:900])
PY
