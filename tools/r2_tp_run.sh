#!/bin/bash
# Round-2 tensor-parallel run (N GPUs of one box; charged N x): parity of every exchange mode on a small model, the NCCL
# baseline latency, then bench lines for the default path vs the persistent kernel with its exchange modes.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/r2_tp_run.sh 2'
#   gpurun --gpus 4 --timeout 1200 -- 'bash tools/r2_tp_run.sh 4 llama2-70b "3"'
# Every step has its own timeout: a cross-GPU spin bug traps after 20 s (SpinGuard), and the step timeout is the backstop.
set -u
N=${1:-2}
WL=${2:-llama2-7b}
MODES=${3:-"0 1 2 3"}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
echo "== 1. TP parity (tiny model, world <= $N): default, tc, tp_mega 1/2/3, two-shot"
SSB_EXPERIMENTAL=1 timeout -k 20 600 python -m pytest tests/test_tp_gpu.py -q 2>&1 | tail -12 | tee $O/r2_tp_parity_n$N.log
echo "== 2. NCCL baseline latency"
timeout -k 20 200 $TR tools/nccl_ar_bench.py 2>&1 | tail -1 | tee $O/r2_nccl_ar_n$N.json
echo "== 3. bench: $WL TP$N, tp_mega modes $MODES"
for M in $MODES; do
  timeout -k 20 600 $TR bench.py --gpus $N --steps 3 --warmup 2 --workload $WL --no-batch32 --no-extras --engine-params "{\"tp_mega\": $M}" 2>&1 | tail -1 \
    | tee -a $O/r2_tp_bench_n$N.jsonl | cut -c1-600
done
python - <<PY
import json
for ln in open("$O/r2_tp_bench_n$N.jsonl"):
    try:
        d = json.loads(ln)
    except ValueError:
        print("unparsed:", ln[:200]); continue
    print(f'{d["config"]["workload"][:40]:42s} {json.dumps(d["config"].get("engine_params", {})):18s} {d["value"]:8.1f} tok/s  {d["decode_ms_per_token"]:.3f} ms/token  frac/GPU {d["roofline"]["decode_step"]["frac"]:.3f}  TTFT {d["ttft_ms_p50"]:.1f} ms')
PY
