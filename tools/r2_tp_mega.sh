#!/bin/bash
# Round-2 runbook for the experimental tensor-parallel paths: the persistent kernel under TP ("tp_mega": 1, csrc/mega.h;
# decode tokens/s) and the two-shot prefill allreduce ("tp_two_shot": 1, csrc/tp_twoshot.cu; TTFT).
# NEVER run on hardware before this script: every step is wrapped in a short timeout because a cross-GPU spin-wait
# bug would otherwise hang the box (a gpurun strike).  Usage (charged Nx):
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_tp_mega.sh 2'
#   gpurun --gpus 4 --timeout 1200 -- 'bash tools/r2_tp_mega.sh 4 llama2-70b'
# A variant library applies to every rank when exported: SSB_LIB_VARIANT=fhfma bash tools/r2_tp_mega.sh 4 llama2-70b
# (the mixed-FMA projection loop also serves proj_rows_kernel, i.e. the default TP decode path).
set -u
N=${1:-2}
WL=${2:-llama2-7b}
mkdir -p gpurun_out
echo "== 1. parity: tp_mega and the two-shot prefill allreduce vs TP1 and the oracle (tiny model, $N GPUs)"
SSB_EXPERIMENTAL=1 timeout -k 20 420 python -m pytest tests/test_tp_gpu.py -q -k "exp_" 2>&1 | tail -15 | tee gpurun_out/r2_tp_mega_parity.log
# no -x: every experimental mode reports on its own; a mode that failed is then left out of the A/B runs below by hand
if ! grep -q "passed" gpurun_out/r2_tp_mega_parity.log || grep -q "failed\|error" gpurun_out/r2_tp_mega_parity.log; then
  echo "parity not (all) green: fix or drop the failing mode before benchmarking"; exit 1
fi
echo "== 2. decode tokens/s, default TP path vs tp_mega ($WL, TP$N)"
for P in '{}' '{"tp_mega": 1}' '{"tp_mega": 2}' '{"tp_mega": 3}' '{"tp_two_shot": 1}' '{"tp_mega": 3, "tp_two_shot": 1}'; do
  timeout -k 20 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 3 --warmup 3 --workload $WL --no-batch32 --no-cpu-baseline --engine-params "$P" 2>&1 | tail -1 \
    | tee -a gpurun_out/r2_tp_mega_bench.jsonl
done
