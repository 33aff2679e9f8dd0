#!/bin/bash
# ncu evidence, final kernels (run under gpurun, 1 GPU).  Numbers printed under ncu are never bench values.
TAG=${1:-r01}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch32 --graph 0"
# 1. launch list of batch-32 decode steps (stream-K tcgen05 projections + paged attention)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm|attn_|rmsnorm|argmax|embed|gemv' -s 3000 -c 330 \
    --csv --log-file gpurun_out/launches_b32_$TAG.csv $B --batch 32 > /dev/null 2>&1
# 2. full capture: stream-K decode GEMM (tcgen05, HBM-bound) x4, prefill GEMM (tcgen05, tensor-bound) x4, prefill attention x1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_sk_kernel -s 300 -c 4 -o gpurun_out/prof_tcsk_$TAG $B --batch 32 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|attn_prefill' -s 10 -c 5 -o gpurun_out/prof_prefill_$TAG $B > /dev/null 2>&1
ls -la gpurun_out/*$TAG*
