#!/bin/bash
# Round 2, the last ~2 GPU-minutes: launch list of batch-32 DECODE steps on the final library (the shares behind the 0.62).
O=gpurun_out
mkdir -p $O
NB="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-batch32 --no-extras --graph 0"
timeout -k 5 115 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_gemm_sk_kernel|attn_decode_kernel|rmsnorm_kernel|argmax_kernel|embed_kernel' \
    -s 4000 -c 460 --csv --log-file $O/ncu_r02_launches_decode_b32_final.csv $NB --batch 32 > /dev/null 2>&1
ls -la $O/ncu_r02_launches_decode_b32_final.csv | cut -c30-
