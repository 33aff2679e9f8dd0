#!/bin/bash
# ncu evidence for one decode step (run under gpurun, 1 GPU).  $1 = output tag
set -x
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list: every kernel of a few decode steps with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > gpurun_out/ncu_bench_$TAG.log 2>&1
# full capture of the dominant kernel classes
ncu --set full --clock-control none --import-source on -k regex:proj_rows_kernel -s 800 -c 6 -o gpurun_out/prof_gemv_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 700 -c 2 -o gpurun_out/prof_attn_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --graph 0 > /dev/null 2>&1
