#!/bin/bash
# Round 2, single GPU: the batched owner epilogue of the stream-K projections (TC_SK_EPI_V2) — parity suite first, then the
# A/B against the first version (variant library skepi1) per kernel class and on the batch-32 bench line, the phase stamps
# of the new version, and the driver's default command.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -k 20 900 python -m pytest tests -m gpu -q -x ) > $O/rf7_suite.log 2>&1; tail -3 $O/rf7_suite.log
timeout -k 20 200 python tools/kbench.py '{}' 32 > $O/rf7_kbench_b32_v2.txt 2>&1; cat $O/rf7_kbench_b32_v2.txt
SSB_LIB_VARIANT=skepi1 timeout -k 20 200 python tools/kbench.py '{}' 32 > $O/rf7_kbench_b32_v1.txt 2>&1; cat $O/rf7_kbench_b32_v1.txt
SSB_LIB_VARIANT=skprof timeout -k 20 200 python tools/sk_prof.py 32 > $O/rf7_sk_prof_v2.txt 2>&1; grep -A3 "iters 16\|acc_seen -> epi" $O/rf7_sk_prof_v2.txt | grep "iters 16\|acc_seen ->" 
( time timeout -k 20 600 python bench.py --steps 20 --warmup 5 ) > $O/rf7_bench_full.log 2>&1; tail -1 $O/rf7_bench_full.log | cut -c1-300
