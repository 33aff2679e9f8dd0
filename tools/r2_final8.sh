#!/bin/bash
# Round 2, last single-GPU call: the lane-split epilogue passes (stream-K owner + prefill tiles) — parity suite, the
# driver's default command, per-class times at 32 rows, and the prefill A/B against the 8-token epilogue (variant tcepi1).
set -u
O=gpurun_out
mkdir -p $O
( time timeout -k 20 600 python -m pytest tests -m gpu -q -x ) > $O/rf8_suite.log 2>&1; tail -3 $O/rf8_suite.log
( time timeout -k 20 400 python bench.py --steps 20 --warmup 5 ) > $O/rf8_bench_full.log 2>&1; tail -1 $O/rf8_bench_full.log | cut -c1-300
timeout -k 20 120 python tools/kbench.py '{}' 32 > $O/rf8_kbench_b32.txt 2>&1; cat $O/rf8_kbench_b32.txt
SSB_LIB_VARIANT=tcepi1 timeout -k 20 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $O/rf8_bench_tcepi1.log 2>&1; tail -1 $O/rf8_bench_tcepi1.log | cut -c1-300
SSB_LIB_VARIANT=skprof timeout -k 20 120 python tools/sk_prof.py 32 > $O/rf8_sk_prof.txt 2>&1; grep "iters 16\|acc_seen ->" $O/rf8_sk_prof.txt
