#!/bin/bash
# Round 2, single GPU: GPU suite on the final library, then where the batch-32 stream-K projections spend their ~19 us of
# fixed time (tools/sk_prof.py on the stamped variant library) and the per-class rates at 32 rows (tools/kbench.py).
set -u
O=gpurun_out
mkdir -p $O
( time timeout -k 20 900 python -m pytest tests -m gpu -q ) > $O/rf6_suite.log 2>&1; tail -3 $O/rf6_suite.log
SSB_LIB_VARIANT=skprof timeout -k 20 300 python tools/sk_prof.py 32 > $O/rf6_sk_prof.txt 2>&1; tail -30 $O/rf6_sk_prof.txt
timeout -k 20 300 python tools/kbench.py '{}' 32 > $O/rf6_kbench_b32.txt 2>&1; cat $O/rf6_kbench_b32.txt
