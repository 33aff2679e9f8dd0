#!/bin/bash
# Last single-GPU call of round 2: the GPU suite and the default bench line on the final library (ssb_kv_blocks, all-or-nothing
# block grants, atomic per-device attribute masks), plus the serve host with batching on against a deliberately small KV pool.
set -u
O=gpurun_out
mkdir -p $O
( time timeout -k 20 900 python -m pytest tests -m gpu -x -q ) > $O/rf5_suite.log 2>&1; tail -3 $O/rf5_suite.log
( time timeout -k 20 600 python bench.py --steps 20 --warmup 5 ) > $O/rf5_bench_full.log 2>&1; tail -1 $O/rf5_bench_full.log | cut -c1-300
rm -f $O/r2_load.jsonl
KV_BLOCKS=130 CONCURRENCY=8 timeout -k 20 400 bash tools/r2_load.sh > $O/rf5_serve_kv.log 2>&1; tail -14 $O/rf5_serve_kv.log | cut -c1-400
