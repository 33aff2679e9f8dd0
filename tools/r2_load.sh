#!/bin/bash
# Throughput under load through the real container entrypoint: host/serve with continuous batching on a
# synthetic Llama-2-7B, driven over HTTP by tools/sub_infer.py --bench (512-token id prompts, 128 greedy tokens each).
#   gpurun --timeout 900 -- 'bash tools/r2_load.sh'
# KV_BLOCKS=<n> CONCURRENCY='8' : the same against a small KV pool (admission by blocks: nobody fails, some wait)
set -u
O=gpurun_out
mkdir -p $O /tmp/ssb_load/model
python - <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench, os
json.dump(bench.WORKLOADS["llama2-7b"], open("/tmp/ssb_load/model/config.json", "w"))
p = {"weights": "synthetic", "seed": 0, "batching": 1, "batch_tick": 8, "max_batch": 32, "max_seq_len": 656}
if os.environ.get("KV_BLOCKS"):  # a deliberately small pool (16-token blocks; a request needs 40): admission must queue, not fail
    p["kv_blocks"] = int(os.environ["KV_BLOCKS"])
json.dump(p, open("/tmp/ssb_load/params.json", "w"))
PY
make -C host -j4 > /dev/null 2>&1
PORT=18080 MODEL_DIR=/tmp/ssb_load/model PARAMS_FILE=/tmp/ssb_load/params.json host/serve > $O/r2_load_serve.log 2>&1 &
SRV=$!
for C in ${CONCURRENCY:-1 8 32}; do
  N=$((C * 2)); [ $N -lt 8 ] && N=8
  timeout -k 10 400 python tools/sub_infer.py --url http://127.0.0.1:18080 --wait 120 --bench $N --concurrency $C 2>&1 | tail -1 | tee -a $O/r2_load.jsonl
done
curl -s http://127.0.0.1:18080/metrics | grep -v '^#'
kill $SRV
wait $SRV 2>/dev/null
tail -3 $O/r2_load_serve.log
