"""NCCL baseline for the decode allreduce ("baseline ncclAllReduce on a per-device stream inside the CUDA
graph"): latency of torch.distributed.all_reduce(sum) on the decode-sized message ([1, hidden] fp32 = 16-32 KiB), 640
stream-ordered calls (the host enqueues far ahead of the device, so this is device time), max over ranks.  This is the number the engine's own
exchange (in-kernel LL push, or the one-shot pull kernel) has to beat.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_ar_bench.py"""
import json
import os

import torch
import torch.distributed as dist

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
out = {}
for hidden in (4096, 8192):
    for rows in (1, 32):
        x = torch.ones(rows, hidden, device="cuda", dtype=torch.float32)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        # stream-ordered loop (no graph: NCCL capture needs a warmed-up communicator per stream and hung on this image);
        # 640 back-to-back allreduces with one dependent elementwise kernel in between, like the decode step's residual add
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(640):
            dist.all_reduce(x)
            x.mul_(1.0 / dist.get_world_size())
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 640], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"h{hidden}_rows{rows}"] = round(float(t.item()), 2)
if rank == 0:
    print(json.dumps({"nccl_allreduce_us_incl_one_elementwise_kernel": out, "world": dist.get_world_size(), "nccl": torch.cuda.nccl.version()}))
dist.barrier()
dist.destroy_process_group()
