"""NCCL baseline for the decode allreduce ("baseline ncclAllReduce on a per-device stream inside the CUDA
graph"): latency of torch.distributed.all_reduce(sum) on the decode-sized message ([1, hidden] fp32 = 16-32 KiB), 64
calls captured in ONE CUDA graph (no launch overhead from the host), max over ranks.  This is the number the engine's own
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
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(64):
                    dist.all_reduce(x)
                    x.mul_(1.0 / dist.get_world_size())  # a dependent op between allreduces, like the decode step
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 640], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"h{hidden}_rows{rows}"] = round(float(t.item()), 2)
if rank == 0:
    print(json.dumps({"nccl_allreduce_us_incl_one_elementwise_kernel": out, "world": dist.get_world_size(), "nccl": torch.cuda.nccl.version()}))
dist.barrier()
dist.destroy_process_group()
