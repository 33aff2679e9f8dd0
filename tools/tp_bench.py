"""Under torchrun: per-kernel-class times of one rank's decode step at a TP-sharded workload + allreduce latency."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bench import WORKLOADS
from substratus_b200 import Engine, tp
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
wl = sys.argv[1] if len(sys.argv) > 1 else "llama2-70b"
d = tempfile.mkdtemp()
json.dump(WORKLOADS[wl], open(os.path.join(d, "config.json"), "w"))
e = Engine(d, {"weights": "synthetic", "max_batch": 4, "max_seq_len": 700, "tp_size": world, "tp_rank": rank, "device": lr})
tp.connect(e)
dist.barrier()
for k in ("qkv", "attn", "o", "gate_up", "down", "lm_head"):
    ms, by = e.bench_kernel(k, rows=1, ctx=576, iters=64)
    if rank == 0:
        print(f"{k:10s} {ms*1e3:8.2f} us  {by/ms/1e6:8.1f} GB/s", flush=True)
for rows in (1, 4, 32):
    dist.barrier()
    ms, by = e.bench_kernel("allreduce", rows=min(rows, 4) if rows <= 4 else 4, ctx=576, iters=200)
    t = tp.max_over_ranks(ms)
    if rank == 0:
        print(f"allreduce rows={min(rows,4)} hidden={e.info.hidden_size}: {t*1e3:.2f} us per call (max over ranks)", flush=True)
dist.barrier()
e.close()
dist.destroy_process_group()
