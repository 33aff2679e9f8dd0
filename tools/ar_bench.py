"""Under torchrun: allreduce latency variants (tiny model, hidden 8192 via a fake config)."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from substratus_b200 import Engine, tp
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = dict(model_type="llama", hidden_size=8192, intermediate_size=2048, num_hidden_layers=1, num_attention_heads=64,
           num_key_value_heads=8, vocab_size=1024, max_position_embeddings=256, rms_norm_eps=1e-5, rope_theta=10000.0)
d = tempfile.mkdtemp()
json.dump(cfg, open(os.path.join(d, "config.json"), "w"))
e = Engine(d, {"weights": "synthetic", "max_batch": 4, "max_seq_len": 128, "tp_size": world, "tp_rank": rank, "device": lr})
tp.connect(e)
for name in ("allreduce_ll", "allreduce_lln", "allreduce@v0", "allreduce@v1", "allreduce@v2", "allreduce@v3", "allreduce@v0n", "allreduce@v3n", "allreduce@v4", "allreduce@v6", "allreduce@v7", "allreduce@v7n"):
    dist.barrier()
    ms, _ = e.bench_kernel(name, rows=1, ctx=64, iters=300)
    t = tp.max_over_ranks(ms)
    if rank == 0:
        print(f"{name:16s} {t*1e3:7.2f} us per call", flush=True)
dist.barrier()
e.close()
dist.destroy_process_group()
