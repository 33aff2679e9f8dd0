"""World-size-2 gloo tests (CPU) of the host side of the tensor-parallel path: handle gathering is rank-major and
byte-exact, shard ranges tile the dimension, the max-over-ranks timing reduction works.  The device side (peer
allreduce kernel) is covered by the gpu suite (tests/test_tp_gpu.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from substratus_b200 import tp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = bytes([rank]) * 200 + bytes(range(56))
        allb = tp.gather_blobs(blob)
        ok = len(allb) == 256 * world and all(allb[r * 256] == r and allb[r * 256 + 199] == r for r in range(world))
        ok = ok and allb[200:256] == bytes(range(56))
        mx = tp.max_over_ranks(10.0 + rank)
        q.put((rank, ok, mx))
    finally:
        dist.destroy_process_group()


def test_gather_and_max_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(mx == 11.0 for _, _, mx in res), res


def test_shard_ranges():
    assert tp.shard_ranges(64, 8) == [(i * 8, i * 8 + 8) for i in range(8)]
    with pytest.raises(ValueError):
        tp.shard_ranges(10, 4)


def test_exchange_protocols_survive_random_interleavings():
    """Model (not the CUDA code) of the experimental TP exchange protocols — flags, epochs, parity double-buffering —
    under randomised CTA scheduling: no deadlock, no stale or overwritten read (tools/tp_protocol_check.py)."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("tp_protocol_check", os.path.join(root, "tools", "tp_protocol_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(20)
