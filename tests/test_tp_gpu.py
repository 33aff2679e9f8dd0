"""Tensor-parallel parity on >= 2 GPUs (skipped on a 1-GPU box): TP=2 (and 4 when available) engines, one process
per GPU, allreduce over NVLink peer memory, must reproduce the TP=1 logits up to fp32 summation order and meet the
same oracle tolerance; every rank must produce identical token ids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import llama_ref, synth
from util import rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, model_dir, prompts, ngen, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from substratus_b200 import Engine, tp

        e = Engine(model_dir, dict(mode, tp_size=world, tp_rank=rank, device=rank, max_batch=4, max_seq_len=160))
        tp.connect(e)
        toks, lg = e.generate(prompts, ngen, want_logits=True)
        dist.barrier()
        e.close()
        q.put((rank, toks, lg if rank == 0 else None))
    except Exception as ex:  # surface the failure instead of hanging the parent
        q.put((rank, repr(ex), None))
    finally:
        dist.destroy_process_group()


# Decode exchange modes (csrc/mega.h): "tp_mega" 0 = multi-kernel path with an allreduce kernel per row-parallel projection
# ("tp_ll": 1 LL push kernel for <= 4 rows, default; 0 one-shot pull kernel), 1 / 2 = persistent
# kernel with grid-wide / per-CTA flag + pull, 3 (engine default) = persistent kernel with the 16-byte {value, epoch} push.
# {"tp_two_shot": 1}: reduce-scatter + bf16 gather allreduce for prefill-sized forwards (csrc/tp_twoshot.cu).
# All of them first ran on 2 x B200 in round 2 (profiles/r02_tp_parity_n2.log).
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", [{"gemm_path": "gemv", "tp_mega": 0}, {"gemm_path": "gemv", "tp_mega": 0, "tp_ll": 0}, {"gemm_path": "tc"}, {"gemm_path": "gemv", "tp_mega": 1},
                                  {"gemm_path": "tc", "tp_two_shot": 1, "tp_two_shot_min_rows": 16}, {"gemm_path": "gemv", "tp_mega": 2},
                                  {"gemm_path": "gemv", "tp_mega": 3}, {}],
                         ids=["gemv_multikernel_ll", "gemv_multikernel_pull", "tc", "tp_mega1", "two_shot", "tp_mega2", "tp_mega3", "default"])
def test_tp_matches_tp1_and_oracle(tmp_path, world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from substratus_b200 import Engine

    cfg = dict(synth.TINY_GQA, num_attention_heads=8, num_key_value_heads=4, hidden_size=1024, intermediate_size=2752)
    sd = synth.llama_state_dict(cfg, 17)
    llama_ref.write_hf_dir(str(tmp_path), cfg, sd)
    gen = torch.Generator().manual_seed(5)
    prompts = [torch.randint(0, cfg["vocab_size"], (n,), generator=gen).tolist() for n in (19, 40)]
    ngen = 6
    with Engine(str(tmp_path), dict({k: v for k, v in mode.items() if not k.startswith("tp_")}, max_batch=4, max_seq_len=160)) as e:
        t1, l1 = e.generate(prompts, ngen, want_logits=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), prompts, ngen, mode, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=240) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
    for r, toks, _ in res:
        assert not isinstance(toks, str), f"rank {r}: {toks}"
        assert np.array_equal(toks, res[0][1]), f"rank {r} diverged"
    ltp = res[0][2]
    # TP-N vs TP1 on identical inputs (first step): only the fp32 cross-rank summation order and the bf16 rounding flips it
    # causes differ.  Measured on 2 x B200 in round 1: 6.6e-3; bound = 1.5 x that ("tighten toward the measured value")
    e1 = rel_err(ltp[0], l1[0])
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_tp.txt", "a") as f:
        f.write(f"[tp{world} {mode}] first-step logits rel err vs TP1 {e1:.3e}\n")
    assert e1 < 1e-2, e1
    ref32 = llama_ref.LlamaRef(cfg, sd, torch.float32).forward(torch.tensor([prompts[1]]))[0, -1].numpy()
    refbf = llama_ref.LlamaRef(cfg, sd, torch.bfloat16).forward(torch.tensor([prompts[1]]))[0, -1].float().numpy()
    eg, ec = rel_err(ltp[0, 1], ref32), rel_err(refbf, ref32)
    with open("gpurun_out/parity_tp.txt", "a") as f:
        f.write(f"[tp{world} {mode}] first-token logits vs fp32 oracle: engine {eg:.3e}, CPU bf16 oracle {ec:.3e}\n")
    # ONE position = one draw of the bf16 noise: the single-draw slack of tests/test_parity_gpu.py::_assert_parity (the mean
    # criterion without slack is applied to 12 positions by bench.py's tp_parity object in every multi-GPU line)
    assert eg <= 1.5 * ec + 1e-3, (eg, ec)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("mode", [{"gemm_path": "gemv"}, {"gemm_path": "tc"}], ids=["gemv", "tc"])
def test_falcon_tp_matches_tp1_and_oracle(tmp_path, world, mode):
    """BASELINE config 4 is Falcon-40B at TP2 / TP4: the Falcon layout under tensor parallelism (fused QKV sharded by KV
    group, ONE allreduce per layer for the parallel block: attention and MLP partials are summed locally first) against
    TP1 and the Falcon oracle."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import falcon_ref as fr
    from substratus_b200 import Engine

    cfg = dict(fr.TINY_FALCON, hidden_size=512, num_attention_heads=8, num_kv_heads=4)
    sd = fr.falcon_state_dict(cfg, 29)
    llama_ref.write_hf_dir(str(tmp_path), cfg, sd)
    gen = torch.Generator().manual_seed(6)
    prompts = [torch.randint(0, cfg["vocab_size"], (n,), generator=gen).tolist() for n in (21, 37)]
    ngen = 6
    with Engine(str(tmp_path), dict(mode, max_batch=4, max_seq_len=160)) as e:
        t1, l1 = e.generate(prompts, ngen, want_logits=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), prompts, ngen, mode, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=240) for _ in ps], key=lambda x: x[0])
    for p in ps:
        p.join(timeout=60)
    for r, toks, _ in res:
        assert not isinstance(toks, str), f"rank {r}: {toks}"
        assert np.array_equal(toks, res[0][1]), f"rank {r} diverged"
    ltp = res[0][2]
    e1 = rel_err(ltp[0], l1[0])
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_tp.txt", "a") as f:
        f.write(f"[falcon tp{world} {mode}] first-step logits rel err vs TP1 {e1:.3e}\n")
    assert e1 < 1.5e-2, e1
    ref32 = fr.FalconRef(cfg, sd, torch.float32).forward(torch.tensor([prompts[1]]))[0, -1].numpy()
    refbf = fr.FalconRef(cfg, sd, torch.bfloat16).forward(torch.tensor([prompts[1]]))[0, -1].float().numpy()
    assert rel_err(ltp[0, 1], ref32) <= 1.5 * rel_err(refbf, ref32) + 1e-3


@pytest.mark.parametrize("family", ["llama", "falcon"])
def test_presharded_ranks_reproduce_the_full_checkpoint(tmp_path, family):
    """ranks that load their pre-sharded artifact (tools/tp_shard.py -> ssb_tp2/rank<r>.safetensors, the full
    shards deleted) hold the same weights as ranks that slice the full checkpoint themselves: same ids, same logits.
    (The host side of this — inventory, metadata, partition — runs on CPU in tests/test_loader_cpu.py.)"""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import shutil
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import tp_shard

    if family == "llama":
        cfg = dict(synth.TINY_GQA, num_attention_heads=8, num_key_value_heads=4, hidden_size=1024, intermediate_size=2752)
        sd = synth.llama_state_dict(cfg, 17)
    else:
        from oracle import falcon_ref as fr

        cfg = dict(fr.TINY_FALCON, hidden_size=512, num_attention_heads=8, num_kv_heads=4)
        sd = fr.falcon_state_dict(cfg, 29)
    full, pre = tmp_path / "full", tmp_path / "pre"
    llama_ref.write_hf_dir(str(full), cfg, sd)
    shutil.copytree(full, pre)
    tp_shard.shard(str(pre), world, quiet=True)
    for f in os.listdir(pre):
        if f.endswith(".safetensors"):
            os.remove(pre / f)  # only config.json + ssb_tp2/ are left
    gen = torch.Generator().manual_seed(5)
    prompts = [torch.randint(0, cfg["vocab_size"], (n,), generator=gen).tolist() for n in (19, 40)]
    out = {}
    for tag, d in (("full", full), ("pre", pre)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_worker, args=(r, world, port, str(d), prompts, 6, {}, q)) for r in range(world)]
        for p in ps:
            p.start()
        res = sorted([q.get(timeout=240) for _ in ps], key=lambda x: x[0])
        for p in ps:
            p.join(timeout=60)
        for r, toks, _ in res:
            assert not isinstance(toks, str), f"{tag} rank {r}: {toks}"
        out[tag] = res[0]
    assert np.array_equal(out["full"][1], out["pre"][1])
    assert rel_err(out["pre"][2][0], out["full"][2][0]) < 1e-3  # identical weights; only exchange timing may differ
