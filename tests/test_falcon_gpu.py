"""Falcon (new decoder architecture, the falcon-40b class of BASELINE config 4) parity: engine vs the oracle
restatement (pinned to HF FalconForCausalLM in tests/test_oracle.py), CUDA-core and tensor-core paths, plus TP."""
import numpy as np
import pytest
import torch

from oracle import falcon_ref as fr
from oracle import llama_ref
from util import greedy_agree, rel_err

pytestmark = pytest.mark.gpu


def _tf_logits(cfg, sd, dtype, prompts, forced):
    ref = fr.FalconRef(cfg, sd, dtype)
    out = [ref.forward(torch.tensor(prompts))[:, -1].float()]
    f = torch.tensor(forced)
    for s in range(f.shape[1] - 1):
        out.append(ref.forward(f[:, s:s + 1])[:, -1].float())
    return torch.stack(out, 1).numpy()


@pytest.mark.parametrize("mode", [{"gemm_path": "gemv"}, {"gemm_path": "tc"}, {"gemm_path": "gemv", "use_pdl": 0, "use_graph": 0}])
@pytest.mark.parametrize("kv", [2, 1])
def test_falcon_vs_oracle(tmp_path, mode, kv):
    from substratus_b200 import Engine

    cfg = dict(fr.TINY_FALCON, num_kv_heads=kv)
    sd = fr.falcon_state_dict(cfg, 11)
    llama_ref.write_hf_dir(str(tmp_path), cfg, sd, shards=2)
    gen = torch.Generator().manual_seed(3)
    prompts = [torch.randint(0, cfg["vocab_size"], (23,), generator=gen).tolist() for _ in range(2)]
    ngen = 8
    with Engine(str(tmp_path), dict(mode, max_batch=4, max_seq_len=128)) as e:
        assert e.info.model_type == b"falcon"
        toks, lg = e.generate(prompts, ngen, want_logits=True)
    lg = np.transpose(lg, (1, 0, 2))
    l32 = _tf_logits(cfg, sd, torch.float32, prompts, toks)
    lbf = _tf_logits(cfg, sd, torch.bfloat16, prompts, toks)
    eg = np.array([[rel_err(lg[i, s], l32[i, s]) for s in range(ngen)] for i in range(2)])
    ec = np.array([[rel_err(lbf[i, s], l32[i, s]) for s in range(ngen)] for i in range(2)])
    print(f"falcon kv={kv} {mode}: err_gpu mean {eg.mean():.3e} max {eg.max():.3e} | err_cpu mean {ec.mean():.3e} max {ec.max():.3e}")
    assert eg.mean() <= ec.mean() + 1e-3
    assert eg.max() <= 1.5 * ec.max() + 1e-3
    want, wl = fr.FalconRef(cfg, sd, torch.float32).generate(torch.tensor(prompts), ngen)
    ok, exact, msg = greedy_agree(toks, want.numpy(), wl.numpy(), 4 * float(np.abs(lbf - l32).max()))
    assert ok and exact >= ngen, msg


def test_falcon_synthetic_matches_file(tmp_path):
    from substratus_b200 import Engine
    import os

    cfg = fr.TINY_FALCON
    sd = fr.falcon_state_dict(cfg, 4)
    d1, d2 = tmp_path / "f", tmp_path / "s"
    llama_ref.write_hf_dir(str(d1), cfg, sd)
    llama_ref.write_hf_dir(str(d2), cfg, {})
    os.remove(d2 / "model.safetensors")
    ids = [list(range(5, 30))]
    with Engine(str(d1), {"max_batch": 2, "max_seq_len": 64}) as e:
        t1, l1 = e.generate(ids, 3, want_logits=True)
    with Engine(str(d2), {"max_batch": 2, "max_seq_len": 64, "weights": "synthetic", "seed": 4}) as e:
        t2, l2 = e.generate(ids, 3, want_logits=True)
    assert np.array_equal(l1, l2)
