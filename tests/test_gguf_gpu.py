"""GGUF "dequant-on-load -> bf16" (BASELINE config 3, examples/llama2-13b-chat-gguf): the load-time CUDA dequant kernel
must be bit-exact with llama.cpp's gguf-py (`gguf.quants.dequantize`, the first-party block definition) after one RNE
rounding to bf16, and an engine loading a Q4_0 GGUF must match the oracle run on the gguf-py-dequantised weights."""
import numpy as np
import pytest
import torch

from oracle import llama_ref, synth
from util import rel_err

pytestmark = pytest.mark.gpu
gguf = pytest.importorskip("gguf")


def _bf16_bits(x):
    return synth.f32_to_bf16_bits(np.asarray(x, dtype=np.float32))


@pytest.mark.parametrize("qt", ["Q4_0", "Q8_0", "Q4_K", "Q6_K", "F16", "F32"])
def test_dequant_kernel_bit_exact(qt):
    from gguf import GGMLQuantizationType as T
    from gguf import quants

    from substratus_b200.engine import debug_dequant

    rng = np.random.default_rng(0)
    n = 256 * 37
    t = getattr(T, qt)
    if qt in ("Q4_0", "Q8_0", "F16", "F32"):
        raw = quants.quantize(rng.standard_normal(n, dtype=np.float32).reshape(37, 256) * 0.05, t)
    else:  # gguf-py has no K-quant quantiser: random but well-formed blocks (finite fp16 scales)
        bs = {"Q4_K": 144, "Q6_K": 210}[qt]
        raw = rng.integers(0, 256, size=(37, bs), dtype=np.uint8)
        sc = rng.standard_normal((37, 2)).astype(np.float16) * np.float16(0.01)
        if qt == "Q4_K":
            raw[:, 0:4] = sc.view(np.uint8).reshape(37, 4)
        else:
            raw[:, 208:210] = sc[:, :1].view(np.uint8).reshape(37, 2)
    want = _bf16_bits(quants.dequantize(raw, t).reshape(-1))
    got = debug_dequant(int(t), raw, n)
    assert np.array_equal(got, want), (qt, int((got != want).sum()))


def _write_gguf(path, cfg, sd, qtype):
    from gguf import GGMLQuantizationType as T
    from gguf import GGUFWriter, quants

    w = GGUFWriter(path, "llama")
    h, nh, nkv = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_key_value_heads"]
    w.add_context_length(cfg["max_position_embeddings"])
    w.add_embedding_length(h)
    w.add_block_count(cfg["num_hidden_layers"])
    w.add_feed_forward_length(cfg["intermediate_size"])
    w.add_head_count(nh)
    w.add_head_count_kv(nkv)
    w.add_layer_norm_rms_eps(cfg["rms_norm_eps"])
    w.add_rope_freq_base(cfg["rope_theta"])

    def permute(t, heads):  # llama.cpp convert_hf_to_gguf.py LlamaModel.permute
        return t.reshape(heads, 2, t.shape[0] // heads // 2, *t.shape[1:]).swapaxes(1, 2).reshape(t.shape)

    deq = {}
    names = {"self_attn.q_proj": "attn_q", "self_attn.k_proj": "attn_k", "self_attn.v_proj": "attn_v",
             "self_attn.o_proj": "attn_output", "mlp.gate_proj": "ffn_gate", "mlp.up_proj": "ffn_up",
             "mlp.down_proj": "ffn_down", "input_layernorm": "attn_norm", "post_attention_layernorm": "ffn_norm"}
    for k, v in sd.items():
        a = v.float().numpy()
        if k == "model.embed_tokens.weight":
            g = "token_embd.weight"
        elif k == "model.norm.weight":
            g = "output_norm.weight"
        elif k == "lm_head.weight":
            g = "output.weight"
        else:
            _, _, l, rest = k.split(".", 3)
            g = f"blk.{l}." + names[rest.rsplit(".", 1)[0]] + ".weight"
        if "q_proj" in k:
            a = permute(a, nh)
        if "k_proj" in k:
            a = permute(a, nkv)
        if a.ndim == 2 and a.shape[1] % 32 == 0:
            raw = quants.quantize(a, qtype)
            w.add_tensor(g, raw, raw_dtype=qtype)
            d = quants.dequantize(raw, qtype)
        else:
            w.add_tensor(g, a.astype(np.float32))
            d = a
        if "q_proj" in k or "k_proj" in k:  # undo the permutation for the HF-layout oracle
            heads = nh if "q_proj" in k else nkv
            d = d.reshape(heads, d.shape[0] // heads // 2, 2, *d.shape[1:]).swapaxes(1, 2).reshape(d.shape)
        deq[k] = torch.from_numpy(np.ascontiguousarray(d)).to(torch.bfloat16)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return deq


def test_engine_loads_q4_0_gguf(tmp_path):
    from gguf import GGMLQuantizationType as T

    from substratus_b200 import Engine

    cfg = synth.TINY_GQA
    sd = synth.llama_state_dict(cfg, 9)
    deq = _write_gguf(str(tmp_path / "model.bin"), cfg, sd, T.Q4_0)  # the loader image stores `files: model.bin`
    ids = torch.randint(0, cfg["vocab_size"], (1, 21), generator=torch.Generator().manual_seed(2))
    with Engine(str(tmp_path), {"max_batch": 2, "max_seq_len": 64}) as e:
        assert e.info.n_layers == cfg["num_hidden_layers"] and e.info.vocab_size == cfg["vocab_size"]
        toks, lg = e.generate(ids.tolist(), 4, want_logits=True)
    l32 = llama_ref.LlamaRef(cfg, deq, torch.float32).forward(ids)[0, -1].numpy()
    lbf = llama_ref.LlamaRef(cfg, deq, torch.bfloat16).forward(ids)[0, -1].float().numpy()
    assert rel_err(lg[0, 0], l32) <= 1.5 * rel_err(lbf, l32) + 1e-3
