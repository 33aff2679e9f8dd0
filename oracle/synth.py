"""TEST INFRASTRUCTURE — synthetic, seeded weights shared by the oracle and the engine.

No checkpoints exist offline, so every config runs on seeded
synthetic weights at the real architecture shapes.  To let the GPU engine
create 13-137 GB of weights on the device AND let the CPU oracle create the
very same values for small shapes, weights come from a counter-based hash
(splitmix64 finaliser) of (seed, tensor id, element index in the HF logical
``[out_features, in_features]`` layout):

    z  = idx + tid*0x9E3779B97F4A7C15 + seed*0xBF58476D1CE4E5B9      (mod 2^64)
    z  = (z ^ z>>30) * 0xBF58476D1CE4E5B9 ; z = (z ^ z>>27) * 0x94D049BB133111EB ; z ^= z>>31
    u  = (float(z>>41) - 2^22 + 0.5) * 2^-22                          (exact in fp32, uniform (-1,1))
    w  = bf16_rne(base + u * amp)                                     (one fp32 mul + one fp32 add)

The CUDA twin is ``ssb_synth_fill_kernel`` in ``substratus_b200/csrc/kernels.cu``;
``tests/test_parity_gpu.py::test_synth_matches_oracle`` checks them bit-for-bit.

Tensor ids: ``layer*16 + kind`` (kind: 0 q, 1 k, 2 v, 3 o, 4 gate, 5 up, 6 down,
7 input_layernorm, 8 post_attention_layernorm, 9 falcon dense_h_to_4h,
10 falcon dense_4h_to_h, 11 falcon fused query_key_value, 12/13 ln bias);
globals: ``GLOBAL + {0 embed, 1 final norm, 2 lm_head, 3 final norm bias}``.
"""
from __future__ import annotations

import numpy as np

GLOBAL = 1 << 20
K_Q, K_K, K_V, K_O, K_GATE, K_UP, K_DOWN, K_LN1, K_LN2 = range(9)
K_FC1, K_FC2, K_QKV, K_LN1_B, K_LN2_B = 9, 10, 11, 12, 13
G_EMBED, G_NORM, G_LMHEAD, G_NORM_B = 0, 1, 2, 3

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def synth_f32(seed: int, tid: int, n: int, amp: float, base: float = 0.0, start: int = 0) -> np.ndarray:
    """fp32 values (before bf16 rounding) for element indices [start, start+n)."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.uint64)
        z = idx + np.uint64(tid) * _M1 + np.uint64(seed) * _M2
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(41)).astype(np.float32) - np.float32(4194304.0) + np.float32(0.5)) * np.float32(2.0 ** -22)
    return (np.float32(base) + u * np.float32(amp)).astype(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 (returned as uint16 bit patterns)."""
    b = x.astype(np.float32).view(np.uint32)
    rnd = ((b >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    return ((b + rnd) >> np.uint32(16)).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


_fast_fill = None  # optional (seed, tid, start, n, amp, base) -> uint16 bits; tests plug in the engine's C host twin


def set_fast_fill(fn) -> None:
    """Full-width parity tests generate billions of synthetic weights; the numpy hash above manages ~8 M/s.  They may
    plug in ``substratus_b200.engine.synth_fill_host`` (the C twin of the same generator, checked bit-for-bit against
    this module in tests/test_oracle.py) to fill large tensors.  Values are identical; only the speed differs."""
    global _fast_fill
    _fast_fill = fn


def synth_bits(seed: int, tid: int, n: int, amp: float, base: float = 0.0) -> np.ndarray:
    if _fast_fill is not None and n >= (1 << 16):
        from concurrent.futures import ThreadPoolExecutor
        import os

        nthr = max(1, min(32, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
        step = -(-n // nthr)
        parts = [(s0, min(step, n - s0)) for s0 in range(0, n, step)]
        with ThreadPoolExecutor(nthr) as ex:  # ctypes releases the GIL inside the C call
            outs = list(ex.map(lambda p: _fast_fill(seed, tid, p[0], p[1], amp, base), parts))
        return np.concatenate(outs) if len(outs) > 1 else outs[0]
    return f32_to_bf16_bits(synth_f32(seed, tid, n, amp, base))


def synth_bf16(seed: int, tid: int, shape, amp: float, base: float = 0.0):
    """torch.bfloat16 tensor of ``shape`` with the engine's synthetic values."""
    import torch

    n = int(np.prod(shape))
    bits = synth_bits(seed, tid, n, amp, base)
    return torch.from_numpy(bits.view(np.int16).copy()).view(torch.bfloat16).reshape(*shape)


# amplitude conventions (the engine uses the same constants, csrc/engine.cu: synth_*):
#   matrices: uniform with std 0.02  -> amp = 0.02*sqrt(3)
#   lm_head : amp * LMHEAD_GAIN so that greedy argmax has a visible margin
#   norms   : 1 + 0.1*u
W_AMP = 0.02 * 3.0 ** 0.5
LMHEAD_GAIN = 4.0
NORM_AMP = 0.1


def llama_tensor_specs(cfg: dict) -> dict:
    """HF tensor name -> (tensor id, shape, amp, base) for a Llama-family config dict (HF ``config.json`` keys)."""
    h = cfg["hidden_size"]
    nh = cfg["num_attention_heads"]
    nkv = cfg.get("num_key_value_heads", nh)
    d = cfg.get("head_dim") or h // nh
    inter = cfg["intermediate_size"]
    v = cfg["vocab_size"]
    sp = {"model.embed_tokens.weight": (GLOBAL + G_EMBED, (v, h), W_AMP, 0.0)}
    for l in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{l}."
        t = l * 16
        sp[p + "self_attn.q_proj.weight"] = (t + K_Q, (nh * d, h), W_AMP, 0.0)
        sp[p + "self_attn.k_proj.weight"] = (t + K_K, (nkv * d, h), W_AMP, 0.0)
        sp[p + "self_attn.v_proj.weight"] = (t + K_V, (nkv * d, h), W_AMP, 0.0)
        sp[p + "self_attn.o_proj.weight"] = (t + K_O, (h, nh * d), W_AMP, 0.0)
        sp[p + "mlp.gate_proj.weight"] = (t + K_GATE, (inter, h), W_AMP, 0.0)
        sp[p + "mlp.up_proj.weight"] = (t + K_UP, (inter, h), W_AMP, 0.0)
        sp[p + "mlp.down_proj.weight"] = (t + K_DOWN, (h, inter), W_AMP, 0.0)
        sp[p + "input_layernorm.weight"] = (t + K_LN1, (h,), NORM_AMP, 1.0)
        sp[p + "post_attention_layernorm.weight"] = (t + K_LN2, (h,), NORM_AMP, 1.0)
    sp["model.norm.weight"] = (GLOBAL + G_NORM, (h,), NORM_AMP, 1.0)
    sp["lm_head.weight"] = (GLOBAL + G_LMHEAD, (v, h), W_AMP * LMHEAD_GAIN, 0.0)
    return sp


def llama_state_dict(cfg: dict, seed: int) -> dict:
    """HF-named state dict (bf16) for a Llama-family config dict."""
    return {k: synth_bf16(seed, tid, shape, amp, base) for k, (tid, shape, amp, base) in llama_tensor_specs(cfg).items()}


class LazyStateDict:
    """Mapping that generates each synthetic tensor on access (in ``dtype``) and keeps nothing: a full-depth 7B oracle in
    fp32 would otherwise hold 27 GB.  ``oracle.llama_ref.LlamaRef`` / ``falcon_ref.FalconRef`` take it in place of a dict."""

    lazy = True

    def __init__(self, specs: dict, seed: int, dtype):
        self.specs, self.seed, self.dtype = specs, seed, dtype

    def __getitem__(self, k):
        tid, shape, amp, base = self.specs[k]
        return synth_bf16(self.seed, tid, shape, amp, base).to(self.dtype)

    def keys(self):
        return self.specs.keys()


LLAMA2_7B = dict(model_type="llama", hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=32, vocab_size=32000,
                 max_position_embeddings=4096, rms_norm_eps=1e-5, rope_theta=10000.0,
                 tie_word_embeddings=False, torch_dtype="bfloat16")
LLAMA2_13B = dict(LLAMA2_7B, hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                  num_attention_heads=40, num_key_value_heads=40)
LLAMA2_70B = dict(LLAMA2_7B, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                  num_attention_heads=64, num_key_value_heads=8)
TINY_MHA = dict(LLAMA2_7B, hidden_size=256, intermediate_size=688, num_hidden_layers=2,
                num_attention_heads=2, num_key_value_heads=2, vocab_size=512, max_position_embeddings=512)
TINY_GQA = dict(LLAMA2_7B, hidden_size=512, intermediate_size=1376, num_hidden_layers=3,
                num_attention_heads=4, num_key_value_heads=2, vocab_size=1008, max_position_embeddings=512)
