"""TEST INFRASTRUCTURE — plain-torch CPU restatement of the Falcon (new_decoder_architecture, e.g. falcon-40b) forward.

Pinned against HF ``transformers==5.5.0`` ``FalconForCausalLM(attn_implementation="eager")`` in tests/test_oracle.py.
HF = site-packages/transformers/models/falcon/modeling_falcon.py:
  fused QKV split   :259-280 (_split_heads, grouped layout [kv_group, (G query heads, k, v), head_dim])
  rotary            :69-100 (same half-split pairing as Llama), FalconRotaryEmbedding
  attention (eager) :365-376 (scores /= sqrt(d) in the input dtype, softmax in the INPUT dtype, then P @ V, dense)
  MLP               :531-544 (dense_h_to_4h -> GELU(erf) -> dense_4h_to_h)
  decoder layer     :572-636 (ln_attn / ln_mlp on the same input, parallel block: out = residual + (mlp + attn))
"""
from __future__ import annotations

import math

import torch

from . import synth
from .llama_ref import apply_rope, rope_tables


def falcon_tensor_specs(cfg: dict) -> dict:
    """HF tensor name -> (tensor id, shape, amp, base)."""
    h, nh, nkv = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_kv_heads"]
    d = h // nh
    ffn = cfg.get("ffn_hidden_size") or 4 * h
    v = cfg["vocab_size"]
    sp = {"transformer.word_embeddings.weight": (synth.GLOBAL + synth.G_EMBED, (v, h), synth.W_AMP, 0.0)}
    for l in range(cfg["num_hidden_layers"]):
        p, t = f"transformer.h.{l}.", l * 16
        sp[p + "self_attention.query_key_value.weight"] = (t + synth.K_QKV, ((nh + 2 * nkv) * d, h), synth.W_AMP, 0.0)
        sp[p + "self_attention.dense.weight"] = (t + synth.K_O, (h, h), synth.W_AMP, 0.0)
        sp[p + "mlp.dense_h_to_4h.weight"] = (t + synth.K_FC1, (ffn, h), synth.W_AMP, 0.0)
        sp[p + "mlp.dense_4h_to_h.weight"] = (t + synth.K_FC2, (h, ffn), synth.W_AMP, 0.0)
        sp[p + "ln_attn.weight"] = (t + synth.K_LN1, (h,), synth.NORM_AMP, 1.0)
        sp[p + "ln_attn.bias"] = (t + synth.K_LN1_B, (h,), synth.NORM_AMP, 0.0)
        sp[p + "ln_mlp.weight"] = (t + synth.K_LN2, (h,), synth.NORM_AMP, 1.0)
        sp[p + "ln_mlp.bias"] = (t + synth.K_LN2_B, (h,), synth.NORM_AMP, 0.0)
    sp["transformer.ln_f.weight"] = (synth.GLOBAL + synth.G_NORM, (h,), synth.NORM_AMP, 1.0)
    sp["transformer.ln_f.bias"] = (synth.GLOBAL + synth.G_NORM_B, (h,), synth.NORM_AMP, 0.0)
    sp["lm_head.weight"] = (synth.GLOBAL + synth.G_LMHEAD, (v, h), synth.W_AMP * synth.LMHEAD_GAIN, 0.0)
    return sp


def falcon_state_dict(cfg: dict, seed: int) -> dict:
    return {k: synth.synth_bf16(seed, tid, shape, amp, base) for k, (tid, shape, amp, base) in falcon_tensor_specs(cfg).items()}


FALCON_40B = dict(model_type="falcon", hidden_size=8192, num_hidden_layers=60, num_attention_heads=128, num_kv_heads=8,
                  vocab_size=65024, layer_norm_epsilon=1e-5, new_decoder_architecture=True, parallel_attn=True, bias=False,
                  alibi=False, multi_query=True, rope_theta=10000.0, max_position_embeddings=2048,
                  tie_word_embeddings=False, torch_dtype="bfloat16")
TINY_FALCON = dict(FALCON_40B, hidden_size=512, num_hidden_layers=2, num_attention_heads=8, num_kv_heads=2, vocab_size=768,
                   max_position_embeddings=512)


class FalconRef:
    def __init__(self, cfg: dict, sd: dict, dtype=torch.bfloat16):
        self.cfg, self.dtype = cfg, dtype
        self.sd = sd if getattr(sd, "lazy", False) else {k: v.to(dtype) for k, v in sd.items()}
        self.h, self.nh, self.nkv = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_kv_heads"]
        self.d = self.h // self.nh
        self.L = cfg["num_hidden_layers"]
        self.eps = cfg.get("layer_norm_epsilon", 1e-5)
        self.theta = cfg.get("rope_theta", 10000.0)
        self.reset()

    def reset(self):
        self.kcache = [None] * self.L
        self.vcache = [None] * self.L

    @torch.no_grad()
    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        sd, dt, F = self.sd, self.dtype, torch.nn.functional
        b, t = ids.shape
        past = 0 if self.kcache[0] is None else self.kcache[0].shape[2]
        x = F.embedding(ids, sd["transformer.word_embeddings.weight"])
        cos, sin = rope_tables(torch.arange(past, past + t), self.d, self.theta, dt)
        g = self.nh // self.nkv
        for l in range(self.L):
            p = f"transformer.h.{l}."
            res = x
            a_in = F.layer_norm(x, (self.h,), sd[p + "ln_attn.weight"], sd[p + "ln_attn.bias"], self.eps)
            m_in = F.layer_norm(x, (self.h,), sd[p + "ln_mlp.weight"], sd[p + "ln_mlp.bias"], self.eps)
            fused = a_in @ sd[p + "self_attention.query_key_value.weight"].T
            qkv = fused.view(b, t, self.nkv, g + 2, self.d)
            shp = (b, self.nh, t, self.d)  # HF reshapes after the transpose (contiguous copies; same bf16 matmul kernels)
            q = qkv[:, :, :, :-2].flatten(2, 3).transpose(1, 2).reshape(shp)
            k = torch.broadcast_to(qkv[:, :, :, [-2]], qkv[:, :, :, :-2].shape).flatten(2, 3).transpose(1, 2).reshape(shp)
            v = torch.broadcast_to(qkv[:, :, :, [-1]], qkv[:, :, :, :-2].shape).flatten(2, 3).transpose(1, 2).reshape(shp)
            q, k = apply_rope(q, k, cos, sin)
            if self.kcache[l] is not None:
                k = torch.cat((self.kcache[l], k), dim=2)
                v = torch.cat((self.vcache[l], v), dim=2)
            self.kcache[l], self.vcache[l] = k, v
            s = q @ k.transpose(-1, -2)
            s = s / math.sqrt(self.d)
            kv_len = k.shape[2]
            qpos = torch.arange(t)[:, None] + past
            mask = torch.where(torch.arange(kv_len)[None, :] <= qpos, 0.0, torch.finfo(dt).min).to(dt)
            pr = F.softmax(s + mask[None, None], dim=-1, dtype=dt)
            ao = (pr @ v).view(b, self.nh, t, self.d).permute(0, 2, 1, 3).reshape(b, t, self.h)
            attn_out = ao @ sd[p + "self_attention.dense.weight"].T
            mlp = F.gelu(m_in @ sd[p + "mlp.dense_h_to_4h.weight"].T) @ sd[p + "mlp.dense_4h_to_h.weight"].T
            mlp = mlp + attn_out
            x = mlp + res
        x = F.layer_norm(x, (self.h,), sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], self.eps)
        return x @ sd["lm_head.weight"].T

    @torch.no_grad()
    def generate(self, prompt: torch.Tensor, max_new_tokens: int):
        self.reset()
        logits = self.forward(prompt)[:, -1, :].float()
        toks, lg = [], []
        for i in range(max_new_tokens):
            nxt = logits.argmax(dim=-1)
            toks.append(nxt)
            lg.append(logits)
            if i + 1 < max_new_tokens:
                logits = self.forward(nxt[:, None])[:, -1, :].float()
        return torch.stack(toks, 1), torch.stack(lg, 1)


def hf_model(cfg: dict, sd: dict, dtype=torch.bfloat16):
    from transformers import FalconConfig, FalconForCausalLM

    keys = {k: v for k, v in cfg.items() if k not in ("model_type", "torch_dtype", "rope_theta")}
    hcfg = FalconConfig(**keys, rope_parameters={"rope_type": "default", "rope_theta": cfg.get("rope_theta", 10000.0)},
                        attn_implementation="eager")
    m = FalconForCausalLM(hcfg).to(dtype)
    m.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=True)
    # .to(bf16) also rounded the non-persistent fp32 inv_freq buffer; from_pretrained keeps it fp32 — rebuild it
    from transformers.models.falcon.modeling_falcon import FalconRotaryEmbedding

    m.transformer.rotary_emb = FalconRotaryEmbedding(hcfg)
    return m.eval()
