"""TEST INFRASTRUCTURE — plain-torch CPU restatement of the Llama decode path.

Parity status: the reference pins nothing for this path (see oracle/__init__.py:
"parity unpinned" upstream); this restatement is pinned against HF
``transformers==5.5.0`` itself in ``tests/test_oracle.py`` (bit-exact logits
in bf16 and fp32 on CPU) and against ``tests/golden/*.json``.

Each function cites the HF file:line it follows (HF = site-packages/transformers):
  rms_norm        HF:models/llama/modeling_llama.py:62-67
  rope tables     HF:models/llama/modeling_llama.py:96-113 (inv_freq), :124-135 (cos/sin cast)
  apply_rope      HF:models/llama/modeling_llama.py:138-168 (rotate_half, half-split pairing)
  attention       HF:models/llama/modeling_llama.py:187-221 (repeat_kv, eager_attention_forward)
  mlp             HF:models/llama/modeling_llama.py:177-183
  decoder layer   HF:models/llama/modeling_llama.py:292-332
  model / lm_head HF:models/llama/modeling_llama.py:355-500
  greedy loop     HF:generation/utils.py:2658-2800 (logits[:, -1].float() -> argmax)
"""
from __future__ import annotations

import math

import torch


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)  # cast to the input dtype BEFORE the weight multiply


def rope_tables(positions: torch.Tensor, head_dim: int, theta: float, dtype) -> tuple[torch.Tensor, torch.Tensor]:
    """cos/sin of shape [T, head_dim]; computed in fp32, stored in ``dtype``."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float32) / head_dim))
    freqs = (inv_freq[:, None].float() @ positions[None, :].float()).transpose(0, 1)  # [T, d/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """q,k: [B, heads, T, d]; cos/sin: [T, d] (broadcast over batch and heads)."""
    cos = cos[None, None]
    sin = sin[None, None]
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def eager_attention(q, k, v, n_rep: int, causal_offset: int) -> torch.Tensor:
    """q: [B,H,T,d]; k,v: [B,KV,S,d] (S = past + T).  Returns [B,T,H*d]."""
    b, kvh, s, d = k.shape
    if n_rep > 1:
        k = k[:, :, None].expand(b, kvh, n_rep, s, d).reshape(b, kvh * n_rep, s, d)
        v = v[:, :, None].expand(b, kvh, n_rep, s, d).reshape(b, kvh * n_rep, s, d)
    scaling = d ** -0.5
    w = torch.matmul(q, k.transpose(2, 3)) * scaling
    t = q.shape[2]
    qpos = torch.arange(t)[:, None] + causal_offset
    kpos = torch.arange(s)[None, :]
    mask = torch.where(kpos <= qpos, 0.0, torch.finfo(q.dtype).min).to(q.dtype)
    w = w + mask[None, None]
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v)  # [B,H,T,d]
    return o.transpose(1, 2).reshape(b, t, -1).contiguous()


class LlamaRef:
    """Functional Llama forward with an explicit (dense) KV cache per sequence batch."""

    def __init__(self, cfg: dict, sd: dict, dtype=torch.bfloat16):
        self.cfg = cfg
        self.dtype = dtype
        self.sd = sd if getattr(sd, "lazy", False) else {k: v.to(dtype) for k, v in sd.items()}
        self.h = cfg["hidden_size"]
        self.nh = cfg["num_attention_heads"]
        self.nkv = cfg.get("num_key_value_heads", self.nh)
        self.d = cfg.get("head_dim") or self.h // self.nh
        self.L = cfg["num_hidden_layers"]
        self.eps = cfg.get("rms_norm_eps", 1e-6)
        self.theta = cfg.get("rope_theta", 10000.0)
        self.kcache = [None] * self.L
        self.vcache = [None] * self.L
        self.taps: dict = {}

    def reset(self):
        self.kcache = [None] * self.L
        self.vcache = [None] * self.L

    @torch.no_grad()
    def forward(self, ids: torch.Tensor, tap: bool = False) -> torch.Tensor:
        """ids: [B, T] new tokens appended to the cache.  Returns logits [B, T, V] in self.dtype."""
        sd, dt = self.sd, self.dtype
        b, t = ids.shape
        past = 0 if self.kcache[0] is None else self.kcache[0].shape[2]
        x = torch.nn.functional.embedding(ids, sd["model.embed_tokens.weight"])
        cos, sin = rope_tables(torch.arange(past, past + t), self.d, self.theta, dt)
        for l in range(self.L):
            p = f"model.layers.{l}."
            res = x
            xn = rms_norm(x, sd[p + "input_layernorm.weight"], self.eps)
            q = torch.nn.functional.linear(xn, sd[p + "self_attn.q_proj.weight"]).view(b, t, self.nh, self.d).transpose(1, 2)
            k = torch.nn.functional.linear(xn, sd[p + "self_attn.k_proj.weight"]).view(b, t, self.nkv, self.d).transpose(1, 2)
            v = torch.nn.functional.linear(xn, sd[p + "self_attn.v_proj.weight"]).view(b, t, self.nkv, self.d).transpose(1, 2)
            q, k = apply_rope(q, k, cos, sin)
            if self.kcache[l] is not None:
                k = torch.cat((self.kcache[l], k), dim=2)
                v = torch.cat((self.vcache[l], v), dim=2)
            self.kcache[l], self.vcache[l] = k, v
            a = eager_attention(q, k, v, self.nh // self.nkv, past)
            if tap and l == 0:
                self.taps.update(xn0=xn.clone(), q0=q.clone(), k0=k.clone(), attn0=a.clone())
            x = res + torch.nn.functional.linear(a, sd[p + "self_attn.o_proj.weight"])
            res = x
            xn = rms_norm(x, sd[p + "post_attention_layernorm.weight"], self.eps)
            g = torch.nn.functional.linear(xn, sd[p + "mlp.gate_proj.weight"])
            u = torch.nn.functional.linear(xn, sd[p + "mlp.up_proj.weight"])
            x = res + torch.nn.functional.linear(torch.nn.functional.silu(g) * u, sd[p + "mlp.down_proj.weight"])
            if tap and l == 0:
                self.taps.update(h0=x.clone())
        x = rms_norm(x, sd["model.norm.weight"], self.eps)
        return torch.nn.functional.linear(x, sd["lm_head.weight"])

    @torch.no_grad()
    def generate(self, prompt: torch.Tensor, max_new_tokens: int):
        """Greedy.  Returns (token ids [B, n], fp32 logits used for each pick [B, n, V])."""
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
    """The real HF model (the library the reference's Basaran image wraps), eager attention."""
    from transformers import LlamaConfig, LlamaForCausalLM

    keys = {k: v for k, v in cfg.items() if k not in ("model_type", "torch_dtype", "rope_theta")}
    hcfg = LlamaConfig(**keys, rope_parameters={"rope_type": "default", "rope_theta": cfg.get("rope_theta", 10000.0)},
                       attn_implementation="eager")
    with torch.device("meta"):
        m = LlamaForCausalLM(hcfg)
    m = m.to_empty(device="cpu").to(dtype)
    missing = m.load_state_dict({k: v.to(dtype) for k, v in sd.items()}, strict=True, assign=True)
    del missing
    # rotary inv_freq is a non-persistent buffer: rebuild it after to_empty()
    from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding

    m.model.rotary_emb = LlamaRotaryEmbedding(hcfg)
    return m.eval()


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """SURVEY.md §7 metric: max|a-ref| / max|ref| (per call; caller slices positions)."""
    return float((a.double() - ref.double()).abs().max() / ref.double().abs().max())


def write_hf_dir(path: str, cfg: dict, sd: dict, shards: int = 1) -> None:
    """Write an HF snapshot dir (config.json + [sharded] safetensors + index) as the loader image would."""
    import json
    import os

    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    c = dict(cfg)
    c.setdefault("architectures", ["LlamaForCausalLM"])
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(c, f)
    names = list(sd.keys())
    if shards <= 1:
        save_file({k: sd[k].contiguous() for k in names}, os.path.join(path, "model.safetensors"))
        return
    per = math.ceil(len(names) / shards)
    wm = {}
    for s in range(shards):
        fn = f"model-{s + 1:05d}-of-{shards:05d}.safetensors"
        part = names[s * per:(s + 1) * per]
        save_file({k: sd[k].contiguous() for k in part}, os.path.join(path, fn))
        wm.update({k: fn for k in part})
    with open(os.path.join(path, "model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": wm}, f)
