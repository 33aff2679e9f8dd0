"""TEST INFRASTRUCTURE — generates tests/golden/*.json by running the REAL HF transformers model
(``LlamaForCausalLM``, ``attn_implementation="eager"``, the library the reference's Basaran image wraps)
on CPU over seeded synthetic weights.  Run here (HF is importable in this image); the vectors are committed
so the oracle restatement and the CUDA path are pinned to them on any box.

    python -m oracle.make_fixtures
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import llama_ref, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("tiny_mha", synth.TINY_MHA, 7, 2, 24, 16), ("tiny_gqa", synth.TINY_GQA, 3, 2, 20, 12)]


def prompts_for(cfg, nseq, plen, seed=1234):
    """The synthetic request of SURVEY.md §8d: ids = randint(0, V) with seed 1234+i."""
    rows = []
    for i in range(nseq):
        g = torch.Generator().manual_seed(seed + i)
        rows.append(torch.randint(0, cfg["vocab_size"], (plen,), generator=g))
    return torch.stack(rows)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, cfg, seed, nseq, plen, ngen in CASES:
        sd = synth.llama_state_dict(cfg, seed)
        ids = prompts_for(cfg, nseq, plen)
        rec = {"name": name, "config": cfg, "weight_seed": seed, "prompt": ids.tolist(), "max_new_tokens": ngen,
               "transformers": __import__("transformers").__version__, "torch": torch.__version__,
               "attn_implementation": "eager",
               "weights_checksum": int(sum(int(v.view(torch.int16).to(torch.int64).sum()) for v in sd.values()))}
        for tag, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
            m = llama_ref.hf_model(cfg, sd, dt)
            with torch.no_grad():
                out = m.generate(ids, max_new_tokens=ngen, do_sample=False, pad_token_id=0)
                lg = m(ids).logits[:, -1, :].float()
            rec[f"tokens_{tag}"] = out[:, plen:].tolist()
            rec[f"first_logits_{tag}"] = [[float(x) for x in row] for row in lg.numpy()]
        top2 = torch.tensor(rec["first_logits_fp32"]).topk(2, dim=-1).values
        rec["first_margin_fp32"] = [float(x) for x in (top2[:, 0] - top2[:, 1])]
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(rec, f)
        print(name, "tokens bf16", rec["tokens_bf16"][0][:8], "margin", rec["first_margin_fp32"])


if __name__ == "__main__":
    main()
