"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_gold(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        return json.load(f)


def greedy_agree(tok_a, tok_b, logits_b, eps):
    """Greedy sequences tok_a / tok_b ([nseq, n]) agree if, per sequence, they are identical up to the first step
    where they differ AND at that step the reference logits (logits_b [nseq, n, V], the logits tok_b was picked
    from) separate the two candidates by less than eps (a rounding-level tie); after a tie the sequences follow
    different contexts and are not compared.  Returns (ok, n_exact_steps_total, message)."""
    tok_a = np.asarray(tok_a)
    tok_b = np.asarray(tok_b)
    exact = 0
    for i in range(tok_a.shape[0]):
        for s in range(tok_a.shape[1]):
            if tok_a[i, s] == tok_b[i, s]:
                exact += 1
                continue
            gap = float(logits_b[i, s, tok_b[i, s]] - logits_b[i, s, tok_a[i, s]])
            if gap >= eps:
                return False, exact, f"seq {i} step {s}: {tok_a[i, s]} vs {tok_b[i, s]} with margin {gap:.4g} >= {eps}"
            break
    return True, exact, "ok"


def rel_err(a, ref):
    """SURVEY.md §7 metric: max|a-ref| / max|ref|."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / np.abs(ref).max())


def assert_greedy_valid(cfg, sd, prompts, toks, what=""):
    """Every id in toks[i] must be a valid greedy pick for ITS OWN context: the fp32 oracle, teacher-forced with the ids
    under test, must either pick the same id or separate the two candidates by less than the bf16 noise (4 x the largest
    bf16-vs-fp32 logit difference of the oracle itself).  Unlike comparing with a solo run, this holds however batch-mates
    change the GEMM schedule.  Returns the number of exact matches."""
    import torch

    from oracle import llama_ref

    exact = 0
    r32 = llama_ref.LlamaRef(cfg, sd, torch.float32)
    rbf = llama_ref.LlamaRef(cfg, sd, torch.bfloat16)
    for p, t in zip(prompts, toks):
        ids = torch.tensor([list(p) + [int(x) for x in t[:-1]]])
        r32.reset()
        rbf.reset()
        l32 = r32.forward(ids)[0, len(p) - 1:].float().numpy()
        lbf = rbf.forward(ids)[0, len(p) - 1:].float().numpy()
        noise = 4 * float(np.abs(lbf - l32).max())
        for s_, tok in enumerate(t):
            want = int(l32[s_].argmax())
            if want == int(tok):
                exact += 1
                continue
            gap = float(l32[s_, want] - l32[s_, int(tok)])
            assert gap < noise, f"{what}: prompt len {len(p)} step {s_}: got {tok}, fp32 oracle picks {want} with margin {gap:.4g} >= noise {noise:.4g}"
    return exact
