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
