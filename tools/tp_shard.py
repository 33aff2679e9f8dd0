"""Write tensor-parallel PRE-SHARDED artifacts next to an HF snapshot (the Model loader's output layout).

    python tools/tp_shard.py <model_dir> --tp 4        ->  <model_dir>/ssb_tp4/rank{0..3}.safetensors

Each rank file is one safetensors container with exactly what that rank's engine uploads: its rows of the column-parallel
projections (q / k / v, gate / up; Falcon: its KV groups of the fused query_key_value and its rows of dense_h_to_4h), its
columns of the row-parallel ones (o_proj / down_proj; Falcon dense / dense_4h_to_h), and the replicated tensors (embedding,
lm_head, norms) — under the original HF tensor names, with `__metadata__` {"format": "ssb-tp", "tp_size", "tp_rank"}.
An engine created with tp_size = N picks `ssb_tp<N>/rank<r>.safetensors` up by itself (include/ssb.h, "tp_presharded") and
then reads ~1/N of the projection bytes at pod start instead of mapping the whole checkpoint on every rank.
The partition is the one csrc/engine.cu::fill_weights applies to a full checkpoint (heads, KV heads and intermediate
columns in contiguous blocks per rank); the same ids come out either way (tests/test_tp_gpu.py::test_presharded_*).
Sources: *.safetensors snapshots (Llama family, Falcon new_decoder_architecture).  No GPU needed."""
import argparse
import glob
import json
import os
import sys


def plan(cfg, tp):
    """name -> ("rows" | "cols" | "full", start, stop) per rank, as a function rank -> dict, plus a name filter."""
    mt = cfg.get("model_type", "llama")
    h = cfg["hidden_size"]
    heads = cfg["num_attention_heads"]
    if mt == "falcon":
        if not cfg.get("new_decoder_architecture", False):
            raise SystemExit("only Falcon checkpoints with new_decoder_architecture (40B / 180B layout) are supported")
        kvh = cfg.get("num_kv_heads", heads)
        inter = cfg.get("ffn_hidden_size", 4 * h)
    else:
        kvh = cfg.get("num_key_value_heads", heads)
        inter = cfg["intermediate_size"]
    d = cfg.get("head_dim", h // heads)
    if kvh % tp or inter % (8 * tp):
        raise SystemExit(f"tp {tp} must divide the KV heads ({kvh}) and intermediate_size / 8 ({inter})")
    hl, kvl, il = heads // tp, kvh // tp, inter // tp

    def for_rank(r):
        def rule(name):
            if mt == "falcon":
                gs = (heads // kvh + 2) * d  # rows of one KV group in the fused matrix: [G query heads | k | v]
                if name.endswith("self_attention.query_key_value.weight"):
                    return ("rows", r * kvl * gs, (r + 1) * kvl * gs)
                if name.endswith("self_attention.dense.weight"):
                    return ("cols", r * hl * d, (r + 1) * hl * d)
                if name.endswith("mlp.dense_h_to_4h.weight"):
                    return ("rows", r * il, (r + 1) * il)
                if name.endswith("mlp.dense_4h_to_h.weight"):
                    return ("cols", r * il, (r + 1) * il)
                return ("full", 0, 0)
            if name.endswith("self_attn.q_proj.weight"):
                return ("rows", r * hl * d, (r + 1) * hl * d)
            if name.endswith("self_attn.k_proj.weight") or name.endswith("self_attn.v_proj.weight"):
                return ("rows", r * kvl * d, (r + 1) * kvl * d)
            if name.endswith("self_attn.o_proj.weight"):
                return ("cols", r * hl * d, (r + 1) * hl * d)
            if name.endswith("mlp.gate_proj.weight") or name.endswith("mlp.up_proj.weight"):
                return ("rows", r * il, (r + 1) * il)
            if name.endswith("mlp.down_proj.weight"):
                return ("cols", r * il, (r + 1) * il)
            return ("full", 0, 0)
        return rule
    return for_rank


def shard(model_dir, tp, out_dir=None, quiet=False):
    from safetensors import safe_open
    from safetensors.torch import save_file

    cfg = json.load(open(os.path.join(model_dir, "config.json")))
    files = sorted(glob.glob(os.path.join(model_dir, "*.safetensors")))
    if not files:
        raise SystemExit(f"no *.safetensors in {model_dir} (convert .bin / GGUF snapshots first)")
    out_dir = out_dir or os.path.join(model_dir, f"ssb_tp{tp}")
    os.makedirs(out_dir, exist_ok=True)
    for_rank = plan(cfg, tp)
    written = []
    for r in range(tp):  # one pass over the source per rank: peak memory = one rank's tensors
        rule, out = for_rank(r), {}
        for fn in files:
            with safe_open(fn, framework="pt") as f:
                for name in f.keys():
                    kind, a, b = rule(name)
                    if kind == "full":
                        out[name] = f.get_tensor(name)
                    elif kind == "rows":
                        out[name] = f.get_slice(name)[a:b].contiguous()
                    else:
                        out[name] = f.get_slice(name)[:, a:b].contiguous()
        path = os.path.join(out_dir, f"rank{r}.safetensors")
        save_file(out, path, metadata={"format": "ssb-tp", "tp_size": str(tp), "tp_rank": str(r),
                                       "source": os.path.basename(os.path.abspath(model_dir))})
        written.append(path)
        if not quiet:
            print(f"{path}: {len(out)} tensors, {os.path.getsize(path) / 1e6:.1f} MB", flush=True)
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("model_dir")
    ap.add_argument("--tp", type=int, required=True, help="tensor-parallel size (2, 4 or 8)")
    ap.add_argument("--out", default=None, help="output directory (default <model_dir>/ssb_tp<N>, where the engine looks)")
    a = ap.parse_args(argv)
    if a.tp < 2 or a.tp > 8:
        ap.error("--tp must be 2..8")
    shard(a.model_dir, a.tp, a.out)


if __name__ == "__main__":
    sys.exit(main())
