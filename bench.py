#!/usr/bin/env python
"""bench.py — decode tokens/s + p50 TTFT of the Server-CRD inference hot path (BASELINE.json metric).

A "step" is one synthetic request batch through the hot path: B prompts of 512 token ids
(randint(0,V), seed 1234+i) -> prefill -> 127 further greedy decode steps (128 new tokens).
  value  = decode tokens/s, whole job, device-timed (CUDA events on the engine stream), inputs resident in HBM
           (the decode loop feeds tokens back on the device);
  e2e    = the same metric through the public C-ABI call with HOST buffers (ssb_prefill / ssb_decode take host
           token ids and return host token ids; host<->device copies inside the timed region, wall clock);
  ttft   = p50 wall time of ssb_prefill (host ids in -> first token id on the host).
Weights are seeded synthetic values at the real Llama-2-7B shapes (no checkpoints exist offline).
Inputs are larger than L2 (13.2 GB of weights stream per decode step vs 126 MB L2) so no L2 flush is needed.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROMPT_LEN, NEW_TOKENS = 512, 128
TTFT_SAMPLES = 20  # p50 TTFT over >= 20 requests

WORKLOADS = {
    # "tiny" exists for the CPU contract test of the reference arm (tests/test_bench_contract.py), not for benchmarking
    "tiny": dict(model_type="llama", hidden_size=256, intermediate_size=688, num_hidden_layers=8, num_attention_heads=2,
                 num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, rms_norm_eps=1e-5, rope_theta=10000.0,
                 tie_word_embeddings=False, torch_dtype="bfloat16"),
    "llama2-7b": dict(model_type="llama", hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, vocab_size=32000, max_position_embeddings=4096,
                      rms_norm_eps=1e-5, rope_theta=10000.0, tie_word_embeddings=False, torch_dtype="bfloat16"),
    "llama2-13b": dict(model_type="llama", hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                       num_attention_heads=40, num_key_value_heads=40, vocab_size=32000, max_position_embeddings=4096,
                       rms_norm_eps=1e-5, rope_theta=10000.0, tie_word_embeddings=False, torch_dtype="bfloat16"),
    # BASELINE config 4 (examples/falcon-40b): new_decoder_architecture, 128 query / 8 KV heads of 64, GELU MLP 4h, parallel block
    "falcon-40b": dict(model_type="falcon", hidden_size=8192, num_hidden_layers=60, num_attention_heads=128, num_kv_heads=8,
                       vocab_size=65024, layer_norm_epsilon=1e-5, new_decoder_architecture=True, parallel_attn=True, bias=False,
                       alibi=False, multi_query=True, rope_theta=10000.0, max_position_embeddings=2048,
                       tie_word_embeddings=False, torch_dtype="bfloat16"),
    "tiny-falcon": dict(model_type="falcon", hidden_size=256, num_hidden_layers=6, num_attention_heads=4, num_kv_heads=2,
                        vocab_size=512, layer_norm_epsilon=1e-5, new_decoder_architecture=True, parallel_attn=True, bias=False,
                        alibi=False, multi_query=True, rope_theta=10000.0, max_position_embeddings=512,
                        tie_word_embeddings=False, torch_dtype="bfloat16"),
    "llama2-70b": dict(model_type="llama", hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                       num_attention_heads=64, num_key_value_heads=8, vocab_size=32000, max_position_embeddings=4096,
                       rms_norm_eps=1e-5, rope_theta=10000.0, tie_word_embeddings=False, torch_dtype="bfloat16"),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def synthetic_prompts(vocab, batch, plen):
    import torch

    rows = []
    for i in range(batch):
        g = torch.Generator().manual_seed(1234 + i)
        rows.append(torch.randint(0, vocab, (plen,), generator=g).tolist())
    return rows


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def _hf_cpu_timed(cfg, layers, prompt_len, new_tokens, batch, dtype):
    """Build a `layers`-deep model at cfg's shapes (random init) and time prefill + greedy decode on the CPU."""
    import torch

    falcon = cfg.get("model_type") == "falcon"
    if falcon:
        from transformers import FalconConfig as Config, FalconForCausalLM as Model
        from transformers.models.falcon.modeling_falcon import FalconRotaryEmbedding as Rotary
    else:
        from transformers import LlamaConfig as Config, LlamaForCausalLM as Model
        from transformers.models.llama.modeling_llama import LlamaRotaryEmbedding as Rotary

    keys = {k: v for k, v in cfg.items() if k not in ("model_type", "torch_dtype", "rope_theta")}
    keys["num_hidden_layers"] = layers
    hcfg = Config(**keys, rope_parameters={"rope_type": "default", "rope_theta": cfg.get("rope_theta", 10000.0)},
                  attn_implementation="eager")
    with torch.device("meta"):
        m = Model(hcfg)
    m = m.to_empty(device="cpu").to(dtype)
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() >= 2:
                p.uniform_(-0.0346, 0.0346, generator=g)
            else:
                p.fill_(1.0)
    (m.transformer if falcon else m.model).rotary_emb = Rotary(hcfg)  # to_empty() dropped the inv_freq buffer
    m.eval()
    ids = torch.tensor(synthetic_prompts(cfg["vocab_size"], batch, prompt_len))
    with torch.no_grad():
        t1 = time.time()
        out = m(ids, use_cache=True)
        nxt = out.logits[:, -1].float().argmax(-1)
        ttft = time.time() - t1
        past = out.past_key_values
        t2 = time.time()
        for _ in range(new_tokens - 1):
            out = m(nxt[:, None], past_key_values=past, use_cache=True)
            past = out.past_key_values
            nxt = out.logits[:, -1].float().argmax(-1)
        dec = (time.time() - t2) / max(1, new_tokens - 1)
    del m
    return ttft, dec


def hf_cpu_generate(cfg, prompt_len, new_tokens, batch=1, threads=None, dtype="float32"):
    """The reference's CPU serving path restated: the library the Basaran image wraps — HF transformers
    on host cores, greedy, eager attention, random-init weights at the real shapes.  BOUNDED SAMPLE: the decoder stack
    is timed at 2 and at 4 layers and extrapolated linearly to the config's depth (per-token time = fixed (embedding,
    final norm, lm_head) + L * per-layer), because a full-depth CPU pass of a 7B model costs minutes per request
    (measured on the GPU box's host: 38 s per token in bf16).  dtype float32 = HF's default for CPU serving."""
    import torch

    # cores this process may actually use (cgroup / affinity), capped: torch's CPU GEMV collapses when oversubscribed
    # (measured on the GPU box: 128 threads -> 0.04 tok/s, 25 s per token)
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = threads or max(1, min(avail, 32))
    torch.set_num_threads(threads)
    dt = getattr(torch, dtype)
    t0 = time.time()
    L = cfg["num_hidden_layers"]
    ttft2, dec2 = _hf_cpu_timed(cfg, 2, prompt_len, new_tokens, batch, dt)
    ttft4, dec4 = _hf_cpu_timed(cfg, 4, prompt_len, new_tokens, batch, dt)
    per_layer_dec = max((dec4 - dec2) / 2.0, 1e-9)
    per_layer_ttft = max((ttft4 - ttft2) / 2.0, 1e-9)
    dec = max(dec2 - 2 * per_layer_dec, 0.0) + L * per_layer_dec
    ttft = max(ttft2 - 2 * per_layer_ttft, 0.0) + L * per_layer_ttft
    return {"decode_tok_s": batch / dec, "ttft_s": ttft, "cores": os.cpu_count(), "threads": threads,
            "build_s": time.time() - t0, "dtype": dtype, "per_layer_decode_ms": per_layer_dec * 1e3,
            "sample": f"{batch}x({prompt_len}-token prompt + {new_tokens} greedy tokens) timed at 2 and 4 decoder layers of the "
                      f"{L}-layer model and extrapolated linearly to {L} layers; HF transformers {dtype} eager on {threads} host threads"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU serving path (HF transformers; see hf_cpu_generate) timed on the
    host cores on a BOUNDED sample of the same workload.  Rank 0 only."""
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    plen, ntok = args.ref_prompt_len, args.ref_new_tokens
    # one "step" of this arm = one bounded CPU sample (see hf_cpu_generate); a full-depth CPU request would take minutes,
    # so at most two samples are timed however large --steps is, and --warmup is not needed on the CPU
    vals, ms = [], []
    r = None
    for _ in range(max(1, min(args.steps, 2))):
        t0 = time.time()
        r = hf_cpu_generate(cfg, plen, ntok, batch=args.batch)
        vals.append(r["decode_tok_s"])
        ms.append((time.time() - t0) * 1e3)
    v = statistics.mean(vals)
    sample = r["sample"]
    line = {"impl": "reference", "metric": "decode_tokens_per_sec", "value": v, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": statistics.mean(ms), "samples_timed": len(vals),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} decode, batch {args.batch}, CPU reference path", "sample": sample},
            "ttft_ms_p50": r["ttft_s"] * 1e3,
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": r["threads"], "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--workload", default="llama2-7b", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch32", action="store_true")
    ap.add_argument("--ref-prompt-len", type=int, default=16)
    ap.add_argument("--ref-new-tokens", type=int, default=5)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--engine-params", default="", help="JSON object merged into the engine's params.json (experiments, e.g. "
                    "'{\"tp_mega\": 1}'); echoed in config.engine_params so an overridden run is never mistaken for the default")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    import torch

    from substratus_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the serving path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = WORKLOADS[args.workload]
    peak_gbs, peak_src = load_peaks()
    tmp = tempfile.mkdtemp(prefix="ssb_bench_")
    with open(os.path.join(tmp, "config.json"), "w") as f:
        json.dump(cfg, f)
    params = {"weights": "synthetic", "seed": 0, "max_batch": max(args.batch, 32), "max_seq_len": PROMPT_LEN + NEW_TOKENS + 16,
              "use_pdl": args.pdl, "use_graph": args.graph, "device": local_rank, "tp_size": world, "tp_rank": rank}
    extra_params = json.loads(args.engine_params) if args.engine_params else {}
    params.update(extra_params)
    t_load = time.time()
    eng = Engine(tmp, params)
    if world > 1:
        from substratus_b200 import tp

        tp.connect(eng)  # all-gather the CUDA-IPC handles of the exchange buffers; allreduce then runs over NVLink
    t_load = time.time() - t_load
    info = eng.info
    ctx_mean = PROMPT_LEN + (NEW_TOKENS - 1) / 2.0 + 0.5

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def measure(B, steps, warmup):
        """K timed request batches of size B after W warm-ups; device times are MAX over ranks."""
        prompts = synthetic_prompts(cfg["vocab_size"], B, PROMPT_LEN)

        def one_request():
            sids = [eng.seq_create() for _ in range(B)]
            w0 = time.perf_counter()
            first, _ = eng.prefill(sids, prompts)  # host ids in (H2D inside), first token id back on the host
            w1 = time.perf_counter()
            pre_ms = eng.timing().prefill_ms
            eng.decode(sids, first, NEW_TOKENS - 1)  # 127 device-resident steps, ids back on the host (D2H inside)
            w2 = time.perf_counter()
            dec_ms = eng.timing().decode_ms
            for s in sids:
                eng.seq_free(s)
            return dict(ttft_wall_ms=(w1 - w0) * 1e3, prefill_dev_ms=pre_ms, decode_dev_ms=dec_ms,
                        decode_wall_ms=(w2 - w1) * 1e3, total_wall_ms=(w2 - w0) * 1e3)

        for _ in range(warmup):
            one_request()
        eng.timing_reset()
        sampler = ClockSampler(local_rank)
        sampler.start()
        barrier()
        t0 = time.perf_counter()
        recs = [one_request() for _ in range(steps)]
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        if world > 1:
            from substratus_b200 import tp

            wall = tp.max_over_ranks(wall)
            for r in recs:
                for k in r:
                    r[k] = tp.max_over_ranks(r[k])
        tm = eng.timing()
        # SURVEY.md 8d quotes p50 TTFT over >= 20 requests after the warm-ups: top the K timed requests up with prefill-only
        # requests.  They run after the timed region and after the counters were read, so value / e2e / launches / bytes
        # are those of the K steps alone.
        ttfts = [r["ttft_wall_ms"] for r in recs]
        pre_devs = [r["prefill_dev_ms"] for r in recs]
        for _ in range(max(0, TTFT_SAMPLES - steps)):
            sids = [eng.seq_create() for _ in range(B)]
            w0 = time.perf_counter()
            eng.prefill(sids, prompts)
            t_ms = (time.perf_counter() - w0) * 1e3
            d_ms = eng.timing().prefill_ms
            for sid in sids:
                eng.seq_free(sid)
            if world > 1:
                t_ms, d_ms = tp.max_over_ranks(t_ms), tp.max_over_ranks(d_ms)
            ttfts.append(t_ms)
            pre_devs.append(d_ms)
        ntok = B * (NEW_TOKENS - 1)
        dec_dev_s = sum(r["decode_dev_ms"] for r in recs) / 1e3
        dec_wall_s = sum(r["decode_wall_ms"] for r in recs) / 1e3
        bytes_step = info.weight_bytes_per_step + B * ctx_mean * info.kv_bytes_per_token
        step_ms = dec_dev_s * 1e3 / (steps * (NEW_TOKENS - 1))
        step_gbs = bytes_step / (step_ms * 1e-3) / 1e9
        return dict(value=steps * ntok / dec_dev_s, e2e=steps * ntok / dec_wall_s, wall=wall, clocks=clocks, tm=tm,
                    ttft_ms_p50=statistics.median(ttfts), prefill_device_ms_p50=statistics.median(pre_devs), ttft_samples=len(ttfts),
                    request_tok_s=steps * B * NEW_TOKENS / (sum(r["total_wall_ms"] for r in recs) / 1e3),
                    step_ms=step_ms, bytes_step=bytes_step, step_gbs=step_gbs)

    B = args.batch
    m = measure(B, args.steps, args.warmup)
    launches = int(m["tm"].kernel_launches)
    h2d, d2h = m["tm"].h2d_bytes // args.steps, m["tm"].d2h_bytes // args.steps

    # Dominant kernel.  At batch <= 4 on one GPU the whole decode step IS one kernel (decode_mega_kernel): its
    # achieved bandwidth is the step's algorithmic bytes over its CUDA-event duration.  Otherwise the dominant kernel is
    # the gate/up projection (44% of the weight bytes), timed live per launch with ssb_bench_kernel.
    kern = {}
    for k in ("gate_up", "qkv", "down", "o", "lm_head", "attn"):
        ms, by = eng.bench_kernel(k, rows=B, ctx=int(ctx_mean), iters=64)
        kern[k] = {"ms": ms, "bytes": by, "gbs": by / (ms * 1e-3) / 1e9}
    # the persistent kernel serves Llama-family decode at batch <= 4 on one GPU (and under TP only with the experimental
    # "tp_mega" engine param); Falcon always runs the multi-kernel path
    mega = B <= 4 and cfg.get("model_type") != "falcon" and (world == 1 or bool(extra_params.get("tp_mega")))
    if mega:
        roofline = {"bound": "hbm", "kernel": "decode_mega_kernel (persistent single-kernel decode step: all projections, attention, pick)",
                    "achieved": m["step_gbs"], "peak": peak_gbs, "unit": "GB/s", "frac": m["step_gbs"] / peak_gbs,
                    "traffic": MEGA_TRAFFIC_BYTES.get((args.workload, B)), "algorithmic_bytes_per_launch": m["bytes_step"],
                    "ms_per_launch": m["step_ms"]}
    else:
        dom = kern["gate_up"]
        roofline = {"bound": "hbm", "kernel": "gate/up projection (gemv_kernel for <= 4 rows, tc_gemm_sk_kernel tcgen05 stream-K above)",
                    "achieved": dom["gbs"], "peak": peak_gbs, "unit": "GB/s", "frac": dom["gbs"] / peak_gbs, "traffic": None,
                    "algorithmic_bytes_per_launch": dom["bytes"], "ms_per_launch": dom["ms"]}
    roofline.update({"peak_source": peak_src, "per_kernel_class_gbs": {k: round(v["gbs"], 1) for k, v in kern.items()},
                     "decode_step": {"bytes": m["bytes_step"], "ms": m["step_ms"], "achieved": m["step_gbs"],
                                     "frac": m["step_gbs"] / peak_gbs, "roofline_tok_s": B * peak_gbs * 1e9 / m["bytes_step"]}})

    line = {
        "metric": "decode_tokens_per_sec", "value": m["value"], "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["wall"] * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload} bf16 decode, batch {B}, {PROMPT_LEN}-token prompt + {NEW_TOKENS} new tokens, {world}xB200",
                   "batch": B, "prompt_len": PROMPT_LEN, "new_tokens": NEW_TOKENS, "parallelism": f"tp{world}",
                   "allreduce": "one-shot over NVLink peer memory (CUDA IPC)" if world > 1 else "none",
                   "l2": "inputs larger than L2 (weights streamed per decode step >> 126 MB)", "pdl": args.pdl, "graph": args.graph},
        "ttft_ms_p50": m["ttft_ms_p50"], "ttft_samples": m["ttft_samples"], "prefill_device_ms_p50": m["prefill_device_ms_p50"],
        "decode_ms_per_token": m["step_ms"],
        "e2e": {"value": m["e2e"], "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "request_tokens_per_sec": m["request_tok_s"]},
        "gpu_launches": launches, "clocks": m["clocks"], "roofline": roofline, "load_s": t_load,
        "hbm_gb": info.hbm_bytes_allocated / 1e9,
    }
    if extra_params:
        line["config"]["engine_params"] = extra_params
    from substratus_b200.engine import load_library

    line["engine"] = load_library().ssb_version().decode()  # names the SSB_LIB_VARIANT build, if one was selected
    if B == 1 and not args.no_batch32:  # the metric is quoted at batch 1 AND 32: same engine, second measurement
        try:
            m32 = measure(32, max(2, args.steps // 2), 2)
            line["batch32"] = {"value": m32["value"], "unit": "tokens/s", "e2e": m32["e2e"], "ttft_ms_p50": m32["ttft_ms_p50"],
                               "decode_ms_per_step": m32["step_ms"], "bytes_per_step": m32["bytes_step"],
                               "hbm_roofline_frac": m32["step_gbs"] / peak_gbs}
        except Exception as ex:  # the headline line above must still be printed; the failure is reported, not hidden
            line["batch32"] = {"error": repr(ex)}
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not args.no_cpu_baseline:
        try:
            r = hf_cpu_generate(cfg, args.ref_prompt_len, args.ref_new_tokens, batch=1)
            line["cpu_baseline"] = {"value": r["decode_tok_s"], "unit": "tokens/s", "cores": r["threads"], "kind": "reference",
                                    "sample": r["sample"] + " (the library the reference's Basaran image wraps)",
                                    "ttft_ms": r["ttft_s"] * 1e3, "sample_wall_s": r["build_s"]}
        except Exception as ex:  # host RAM etc.
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"failed: {type(ex).__name__}: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of one decode_mega_kernel launch (ncu --set full, profiles/)
MEGA_TRAFFIC_BYTES = {("llama2-7b", 1): 13.494e9}


if __name__ == "__main__":
    main()
