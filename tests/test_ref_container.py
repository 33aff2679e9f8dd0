"""BASELINE config 1 (plumbing, no GPU): the reference's CPU serving container restated (oracle/ref_server.py, HF
transformers behind the container contract) serves an OPT-125m-architecture model; one greedy request through
`/v1/completions` returns exactly HF `generate(do_sample=False)`'s ids.  Same client code as tests/test_serve_host.py."""
import json
import threading
import time
import urllib.request

import torch

from oracle import ref_server

OPT_TINY = dict(model_type="opt", architectures=["OPTForCausalLM"], hidden_size=96, ffn_dim=384, num_hidden_layers=2,
                num_attention_heads=4, vocab_size=512, max_position_embeddings=128, word_embed_proj_dim=96,
                do_layer_norm_before=True, activation_function="relu", pad_token_id=1, bos_token_id=2, eos_token_id=2)


def _req(url, data=None):
    r = urllib.request.Request(url, data=json.dumps(data).encode() if data is not None else None,
                               headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(r, timeout=60) as f:
        return f.status, json.loads(f.read())


def test_opt_cpu_reference_container(tmp_path):
    (tmp_path / "config.json").write_text(json.dumps(OPT_TINY))
    srv = ref_server.serve(str(tmp_path), 0, "float32")
    port = srv.server_address[1]
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        for _ in range(300):
            try:
                if _req(f"http://127.0.0.1:{port}/")[0] == 200:
                    break
            except Exception:
                pass
            time.sleep(0.1)
        assert ref_server.State.ready
        prompt = torch.randint(3, 512, (1, 12), generator=torch.Generator().manual_seed(1234))
        st, r = _req(f"http://127.0.0.1:{port}/v1/completions", {"prompt": prompt[0].tolist(), "max_tokens": 3})
        assert st == 200 and r["usage"] == {"prompt_tokens": 12, "completion_tokens": 3}
        with torch.no_grad():
            want = ref_server.State.model.generate(prompt, max_new_tokens=3, do_sample=False, pad_token_id=1)[0, 12:]
        assert r["choices"][0]["tokens"] == want.tolist()
    finally:
        srv.shutdown()


def test_opt_125m_real_shapes_through_the_cpu_container(tmp_path):
    """BASELINE config 1 at its REAL shapes (examples/facebook-opt-125m/base-server.yaml:1-8: no resources => CPU): h768 L12
    H12 ffn3072 V50272, random init (no checkpoints offline), one greedy request through `/v1/completions` equals HF
    generate() id for id, and the response carries the timing fields bench.py's opt_125m_cpu_container sub-object reports."""
    import sys, os

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench

    (tmp_path / "config.json").write_text(json.dumps(bench.WORKLOADS["opt-125m"]))
    ref_server.State.ready = False
    srv = ref_server.serve(str(tmp_path), 0, "float32")
    port = srv.server_address[1]
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        for _ in range(1200):
            try:
                if _req(f"http://127.0.0.1:{port}/")[0] == 200:
                    break
            except Exception:
                pass
            time.sleep(0.1)
        assert ref_server.State.ready and sum(p.numel() for p in ref_server.State.model.parameters()) > 120e6
        prompt = torch.tensor([bench.synthetic_prompts(50272, 1, 16)[0]])
        st, r = _req(f"http://127.0.0.1:{port}/v1/completions", {"prompt": prompt[0].tolist(), "max_tokens": 8})
        assert st == 200 and r["usage"] == {"prompt_tokens": 16, "completion_tokens": 8}
        assert r["decode_tokens_per_sec"] > 0 and r["ttft_ms"] > 0
        with torch.no_grad():
            want = ref_server.State.model.generate(prompt, max_new_tokens=8, do_sample=False, pad_token_id=1)[0, 16:]
        assert r["choices"][0]["tokens"] == want.tolist()
    finally:
        srv.shutdown()
