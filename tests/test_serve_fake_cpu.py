"""The serve host's HTTP surface on a CPU-only box: host/serve.cpp built against tests/fake_ssb/fake_ssb.cpp (a GPU-free
stand-in for the C ABI; see its header) instead of libsubstratus_b200.so.  Covers what the reference's container contract
and its one request shape need from the HOST (docs/container-contract.md:50-55, server_controller.go:156-172,
test/system.sh:73-78): readiness 503 -> 200, request validation, /generate and /v1/completions, "stream": true (SSE),
opt-in continuous batching, the in-container tensor-parallel rank threads, error propagation.  The same requests run
against the real engine in tests/test_serve_host.py (-m gpu)."""
import concurrent.futures as cf
import json
import os
import socket
import subprocess
import time
import urllib.error
import urllib.request

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = (1 << 64) - 1


def fold(s, tok):
    z = (s + 0x9E3779B97F4A7C15 + (tok & 0xFFFFFFFF)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def fake_generate(prompt, n, vocab):
    s = 0
    for t in prompt:
        s = fold(s, t)
    out = [s % vocab]
    while len(out) < n:
        s = fold(s, out[-1])
        out.append(s % vocab)
    return out


@pytest.fixture(scope="module")
def serve_fake(tmp_path_factory):
    if os.environ.get("SERVE_FAKE_EXE"):  # e.g. an -fsanitize=address,undefined build of the same sources
        return os.environ["SERVE_FAKE_EXE"]
    exe = str(tmp_path_factory.mktemp("fake") / "serve_fake")
    src = [os.path.join(ROOT, p) for p in ("host/serve.cpp", "tests/fake_ssb/fake_ssb.cpp", "substratus_b200/csrc/tokenizer.cpp",
                                           "substratus_b200/csrc/loader.cpp", "substratus_b200/csrc/torch_zip.cpp")]
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-Wall", "-o", exe] + src, check=True, cwd=os.path.join(ROOT, "host"))
    return exe


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _req(port, path, body=None, timeout=20):
    data = None if body is None else (body if isinstance(body, bytes) else json.dumps(body).encode())
    try:
        with urllib.request.urlopen(urllib.request.Request(f"http://127.0.0.1:{port}{path}", data=data), timeout=timeout) as r:
            raw = r.read()
            return r.status, (json.loads(raw) if r.headers.get_content_type() == "application/json" else raw.decode())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read() or b"{}")


def _sse(port, path, body, timeout=30):
    """-> list of decoded `data:` payloads (the last one is the string "[DONE]")"""
    with urllib.request.urlopen(urllib.request.Request(f"http://127.0.0.1:{port}{path}", data=json.dumps(body).encode()), timeout=timeout) as r:
        assert r.status == 200 and r.headers.get_content_type() == "text/event-stream"
        raw = r.read()
    events = []
    for block in raw.split(b"\n\n"):
        if block:
            block = block.decode("utf-8")  # every event on its own must be valid UTF-8 (no character cut between events)
            assert block.startswith("data: "), block
            events.append(block[6:] if block[6:] == "[DONE]" else json.loads(block[6:]))
    return events


class Server:
    def __init__(self, exe, tmp_path, params, tokenizer=None):
        self.port = _free_port()
        d = tmp_path / f"model{self.port}"
        d.mkdir()
        if tokenizer:
            tokenizer(str(d / "tokenizer.json"))
        pf = d / "params.json"
        pf.write_text(json.dumps(params))
        env = {k: v for k, v in os.environ.items() if not k.startswith("PARAM_")}
        env.update(PORT=str(self.port), PARAMS_FILE=str(pf), MODEL_DIR=str(d))
        self.p = subprocess.Popen([exe], env=env, stderr=subprocess.PIPE, text=True)

    def wait_ready(self, timeout=20):
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if _req(self.port, "/")[0] == 200:
                    return
            except (urllib.error.URLError, ConnectionError):
                pass
            assert self.p.poll() is None, self.p.stderr.read()
            time.sleep(0.02)
        raise AssertionError("server did not become ready")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.p.kill()
        self.p.wait()


def test_readiness_503_until_loaded_then_200(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"fake_load_ms": 1500}) as s:
        t0 = time.time()
        saw_503 = False
        while time.time() - t0 < 10:
            try:
                code, body = _req(s.port, "/")
            except (urllib.error.URLError, ConnectionError):
                time.sleep(0.02)
                continue
            if code == 503:
                saw_503 = True
                assert body == {"status": "loading"}
                assert _req(s.port, "/generate", {"tokens": [1], "max_new_tokens": 2})[0] == 503  # not ready: refuse work
                time.sleep(0.1)
            else:
                assert code == 200 and body["status"] == "ready"
                break
        assert saw_503, "the probe must see 503 (not a refused connection) while the model loads"
        assert _req(s.port, "/healthz")[0] == 200


def test_load_failure_exits_nonzero(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"fake_load_error": 1}) as s:
        assert s.p.wait(timeout=10) == 1  # Deployment restarts the pod (server_controller.go:280-296)
        assert "engine create failed" in s.p.stderr.read()


def test_generate_and_completions_ids_and_validation(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"fake_vocab": 321, "max_seq_len": 64}) as s:
        s.wait_ready()
        prompt = [5, 17, 300, 2]
        want = fake_generate(prompt, 9, 321)
        code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 9})
        assert code == 200 and r["tokens"] == want and r["text"] == ""
        code, r = _req(s.port, "/v1/completions", {"prompt": prompt, "max_tokens": 9})
        assert code == 200 and r["choices"][0]["tokens"] == want and r["choices"][0]["finish_reason"] == "length"
        assert r["usage"] == {"prompt_tokens": 4, "completion_tokens": 9} and r["object"] == "text_completion"
        assert _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 1})[1]["tokens"] == want[:1]
        # validation: 400 with an error string, the engine is never called
        for bad in ({"tokens": [], "max_new_tokens": 2}, {"tokens": [321], "max_new_tokens": 2}, {"tokens": [1.5]}, {"tokens": [-1]},
                    {"tokens": "x"}, {"tokens": [1] * 60, "max_new_tokens": 10}, {"tokens": [1], "max_new_tokens": 0}, {}):
            code, r = _req(s.port, "/generate", bad)
            assert code == 400 and r["error"], bad
        assert _req(s.port, "/generate", b"{not json")[0] == 400
        code, r = _req(s.port, "/v1/completions", {"prompt": "text", "max_tokens": 2})
        assert code == 400 and "tokenizer.json" in r["error"]  # no tokenizer in the model dir: ids only, said clearly
        assert _req(s.port, "/nope")[0] == 404
        code, m = _req(s.port, "/metrics")
        assert code == 200 and "ssb_requests_total" in m and "ssb_generated_tokens_total 19" in m


@pytest.mark.parametrize("chunk", [1, 4])
def test_stream_ids_match_plain_request(serve_fake, tmp_path, chunk):
    with Server(serve_fake, tmp_path, {"fake_vocab": 500, "stream_chunk": chunk}) as s:
        s.wait_ready()
        prompt = [9, 8, 7]
        want = fake_generate(prompt, 11, 500)
        ev = _sse(s.port, "/v1/completions", {"prompt": prompt, "max_tokens": 11, "stream": True})
        assert ev[-1] == "[DONE]"
        body, last = ev[:-2], ev[-2]
        assert [len(e["choices"][0]["tokens"]) for e in body] == [1] + [chunk] * (10 // chunk) + ([10 % chunk] if 10 % chunk else [])
        assert sum((e["choices"][0]["tokens"] for e in body), []) == want
        assert all(e["choices"][0]["finish_reason"] is None for e in body)
        assert last["choices"][0]["finish_reason"] == "length" and last["usage"] == {"prompt_tokens": 3, "completion_tokens": 11}
        ev = _sse(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 11, "stream": 1})
        assert sum((e["tokens"] for e in ev[:-1]), []) == want and ev[-2]["done"] is True and ev[0]["done"] is False


def test_stream_text_is_the_plain_text_cut_at_utf8_boundaries(serve_fake, tmp_path):
    from test_tokenizer import _llama_like

    vocab = {}

    def make(path):
        vocab["n"] = _llama_like(path).get_vocab_size()

    # ids are pseudo-random over the whole vocabulary, so byte-fallback pieces (<0xNN>) and multi-byte characters split
    # across events are common: every event must still be valid UTF-8 and the pieces must add up to the plain answer
    with Server(serve_fake, tmp_path, {"fake_vocab": 700}, tokenizer=make) as s:
        s.wait_ready()
        assert vocab["n"] <= 700
    with Server(serve_fake, tmp_path, {"fake_vocab": vocab["n"]}, tokenizer=make) as s:
        s.wait_ready()
        for text in ["Hello world", "naïve 你好 🙂", "the quick brown fox"]:
            code, plain = _req(s.port, "/v1/completions", {"prompt": text, "max_tokens": 60})
            assert code == 200, plain
            ev = _sse(s.port, "/v1/completions", {"prompt": text, "max_tokens": 60, "stream": True})
            ids = sum((e["choices"][0]["tokens"] for e in ev[:-1]), [])
            pieces = [e["choices"][0]["text"] for e in ev[:-1]]
            assert ids == plain["choices"][0]["tokens"]
            assert "".join(pieces) == plain["choices"][0]["text"]
            assert all("�" not in p for p in pieces[:-1]) or "�" in plain["choices"][0]["text"]


def test_decode_failure_is_reported(serve_fake, tmp_path):
    # a per-request failure (here: the engine says ENOMEM, e.g. the KV pool is exhausted) is answered and the pod lives on
    with Server(serve_fake, tmp_path, {"fake_fail_after": 2, "fake_fail_code": -4}) as s:
        s.wait_ready()
        ev = _sse(s.port, "/generate", {"tokens": [1, 2], "max_new_tokens": 8, "stream": True})
        assert ev[-1] == "[DONE]" and "decode failure" in ev[-2]["error"] and len(ev) == 3 + 2  # first token + 2 good chunks
        code, r = _req(s.port, "/generate", {"tokens": [1, 2], "max_new_tokens": 8})
        assert code == 500 and "decode failure" in r["error"]
        assert "ssb_errors_total 2" in _req(s.port, "/metrics")[1]
        assert _req(s.port, "/")[0] == 200 and s.p.poll() is None


@pytest.mark.parametrize("params", [{}, {"tp_size": 2}, {"batching": 1}])
def test_device_failure_answers_then_exits_nonzero(serve_fake, tmp_path, params):
    """a sticky device error (SSB_ECUDA) must not leave a ready-looking pod behind: the request is answered with
    the error, then the process exits non-zero so the Deployment restarts it (server_controller.go:280-296).  Same with
    the in-container tensor-parallel rank threads, where a healthy rank would otherwise wait for the dead one."""
    with Server(serve_fake, tmp_path, dict(params, fake_fail_after=0)) as s:
        s.wait_ready()
        try:
            code, r = _req(s.port, "/generate", {"tokens": [1, 2], "max_new_tokens": 8, "stop_at_eos": False, "temperature": 0})
            assert code == 500 and "decode failure" in r["error"]
        except (urllib.error.URLError, ConnectionError, json.JSONDecodeError):
            pass  # TP: a rank thread may take the process down before the response is written
        assert s.p.wait(timeout=10) == 1
        assert "fatal engine error" in s.p.stderr.read()


def test_slow_and_oversized_clients_are_bounded(serve_fake, tmp_path):
    """body capped from max_seq_len (413), a header that merely CONTAINS 'content-length:' is not the length,
    params.json eos ids may be strings (IntOrString)."""
    with Server(serve_fake, tmp_path, {"fake_vocab": 100, "max_seq_len": 64, "eos_token_id": "7", "stop_at_eos": 0}) as s:
        s.wait_ready()
        c = socket.create_connection(("127.0.0.1", s.port))  # announce 3 MiB, send nothing: refused from the header alone
        c.sendall(b"POST /generate HTTP/1.1\r\nHost: x\r\nContent-Length: 3145728\r\n\r\n")
        assert c.recv(65536).startswith(b"HTTP/1.1 413")
        c.close()
        body = json.dumps({"tokens": [1, 2, 3], "max_new_tokens": 4}).encode()
        c = socket.create_connection(("127.0.0.1", s.port))
        c.sendall(b"POST /generate HTTP/1.1\r\nHost: x\r\nX-Content-Length: 3\r\nContent-Length: " + str(len(body)).encode() + b"\r\n\r\n" + body)
        resp = b""
        while True:
            chunk = c.recv(65536)
            if not chunk:
                break
            resp += chunk
        c.close()
        assert resp.startswith(b"HTTP/1.1 200") and json.loads(resp.split(b"\r\n\r\n", 1)[1])["tokens"] == fake_generate([1, 2, 3], 4, 100)
        # a string-valued eos id in params.json is honoured: stop_at_eos requests are accepted (400 before the fix)
        code, r = _req(s.port, "/generate", {"tokens": [1, 2, 3], "max_new_tokens": 20, "stop_at_eos": True})
        assert code == 200 and 7 not in r["tokens"]


@pytest.mark.parametrize("stream", [False, True])
def test_continuous_batching_keeps_every_request_its_own_ids(serve_fake, tmp_path, stream):
    with Server(serve_fake, tmp_path, {"fake_vocab": 999, "batching": 1, "batch_tick": 4, "max_batch": 4, "fake_step_us": 1500}) as s:
        s.wait_ready()

        def one(i):
            prompt = [i + 1, 2 * i + 3, 7]
            n = 5 + 3 * (i % 5)
            if stream:
                ev = _sse(s.port, "/generate", {"tokens": prompt, "max_new_tokens": n, "stream": True})
                got = sum((e["tokens"] for e in ev[:-1]), [])
            else:
                code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": n})
                assert code == 200, r
                got = r["tokens"]
            return got == fake_generate(prompt, n, 999)

        with cf.ThreadPoolExecutor(10) as ex:  # more clients than slots: the rest wait their turn
            assert all(ex.map(one, range(10)))


def test_batching_admits_by_kv_blocks_instead_of_failing_the_batch(serve_fake, tmp_path):
    """ADVICE r1 (scheduler.h): the pool holds 8 blocks of 16 tokens; each request needs 3 (8 + 40 tokens), so two run at a
    time and the others queue — nobody gets a 500 for somebody else's appetite.  A request larger than the whole pool is
    the one that fails, alone."""
    with Server(serve_fake, tmp_path, {"fake_vocab": 999, "batching": 1, "batch_tick": 4, "max_batch": 8, "fake_kv_blocks": 8,
                                      "fake_step_us": 500}) as s:
        s.wait_ready()

        def one(i):
            prompt = [i + 1] * 8
            n = 300 if i == 3 else 40
            code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": n})
            if i == 3:
                return code == 500 and "exhausted" in r["error"]
            return code == 200 and r["tokens"] == fake_generate(prompt, n, 999)

        with cf.ThreadPoolExecutor(8) as ex:
            assert all(ex.map(one, range(8)))
        assert _req(s.port, "/")[0] == 200 and s.p.poll() is None


def test_tensor_parallel_ranks_in_one_container(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"tp_size": 4, "fake_vocab": 100}) as s:
        s.wait_ready()
        want = fake_generate([3, 4], 6, 100)
        assert _req(s.port, "/generate", {"tokens": [3, 4], "max_new_tokens": 6})[1]["tokens"] == want
        ev = _sse(s.port, "/generate", {"tokens": [3, 4], "max_new_tokens": 6, "stream": True})
        assert sum((e["tokens"] for e in ev[:-1]), []) == want


def test_stream_client_disconnect_frees_the_engine(serve_fake, tmp_path):
    # 4000 steps x 2 ms = 8 s if the server kept generating for a client that left
    with Server(serve_fake, tmp_path, {"fake_step_us": 2000, "max_seq_len": 8192}) as s:
        s.wait_ready()
        c = socket.create_connection(("127.0.0.1", s.port))
        body = json.dumps({"tokens": [1], "max_new_tokens": 4000, "stream": True}).encode()
        c.sendall(b"POST /generate HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % len(body) + body)
        got = b""
        while b"data: " not in got:
            got += c.recv(4096)
        c.close()
        t0 = time.time()
        code, r = _req(s.port, "/generate", {"tokens": [2], "max_new_tokens": 3})
        assert code == 200 and r["tokens"] == fake_generate([2], 3, 1000)
        assert time.time() - t0 < 4.0, "the abandoned stream kept the engine busy"


def _prompt_with_eos_at(vocab, eos, lo, hi, n):
    """a prompt whose greedy continuation first emits `eos` at an index in [lo, hi)"""
    for a in range(vocab * vocab):
        prompt = [a // vocab, a % vocab, 3]
        want = fake_generate(prompt, n, vocab)
        if eos in want and lo <= want.index(eos) < hi:
            return prompt, want
    raise AssertionError("no such prompt")


@pytest.mark.parametrize("mode", ["plain", "batching", "tp"])
def test_stop_at_eos(serve_fake, tmp_path, mode):
    extra = {"plain": {}, "batching": {"batching": 1, "batch_tick": 5}, "tp": {"tp_size": 2}}[mode]
    with Server(serve_fake, tmp_path, dict({"fake_vocab": 40, "eos_token_id": [7, 39], "eos_check_every": 4}, **extra)) as s:
        s.wait_ready()
        for lo, hi in ((1, 4), (4, 12), (17, 30), (0, 1)):
            prompt, want = _prompt_with_eos_at(40, 7, lo, hi, 32)
            k = min(want.index(7), want.index(39) if 39 in want else 99)
            code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 32})
            assert code == 200 and r["tokens"] == want and r["finish_reason"] == "length"  # default: the benchmark request never stops
            code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 32, "stop_at_eos": True})
            assert code == 200 and r["tokens"] == want[:k] and r["finish_reason"] == "stop", (lo, r, want)
            code, r = _req(s.port, "/v1/completions", {"prompt": prompt, "max_tokens": 32, "stop_at_eos": True})
            assert r["choices"][0]["tokens"] == want[:k] and r["choices"][0]["finish_reason"] == "stop"
            assert r["usage"]["completion_tokens"] == k
            ev = _sse(s.port, "/v1/completions", {"prompt": prompt, "max_tokens": 32, "stop_at_eos": True, "stream": True})
            assert sum((e["choices"][0]["tokens"] for e in ev[:-1]), []) == want[:k]
            assert ev[-2]["choices"][0]["finish_reason"] == "stop" and ev[-1] == "[DONE]"
        # a request that never meets EOS inside max_tokens ends with "length"
        prompt, want = _prompt_with_eos_at(40, 7, 20, 32, 32)
        k = min(want.index(7), want.index(39) if 39 in want else 99)
        code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": k, "stop_at_eos": True})
        assert r["tokens"] == want[:k] and r["finish_reason"] == "length"


def test_eos_id_comes_from_the_model_dir_and_can_be_the_default(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"fake_vocab": 40}) as s:  # no eos anywhere: asking for it is a client error
        s.wait_ready()
        code, r = _req(s.port, "/generate", {"tokens": [1], "max_new_tokens": 4, "stop_at_eos": True})
        assert code == 400 and "eos_token_id" in r["error"]

    def with_config(path):
        with open(os.path.join(os.path.dirname(path), "config.json"), "w") as f:
            json.dump({"eos_token_id": 7}, f)

    with Server(serve_fake, tmp_path, {"fake_vocab": 40, "stop_at_eos": 1}, tokenizer=lambda p: with_config(p)) as s:
        s.wait_ready()
        prompt, want = _prompt_with_eos_at(40, 7, 3, 20, 32)
        code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 32})
        assert r["tokens"] == want[:want.index(7)] and r["finish_reason"] == "stop"
        code, r = _req(s.port, "/generate", {"tokens": prompt, "max_new_tokens": 32, "stop_at_eos": False})
        assert r["tokens"] == want


def test_sub_infer_client_and_load_generator(serve_fake, tmp_path, capsys):
    """tools/sub_infer.py (the `sub infer` the reference only stubs, internal/cli/infer.go) against the host: single
    request, and the SURVEY §8d synthetic load through HTTP with concurrent clients sharing decode calls."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("sub_infer", os.path.join(ROOT, "tools", "sub_infer.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    with Server(serve_fake, tmp_path, {"fake_vocab": 777, "batching": 1, "max_batch": 8, "fake_step_us": 300, "fake_load_ms": 400}) as s:
        url = f"http://127.0.0.1:{s.port}"
        assert cli.main(["--url", url, "--wait", "20", "--ids", "5,6,7", "--max-tokens", "6"]) == 0
        assert capsys.readouterr().out.split() == [str(t) for t in fake_generate([5, 6, 7], 6, 777)]
        line = cli.bench(url, 12, 6, 777, 64, 20, 60)
        capsys.readouterr()
        assert line["failed"] == 0 and line["generated_tokens"] == 12 * 20 and line["ttft_ms_p50"] > 0 and line["tokens_per_sec"] > 0
        r = cli.stream_completion(url, cli.synthetic_prompt(777, 3, 64), 20)
        assert r["tokens"] == fake_generate(cli.synthetic_prompt(777, 3, 64), 20, 777) and r["finish_reason"] == "length"
        assert cli.main(["--url", url, "text prompt without a tokenizer"]) == 1  # the server's 400 is reported, not swallowed


@pytest.mark.parametrize("mode", ["plain", "tp"])
def test_sampling_is_opt_in_seeded_and_filters_like_hf(serve_fake, tmp_path, mode):
    """temperature > 0: the host draws each token from the step's logits (host/sampler.h: temperature -> top_k -> top_p,
    HF order).  The fake model puts ~0.3 of the mass on its greedy id at T = 1 and spreads the rest."""
    V = 50
    with Server(serve_fake, tmp_path, dict({"fake_vocab": V}, **({"tp_size": 2} if mode == "tp" else {}))) as s:
        s.wait_ready()
        prompt, n = [4, 5, 6], 40
        greedy = fake_generate(prompt, n, V)

        def gen(**kw):
            code, r = _req(s.port, "/generate", dict({"tokens": prompt, "max_new_tokens": n}, **kw))
            assert code == 200, r
            return r["tokens"]

        assert gen() == greedy and gen(temperature=0) == greedy
        assert gen(temperature=1.0, top_k=1, seed=1) == greedy           # top_k = 1 leaves only the argmax
        assert gen(temperature=1.0, top_p=0.05, seed=2) == greedy        # a nucleus smaller than the top token's mass
        assert gen(temperature=0.05, seed=3) == greedy                   # cold: e^(1/0.05) ratio to the runner-up
        a, b, c = gen(temperature=1.0, seed=7), gen(temperature=1.0, seed=7), gen(temperature=1.0, seed=8)
        assert a == b and a != c and a != greedy and len(a) == n          # seeded, and really sampling
        assert all(0 <= t < V for t in a)
        # the sampled ids feed back into the model: after the first non-greedy draw the fake's continuation follows it
        k = next(i for i, (x, y) in enumerate(zip(a, greedy)) if x != y)
        assert a[:k] == greedy[:k]
        ev = _sse(s.port, "/v1/completions", {"prompt": prompt, "max_tokens": n, "stream": True, "temperature": 1.0, "seed": 7})
        assert sum((e["choices"][0]["tokens"] for e in ev[:-1]), []) == a  # same seed, streamed
        hot = gen(temperature=50.0, seed=5)                               # near-uniform: the greedy id is rarely drawn
        assert sum(x == y for x, y in zip(hot, fake_generate(prompt, n, V))) < n // 2
        for bad in ({"temperature": -1}, {"temperature": 1, "top_p": 0}, {"temperature": 1, "top_p": 1.5}, {"temperature": 1, "top_k": -2}):
            assert _req(s.port, "/generate", dict({"tokens": prompt, "max_new_tokens": 4}, **bad))[0] == 400


def test_sampling_is_refused_under_batching(serve_fake, tmp_path):
    with Server(serve_fake, tmp_path, {"fake_vocab": 50, "batching": 1}) as s:
        s.wait_ready()
        code, r = _req(s.port, "/generate", {"tokens": [1], "max_new_tokens": 4, "temperature": 0.7})
        assert code == 400 and "batching" in r["error"]
        assert _req(s.port, "/generate", {"tokens": [1], "max_new_tokens": 4})[0] == 200


def test_sigterm_drains_in_flight_requests_then_exits_zero(serve_fake, tmp_path):
    """Pod deletion: SIGTERM -> no new connections, the request in flight completes, exit code 0 (no restart-loop noise)."""
    import signal
    import threading

    with Server(serve_fake, tmp_path, {"fake_step_us": 3000}) as s:
        s.wait_ready()
        out = {}

        def long_request():
            out["r"] = _req(s.port, "/generate", {"tokens": [1, 2], "max_new_tokens": 600}, timeout=60)  # ~1.8 s of decode

        t = threading.Thread(target=long_request)
        t.start()
        deadline = time.time() + 10
        while "ssb_requests_total 1" not in _req(s.port, "/metrics")[1] and time.time() < deadline:  # the request is in the engine
            time.sleep(0.02)
        s.p.send_signal(signal.SIGTERM)
        t.join(timeout=30)
        assert out["r"][0] == 200 and out["r"][1]["tokens"] == fake_generate([1, 2], 600, 1000)
        assert s.p.wait(timeout=30) == 0
        assert "draining" in s.p.stderr.read()
        with pytest.raises((urllib.error.URLError, ConnectionError)):
            _req(s.port, "/")


def test_malformed_requests_do_not_take_the_server_down(serve_fake, tmp_path):
    """Garbage on the wire (the Service is reachable by anything in the cluster) gets an error or a closed connection;
    the server keeps serving.  Includes a body nested 200 000 levels deep (the JSON parser has a nesting guard)."""
    import random

    with Server(serve_fake, tmp_path, {"fake_vocab": 100}) as s:
        s.wait_ready()
        rng = random.Random(9)
        good = json.dumps({"tokens": [1, 2], "max_new_tokens": 3}).encode()
        payloads = [
            b"", b"\r\n\r\n", b"GET", b"GET / HTTP/1.1", b"POST /generate HTTP/1.1\r\nContent-Length: -5\r\n\r\n{}",
            b"POST /generate HTTP/1.1\r\nContent-Length: 99999999999999999999\r\n\r\n{}",
            b"POST /generate HTTP/1.1\r\nContent-Length: 10\r\n\r\n{}",  # shorter body than announced, then close
            b"POST /generate HTTP/1.1\r\nContent-Length: 200000\r\n\r\n" + b"[" * 200000,
            b"POST /generate HTTP/1.1\r\nContent-Length: 9\r\n\r\n{\"a\":\xff\xfe}",
            b"POST /v1/completions HTTP/1.1\r\nContent-Length: 2\r\n\r\n[]",
            b"\x00" * 5000, b"A" * 70000 + b"\r\n\r\n", b"POST  HTTP/1.1\r\n\r\n", b"POST /generate\r\n\r\n",
        ]
        for _ in range(60):
            m = bytearray(b"POST /generate HTTP/1.1\r\nContent-Length: %d\r\n\r\n" % len(good) + good)
            for _ in range(rng.randint(1, 5)):
                m[rng.randrange(len(m))] = rng.randrange(256)
            payloads.append(bytes(m))
        for pl in payloads:
            c = socket.create_connection(("127.0.0.1", s.port), timeout=10)
            try:
                c.sendall(pl)
                c.shutdown(socket.SHUT_WR)
                c.settimeout(10)
                while c.recv(65536):
                    pass
            except OSError:
                pass
            finally:
                c.close()
            assert s.p.poll() is None, (pl[:60], s.p.stderr.read()[-400:])
        code, r = _req(s.port, "/generate", {"tokens": [1, 2], "max_new_tokens": 3})
        assert code == 200 and r["tokens"] == fake_generate([1, 2], 3, 100)
