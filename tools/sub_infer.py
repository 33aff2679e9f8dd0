#!/usr/bin/env python
"""sub_infer.py — client for a running `serve` container: the first in-repo caller of the Server endpoint.

The reference's `sub infer` (internal/cli/infer.go:17-66, internal/tui/infer_chat.go) is a chat stub that never sends a
request; its system test reaches the server with `kubectl port-forward service/<name>-server 8080:8080` and one
`POST /v1/completions` (test/system.sh:60-78).  This is that client, against the same URL:

  sub_infer.py "Who was the first president?"            one completion, streamed to the terminal
  sub_infer.py --ids 1,15043,3186 --max-tokens 32        token-id prompt (works without a tokenizer.json in the model dir)
  sub_infer.py --chat                                     line-based loop: every line is sent as a prompt
  sub_infer.py --bench 64 --concurrency 32                SURVEY.md §8d synthetic requests (512 random ids, seeds 1234+i,
                                                          128 greedy tokens) through HTTP: p50 TTFT, aggregate tokens/s

Standard library only (urllib + threads), so it runs in any pod or laptop."""
import argparse
import json
import random
import statistics
import sys
import threading
import time
import urllib.error
import urllib.request


def _post(url, body, timeout):
    return urllib.request.urlopen(urllib.request.Request(url, data=json.dumps(body).encode(), headers={"Content-Type": "application/json"}),
                                  timeout=timeout)


def wait_ready(base, timeout):
    """GET / until 200 (the readiness contract: 503 while the model loads)."""
    t0 = time.time()
    while True:
        try:
            with urllib.request.urlopen(base + "/", timeout=5) as r:
                if r.status == 200:
                    return json.loads(r.read() or b"{}")
        except (urllib.error.URLError, ConnectionError, OSError):
            pass
        if time.time() - t0 > timeout:
            raise SystemExit(f"{base}/ did not return 200 within {timeout:.0f} s (model still loading or server down)")
        time.sleep(0.5)


def stream_completion(base, prompt, max_tokens, stop_at_eos=False, timeout=600, on_piece=None):
    """POST /v1/completions with "stream": true.  Returns dict(tokens, text, ttft_s, total_s, finish_reason, usage)."""
    body = {"prompt": prompt, "max_tokens": max_tokens, "stream": True}
    if stop_at_eos:
        body["stop_at_eos"] = True
    t0 = time.perf_counter()
    out = {"tokens": [], "text": "", "ttft_s": None, "finish_reason": None, "usage": None, "error": None}
    try:
        resp = _post(base + "/v1/completions", body, timeout)
    except urllib.error.HTTPError as e:
        out["error"] = (json.loads(e.read() or b"{}").get("error") or str(e))
        out["total_s"] = time.perf_counter() - t0
        return out
    with resp:
        buf = b""
        while True:
            chunk = resp.read1(65536) if hasattr(resp, "read1") else resp.read(4096)
            if not chunk:
                break
            buf += chunk
            while b"\n\n" in buf:
                block, buf = buf.split(b"\n\n", 1)
                if not block.startswith(b"data: "):
                    continue
                payload = block[6:].decode("utf-8")
                if payload == "[DONE]":
                    continue
                ev = json.loads(payload)
                if "error" in ev:
                    out["error"] = ev["error"]
                    continue
                ch = ev["choices"][0]
                if ch["tokens"] and out["ttft_s"] is None:
                    out["ttft_s"] = time.perf_counter() - t0
                out["tokens"] += ch["tokens"]
                out["text"] += ch["text"]
                if on_piece and (ch["text"] or ch["tokens"]):
                    on_piece(ch["text"], ch["tokens"])
                if ch.get("finish_reason"):
                    out["finish_reason"] = ch["finish_reason"]
                    out["usage"] = ev.get("usage")
                    out["server"] = {k: ev[k] for k in ("ttft_ms", "decode_ms", "decode_tokens_per_sec") if k in ev}
    out["total_s"] = time.perf_counter() - t0
    return out


def synthetic_prompt(vocab, i, n=512):
    rng = random.Random(1234 + i)
    return [rng.randrange(vocab) for _ in range(n)]


def bench(base, n_requests, concurrency, vocab, prompt_len, max_tokens, timeout):
    results, lock, nxt = [], threading.Lock(), [0]

    def worker():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= n_requests:
                return
            r = stream_completion(base, synthetic_prompt(vocab, i, prompt_len), max_tokens, timeout=timeout)
            with lock:
                results.append(r)

    t0 = time.perf_counter()
    th = [threading.Thread(target=worker) for _ in range(concurrency)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    ok = [r for r in results if not r["error"] and r["ttft_s"] is not None]
    toks = sum(len(r["tokens"]) for r in ok)
    line = {"requests": n_requests, "failed": len(results) - len(ok), "concurrency": concurrency, "prompt_len": prompt_len, "max_tokens": max_tokens,
            "wall_s": wall, "generated_tokens": toks, "tokens_per_sec": toks / wall if wall > 0 else 0.0,
            "ttft_ms_p50": statistics.median(r["ttft_s"] for r in ok) * 1e3 if ok else None,
            "request_s_p50": statistics.median(r["total_s"] for r in ok) if ok else None,
            "errors": sorted({r["error"] for r in results if r["error"]})[:3]}
    print(json.dumps(line))
    return line


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("prompt", nargs="?", help="text prompt (needs tokenizer.json in the server's model dir)")
    ap.add_argument("--url", default="http://127.0.0.1:8080", help="server base URL (after kubectl port-forward service/<name>-server 8080)")
    ap.add_argument("--ids", help="comma-separated token ids instead of text")
    ap.add_argument("--max-tokens", type=int, default=64)
    ap.add_argument("--stop-at-eos", action="store_true")
    ap.add_argument("--chat", action="store_true", help="read prompts line by line from stdin")
    ap.add_argument("--bench", type=int, default=0, metavar="N", help="send N synthetic benchmark requests and print one JSON line")
    ap.add_argument("--concurrency", type=int, default=1)
    ap.add_argument("--vocab", type=int, default=32000, help="--bench: ids are drawn from [0, vocab)")
    ap.add_argument("--prompt-len", type=int, default=512)
    ap.add_argument("--wait", type=float, default=0.0, help="wait up to this many seconds for GET / to return 200 first")
    ap.add_argument("--timeout", type=float, default=600.0)
    a = ap.parse_args(argv)
    base = a.url.rstrip("/")
    if a.wait > 0:
        wait_ready(base, a.wait)
    if a.bench:
        line = bench(base, a.bench, max(1, a.concurrency), a.vocab, a.prompt_len, a.max_tokens if a.max_tokens != 64 else 128, a.timeout)
        return 1 if line["failed"] else 0

    def one(prompt):
        def show(text, ids):
            sys.stdout.write(text if text else (" ".join(map(str, ids)) + " "))
            sys.stdout.flush()

        r = stream_completion(base, prompt, a.max_tokens, a.stop_at_eos, a.timeout, on_piece=show)
        sys.stdout.write("\n")
        if r["error"]:
            print("error:", r["error"], file=sys.stderr)
            return 1
        srv = r.get("server") or {}
        print(f"[{len(r['tokens'])} tokens, finish={r['finish_reason']}, ttft {1e3 * (r['ttft_s'] or 0):.1f} ms, "
              f"{srv.get('decode_tokens_per_sec', 0):.1f} tok/s on the device]", file=sys.stderr)
        return 0

    if a.chat:
        rc = 0
        for ln in sys.stdin:
            ln = ln.rstrip("\n")
            if ln:
                rc |= one(ln)
        return rc
    if a.ids:
        return one([int(x) for x in a.ids.split(",")])
    if a.prompt is None:
        ap.error("give a prompt, --ids, --chat or --bench N")
    return one(a.prompt)


if __name__ == "__main__":
    sys.exit(main())
