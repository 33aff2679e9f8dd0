// fake_ssb.cpp — TEST INFRASTRUCTURE ONLY.  A GPU-free stand-in for libsubstratus_b200.so that implements the entry
// points of include/ssb.h the serve host calls, so that host/serve.cpp's HTTP surface (readiness 503->200, request
// validation, /generate, /v1/completions, SSE streaming, continuous batching, the in-container tensor-parallel rank
// threads) can be exercised on a CPU-only box (tests/test_serve_fake_cpu.py builds `serve` against this file in a temp
// dir).  It is never linked into the product: host/Makefile links the real library, which has no CPU fallback.
//
// "Model": every sequence carries a 64-bit state folded over all its ids; the greedy next id is state % vocab.  The
// ids of a request therefore depend only on its own prompt — the property the real engine has under greedy decoding —
// and the tests recompute them in Python (fake_next in the test file).
#include <chrono>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ssb.h"
#include "../../substratus_b200/csrc/json.h"
#include "../../substratus_b200/csrc/tokenizer.h"

namespace {
thread_local std::string g_err;
inline uint64_t fold(uint64_t s, int32_t tok) {
  uint64_t z = s + 0x9E3779B97F4A7C15ull + (uint64_t)(uint32_t)tok;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
}  // namespace

// fp32 "logits" of the fake model in state s: the greedy id (s % vocab) gets 9, every other id a hash value in [0, 8)
static void fake_logits(uint64_t s, int vocab, float* out) {
  for (int v = 0; v < vocab; ++v) out[v] = 8.0f * (float)((fold(s, v) >> 40) & 0xFFFF) / 65536.0f;
  out[s % (uint64_t)vocab] = 9.0f;
}

struct ssb_engine {
  int vocab = 1000, max_batch = 32, max_seq_len = 4096, tp_size = 1, tp_rank = 0;
  int fail_code = SSB_ECUDA;  // "fake_fail_code": what the failing decode returns (SSB_ECUDA = the device is gone)
  int kv_blocks = 1 << 20;  // "fake_kv_blocks": pool of 16-token blocks; prefill / decode beyond it return SSB_ENOMEM
  int step_us = 0, fail_after = -1;  // "fake_step_us": sleep per decode step; "fake_fail_after": decode calls before an error
  bool connected = false;
  std::mutex mu;
  std::map<int, std::pair<uint64_t, int>> seqs;  // id -> (state, length)
  int next_id = 0;
  long long decode_calls = 0, max_rows = 0;
};
constexpr int kFakeBlock = 16;
// blocks the sequences hold, with `extra[i]` more tokens on the listed ones (caller holds e->mu)
static int fake_blocks_used(ssb_engine* e, const int* ids, const int* extra, int n, int extra_all) {
  int used = 0;
  for (auto& kv : e->seqs) {
    int len = kv.second.second;
    for (int i = 0; i < n; ++i)
      if (ids[i] == kv.first) len += extra ? extra[i] : extra_all;
    used += (len + kFakeBlock - 1) / kFakeBlock;
  }
  return used;
}
struct ssb_tokenizer {
  ssb::Tokenizer impl;
};

extern "C" {

int ssb_engine_create(const char* model_dir, const char* params_json, ssb_engine** out) {
  if (!model_dir || !out) return SSB_EINVAL;
  *out = nullptr;
  ssb_engine* e = new ssb_engine();
  try {
    const ssb::Json pj = ssb::json_parse(params_json && *params_json ? params_json : "{}");
    e->vocab = (int)pj.get_int("fake_vocab", 1000);
    e->max_batch = (int)pj.get_int("max_batch", 32);
    e->max_seq_len = (int)pj.get_int("max_seq_len", 4096);
    e->tp_size = (int)pj.get_int("tp_size", 1);
    e->tp_rank = (int)pj.get_int("tp_rank", 0);
    e->step_us = (int)pj.get_int("fake_step_us", 0);
    e->fail_after = (int)pj.get_int("fake_fail_after", -1);
    e->fail_code = (int)pj.get_int("fake_fail_code", SSB_ECUDA);
    e->kv_blocks = (int)pj.get_int("fake_kv_blocks", 1 << 20);
    const int load_ms = (int)pj.get_int("fake_load_ms", 0);
    if (load_ms > 0) std::this_thread::sleep_for(std::chrono::milliseconds(load_ms));
    if (pj.get_int("fake_load_error", 0) != 0) {
      g_err = "fake: load error requested";
      delete e;
      return SSB_EIO;
    }
  } catch (std::exception& ex) {
    g_err = std::string("params.json: ") + ex.what();
    delete e;
    return SSB_EINVAL;
  }
  *out = e;
  return SSB_OK;
}
void ssb_engine_destroy(ssb_engine* e) { delete e; }
int ssb_engine_info(ssb_engine* e, ssb_info* out) {
  if (!e || !out) return SSB_EINVAL;
  memset(out, 0, sizeof *out);
  out->vocab_size = e->vocab;
  out->max_batch = e->max_batch;
  out->max_seq_len = e->max_seq_len;
  out->tp_size = e->tp_size;
  out->tp_rank = e->tp_rank;
  out->n_layers = 2;
  out->kv_block_size = kFakeBlock;
  snprintf(out->model_type, sizeof out->model_type, "fake");
  snprintf(out->dtype, sizeof out->dtype, "u64");
  return SSB_OK;
}
int ssb_seq_create(ssb_engine* e, int* seq_id) {
  if (!e || !seq_id) return SSB_EINVAL;
  std::lock_guard<std::mutex> lk(e->mu);
  if ((int)e->seqs.size() >= e->max_batch) {
    g_err = "no free sequence slot";
    return SSB_ESTATE;
  }
  *seq_id = e->next_id++;
  e->seqs[*seq_id] = {0, 0};
  return SSB_OK;
}
int ssb_seq_free(ssb_engine* e, int seq_id) {
  if (!e) return SSB_EINVAL;
  std::lock_guard<std::mutex> lk(e->mu);
  return e->seqs.erase(seq_id) ? SSB_OK : SSB_EINVAL;
}
int ssb_kv_blocks(ssb_engine* e, int* total, int* free_now) {
  if (!e || !total || !free_now) return SSB_EINVAL;
  std::lock_guard<std::mutex> lk(e->mu);
  *total = e->kv_blocks;
  *free_now = e->kv_blocks - fake_blocks_used(e, nullptr, nullptr, 0, 0);
  return SSB_OK;
}
int ssb_prefill(ssb_engine* e, const int* seq_ids, const int32_t* tokens, const int* lens, int nseq, int32_t* next_tok, float* logits) {
  if (!e || !seq_ids || !tokens || !lens || !next_tok || nseq < 1) return SSB_EINVAL;
  if (e->tp_size > 1 && !e->connected) {
    g_err = "tensor-parallel engine used before ssb_tp_connect";
    return SSB_ESTATE;
  }
  std::lock_guard<std::mutex> lk(e->mu);
  if (fake_blocks_used(e, seq_ids, lens, nseq, 0) > e->kv_blocks) {
    g_err = "KV block pool exhausted";
    return SSB_ENOMEM;
  }
  size_t off = 0;
  for (int i = 0; i < nseq; ++i) {
    auto it = e->seqs.find(seq_ids[i]);
    if (it == e->seqs.end() || lens[i] < 1 || it->second.second + lens[i] >= e->max_seq_len) {
      g_err = "bad sequence id or prompt too long";
      return SSB_EINVAL;
    }
    for (int t = 0; t < lens[i]; ++t) it->second.first = fold(it->second.first, tokens[off + t]);
    it->second.second += lens[i];
    off += (size_t)lens[i];
    next_tok[i] = (int32_t)(it->second.first % (uint64_t)e->vocab);
    if (logits) fake_logits(it->second.first, e->vocab, logits + (size_t)i * e->vocab);
  }
  return SSB_OK;
}
int ssb_decode(ssb_engine* e, const int* seq_ids, const int32_t* last_tok, int nseq, int nsteps, int32_t* out_tok, float* logits) {
  if (!e || !seq_ids || !last_tok || !out_tok || nseq < 1 || nsteps < 1) return SSB_EINVAL;
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->fail_after >= 0 && e->decode_calls >= e->fail_after) {
    g_err = "fake: decode failure requested";
    return e->fail_code;
  }
  if (fake_blocks_used(e, seq_ids, nullptr, nseq, nsteps) > e->kv_blocks) {
    g_err = "KV block pool exhausted";
    return SSB_ENOMEM;
  }
  ++e->decode_calls;
  for (int i = 0; i < nseq; ++i) {
    auto it = e->seqs.find(seq_ids[i]);
    if (it == e->seqs.end() || it->second.second + nsteps > e->max_seq_len) {
      g_err = "bad sequence id or max_seq_len exceeded";
      return SSB_EINVAL;
    }
    int32_t tok = last_tok[i];
    for (int s = 0; s < nsteps; ++s) {
      it->second.first = fold(it->second.first, tok);
      ++it->second.second;
      tok = (int32_t)(it->second.first % (uint64_t)e->vocab);
      out_tok[(size_t)i * nsteps + s] = tok;
      if (logits) fake_logits(it->second.first, e->vocab, logits + ((size_t)s * nseq + i) * e->vocab);  // [nsteps][nseq][V]
    }
  }
  if (e->step_us > 0) std::this_thread::sleep_for(std::chrono::microseconds((long long)e->step_us * nsteps));
  return SSB_OK;
}
int ssb_tp_handle_size(void) { return 256; }
int ssb_tp_export(ssb_engine* e, void* handle) {
  if (!e || !handle) return SSB_EINVAL;
  memset(handle, 0, 256);
  memcpy(handle, &e->tp_rank, sizeof(int));
  return SSB_OK;
}
int ssb_tp_connect(ssb_engine* e, const void* handles, int n) {
  if (!e || !handles || n != e->tp_size) {
    g_err = "ssb_tp_connect: wrong number of handles";
    return SSB_EINVAL;
  }
  for (int r = 0; r < n; ++r) {
    int got = -1;
    memcpy(&got, (const char*)handles + (size_t)r * 256, sizeof(int));
    if (got != r) {
      g_err = "ssb_tp_connect: handles out of rank order";
      return SSB_EINVAL;
    }
  }
  e->connected = true;
  return SSB_OK;
}

int ssb_tok_load(const char* path, ssb_tokenizer** out) {
  if (!path || !out) return SSB_EINVAL;
  *out = nullptr;
  ssb_tokenizer* t = new ssb_tokenizer();
  std::string err;
  bool ok = false;
  try {
    ok = t->impl.load(path, &err);
  } catch (std::exception& ex) {
    err = ex.what();
  }
  if (!ok) {
    delete t;
    g_err = err;
    return SSB_EINVAL;
  }
  *out = t;
  return SSB_OK;
}
void ssb_tok_free(ssb_tokenizer* t) { delete t; }
int ssb_tok_encode(ssb_tokenizer* t, const char* text, int add_special, int32_t* ids, int cap, int* n_out) {
  if (!t || !text || !n_out) return SSB_EINVAL;
  const std::vector<int32_t> v = t->impl.encode(text, add_special != 0);
  *n_out = (int)v.size();
  if ((int)v.size() > cap) return SSB_ENOMEM;
  if (!v.empty()) memcpy(ids, v.data(), v.size() * sizeof(int32_t));
  return SSB_OK;
}
int ssb_tok_decode(ssb_tokenizer* t, const int32_t* ids, int n, int skip_special, char* buf, int cap, int* len_out) {
  if (!t || !len_out) return SSB_EINVAL;
  const std::string s = t->impl.decode(std::vector<int32_t>(ids, ids + n), skip_special != 0);
  *len_out = (int)s.size();
  if ((int)s.size() > cap) return SSB_ENOMEM;
  if (!s.empty()) memcpy(buf, s.data(), s.size());
  return SSB_OK;
}
const char* ssb_last_error(void) { return g_err.c_str(); }
const char* ssb_version(void) { return "fake_ssb (tests only, no GPU)"; }

}  // extern "C"
