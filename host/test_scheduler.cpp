// CPU unit test of host/scheduler.h with a fake engine (no GPU, no libsubstratus_b200): every request's greedy ids
// must be the ids it would get alone, whatever shared its batches; decode calls must carry more than one sequence when
// clients overlap; slots must all be returned.  Built and run by tests/test_scheduler_cpu.py.
#include <atomic>
#include <cstdio>
#include <map>
#include <random>

#include "scheduler.h"

struct FakeEngine {
  // a "model" whose next token depends on the last token and the position only: deterministic, batch-independent
  static int32_t step(int32_t last, int pos) { return (int32_t)((last * 31u + pos * 7u + 3u) % 997u); }
  std::mutex mu;
  std::map<int, int> len;  // sid -> cached length
  int next_sid = 0, max_slots = 8;
  int max_len = 1 << 30;    // a sequence that would grow past it makes the call fail with -1 ("bad argument"), state untouched
  int kv_budget = 1 << 30;  // total cached tokens the fake "KV pool" holds: calls that would exceed it fail with -4, state untouched
  std::string err;
  std::atomic<int> decode_calls{0};
  int seq_create(int* sid) {
    std::lock_guard<std::mutex> lk(mu);
    if ((int)len.size() >= max_slots) {
      err = "no slots";
      return -4;
    }
    *sid = next_sid++;
    len[*sid] = 0;
    return 0;
  }
  int seq_free(int sid) {
    std::lock_guard<std::mutex> lk(mu);
    return len.erase(sid) ? 0 : -1;
  }
  int total() {
    int t = 0;
    for (auto& kv : len) t += kv.second;
    return t;
  }
  int prefill(const int* sids, const int32_t* toks, const int* lens, int nseq, int32_t* next) {
    std::lock_guard<std::mutex> lk(mu);
    int add = 0;
    for (int i = 0; i < nseq; ++i) add += lens[i];
    if (total() + add > kv_budget) {
      err = "KV block pool exhausted";
      return -4;
    }
    for (int i = 0, o = 0; i < nseq; o += lens[i], ++i)
      if (toks[o] == 666 || len[sids[i]] + lens[i] > max_len) {  // validation first, like the engine
        err = "bad prompt";
        return -1;
      }
    for (int i = 0, o = 0; i < nseq; o += lens[i], ++i) {
      len[sids[i]] += lens[i];
      next[i] = step(toks[o + lens[i] - 1], len[sids[i]]);
    }
    std::this_thread::sleep_for(std::chrono::microseconds(300));
    return 0;
  }
  int decode(const int* sids, const int32_t* last, int nseq, int nsteps, int32_t* out) {
    std::lock_guard<std::mutex> lk(mu);
    ++decode_calls;
    if (total() + nseq * nsteps > kv_budget) {
      err = "KV block pool exhausted";
      return -4;
    }
    for (int i = 0; i < nseq; ++i)
      if (len[sids[i]] + nsteps > max_len) {
        err = "sequence would exceed max_seq_len";
        return -1;
      }
    for (int i = 0; i < nseq; ++i) {
      int32_t t = last[i];
      for (int s = 0; s < nsteps; ++s) {
        len[sids[i]] += 1;
        t = step(t, len[sids[i]]);
        out[(size_t)i * nsteps + s] = t;
      }
    }
    std::this_thread::sleep_for(std::chrono::microseconds(200 * nsteps));  // a step costs the same for 1 or 8 sequences
    return 0;
  }
  std::string last_error() { return err; }
};

static std::vector<int32_t> alone(const std::vector<int32_t>& prompt, int max_new) {
  std::vector<int32_t> out;
  int pos = (int)prompt.size();
  int32_t t = FakeEngine::step(prompt.back(), pos);
  out.push_back(t);
  while ((int)out.size() < max_new) {
    ++pos;
    t = FakeEngine::step(t, pos);
    out.push_back(t);
  }
  return out;
}

int main() {
  FakeEngine eng;
  int failures = 0;
  {
    ssbhost::BatchScheduler<FakeEngine> sched(&eng, 8, 4);
    const int kClients = 24;
    std::vector<std::thread> th;
    std::vector<ssbhost::Request> reqs(kClients);
    std::mt19937 rng(7);
    for (int c = 0; c < kClients; ++c) {
      reqs[c].prompt.resize(1 + rng() % 40);
      for (auto& t : reqs[c].prompt) t = (int32_t)(rng() % 997);
      reqs[c].max_new = 1 + rng() % 33;
    }
    for (int c = 0; c < kClients; ++c)
      th.emplace_back([&, c] {
        std::this_thread::sleep_for(std::chrono::microseconds(150 * (c % 6)));
        sched.submit(&reqs[c]);
      });
    for (auto& t : th) t.join();
    for (int c = 0; c < kClients; ++c) {
      if (!reqs[c].error.empty() || reqs[c].tokens != alone(reqs[c].prompt, reqs[c].max_new)) {
        printf("FAIL client %d err='%s' n=%zu want %d\n", c, reqs[c].error.c_str(), reqs[c].tokens.size(), reqs[c].max_new);
        ++failures;
      }
    }
    ssbhost::Request bad;
    bad.max_new = 3;  // empty prompt
    sched.submit(&bad);
    if (bad.error.empty()) {
      printf("FAIL empty prompt accepted\n");
      ++failures;
    }
    printf("decode calls %d, rows/call %.2f, max rows %d\n", eng.decode_calls.load(), (double)sched.step_rows() / std::max(1LL, sched.steps()),
           sched.max_rows_seen());
    if (sched.max_rows_seen() < 2) {
      printf("FAIL overlapping clients never shared a decode call\n");
      ++failures;
    }
    if (sched.max_rows_seen() > 8) {
      printf("FAIL batch exceeded max_batch\n");
      ++failures;
    }
  }
  if (!eng.len.empty()) {
    printf("FAIL %zu sequence slots leaked\n", eng.len.size());
    ++failures;
  }
  {
    // resource exhaustion must fail ONE request, not its batch mates: budget for ~3 of 6 long requests
    FakeEngine small;
    small.kv_budget = 200;
    ssbhost::BatchScheduler<FakeEngine> sched(&small, 8, 4);
    const int kClients = 6;
    std::vector<ssbhost::Request> reqs(kClients);
    std::vector<std::thread> th;
    for (int c = 0; c < kClients; ++c) {
      reqs[c].prompt.assign(30, (int32_t)(c + 1));
      reqs[c].max_new = 30;
    }
    for (int c = 0; c < kClients; ++c) th.emplace_back([&, c] { sched.submit(&reqs[c]); });
    for (auto& t : th) t.join();
    int okc = 0, failed = 0;
    for (int c = 0; c < kClients; ++c) {
      if (reqs[c].error.empty()) {
        ++okc;
        if (reqs[c].tokens != alone(reqs[c].prompt, reqs[c].max_new)) {
          printf("FAIL survivor %d got wrong ids\n", c);
          ++failures;
        }
      } else {
        ++failed;
        if (reqs[c].error.find("exhausted") == std::string::npos) {
          printf("FAIL unexpected error '%s'\n", reqs[c].error.c_str());
          ++failures;
        }
      }
    }
    printf("KV budget test: %d served, %d refused\n", okc, failed);
    if (okc < 2 || failed < 1) {
      printf("FAIL exhaustion should refuse some requests and serve the rest (served %d, refused %d)\n", okc, failed);
      ++failures;
    }
    if (!small.len.empty()) {
      printf("FAIL slots leaked after exhaustion\n");
      ++failures;
    }
  }
  {
    // an argument the engine rejects for ONE sequence (a poisoned prompt at prefill, a sequence that outgrows max_len in
    // decode) costs that request only: the others finish with the ids they would get alone
    FakeEngine picky;
    picky.max_len = 50;
    ssbhost::BatchScheduler<FakeEngine> sched(&picky, 8, 4);
    const int kClients = 7;
    std::vector<ssbhost::Request> reqs(kClients);
    std::vector<std::thread> th;
    for (int c = 0; c < kClients; ++c) {
      reqs[c].prompt.assign(10 + c, (int32_t)(c + 1));
      reqs[c].max_new = 20;
    }
    reqs[2].prompt[0] = 666;  // rejected at prefill
    reqs[5].max_new = 60;     // 15 + 60 > 50: rejected by a decode call some ticks in
    for (int c = 0; c < kClients; ++c) th.emplace_back([&, c] { sched.submit(&reqs[c]); });
    for (auto& t : th) t.join();
    for (int c = 0; c < kClients; ++c) {
      const bool should_fail = c == 2 || c == 5;
      if (should_fail != !reqs[c].error.empty() || (should_fail && reqs[c].error_code != -1)) {
        printf("FAIL isolation: request %d error '%s' code %d\n", c, reqs[c].error.c_str(), reqs[c].error_code);
        ++failures;
      }
      if (!should_fail && reqs[c].tokens != alone(reqs[c].prompt, reqs[c].max_new)) {
        printf("FAIL isolation: survivor %d got wrong ids\n", c);
        ++failures;
      }
    }
    if (!picky.len.empty()) {
      printf("FAIL slots leaked after per-sequence errors\n");
      ++failures;
    }
    printf("argument-error isolation test done\n");
  }
  {
    // same pool, but the scheduler knows its size (ssb_kv_blocks): requests queue for blocks instead of failing
    FakeEngine small;
    small.kv_budget = 200;
    ssbhost::BatchScheduler<FakeEngine> sched(&small, 8, 4, /*kv_total_blocks=*/12, /*kv_block_size=*/16);
    const int kClients = 6;
    std::vector<ssbhost::Request> reqs(kClients);
    std::vector<std::thread> th;
    for (int c = 0; c < kClients; ++c) {
      reqs[c].prompt.assign(30, (int32_t)(c + 1));
      reqs[c].max_new = 30;  // 60 tokens -> 4 blocks: three requests fit at a time
    }
    for (int c = 0; c < kClients; ++c) th.emplace_back([&, c] { sched.submit(&reqs[c]); });
    for (auto& t : th) t.join();
    for (int c = 0; c < kClients; ++c) {
      if (!reqs[c].error.empty() || reqs[c].tokens != alone(reqs[c].prompt, reqs[c].max_new)) {
        printf("FAIL admission-controlled request %d: '%s'\n", c, reqs[c].error.c_str());
        ++failures;
      }
    }
    printf("KV admission test: max rows %d, deferred passes %lld\n", sched.max_rows_seen(), sched.deferred());
    if (sched.max_rows_seen() > 3 || sched.deferred() < 1) {
      printf("FAIL admission should cap the batch at 3 rows and defer the rest\n");
      ++failures;
    }
  }
  printf(failures ? "SCHEDULER TEST FAILED\n" : "SCHEDULER TEST OK\n");
  return failures ? 1 : 0;
}
