// test_sampler.cpp — host/sampler.h against the exact distribution it must draw from: softmax(logits / T) restricted to
// top_k, then to the smallest prefix (by descending probability) whose mass reaches top_p, renormalised
// (HF:generation/logits_process.py TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper, then multinomial).
// Built and run by tests/test_scheduler_cpu.py; no GPU, no engine.
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "sampler.h"

using namespace ssbhost;

static std::vector<double> expected(const std::vector<float>& lg, const Sampling& p) {
  const int V = (int)lg.size();
  std::vector<int> idx(V);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return lg[a] != lg[b] ? lg[a] > lg[b] : a < b; });
  int n = (p.top_k > 0 && p.top_k < V) ? p.top_k : V;
  std::vector<double> pr(V, 0.0);
  double Z = 0;
  for (int i = 0; i < n; ++i) Z += pr[idx[i]] = std::exp(((double)lg[idx[i]] - lg[idx[0]]) / p.temperature);
  double acc = 0;
  int keep = n;
  if (p.top_p < 1.f)
    for (int i = 0; i < n; ++i) {
      acc += pr[idx[i]];
      if (acc >= (double)p.top_p * Z) {
        keep = i + 1;
        break;
      }
    }
  double Z2 = 0;
  for (int i = 0; i < n; ++i) {
    if (i >= keep) pr[idx[i]] = 0;
    Z2 += pr[idx[i]];
  }
  for (auto& x : pr) x /= Z2;
  return pr;
}

static int fails = 0;
static void check(const char* name, const std::vector<float>& lg, const Sampling& p, int draws) {
  const int V = (int)lg.size();
  const std::vector<double> want = expected(lg, p);
  std::vector<long> got(V, 0);
  Rng rng{p.seed};
  std::vector<std::pair<float, int>> cand;
  for (int i = 0; i < draws; ++i) ++got[sample_token(lg.data(), V, p, rng, cand)];
  double chi2 = 0;
  int dof = -1;
  for (int v = 0; v < V; ++v) {
    if (want[v] == 0) {
      if (got[v]) {
        printf("FAIL %s: id %d drawn %ld times but has probability 0\n", name, v, got[v]);
        ++fails;
        return;
      }
      continue;
    }
    const double e = want[v] * draws;
    if (e < 5) continue;  // standard chi-square validity rule
    chi2 += (got[v] - e) * (got[v] - e) / e;
    ++dof;
  }
  // P(chi2 > dof + 5*sqrt(2*dof)) is < 1e-5 for the dof used here; the seeds are fixed, so this is deterministic anyway
  const double limit = dof + 5.0 * std::sqrt(2.0 * std::max(dof, 1));
  printf("%-28s dof %3d chi2 %8.1f (limit %.1f)\n", name, dof, chi2, limit);
  if (chi2 > limit) {
    printf("FAIL %s\n", name);
    ++fails;
  }
}

int main() {
  std::vector<float> lg(300);
  uint64_t s = 12345;
  for (auto& x : lg) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    x = (float)((s >> 33) % 8000) / 1000.0f;  // 0 .. 8
  }
  lg[17] = 9.5f;
  lg[250] = 9.5f;  // a tie at the top: lower id first
  check("T=1", lg, Sampling{1.0f, 1.0f, 0, 1}, 400000);
  check("T=0.7 top_k=20", lg, Sampling{0.7f, 1.0f, 20, 2}, 400000);
  check("T=1 top_p=0.6", lg, Sampling{1.0f, 0.6f, 0, 3}, 400000);
  check("T=1.3 top_k=50 top_p=0.9", lg, Sampling{1.3f, 0.9f, 50, 4}, 400000);
  check("T=2 top_p=0.999 (wide nucleus)", lg, Sampling{2.0f, 0.999f, 0, 5}, 400000);
  {  // top_k = 1 and a nucleus below the top probability are the argmax (lowest id on ties), whatever the seed
    Rng rng{99};
    std::vector<std::pair<float, int>> cand;
    for (int i = 0; i < 1000; ++i) {
      if (sample_token(lg.data(), 300, Sampling{1.0f, 1.0f, 1, 0}, rng, cand) != 17) ++fails;
      if (sample_token(lg.data(), 300, Sampling{0.01f, 0.3f, 0, 0}, rng, cand) != 17) ++fails;
    }
  }
  if (fails) {
    printf("SAMPLER TEST FAILED (%d)\n", fails);
    return 1;
  }
  printf("SAMPLER TEST OK\n");
  return 0;
}
