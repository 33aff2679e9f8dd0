// sampler.h — host-side token sampling for requests that ask for it ("temperature" > 0).
//
// The engine's loop is greedy and device-resident (north_star; HF `generate(do_sample=False)`); the serving images the
// reference points at also accept OpenAI-style sampling fields on /v1/completions, so a request MAY opt out of greedy:
// the host then drives the engine one step at a time (`ssb_decode(nsteps=1, logits)`), reads the fp32 logits back
// (V floats per step) and picks the token here.  Order of the filters follows HF's logits processors
// (HF:generation/utils.py `_get_logits_processor`: temperature -> top_k -> top_p), then one multinomial draw.
// Deterministic for a given seed: tensor-parallel ranks see bit-identical logits and therefore draw identical ids.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace ssbhost {

struct Sampling {
  float temperature = 0.f;  // 0 = greedy (the engine's own pick is used, no logits leave the device)
  float top_p = 1.f;
  int top_k = 0;            // 0 = off
  uint64_t seed = 0;
  bool on() const { return temperature > 0.f; }
};

struct Rng {  // splitmix64
  uint64_t s;
  double next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
  }
};

// One draw from softmax(logits / T) restricted to top_k, then to the top_p nucleus.  `cand` is scratch.
inline int sample_token(const float* logits, int V, const Sampling& p, Rng& rng, std::vector<std::pair<float, int>>& cand) {
  cand.resize((size_t)V);
  for (int i = 0; i < V; ++i) cand[(size_t)i] = {logits[i], i};
  auto better = [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; };
  size_t n = cand.size();
  if (p.top_k > 0 && (size_t)p.top_k < n) {
    std::nth_element(cand.begin(), cand.begin() + p.top_k, cand.end(), better);
    n = (size_t)p.top_k;
  }
  const float mx = std::max_element(cand.begin(), cand.begin() + n, [](auto& a, auto& b) { return a.first < b.first; })->first;
  const double invT = 1.0 / (double)p.temperature;
  double Z = 0;
  for (size_t i = 0; i < n; ++i) Z += std::exp(((double)cand[i].first - mx) * invT);
  // nucleus: smallest prefix (by descending probability) whose mass reaches top_p; sort only as much as needed
  size_t keep = n;
  double mass = Z;
  if (p.top_p < 1.f) {
    size_t win = std::min<size_t>(n, 64);
    for (;;) {
      std::partial_sort(cand.begin(), cand.begin() + win, cand.begin() + n, better);
      double acc = 0;
      size_t k = 0;
      while (k < win) {
        acc += std::exp(((double)cand[k].first - mx) * invT);
        ++k;
        if (acc >= (double)p.top_p * Z) break;
      }
      if (acc >= (double)p.top_p * Z || win == n) {
        keep = k;
        mass = acc;
        break;
      }
      win = std::min(n, win * 4);
    }
  }
  const double u = rng.next() * mass;
  double acc = 0;
  for (size_t i = 0; i < keep; ++i) {
    acc += std::exp(((double)cand[i].first - mx) * invT);
    if (u < acc) return cand[i].second;
  }
  return cand[keep - 1].second;
}

}  // namespace ssbhost
