// Host-side Goldilocks helpers used only to precompute per-circuit tables (roots of unity,
// coset shifts and their inverses) when a circuit is created.  Not a verifier: all proof
// arithmetic runs on the GPU.  Constants: src/Algebra/Goldilocks.hs:48-74,126-135.
#pragma once
#include <cstdint>

namespace p2vhost {
static const uint64_t HGL_P = 0xFFFFFFFF00000001ULL;
static const uint64_t HGL_MUL_GEN = 0xc65c18b67785d900ULL;
static const uint64_t HGL_TWO_ADIC_GEN = 0x64fdd1a46201e246ULL;

inline uint64_t hmul(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % HGL_P); }
inline uint64_t hpow(uint64_t x, uint64_t e) {
  uint64_t acc = 1;
  while (e) {
    if (e & 1) acc = hmul(acc, x);
    x = hmul(x, x);
    e >>= 1;
  }
  return acc;
}
inline uint64_t hinv(uint64_t x) { return hpow(x, HGL_P - 2); }
// rootsOfUnity ! k : order 2^k
inline uint64_t hroot(int k) {
  uint64_t x = HGL_TWO_ADIC_GEN;
  for (int i = 0; i < 32 - k; i++) x = hmul(x, x);
  return x;
}
}  // namespace p2vhost
