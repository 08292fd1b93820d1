// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// All arrays use the same SoA conventions as include/p2v.h so the parity tests can feed both
// sides the same buffers.
#include <cstring>
#include <thread>
#include <vector>
#include "hash.hpp"

using namespace orc;

extern "C" {

// which: 0 = `permutation` (dense MDS, Hash/Poseidon.hs:42), 1 = fast-partial schedule
// (Gate/Custom/Poseidon.hs:92-137), 2 = bulk (lazy u128) form.  in/out SoA [12][n].
int orc_poseidon_permute(const uint64_t *in, uint64_t *out, size_t n, int which) {
  for (size_t t = 0; t < n; t++) {
    State s;
    for (int i = 0; i < 12; i++) s[i] = F(in[i * n + t]);
    State r = which == 0 ? permutation(s) : which == 1 ? permutationFast(s) : permutationBulk(s);
    for (int i = 0; i < 12; i++) out[i * n + t] = r[i].v;
  }
  return 0;
}

// `sponge`, Hash/Sponge.hs:26.  leaves SoA [w][n] -> digests SoA [4][n]
int orc_hash_leaves(const uint64_t *leaves, uint32_t w, size_t n, uint64_t *digests) {
  for (size_t t = 0; t < n; t++) {
    std::vector<F> leaf(w);
    for (uint32_t j = 0; j < w; j++) leaf[j] = F(leaves[(size_t)j * n + t]);
    Digest d = sponge(leaf);
    for (int i = 0; i < 4; i++) digests[i * n + t] = d.e[i].v;
  }
  return 0;
}

// `compress`, Hash/Merkle.hs:21
int orc_compress(const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n) {
  for (size_t t = 0; t < n; t++) {
    Digest a, b;
    for (int i = 0; i < 4; i++) { a.e[i] = F(left[i * n + t]); b.e[i] = F(right[i * n + t]); }
    Digest d = compress(a, b);
    for (int i = 0; i < 4; i++) out[i * n + t] = d.e[i].v;
  }
  return 0;
}

// `checkMerkleProof`, Hash/Merkle.hs:39.  Same layout as p2v_merkle_verify; ok[n] bytes.
// Out-of-range cap index ((!!) exception in the reference) is reported as ok = 2.
int orc_merkle_verify(const uint64_t *leaves, uint32_t w, const uint32_t *idx, const uint64_t *siblings,
                      uint32_t path_len, const uint64_t *cap, uint32_t cap_height, size_t n, uint8_t *ok,
                      uint64_t *roots_out) {
  MerkleCap mc;
  for (size_t c = 0; c < ((size_t)1 << cap_height); c++) {
    Digest d;
    for (int i = 0; i < 4; i++) d.e[i] = F(cap[c * 4 + i]);
    mc.roots.push_back(d);
  }
  for (size_t t = 0; t < n; t++) {
    std::vector<F> leaf(w);
    for (uint32_t j = 0; j < w; j++) leaf[j] = F(leaves[(size_t)j * n + t]);
    MerkleProof p;
    for (uint32_t l = 0; l < path_len; l++) {
      Digest d;
      for (int i = 0; i < 4; i++) d.e[i] = F(siblings[(size_t)(l * 4 + i) * n + t]);
      p.siblings.push_back(d);
    }
    auto r = reconstructMerkleRoot_(sponge(leaf), (int)idx[t], p);
    if (roots_out)
      for (int i = 0; i < 4; i++) roots_out[i * n + t] = r.second.e[i].v;
    if ((size_t)r.first >= mc.roots.size()) ok[t] = 2;
    else ok[t] = mc.roots[r.first] == r.second ? 1 : 0;
  }
  return 0;
}

// CPU-baseline timing loop: `threads` workers each iterate the permutation `iters` times on
// their own state (which = as in orc_poseidon_permute).  Returns a checksum.
uint64_t orc_perm_loop(size_t iters, int threads, int which) {
  std::vector<uint64_t> sums(threads, 0);
  std::vector<std::thread> th;
  for (int k = 0; k < threads; k++)
    th.emplace_back([&, k]() {
      State s;
      for (int i = 0; i < 12; i++) s[i] = F((uint64_t)(i + 12 * k));
      for (size_t it = 0; it < iters; it++) s = which == 0 ? permutation(s) : which == 1 ? permutationFast(s) : permutationBulk(s);
      sums[k] = s[0].v;
    });
  for (auto &t : th) t.join();
  uint64_t x = 0;
  for (auto v : sums) x ^= v;
  return x;
}

}  // extern "C"
