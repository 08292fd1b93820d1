// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// All arrays use the same SoA conventions as include/p2v.h so the parity tests can feed both
// sides the same buffers.
#include <cstring>
#include <thread>
#include <vector>
#include "plonk.hpp"

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


// ---- batch verification (restatement of verifyProof & friends on flat blobs) -----------------
// shape/vkey/blobs as in include/p2v.h.  Any output pointer may be NULL.
//   challenges SoA [cw][n]; combined SoA [2r][n]; eqmask [n]; status [n] = verifyProof verdict;
//   fri_status [n] = checkFRIProof verdict alone; qstatus [n][Q]; folded SoA [2][n*Q] (index p*Q+q).
// fast != 0 switches the permutation to the bulk form (bit-identical, tests assert it).
// A C++ exception from the restatement (an `error` site that the fixed shape should make
// unreachable) is reported as status 0xEE.
int orc_verify_batch(const p2v_shape *shape, const uint64_t *vkey_words, const uint64_t *blobs, size_t n, int threads, int fast,
                     uint64_t *challenges, uint64_t *combined, uint8_t *eqmask, uint32_t *status, uint32_t *fri_status,
                     uint32_t *qstatus, uint64_t *folded, unsigned long long *perm_count) {
  CommonCircuitData c = commonFromShape(*shape);
  VerifierOnlyCircuitData vk = vkeyFromWords(c, vkey_words);
  activePermutation() = fast ? permutationBulk : permutation;
  size_t bw = 0;
  {
    p2v_layout dummy;
    (void)dummy;
    ProofWithPublicInputs probe;  // blob size from a round trip of the first proof
    bw = proofToBlob(proofFromBlob(c, blobs)).size();
  }
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  std::vector<unsigned long long> counts(threads, 0);
  int Q = c.fri_config.num_query_rounds, r = c.num_challenges;
  for (int k = 0; k < threads; k++)
    th.emplace_back([&, k]() {
      permCounter() = 0;
      for (size_t p = k; p < n; p += threads) {
        try {
          ProofWithPublicInputs pw = proofFromBlob(c, blobs + p * bw);
          VerifyTrace tr;
          uint32_t st = verifyProofStatus(c, vk, pw, &tr);
          if (status) status[p] = st;
          if (fri_status) fri_status[p] = tr.fri_status;
          if (challenges) {
            std::vector<uint64_t> w = challengesToWords(tr.challenges);
            for (size_t i = 0; i < w.size(); i++) challenges[i * n + p] = w[i];
          }
          if (combined)
            for (int j = 0; j < r; j++) { combined[(2 * j) * n + p] = tr.combined[j].r.v; combined[(2 * j + 1) * n + p] = tr.combined[j].i.v; }
          if (eqmask) {
            uint8_t m = 0;
            for (size_t j = 0; j < tr.eqs_ok.size(); j++) if (tr.eqs_ok[j]) m |= (uint8_t)(1u << j);
            eqmask[p] = m;
          }
          for (int q = 0; q < Q && q < (int)tr.queries.size(); q++) {
            const QueryTrace &qt = tr.queries[q];
            if (qstatus) qstatus[p * Q + q] = qt.status == 0 ? 0u : ((uint32_t)qt.status | ((uint32_t)q << 8) | ((uint32_t)qt.detail << 16));
            if (folded) { folded[p * Q + q] = qt.folded.r.v; folded[n * Q + p * Q + q] = qt.folded.i.v; }
          }
        } catch (const std::exception &e) {
          if (status) status[p] = 0xEE;
          if (fri_status) fri_status[p] = 0xEE;
        }
      }
      counts[k] = permCounter();
    });
  for (auto &t : th) t.join();
  if (perm_count) { *perm_count = 0; for (auto v : counts) *perm_count += v; }
  activePermutation() = permutation;
  return 0;
}

// unfiltered constraint vector of one gate on given openings (for gate-level differential tests):
// wires [num_wires][2], consts [num_consts][2], pih[4]; out [max_out][2]; returns the number of constraints
int orc_gate_constraints(const p2v_shape *shape, int gate_index, const uint64_t *wires, int num_wires, const uint64_t *consts,
                         int num_consts, const uint64_t *pih, uint64_t *out, int max_out) {
  CommonCircuitData c = commonFromShape(*shape);
  EvaluationVars ev;
  for (int i = 0; i < num_wires; i++) ev.local_wires.push_back(FExt(F(wires[2 * i]), F(wires[2 * i + 1])));
  for (int i = 0; i < num_consts; i++) ev.local_constants.push_back(FExt(F(consts[2 * i]), F(consts[2 * i + 1])));
  for (int i = 0; i < 4; i++) ev.public_inputs_hash.push_back(F(pih[i]));
  try {
    std::vector<FExt> v = gateConstraints(c.gates.at(gate_index), ev);
    for (int i = 0; i < (int)v.size() && i < max_out; i++) { out[2 * i] = v[i].r.v; out[2 * i + 1] = v[i].i.v; }
    return (int)v.size();
  } catch (const std::exception &e) {
    return -1;
  }
}

// ---- small algebra / transcript probes for the known-answer tests -----------------
// op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 a^b (b as a signed 64-bit exponent), 5 subgroupGenerator(a), 6 div
uint64_t orc_field_op(int op, uint64_t a, uint64_t b) {
  switch (op) {
    case 0: return (F(a) * F(b)).v;
    case 1: return (F(a) + F(b)).v;
    case 2: return (F(a) - F(b)).v;
    case 3: return inv(F(a)).v;
    case 4: return powi(F(a), (long long)b).v;
    case 5: return subgroupGenerator((int)a).v;
    case 6: return (F(a) / F(b)).v;
  }
  return 0;
}
// op: 0 mul, 1 inv(x), 2 div, 3 x^e (e in y[0], signed)
void orc_ext_op(int op, const uint64_t *x, const uint64_t *y, uint64_t *out) {
  FExt a = FExt(F(x[0]), F(x[1]));
  FExt b = y ? FExt(F(y[0]), F(y[1])) : FExt();
  FExt r;
  switch (op) {
    case 0: r = a * b; break;
    case 1: r = invExt(a); break;
    case 2: r = a / b; break;
    default: r = powExtI(a, (long long)y[0]); break;
  }
  out[0] = r.r.v; out[1] = r.i.v;
}
// Duplex script (Challenge/Pure.hs): ops[i] >= 0 absorbs that many of the next inputs, ops[i] < 0 squeezes
// -ops[i] field elements; returns the number of squeezed outputs.
int orc_duplex_script(const int *ops, int nops, const uint64_t *inputs, uint64_t *outputs, unsigned long long *perms) {
  Duplex dx(zeroState());
  permCounter() = 0;
  int in = 0, out = 0;
  for (int i = 0; i < nops; i++) {
    if (ops[i] >= 0) for (int k = 0; k < ops[i]; k++) dx.absorb(F(inputs[in++]));
    else for (int k = 0; k < -ops[i]; k++) outputs[out++] = dx.squeezeFelt().v;
  }
  if (perms) *perms = permCounter();
  return out;
}
// reverseBitsInt / evalLagrange0 / foldCosetWith probes
uint64_t orc_reverse_bits(int n, uint64_t w) { return reverseBits(n, w); }

// layout implied by a shape, computed by the oracle's own reader/writer (for comparison with p2v_shape_layout)
size_t orc_blob_words(const p2v_shape *shape, const uint64_t *blob) {
  CommonCircuitData c = commonFromShape(*shape);
  return proofToBlob(proofFromBlob(c, blob)).size();
}

}  // extern "C"
