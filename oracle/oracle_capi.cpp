// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
// extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
// All arrays use the same SoA conventions as include/p2v.h so the parity tests can feed both
// sides the same buffers.
#include <cstring>
#include <thread>
#include <vector>
#include "plonk.hpp"
#include "json_reader.hpp"

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
static int verifyBatchCore(const CommonCircuitData &c, const VerifierOnlyCircuitData &vk, const uint64_t *blobs, size_t n, int threads, int fast,
                           uint64_t *challenges, uint64_t *combined, uint8_t *eqmask, uint32_t *status, uint32_t *fri_status,
                           uint32_t *qstatus, uint64_t *folded, unsigned long long *perm_count);

int orc_verify_batch(const p2v_shape *shape, const uint64_t *vkey_words, const uint64_t *blobs, size_t n, int threads, int fast,
                     uint64_t *challenges, uint64_t *combined, uint8_t *eqmask, uint32_t *status, uint32_t *fri_status,
                     uint32_t *qstatus, uint64_t *folded, unsigned long long *perm_count) {
  CommonCircuitData c = commonFromShape(*shape);
  VerifierOnlyCircuitData vk = vkeyFromWords(c, vkey_words);
  return verifyBatchCore(c, vk, blobs, n, threads, fast, challenges, combined, eqmask, status, fri_status, qstatus, folded, perm_count);
}

static int verifyBatchCore(const CommonCircuitData &c, const VerifierOnlyCircuitData &vk, const uint64_t *blobs, size_t n, int threads, int fast,
                           uint64_t *challenges, uint64_t *combined, uint8_t *eqmask, uint32_t *status, uint32_t *fri_status,
                           uint32_t *qstatus, uint64_t *folded, unsigned long long *perm_count) {
  if (n == 0) { if (perm_count) *perm_count = 0; return 0; }
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


// ---- the oracle's own way in: the reference's JSON files, read by oracle/json_reader.hpp (no product code) --------
struct OrcCircuit {
  CommonCircuitData c;
  VerifierOnlyCircuitData vk;
  size_t blob_words = 0;
};
static thread_local std::string orc_error;
const char *orc_last_error() { return orc_error.c_str(); }

// `decode text_common`, `decode text_vkey` (src/testmain.hs:35-36) -> handle, NULL on a decoding error
void *orc_circuit_from_json(const char *common_json, size_t common_len, const char *vkey_json, size_t vkey_len) {
  try {
    auto *h = new OrcCircuit();
    h->c = js::commonFromJson(common_json, common_len);
    h->vk = js::vkeyFromJson(vkey_json, vkey_len);
    return h;
  } catch (const std::exception &e) {
    orc_error = e.what();
    return nullptr;
  }
}
void orc_circuit_free(void *h) { delete (OrcCircuit *)h; }

// `decode text_proof` (src/testmain.hs:37) flattened in declaration order (Types.hs:251-279): returns the number of
// words (what the product calls blob_words) or -1; out may be NULL to ask for the size only.
long long orc_proof_from_json(void *h, const char *proof_json, size_t len, uint64_t *out, size_t cap_words) {
  try {
    (void)h;
    std::vector<uint64_t> b = proofToBlob(js::proofFromJson(proof_json, len));
    if (out) {
      if (b.size() > cap_words) throw std::runtime_error("output buffer too small");
      memcpy(out, b.data(), b.size() * 8);
    }
    return (long long)b.size();
  } catch (const std::exception &e) {
    orc_error = e.what();
    return -1;
  }
}
// vkey as words (cap ++ circuit_digest), the form p2v_circuit_create takes
long long orc_vkey_words(void *h, uint64_t *out, size_t cap_words) {
  auto *oc = (OrcCircuit *)h;
  std::vector<uint64_t> b;
  for (auto &d : oc->vk.constants_sigmas_cap.roots) blobPutD(b, d);
  blobPutD(b, oc->vk.circuit_digest);
  if (out) {
    if (b.size() > cap_words) return -1;
    memcpy(out, b.data(), b.size() * 8);
  }
  return (long long)b.size();
}
// a few numbers of the circuit for callers that must size buffers without the product's p2v_shape:
// out[0..7] = num_challenges, num_queries, num_steps, num_lookup_polys, degree_bits, rate_bits, cap_height, num_public_inputs
void orc_circuit_info(void *h, int *out) {
  auto *oc = (OrcCircuit *)h;
  out[0] = oc->c.num_challenges; out[1] = oc->c.fri_config.num_query_rounds; out[2] = (int)oc->c.fri_config.step_arity_bits.size();
  out[3] = oc->c.num_lookup_polys; out[4] = oc->c.degree_bits; out[5] = oc->c.fri_config.rate_bits; out[6] = oc->c.fri_config.cap_height;
  out[7] = oc->c.num_public_inputs;
}
int orc_verify_batch_h(void *h, const uint64_t *blobs, size_t n, int threads, int fast, uint64_t *challenges, uint64_t *combined,
                       uint8_t *eqmask, uint32_t *status, uint32_t *fri_status, uint32_t *qstatus, uint64_t *folded,
                       unsigned long long *perm_count) {
  auto *oc = (OrcCircuit *)h;
  return verifyBatchCore(oc->c, oc->vk, blobs, n, threads, fast, challenges, combined, eqmask, status, fri_status, qstatus, folded, perm_count);
}

// verifyProof against GIVEN challenges (the counterpart of p2v_verify_intermediates' challenges_in test hook): the flat
// words (include/p2v.h order) are turned back into ProofChallenges and fed to the same restated functions
// (checkCombinedPlonkEquations', Plonk/Verifier.hs:35-52; checkFRIProof, Plonk/FRI.hs:358-407), so branches that no honest
// transcript reaches — evalLagrange0 at zeta = 1 (Algebra/Poly.hs:14-17), combineInitial with x = zeta and inv 0 = 0
// (Plonk/FRI.hs:151-207, Algebra/Goldilocks.hs:155) — can be compared.
int orc_verify_with_challenges_h(void *h, const uint64_t *blobs, size_t n, const uint64_t *challenges_in, uint64_t *combined,
                                 uint8_t *eqmask, uint32_t *status, uint32_t *qstatus, uint64_t *folded) {
  auto *oc = (OrcCircuit *)h;
  const CommonCircuitData &c = oc->c;
  if (n == 0) return 0;
  size_t bw = proofToBlob(proofFromBlob(c, blobs)).size();
  int r = c.num_challenges, Q = c.fri_config.num_query_rounds, nsteps = (int)c.fri_config.step_arity_bits.size();
  for (size_t p = 0; p < n; p++) {
    ProofWithPublicInputs pw = proofFromBlob(c, blobs + p * bw);
    size_t w = 0;
    auto next = [&]() { return F(challenges_in[(w++) * n + p]); };
    ProofChallenges ch;
    for (int i = 0; i < r; i++) ch.plonk_betas.push_back(next());
    for (int i = 0; i < r; i++) ch.plonk_gammas.push_back(next());
    for (int i = 0; i < r; i++) ch.plonk_alphas.push_back(next());
    if (c.num_lookup_polys > 0)
      for (int i = 0; i < r; i++) { LookupDelta d; d.lookup_A = next(); d.lookup_B = next(); d.lookup_alpha = next(); d.lookup_delta = next(); ch.plonk_deltas.push_back(d); }
    { F a = next(); F b = next(); ch.plonk_zeta = FExt(a, b); }
    { F a = next(); F b = next(); ch.fri_challenges.fri_alpha = FExt(a, b); }
    for (int s2 = 0; s2 < nsteps; s2++) { F a = next(); F b = next(); ch.fri_challenges.fri_betas.push_back(FExt(a, b)); }
    ch.fri_challenges.fri_pow_response = next();
    for (int q = 0; q < Q; q++) ch.fri_challenges.fri_query_indices.push_back((int)(next().v & (((uint64_t)1 << c.lde_bits()) - 1)));
    try {
      std::vector<FExt> comb;
      std::vector<bool> oks = checkCombinedPlonkEquations_(c, pw, ch, &comb);
      uint32_t mask = 0;
      for (size_t i = 0; i < oks.size(); i++) if (!oks[i]) mask |= 1u << i;
      std::vector<QueryTrace> qt;
      uint32_t fri = checkFRIProofStatus(c, oc->vk, pw.proof, ch, &qt);
      if (status) status[p] = mask ? (uint32_t)(P2V_ST_FALSE_EQS | (mask << 16)) : fri;
      if (eqmask) eqmask[p] = (uint8_t)(~mask & ((1u << r) - 1));
      if (combined) for (int j = 0; j < r; j++) { combined[(2 * j) * n + p] = comb[j].r.v; combined[(2 * j + 1) * n + p] = comb[j].i.v; }
      for (int q = 0; q < Q && q < (int)qt.size(); q++) {
        if (qstatus) qstatus[p * Q + q] = qt[q].status == 0 ? 0u : ((uint32_t)qt[q].status | ((uint32_t)q << 8) | ((uint32_t)qt[q].detail << 16));
        if (folded) { folded[p * Q + q] = qt[q].folded.r.v; folded[n * Q + p * Q + q] = qt[q].folded.i.v; }
      }
    } catch (const std::exception &e) {
      if (status) status[p] = 0xEE;
    }
  }
  return 0;
}

// Recomputed Merkle roots of every opening of every query round (north_star: "recomputed Merkle roots"):
// reconstructMerkleRoot' (Hash/Merkle.hs:30-37) for the four initial oracles (Plonk/FRI.hs:105-117, index = query index) and
// for every folding step (Plonk/FRI.hs:316, index = query_index_rev of the folded index), whatever the verdict is.
// roots_out SoA [(4+steps)*4][n*Q]: word i of tree t of (proof p, query q) at ((t*4+i) * n*Q + p*Q + q).
int orc_fri_roots(void *h, const uint64_t *blobs, size_t n, uint64_t *roots_out) {
  auto *oc = (OrcCircuit *)h;
  const CommonCircuitData &c = oc->c;
  if (n == 0) return 0;
  size_t bw = proofToBlob(proofFromBlob(c, blobs)).size();
  size_t Q = (size_t)c.fri_config.num_query_rounds;
  const std::vector<int> &ar = c.fri_config.step_arity_bits;
  for (size_t p = 0; p < n; p++) {
    ProofWithPublicInputs pw = proofFromBlob(c, blobs + p * bw);
    ProofChallenges ch = proofChallenges(c, oc->vk, pw);
    const FriProof &fp = pw.proof.opening_proof;
    for (size_t q = 0; q < Q; q++) {
      int idx = ch.fri_challenges.fri_query_indices[q];
      const FriQueryRound &rd = fp.query_round_proofs[q];
      auto put = [&](size_t t, const Digest &d) {
        for (int i = 0; i < 4; i++) roots_out[(t * 4 + i) * n * Q + p * Q + q] = d.e[i].v;
      };
      for (int o = 0; o < 4; o++) {
        const auto &ep = rd.initial_trees_proof.evals_proofs[o];
        put(o, reconstructMerkleRoot_(sponge(ep.first), idx, ep.second).second);
      }
      int qidx = idx;
      for (size_t s = 0; s < ar.size(); s++) {
        qidx >>= ar[s];
        put(4 + s, reconstructMerkleRoot_(sponge(flattenExt(rd.steps[s].evals)), qidx, rd.steps[s].merkle_proof).second);
      }
    }
  }
  return 0;
}

// What `testmain` prints for this proof (src/testmain.hs:40-63), line by line, with the reference's Show instances:
// Digest = derived Show over `show`-only Goldilocks ("MkDigest a b c d", Hash/Digest.hs:36-38, Algebra/Goldilocks.hs:90-91),
// Ext = "(re + X*im)" (Algebra/GoldilocksExt.hs:37-38), lists "[a,b]", Bool "True"/"False".  A proof that runs into one of
// the reference's `error` sites ends with the line "testmain: <message>" where GHC would print it on stderr and abort.
long long orc_testmain(void *h, const char *proof_json, size_t len, char *out, size_t cap) {
  auto *oc = (OrcCircuit *)h;
  try {
    ProofWithPublicInputs pw = js::proofFromJson(proof_json, len);
    std::string t;
    auto num = [](F x) { return std::to_string((unsigned long long)x.v); };
    Digest pih = sponge(pw.public_inputs);
    t += "public inputs hash = MkDigest " + num(pih.e[0]) + " " + num(pih.e[1]) + " " + num(pih.e[2]) + " " + num(pih.e[3]) + "\n";
    const OpeningSet &o = pw.proof.openings;
    auto cnt = [&](const char *name, size_t k) {
      std::string nm = std::string("# opening_") + name;
      while (nm.size() < 26) nm += ' ';
      t += nm + " = " + std::to_string(k) + "\n";
    };
    cnt("constants", o.constants.size()); cnt("plonk_sigmas", o.plonk_sigmas.size()); cnt("wires", o.wires.size());
    cnt("plonk_zs", o.plonk_zs.size()); cnt("plonk_zs_next", o.plonk_zs_next.size()); cnt("partial_products", o.partial_products.size());
    cnt("quotient_polys", o.quotient_polys.size()); cnt("lookup_zs", o.lookup_zs.size()); cnt("lookup_zs_next", o.lookup_zs_next.size());
    ProofChallenges ch = proofChallenges(oc->c, oc->vk, pw);
    std::vector<FExt> combined;
    std::vector<bool> oks = checkCombinedPlonkEquations_(oc->c, pw, ch, &combined);
    t += "[";
    for (size_t i = 0; i < combined.size(); i++) t += std::string(i ? "," : "") + "(" + num(combined[i].r) + " + X*" + num(combined[i].i) + ")";
    t += "]\n[";
    for (size_t i = 0; i < oks.size(); i++) t += std::string(i ? "," : "") + (oks[i] ? "True" : "False");
    t += "]\n";
    bool eqs = true;
    for (bool b : oks) eqs = eqs && b;
    if (!eqs) t += "proof verification result = False\n";
    else {
      uint32_t st = checkFRIProofStatus(oc->c, oc->vk, pw.proof, ch, nullptr);
      switch (st & 0xFF) {
        case P2V_ST_ACCEPT: t += "proof verification result = True\n"; break;
        case P2V_ST_FALSE_POW: case P2V_ST_FALSE_FINAL: t += "proof verification result = False\n"; break;
        case P2V_ST_ERR_INIT_MERKLE: t += "proof verification result = testmain: checkInitialTreeProofs: at least one Merkle proof failed\n"; break;
        case P2V_ST_ERR_STEP_MERKLE: t += "proof verification result = testmain: folding step Merkle proof does not check out\n"; break;
        case P2V_ST_ERR_STEP_EVAL: t += "proof verification result = testmain: folding step evaluation does not match the opening\n"; break;
        default: t += "proof verification result = testmain: error site " + std::to_string(st & 0xFF) + "\n";
      }
    }
    if (out) {
      if (t.size() + 1 > cap) throw std::runtime_error("output buffer too small");
      memcpy(out, t.c_str(), t.size() + 1);
    }
    return (long long)t.size();
  } catch (const std::exception &e) {
    orc_error = e.what();
    return -1;
  }
}

}  // extern "C"
