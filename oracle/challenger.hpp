// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// CPU restatement of the Fiat-Shamir layer:
//   src/Challenge/Pure.hs      duplex sponge state machine (:27-107)
//   src/Challenge/FRI.hs       toFriOpenings :46-61, friChallenges :65-104
//   src/Challenge/Verifier.hs  mkLookupDeltaList :36-40, proofChallenges :58-103
// Parity: unpinned by the reference (duplexTest prints but has no expected values); pinned here
// by SURVEY.md App. I vectors, the Python twin, and the 114-permutation count at the standard
// recursion shape (commentary/FRI.md:263).
#pragma once
#include "types.hpp"

namespace orc {

// data DuplexState = Absorbing old inp | Squeezing old out   (Challenge/Pure.hs:27-30)
struct Duplex {
  bool absorbing = true;
  State old;
  std::vector<F> buf;  // Absorbing: pending inputs (<= 8); Squeezing: remaining outputs (front = next)

  explicit Duplex(const State &s) : old(s) {}  // duplexInitialState :32-33

  static State overwrite(const std::vector<F> &inp, State st) {  // :35-36
    for (size_t i = 0; i < inp.size(); i++) st[i] = inp[i];
    return st;
  }
  static std::vector<F> extract(const State &st) {  // reverse $ take 8  :41-43
    std::vector<F> out;
    for (int i = 7; i >= 0; i--) out.push_back(st[i]);
    return out;
  }
  void freshSqueezing(const State &st) { absorbing = false; old = st; buf = extract(st); }  // :45-46

  // absorbFelt :50-58
  void absorbFelt(F x) {
    if (!absorbing) { absorbing = true; buf.clear(); }  // Squeezing old _ -> Absorbing old []
    if (buf.size() < 8) { buf.push_back(x); return; }
    old = perm(overwrite(buf, old));  // duplex inp old
    buf.clear();
    buf.push_back(x);
  }
  // squeezeFelt :60-69
  F squeezeFelt() {
    for (;;) {
      if (!absorbing) {
        if (buf.empty()) { freshSqueezing(perm(old)); continue; }
        F y = buf.front();
        buf.erase(buf.begin());
        return y;
      }
      if (buf.empty()) freshSqueezing(perm(old));
      else freshSqueezing(perm(overwrite(buf, old)));
    }
  }
  // Absorb instances :73-91
  void absorb(F x) { absorbFelt(x); }
  void absorb(const FExt &x) { absorbFelt(x.r); absorbFelt(x.i); }
  void absorb(const Digest &d) { for (int i = 0; i < 4; i++) absorbFelt(d.e[i]); }
  void absorb(const MerkleCap &c) { for (auto &d : c.roots) absorb(d); }
  template <class T> void absorb(const std::vector<T> &xs) { for (auto &x : xs) absorb(x); }
  // Squeeze instances :95-107
  FExt squeezeExt() { F a = squeezeFelt(); F b = squeezeFelt(); return FExt(a, b); }
  std::vector<F> squeezeN(int n) { std::vector<F> v; for (int i = 0; i < n; i++) v.push_back(squeezeFelt()); return v; }
};

// data FriChallenges, Challenge/FRI.hs:24-30
struct FriChallenges {
  FExt fri_alpha;
  std::vector<FExt> fri_betas;
  F fri_pow_response;
  std::vector<int> fri_query_indices;
};
// data LookupDelta, Challenge/Verifier.hs:20-27
struct LookupDelta { F lookup_A, lookup_B, lookup_alpha, lookup_delta; };
// data ProofChallenges, Challenge/Verifier.hs:45-53
struct ProofChallenges {
  std::vector<F> plonk_betas, plonk_gammas, plonk_alphas;
  std::vector<LookupDelta> plonk_deltas;
  FExt plonk_zeta;
  FriChallenges fri_challenges;
};

// toFriOpenings, Challenge/FRI.hs:46-61
struct FriOpenings { std::vector<FExt> batch_this, batch_next; };
inline FriOpenings toFriOpenings(const OpeningSet &o) {
  FriOpenings f;
  for (auto *v : {&o.constants, &o.plonk_sigmas, &o.wires, &o.plonk_zs, &o.partial_products, &o.quotient_polys, &o.lookup_zs})
    f.batch_this.insert(f.batch_this.end(), v->begin(), v->end());
  for (auto *v : {&o.plonk_zs_next, &o.lookup_zs_next}) f.batch_next.insert(f.batch_next.end(), v->begin(), v->end());
  return f;
}

// friChallenges, Challenge/FRI.hs:65-104
inline FriChallenges friChallenges(Duplex &dx, const CommonCircuitData &common, const Proof &proof) {
  FriChallenges fc;
  FriOpenings fo = toFriOpenings(proof.openings);
  dx.absorb(fo.batch_this);
  dx.absorb(fo.batch_next);
  fc.fri_alpha = dx.squeezeExt();
  for (auto &cap : proof.opening_proof.commit_phase_merkle_caps) {
    dx.absorb(cap);
    fc.fri_betas.push_back(dx.squeezeExt());
  }
  dx.absorb(proof.opening_proof.final_poly);
  dx.absorb(proof.opening_proof.pow_witness);
  fc.fri_pow_response = dx.squeezeFelt();
  u64 lde_size = (u64)1 << (common.degree_bits + common.fri_config.rate_bits);
  for (int i = 0; i < common.fri_config.num_query_rounds; i++) {
    F felt = dx.squeezeFelt();
    fc.fri_query_indices.push_back((int)(felt.v % lde_size));  // mod (asInteger felt) lde_size :95-97
  }
  return fc;
}

// mkLookupDeltaList, Challenge/Verifier.hs:36-40
inline std::vector<LookupDelta> mkLookupDeltaList(const std::vector<F> &fs) {
  std::vector<LookupDelta> out;
  for (size_t i = 0; i < fs.size(); i += 4) {
    if (i + 4 > fs.size()) throw std::runtime_error("mkLookupDelta: expecting 4 field elements");
    out.push_back(LookupDelta{fs[i], fs[i + 1], fs[i + 2], fs[i + 3]});
  }
  return out;
}

// proofChallenges, Challenge/Verifier.hs:58-103
inline ProofChallenges proofChallenges(const CommonCircuitData &common, const VerifierOnlyCircuitData &vd,
                                       const ProofWithPublicInputs &pw) {
  ProofChallenges pc;
  int r = common.num_challenges;
  bool has_lookup = common.num_lookup_polys > 0;
  Digest public_inputs_hash = sponge(pw.public_inputs);
  Duplex dx(zeroState());
  dx.absorb(vd.circuit_digest);
  dx.absorb(public_inputs_hash);
  dx.absorb(pw.proof.wires_cap);
  pc.plonk_betas = dx.squeezeN(r);
  pc.plonk_gammas = dx.squeezeN(r);
  if (has_lookup) {
    std::vector<F> deltas = dx.squeezeN(2 * r);
    std::vector<F> all = pc.plonk_betas;
    all.insert(all.end(), pc.plonk_gammas.begin(), pc.plonk_gammas.end());
    all.insert(all.end(), deltas.begin(), deltas.end());
    pc.plonk_deltas = mkLookupDeltaList(all);
  }
  dx.absorb(pw.proof.plonk_zs_partial_products_cap);
  pc.plonk_alphas = dx.squeezeN(r);
  dx.absorb(pw.proof.quotient_polys_cap);
  pc.plonk_zeta = dx.squeezeExt();
  pc.fri_challenges = friChallenges(dx, common, pw.proof);
  return pc;
}

// Flat form shared with p2v_challenges (include/p2v.h): betas[r], gammas[r], alphas[r],
// deltas[4r or 0], zeta[2], fri_alpha[2], fri_betas[2*steps], pow_response, query_indices[Q]
inline std::vector<uint64_t> challengesToWords(const ProofChallenges &pc) {
  std::vector<uint64_t> w;
  for (auto &x : pc.plonk_betas) w.push_back(x.v);
  for (auto &x : pc.plonk_gammas) w.push_back(x.v);
  for (auto &x : pc.plonk_alphas) w.push_back(x.v);
  for (auto &d : pc.plonk_deltas) { w.push_back(d.lookup_A.v); w.push_back(d.lookup_B.v); w.push_back(d.lookup_alpha.v); w.push_back(d.lookup_delta.v); }
  w.push_back(pc.plonk_zeta.r.v); w.push_back(pc.plonk_zeta.i.v);
  w.push_back(pc.fri_challenges.fri_alpha.r.v); w.push_back(pc.fri_challenges.fri_alpha.i.v);
  for (auto &b : pc.fri_challenges.fri_betas) { w.push_back(b.r.v); w.push_back(b.i.v); }
  w.push_back(pc.fri_challenges.fri_pow_response.v);
  for (int i : pc.fri_challenges.fri_query_indices) w.push_back((uint64_t)i);
  return w;
}

}  // namespace orc
