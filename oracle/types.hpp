// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// Records mirroring src/Types.hs (CommonCircuitData :47-70, CircuitConfig :73-87, SelectorsInfo
// :97-101, FriConfig :116-123, FriParams :151-173, FriProof :176-185, FriQueryRound :187-211,
// VerifierOnlyCircuitData :236-240, ProofWithPublicInputs/Proof/OpeningSet :251-279) and the Gate
// ADT of src/Gate/Base.hs:27-45, plus conversion from the flat (shape, blob) form of
// include/p2v.h.  The layout computation here is written independently of the product's
// (plonky2-verifier_b200/csrc/host) and the tests compare the two.
#pragma once
#include <string>
#include "../include/p2v.h"
#include "hash.hpp"

namespace orc {

struct Gate {
  int kind = P2V_GATE_NOOP;
  int p0 = 0, p1 = 0, p2 = 0;
  std::vector<F> weights;  // CosetInterpolationGate barycentric_weights
};

struct Range { int start, end; };  // [start,end)  Misc/Aux.hs:146-155

struct FriConfig {
  int rate_bits, cap_height, proof_of_work_bits, num_query_rounds;
  // reduction strategy already expanded by expandReductionStrategy (Plonk/FRI.hs:337-354)
  std::vector<int> step_arity_bits;
};

struct CommonCircuitData {
  // CircuitConfig
  int num_wires, num_routed_wires, config_num_constants, num_challenges;
  FriConfig fri_config;
  int degree_bits;
  std::vector<Gate> gates;
  std::vector<int> selector_indices;
  std::vector<Range> selector_groups;
  int quotient_degree_factor, num_constants, num_public_inputs;
  std::vector<F> k_is;
  int num_partial_products, num_lookup_polys, num_lookup_selectors;
  std::vector<std::vector<std::pair<F, F>>> luts;
  int nrows() const { return 1 << degree_bits; }
  int lde_bits() const { return degree_bits + fri_config.rate_bits; }
};

struct OpeningSet {
  std::vector<FExt> constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys,
      lookup_zs, lookup_zs_next;
};
struct FriInitialTreeProof { std::vector<std::pair<std::vector<F>, MerkleProof>> evals_proofs; };
struct FriQueryStep { std::vector<FExt> evals; MerkleProof merkle_proof; };
struct FriQueryRound { FriInitialTreeProof initial_trees_proof; std::vector<FriQueryStep> steps; };
struct FriProof {
  std::vector<MerkleCap> commit_phase_merkle_caps;
  std::vector<FriQueryRound> query_round_proofs;
  std::vector<FExt> final_poly;
  F pow_witness;
};
struct Proof {
  MerkleCap wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap;
  OpeningSet openings;
  FriProof opening_proof;
};
struct ProofWithPublicInputs { Proof proof; std::vector<F> public_inputs; };
struct VerifierOnlyCircuitData { MerkleCap constants_sigmas_cap; Digest circuit_digest; };

// ---------------------------------------------------------------------------------------
inline CommonCircuitData commonFromShape(const p2v_shape &s) {
  CommonCircuitData c;
  c.num_wires = s.num_wires;
  c.num_routed_wires = s.num_routed_wires;
  c.config_num_constants = s.num_gate_constants;
  c.num_challenges = s.num_challenges;
  c.fri_config.rate_bits = s.rate_bits;
  c.fri_config.cap_height = s.cap_height;
  c.fri_config.proof_of_work_bits = s.pow_bits;
  c.fri_config.num_query_rounds = s.num_queries;
  for (int i = 0; i < s.num_steps; i++) c.fri_config.step_arity_bits.push_back(s.step_arity_bits[i]);
  c.degree_bits = s.degree_bits;
  for (int g = 0; g < s.num_gates; g++) {
    Gate gt;
    gt.kind = s.gates[g].kind; gt.p0 = s.gates[g].p0; gt.p1 = s.gates[g].p1; gt.p2 = s.gates[g].p2;
    for (int k = 0; k < s.gates[g].weights_len; k++) gt.weights.push_back(F(s.weights[s.gates[g].weights_off + k]));
    c.gates.push_back(gt);
    c.selector_indices.push_back(s.gates[g].group);
  }
  for (int g = 0; g < s.num_groups; g++) c.selector_groups.push_back(Range{s.group_start[g], s.group_end[g]});
  c.quotient_degree_factor = s.quotient_degree_factor;
  c.num_constants = s.num_constants;
  c.num_public_inputs = s.num_public_inputs;
  for (int i = 0; i < s.num_routed_wires; i++) c.k_is.push_back(F(s.k_is[i]));
  c.num_partial_products = s.num_partial_products;
  c.num_lookup_polys = s.num_lookup_polys;
  c.num_lookup_selectors = s.num_lookup_selectors;
  for (int l = 0; l < s.num_luts; l++) {
    std::vector<std::pair<F, F>> t;
    for (int k = s.lut_off[l]; k < s.lut_off[l + 1]; k++) t.push_back({F(s.lut_pairs[2 * k]), F(s.lut_pairs[2 * k + 1])});
    c.luts.push_back(t);
  }
  return c;
}

// oracleWidths, Plonk/FRI.hs:56-65
inline std::array<int, 4> oracleWidths(const CommonCircuitData &c) {
  int r = c.num_challenges;
  return {c.num_constants + c.num_routed_wires, c.num_wires, r * (1 + c.num_partial_products + c.num_lookup_polys),
          r * c.quotient_degree_factor};
}

// Sequential reader over the flat blob (field order of Types.hs:251-279, see p2v_layout doc).
struct BlobReader {
  const uint64_t *p;
  size_t pos = 0;
  explicit BlobReader(const uint64_t *q) : p(q) {}
  F f() { return F(p[pos++]); }
  FExt e() { F a = f(); F b = f(); return FExt(a, b); }
  Digest d() { Digest x; for (int i = 0; i < 4; i++) x.e[i] = f(); return x; }
  MerkleCap cap(int n) { MerkleCap c; for (int i = 0; i < n; i++) c.roots.push_back(d()); return c; }
  std::vector<FExt> exts(int n) { std::vector<FExt> v; for (int i = 0; i < n; i++) v.push_back(e()); return v; }
  std::vector<F> fs(int n) { std::vector<F> v; for (int i = 0; i < n; i++) v.push_back(f()); return v; }
  MerkleProof path(int n) { MerkleProof m; for (int i = 0; i < n; i++) m.siblings.push_back(d()); return m; }
};

inline size_t blobWords(const CommonCircuitData &c);

inline ProofWithPublicInputs proofFromBlob(const CommonCircuitData &c, const uint64_t *blob) {
  BlobReader rd(blob);
  ProofWithPublicInputs pw;
  Proof &p = pw.proof;
  int ncap = 1 << c.fri_config.cap_height;
  int r = c.num_challenges;
  p.wires_cap = rd.cap(ncap);
  p.plonk_zs_partial_products_cap = rd.cap(ncap);
  p.quotient_polys_cap = rd.cap(ncap);
  OpeningSet &o = p.openings;
  o.constants = rd.exts(c.num_constants);
  o.plonk_sigmas = rd.exts(c.num_routed_wires);
  o.wires = rd.exts(c.num_wires);
  o.plonk_zs = rd.exts(r);
  o.plonk_zs_next = rd.exts(r);
  o.partial_products = rd.exts(r * c.num_partial_products);
  o.quotient_polys = rd.exts(r * c.quotient_degree_factor);
  o.lookup_zs = rd.exts(r * c.num_lookup_polys);
  o.lookup_zs_next = rd.exts(r * c.num_lookup_polys);
  FriProof &fp = p.opening_proof;
  int nsteps = (int)c.fri_config.step_arity_bits.size();
  for (int s = 0; s < nsteps; s++) fp.commit_phase_merkle_caps.push_back(rd.cap(ncap));
  int total_arity = 0;
  for (int a : c.fri_config.step_arity_bits) total_arity += a;
  fp.final_poly = rd.exts(1 << (c.degree_bits - total_arity));
  fp.pow_witness = rd.f();
  pw.public_inputs = rd.fs(c.num_public_inputs);
  auto widths = oracleWidths(c);
  int init_len = c.lde_bits() - c.fri_config.cap_height;
  for (int q = 0; q < c.fri_config.num_query_rounds; q++) {
    FriQueryRound qr;
    for (int orc_i = 0; orc_i < 4; orc_i++) {
      std::vector<F> leaf = rd.fs(widths[orc_i]);
      MerkleProof mp = rd.path(init_len);
      qr.initial_trees_proof.evals_proofs.push_back({leaf, mp});
    }
    int bits = c.lde_bits();
    for (int s = 0; s < nsteps; s++) {
      int a = c.fri_config.step_arity_bits[s];
      FriQueryStep st;
      st.evals = rd.exts(1 << a);
      bits -= a;
      int plen = bits - c.fri_config.cap_height;
      if (plen < 0) plen = 0;
      st.merkle_proof = rd.path(plen);
      qr.steps.push_back(st);
    }
    fp.query_round_proofs.push_back(qr);
  }
  return pw;
}

inline void blobPutF(std::vector<uint64_t> &b, F x) { b.push_back(x.v); }
inline void blobPutE(std::vector<uint64_t> &b, const FExt &x) { b.push_back(x.r.v); b.push_back(x.i.v); }
inline void blobPutD(std::vector<uint64_t> &b, const Digest &d) { for (int i = 0; i < 4; i++) b.push_back(d.e[i].v); }
inline std::vector<uint64_t> proofToBlob(const ProofWithPublicInputs &pw) {
  std::vector<uint64_t> b;
  const Proof &p = pw.proof;
  for (auto &d : p.wires_cap.roots) blobPutD(b, d);
  for (auto &d : p.plonk_zs_partial_products_cap.roots) blobPutD(b, d);
  for (auto &d : p.quotient_polys_cap.roots) blobPutD(b, d);
  const OpeningSet &o = p.openings;
  for (auto *v : {&o.constants, &o.plonk_sigmas, &o.wires, &o.plonk_zs, &o.plonk_zs_next, &o.partial_products,
                  &o.quotient_polys, &o.lookup_zs, &o.lookup_zs_next})
    for (auto &x : *v) blobPutE(b, x);
  for (auto &c : p.opening_proof.commit_phase_merkle_caps)
    for (auto &d : c.roots) blobPutD(b, d);
  for (auto &x : p.opening_proof.final_poly) blobPutE(b, x);
  blobPutF(b, p.opening_proof.pow_witness);
  for (auto &x : pw.public_inputs) blobPutF(b, x);
  for (auto &qr : p.opening_proof.query_round_proofs) {
    for (auto &ep : qr.initial_trees_proof.evals_proofs) {
      for (auto &x : ep.first) blobPutF(b, x);
      for (auto &d : ep.second.siblings) blobPutD(b, d);
    }
    for (auto &st : qr.steps) {
      for (auto &x : st.evals) blobPutE(b, x);
      for (auto &d : st.merkle_proof.siblings) blobPutD(b, d);
    }
  }
  return b;
}

inline VerifierOnlyCircuitData vkeyFromWords(const CommonCircuitData &c, const uint64_t *w) {
  BlobReader rd(w);
  VerifierOnlyCircuitData v;
  v.constants_sigmas_cap = rd.cap(1 << c.fri_config.cap_height);
  v.circuit_digest = rd.d();
  return v;
}

}  // namespace orc
