// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// CPU restatement of the gate constraint programs:
//   src/Gate/Constraints.hs:37-128        gateComputation (simple gates, ExponentiationGate)
//   src/Gate/Custom/Poseidon.hs:49-150    PoseidonMdsGate, PoseidonGate
//   src/Gate/Custom/CosetInterp.hs:51-121 CosetInterpolationGate
//   src/Gate/Custom/RandomAccess.hs:47-88 RandomAccessGate
//   src/Gate/Custom/Reducing.hs:28-60     ReducingGate / ReducingExtensionGate
//   src/Gate/Computation.hs:157-211       runStraightLine / evalConstraint (variable lookup)
//   src/Gate/Selector.hs:31-95            selector configuration and filter polynomials
//
// The reference builds a symbolic straight-line program (`Compute` monad -> `StraightLine`) and
// then interprets it over FExt (Gate/Computation.hs:157-164, Algebra/Expr.hs:130-138).  Field
// arithmetic is exact, so evaluating the same formulas directly over FExt gives identical values;
// this file does that, keeping the ORDER of the committed constraints (the order is part of the
// contract).  `Ext (Expr v)` values (wireExt, Gate/Vars.hs:56-57) become EExt.
// Parity: unpinned by the reference (its gate probes are commented out and have no expected
// values, Gate/Constraints.hs:133-150); cross-checked against oracle/pyref.py and by
// honest-witness tests (constraints vanish on correctly computed rows).
#pragma once
#include "types.hpp"

namespace orc {

// EvaluationVars, Gate/Computation.hs:177-184
struct EvaluationVars {
  std::vector<FExt> local_selectors, local_lkp_sels, local_constants, local_wires;
  std::vector<F> public_inputs_hash;
};

struct GateEval {
  const EvaluationVars &v;
  std::vector<FExt> out;  // committed constraints, in order
  explicit GateEval(const EvaluationVars &vars) : v(vars) {}

  // wire / cnst / hash, Gate/Vars.hs:49-53 + evalConstraint :200-211 (array `!` raises out of range)
  FExt wire(int i) const {
    if (i < 0 || (size_t)i >= v.local_wires.size()) throw std::runtime_error("wire index out of range");
    return v.local_wires[i];
  }
  FExt cnst(int i) const {
    if (i < 0 || (size_t)i >= v.local_constants.size()) throw std::runtime_error("constant index out of range");
    return v.local_constants[i];
  }
  FExt hash(int i) const { return fromBase(v.public_inputs_hash.at(i)); }
  EExt wireExt(int i) const { return EExt(wire(i), wire(i + 1)); }  // Gate/Vars.hs:56-57
  static FExt lit(F x) { return fromBase(x); }
  static FExt litI(long long k) { return fromBase(F::fromInt(k)); }

  void commit(const FExt &x) { out.push_back(x); }
  void commitExt(const EExt &x) { commit(x.r); commit(x.i); }  // Gate/Computation.hs:75-76
};

static inline FExt FE1() { return FExt(F(1), F(0)); }
static inline FExt FE0() { return FExt(F(0), F(0)); }

// sbox, Gate/Custom/Poseidon.hs:23-30
inline FExt gate_sbox(const FExt &x) {
  FExt x2 = x * x, x3 = x * x2, x4 = x2 * x2;
  return x3 * x4;
}

// poseidonMdsGateConstraints, Gate/Custom/Poseidon.hs:49-59
inline void poseidonMdsGate(GateEval &g) {
  for (int i = 0; i < 12; i++) {
    EExt result(FE0(), FE0());
    for (int j = 0; j < 12; j++) result = result + scaleEE(GateEval::lit(mdsMatrixCoeff(i, j)), g.wireExt(2 * j));
    g.commitExt(g.wireExt(2 * (i + 12)) - result);
  }
}

// poseidonGateConstraints, Gate/Custom/Poseidon.hs:63-150
inline void poseidonGate(GateEval &g) {
  auto input = [&](int i) { return g.wire(i); };
  auto output = [&](int i) { return g.wire(i + 12); };
  FExt swap_flag = g.wire(24);
  auto delta = [&](int i) { return g.wire(25 + i); };
  auto initial_sbox_in = [&](int r, int i) { return g.wire(29 + 12 * (r - 1) + i); };
  auto partial_sbox_in = [&](int r) { return g.wire(29 + 36 + r); };
  auto final_sbox_in = [&](int r, int i) { return g.wire(29 + 36 + 22 + 12 * r + i); };
  typedef std::vector<FExt> PS;
  auto mds = [&](const PS &st) {
    PS o(12);
    for (int i = 0; i < 12; i++) {
      FExt acc = FE0();
      for (int j = 0; j < 12; j++) acc = acc + GateEval::lit(mdsMatrixCoeff(i, j)) * st[j];
      o[i] = acc;
    }
    return o;
  };
  auto plus_rc = [&](int r, const PS &st) {
    PS o(12);
    for (int i = 0; i < 12; i++) o[i] = st[i] + GateEval::lit(F(ALL_ROUND_CONSTANTS[r][i]));
    return o;
  };
  // merkle swap  :66-76
  g.commit(swap_flag * (swap_flag - FE1()));
  for (int i = 0; i < 4; i++) g.commit(swap_flag * (input(i + 4) - input(i)) - delta(i));
  PS state(12);
  for (int i = 0; i < 4; i++) state[i] = input(i) + delta(i);
  for (int i = 4; i < 8; i++) state[i] = input(i) - delta(i - 4);
  for (int i = 8; i < 12; i++) state[i] = input(i);
  // initial full rounds  :79-89
  for (int r = 0; r < 4; r++) {
    PS s2 = plus_rc(r, state);
    PS s3 = s2;
    if (r != 0) {
      for (int i = 0; i < 12; i++) g.commit(s2[i] - initial_sbox_in(r, i));
      for (int i = 0; i < 12; i++) s3[i] = initial_sbox_in(r, i);
    }
    PS s4(12);
    for (int i = 0; i < 12; i++) s4[i] = gate_sbox(s3[i]);
    state = mds(s4);
  }
  // partial rounds  :92-100
  for (int i = 0; i < 12; i++) state[i] = state[i] + GateEval::lit(F(FAST_PARTIAL_FIRST_RC[i]));
  {  // mdsInitPartial :121-125
    PS t(12);
    t[0] = state[0];
    for (int i = 0; i < 11; i++) {
      FExt acc = FE0();
      for (int j = 0; j < 11; j++) acc = acc + GateEval::lit(partialMdsMatrixCoeff(i, j)) * state[j + 1];
      t[i + 1] = acc;
    }
    state = t;
  }
  for (int r = 0; r < 22; r++) {
    FExt sbox_in = partial_sbox_in(r);
    g.commit(state[0] - sbox_in);
    FExt y = gate_sbox(sbox_in);
    FExt z = r < 21 ? y + GateEval::lit(F(FAST_PARTIAL_RCS[r])) : y;
    state[0] = z;
    // mdsFastPartial :126-131
    FExt d = state[0] * GateEval::lit(mdsMatrixCoeff(0, 0));
    for (int i = 0; i < 11; i++) d = d + state[i + 1] * GateEval::lit(F(FAST_PARTIAL_W_HATS[r][i]));
    PS res(12);
    res[0] = d;
    for (int i = 0; i < 11; i++) res[i + 1] = state[i + 1] + state[0] * GateEval::lit(F(FAST_PARTIAL_VS[r][i]));
    state = res;
  }
  // final full rounds  :103-108
  for (int r = 0; r < 4; r++) {
    PS s2 = plus_rc(r + 26, state);
    for (int i = 0; i < 12; i++) g.commit(s2[i] - final_sbox_in(r, i));
    PS s3(12);
    for (int i = 0; i < 12; i++) s3[i] = gate_sbox(final_sbox_in(r, i));
    state = mds(s3);
  }
  for (int i = 0; i < 12; i++) g.commit(state[i] - output(i));
}

// cosetInterpolationGateConstraints, Gate/Custom/CosetInterp.hs:51-121
inline void cosetInterpolationGate(GateEval &g, int subgroup_bits, int degree, const std::vector<F> &weights) {
  int n_points = 1 << subgroup_bits;
  if (degree < 2) throw std::runtime_error("cosetInterpolationGate: degree < 2");
  int n_intermediates = (n_points - 2) / (degree - 1);
  std::vector<F> domain = enumerateSubgroup(subgroup_bits);
  FExt coset_shift = g.wire(0);
  auto poly_value = [&](int k) { return g.wireExt(1 + 2 * k); };
  EExt eval_loc = g.wireExt(1 + 2 * n_points);
  EExt eval_result = g.wireExt(1 + 2 * n_points + 2);
  auto tmp_eval = [&](int i) { return g.wireExt(1 + 2 * (n_points + 2) + 2 * i); };
  auto tmp_prod = [&](int i) { return g.wireExt(1 + 2 * (n_points + 2) + 2 * (n_intermediates + i)); };
  EExt shifted_loc = g.wireExt(1 + 2 * (n_points + 2) + 4 * n_intermediates);

  g.commitExt(eval_loc - scaleEE(coset_shift, shifted_loc));
  // chunk xs = take degree xs : partition (degree-1) (drop degree xs)   :121
  struct Chunk { int start, len; };
  std::vector<Chunk> chunks;
  auto chunkOf = [&](int total) {
    std::vector<Chunk> cs;
    cs.push_back(Chunk{0, total < degree ? total : degree});
    for (int pos = degree; pos < total; pos += degree - 1) cs.push_back(Chunk{pos, std::min(degree - 1, total - pos)});
    return cs;
  };
  // zip3 chunked_domain chunked_values chunked_weights: truncates to the shortest
  std::vector<Chunk> cd = chunkOf(n_points), cw = chunkOf((int)weights.size());
  size_t nchunks = std::min(cd.size(), cw.size());
  // initials = initial : [(tmp_eval i, tmp_prod i) | i <- [0..n_int-1]]; stuff = zipWith worker initials chunks
  size_t nstuff = std::min(nchunks, (size_t)n_intermediates + 1);
  std::vector<std::pair<EExt, EExt>> stuff;
  for (size_t c = 0; c < nstuff; c++) {
    EExt eval = c == 0 ? EExt(FE0(), FE0()) : tmp_eval((int)c - 1);
    EExt prod = c == 0 ? EExt(FE1(), FE0()) : tmp_prod((int)c - 1);
    int len = std::min(cd[c].len, cw[c].len);  // zipWith scaleExt weights values; zip weighted domain
    for (int k = 0; k < len; k++) {
      int idx = cd[c].start + k;
      EExt val = scaleEE(GateEval::lit(weights[cw[c].start + k]), poly_value(idx));
      EExt term = shifted_loc - fromBaseEE(GateEval::lit(domain[idx]));
      EExt next_eval = term * eval + val * prod;
      EExt next_prod = term * prod;
      eval = next_eval;
      prod = next_prod;
    }
    stuff.push_back({eval, prod});
  }
  if (stuff.empty()) throw std::runtime_error("cosetInterpolationGate: last of empty list");
  for (size_t i = 0; i + 1 < stuff.size(); i++) {
    g.commitExt(tmp_eval((int)i) - stuff[i].first);
    g.commitExt(tmp_prod((int)i) - stuff[i].second);
  }
  g.commitExt(eval_result - stuff.back().first);
}

// randomAccessGateConstraints, Gate/Custom/RandomAccess.hs:47-88
inline void randomAccessGate(GateEval &g, int num_bits, int num_copies, int num_extra) {
  int veclen = 1 << num_bits, width = 2 + veclen;
  int bits_start_at = width * num_copies + num_extra;
  for (int k = 0; k < num_copies; k++) {
    auto bits = [&](int j) { return g.wire(bits_start_at + k * num_bits + j); };
    for (int j = 0; j < num_bits; j++) g.commit(bits(j) * (bits(j) - FE1()));
    // foldr (\b acc -> 2*acc + b) 0 bits
    FExt reconstr = FE0();
    for (int j = num_bits - 1; j >= 0; j--) reconstr = GateEval::litI(2) * reconstr + bits(j);
    g.commit(reconstr - g.wire(k * width + 0));
    std::vector<FExt> vals;
    for (int i = 0; i < veclen; i++) vals.push_back(g.wire(k * width + 2 + i));
    for (int j = 0; j < num_bits; j++) {
      std::vector<FExt> nxt;
      FExt b = bits(j);
      for (size_t i = 0; i + 1 < vals.size(); i += 2) nxt.push_back(vals[i] + b * (vals[i + 1] - vals[i]));
      vals = nxt;
    }
    g.commit(vals[0] - g.wire(k * width + 1));
  }
  for (int j = 0; j < num_extra; j++) g.commit(g.cnst(j) - g.wire(num_copies * width + j));
}

// reducingGateConstraints / reducingExtensionGateConstraints, Gate/Custom/Reducing.hs:28-60
inline void reducingGate(GateEval &g, int n, bool ext) {
  EExt output = g.wireExt(0), alpha = g.wireExt(2), initial = g.wireExt(4);
  auto accum = [&](int i) { return i < n - 1 ? g.wireExt(6 + (ext ? 2 * n : n) + 2 * i) : output; };
  for (int i = 0; i < n; i++) {
    EExt prev = i == 0 ? initial : accum(i - 1);
    EExt coeff = ext ? g.wireExt(6 + 2 * i) : fromBaseEE(g.wire(6 + i));
    g.commitExt(prev * alpha + coeff - accum(i));
  }
}

// exponentiationGateConstraints, Gate/Constraints.hs:114-128
inline void exponentiationGate(GateEval &g, int n) {
  FExt base = g.wire(0);
  auto exp_bit = [&](int i) { return g.wire(i + 1); };
  auto tmp_val = [&](int i) { return g.wire(n + 2 + i); };
  for (int i = 0; i < n; i++) {
    FExt prev = i == 0 ? FE1() : tmp_val(i - 1) * tmp_val(i - 1);
    FExt cur = exp_bit(n - 1 - i);
    g.commit(prev * (cur * base + (FE1() - cur)) - tmp_val(i));
  }
  g.commit(g.wire(n + 1) - tmp_val(n - 1));
}

// gateComputation, Gate/Constraints.hs:40-108: unfiltered constraint vector of one gate
inline std::vector<FExt> gateConstraints(const Gate &gate, const EvaluationVars &vars) {
  GateEval g(vars);
  switch (gate.kind) {
    case P2V_GATE_ARITHMETIC:  // :45-46
      for (int i = 0; i < gate.p0; i++) {
        int j = 4 * i;
        g.commit(g.wire(j + 3) - g.cnst(0) * g.wire(j) * g.wire(j + 1) - g.cnst(1) * g.wire(j + 2));
      }
      break;
    case P2V_GATE_ARITHMETIC_EXT:  // :49-54
      for (int i = 0; i < gate.p0; i++) {
        int j = 8 * i;
        EExt c0 = fromBaseEE(g.cnst(0)), c1 = fromBaseEE(g.cnst(1));
        g.commitExt(g.wireExt(j + 6) - c0 * g.wireExt(j) * g.wireExt(j + 2) - c1 * g.wireExt(j + 4));
      }
      break;
    case P2V_GATE_BASE_SUM: {  // :57-62
      int L = gate.p0, B = gate.p1;
      auto limb = [&](int i) { return g.wire(i + 1); };
      // horner = go 0 where go k = if k < L-1 then limb k + B * go (k+1) else limb k
      FExt h = limb(L - 1);
      for (int k = L - 2; k >= 0; k--) h = limb(k) + GateEval::litI(B) * h;
      g.commit(h - g.wire(0));
      for (int i = 0; i < L; i++) {
        FExt prod = FE1();
        for (int k = 0; k < B; k++) prod = prod * (limb(i) - GateEval::litI(k));
        g.commit(prod);
      }
      break;
    }
    case P2V_GATE_COSET_INTERP: cosetInterpolationGate(g, gate.p0, gate.p1, gate.weights); break;
    case P2V_GATE_CONSTANT:  // :68-69
      for (int i = 0; i < gate.p0; i++) g.commit(g.cnst(i) - g.wire(i));
      break;
    case P2V_GATE_EXPONENTIATION: exponentiationGate(g, gate.p0); break;
    case P2V_GATE_LOOKUP: case P2V_GATE_LOOKUP_TABLE: case P2V_GATE_NOOP: break;  // :76-77,85
    case P2V_GATE_MUL_EXT:  // :80-83
      for (int i = 0; i < gate.p0; i++) {
        int j = 6 * i;
        g.commitExt(g.wireExt(j + 4) - fromBaseEE(g.cnst(0)) * g.wireExt(j) * g.wireExt(j + 2));
      }
      break;
    case P2V_GATE_PUBLIC_INPUT:  // :88-89
      for (int i = 0; i < 4; i++) g.commit(g.wire(i) - g.hash(i));
      break;
    case P2V_GATE_POSEIDON:
      if (gate.p0 != 12) throw std::runtime_error("gateConstraints/PoseidonGate: unsupported width");
      poseidonGate(g);
      break;
    case P2V_GATE_POSEIDON_MDS:
      if (gate.p0 != 12) throw std::runtime_error("gateConstraints/PoseidonMdsGate: unsupported width");
      poseidonMdsGate(g);
      break;
    case P2V_GATE_RANDOM_ACCESS: randomAccessGate(g, gate.p0, gate.p1, gate.p2); break;
    case P2V_GATE_REDUCING: reducingGate(g, gate.p0, false); break;
    case P2V_GATE_REDUCING_EXT: reducingGate(g, gate.p0, true); break;
    default: throw std::runtime_error("gateConstraints: unknown gate");
  }
  return g.out;
}

// ---- selectors, Gate/Selector.hs ---------------------------------------------------------
struct SelectorConfig { int numGateSelectors, numLookupSelectors, numGateConstants, numSigmaColumns; };
// getSelectorConfig :31-47
inline SelectorConfig getSelectorConfig(const CommonCircuitData &c) {
  int nluts = (int)c.luts.size();
  int expected_lookup_sels = nluts == 0 ? 0 : 4 + nluts;
  int num_gate_selectors = (int)c.selector_groups.size();
  if (c.num_lookup_selectors != expected_lookup_sels)
    throw std::runtime_error("getSelectorConfig: fatal: num_lookup_selectors /= (4 + #nluts)");
  if (c.num_constants != num_gate_selectors + c.num_lookup_selectors + c.config_num_constants)
    throw std::runtime_error("getSelectorConfig: fatal: constant columns tally does not add up!");
  return SelectorConfig{num_gate_selectors, c.num_lookup_selectors, c.config_num_constants, c.num_routed_wires};
}
struct ConstantColumns { std::vector<FExt> gateSelectors, lookupSelectors, gateConstants; };
// splitConstantColumns :62-74
inline ConstantColumns splitConstantColumns(const SelectorConfig &sc, const std::vector<FExt> &xs) {
  ConstantColumns cc;
  size_t pos = 0;
  auto take = [&](int n, std::vector<FExt> &dst) {
    for (int i = 0; i < n && pos < xs.size(); i++) dst.push_back(xs[pos++]);
  };
  take(sc.numGateSelectors, cc.gateSelectors);
  take(sc.numLookupSelectors, cc.lookupSelectors);
  take(sc.numGateConstants, cc.gateConstants);
  if (pos != xs.size()) throw std::runtime_error("splitConstantColumns: fatal: numbers do not add up");
  if ((int)cc.gateConstants.size() != sc.numGateConstants)
    throw std::runtime_error("splitConstantColumns: fatal: not enough constant columns");
  return cc;
}
// evalGateSelectorPoly :83-89, evalGateSelectors :93-95
inline std::vector<FExt> evalGateSelectors(const CommonCircuitData &c, const std::vector<FExt> &xs) {
  std::vector<FExt> values;
  FExt unused = fromBase(F((u64)0xFFFFFFFFULL));  // 2^32 - 1
  for (size_t k = 0; k < c.selector_indices.size(); k++) {
    int grp = c.selector_indices[k];
    FExt x = xs.at(grp);
    Range range = c.selector_groups.at(grp);
    FExt value = c.selector_groups.size() > 1 ? unused - x : FE1();
    for (int j = range.start; j < range.end; j++)
      if (j != (int)k) value = value * (fromBase(F::fromInt(j)) - x);
    values.push_back(value);
  }
  return values;
}

}  // namespace orc
