// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// CPU restatement of the Plonk layer:
//   src/Plonk/Vanishing.hs:48-143   evalCombinedPlonkConstraints / evalAllPlonkConstraints
//   src/Plonk/Lookups.hs:45-132     evalLookupEquations
//   src/Plonk/Verifier.hs:31-65     checkCombinedPlonkEquations', verifyProof
//   src/Plonk/FRI.hs:56-407         oracle widths, initial trees, combineInitial, PoW, folding, checkFRIProof
// Verdicts follow SURVEY.md App. E: `verifyProof` is True / False / `error`; the status word
// distinguishes which, in the reference's (lazy) evaluation order.
// Parity: unpinned by the reference (no fixtures, no expected values); cross-checked against the
// independent Python twin (oracle/pyref.py) and self-consistent with the fixture prover.
#pragma once
#include <algorithm>
#include "challenger.hpp"
#include "gates.hpp"

namespace orc {

// ---- Misc/Aux.hs helpers -------------------------------------------------------------------
template <class T> inline std::vector<std::vector<T>> partitionK(int k, const std::vector<T> &xs) {  // :109-113
  std::vector<std::vector<T>> out;
  if (xs.empty()) return out;
  if (k <= 0) throw std::runtime_error("partition: non-positive chunk size on a non-empty list (diverges in the reference)");
  for (size_t pos = 0; pos < xs.size(); pos += k)
    out.emplace_back(xs.begin() + pos, xs.begin() + std::min(xs.size(), pos + (size_t)k));
  return out;
}
inline int divCeil(int n, int k) { return (n + k - 1) / k; }

// ---- Plonk/Lookups.hs ----------------------------------------------------------------------
// lookupSelectorIndex :35-41
enum { LS_TransSre = 0, LS_TransLdc = 1, LS_InitSre = 2, LS_LastLdc = 3, LS_StartEnd = 4 };

// evalLookupEquations :45-132
inline std::vector<FExt> evalLookupEquations(const CommonCircuitData &c, const std::vector<FExt> &lkpSels,
                                             const OpeningSet &o, const ProofChallenges &ch) {
  auto selector = [&](int idx) { return lkpSels.at(idx); };
  std::vector<std::pair<FExt, FExt>> zipped;
  for (size_t i = 0; i < std::min(o.lookup_zs.size(), o.lookup_zs_next.size()); i++)
    zipped.push_back({o.lookup_zs[i], o.lookup_zs_next[i]});
  auto roundChunks = partitionK(c.num_lookup_polys, zipped);
  if (roundChunks.size() != ch.plonk_deltas.size()) throw std::runtime_error("safeZipWith: different input lengths");
  int num_lu_slots = c.num_routed_wires / 2;
  int num_lut_slots = c.num_routed_wires / 3;
  int num_sldc_polys = c.num_lookup_polys - 1;
  int lu_degree = c.quotient_degree_factor - 1;
  int lut_degree = divCeil(num_lut_slots, num_sldc_polys);
  std::vector<FExt> final_;
  for (size_t rnd = 0; rnd < roundChunks.size(); rnd++) {
    const LookupDelta &ld = ch.plonk_deltas[rnd];
    const auto &columns = roundChunks[rnd];
    if (columns.empty()) throw std::runtime_error("roundWorker: irrefutable pattern failed");
    FExt re = columns[0].first, re_next = columns[0].second;
    std::vector<FExt> sldc, sldc_next;
    for (size_t i = 1; i < columns.size(); i++) { sldc.push_back(columns[i].first); sldc_next.push_back(columns[i].second); }
    auto wchunks2 = partitionK(2, o.wires);
    auto wchunks3 = partitionK(3, o.wires);
    std::vector<FExt> lu_combos, lut_combos_A, lut_combos_B;
    for (int i = 0; i < num_lu_slots && i < (int)wchunks2.size(); i++) {
      if (wchunks2[i].size() != 2) continue;  // list-comprehension pattern [inp,out] skips non-matching
      lu_combos.push_back(wchunks2[i][0] + scaleExt(ld.lookup_A, wchunks2[i][1]));
    }
    for (int i = 0; i < num_lut_slots && i < (int)wchunks3.size(); i++) {
      if (wchunks3[i].size() != 3) continue;
      lut_combos_A.push_back(wchunks3[i][0] + scaleExt(ld.lookup_A, wchunks3[i][1]));
      lut_combos_B.push_back(wchunks3[i][0] + scaleExt(ld.lookup_B, wchunks3[i][1]));
    }
    std::vector<FExt> mults;
    for (int i = 0; i < num_lut_slots; i++) mults.push_back(o.wires.at(3 * i + 2));
    auto chunks_lu_combo = partitionK(lu_degree, lu_combos);
    auto chunks_lut_combo = partitionK(lut_degree, lut_combos_A);
    auto chunks_mults = partitionK(lut_degree, mults);

    if (sldc.empty()) throw std::runtime_error("last/head: empty list");
    FExt eq_last_sldc = selector(LS_LastLdc) * sldc.back();
    FExt eq_ini_sum = selector(LS_InitSre) * sldc.front();
    FExt eq_ini_re = selector(LS_InitSre) * re;
    std::vector<FExt> eq_finals_re;
    for (size_t k = 0; k < c.luts.size(); k++) {
      const auto &lut = c.luts[k];
      int lut_nrows = divCeil((int)lut.size(), num_lut_slots);
      int padded_size = lut_nrows * num_lut_slots;
      F cur(0);
      for (int t = 0; t < padded_size; t++) {
        const auto &e = t < (int)lut.size() ? lut[t] : lut.at(0);  // take padded $ lut ++ repeat (head lut)
        cur = ld.lookup_delta * cur + (e.first + ld.lookup_B * e.second);
      }
      eq_finals_re.push_back(selector(LS_StartEnd + (int)k) * (re - fromBase(cur)));
    }
    FExt cur_sum = re_next;
    for (auto &elt : lut_combos_B) cur_sum = scaleExt(ld.lookup_delta, cur_sum) + elt;
    FExt eq_re_trans = selector(LS_TransSre) * (re - cur_sum);
    // prevThisPairs = pairs (last sldc_next : sldc)
    std::vector<FExt> seq;
    seq.push_back(sldc_next.back());
    seq.insert(seq.end(), sldc.begin(), sldc.end());
    std::vector<FExt> eqs_sldc;
    size_t npairs = seq.size() - 1;
    size_t nz = std::min({npairs, chunks_lu_combo.size(), chunks_lut_combo.size(), chunks_mults.size()});
    FExt alpha = fromBase(ld.lookup_alpha);
    for (size_t t = 0; t < nz; t++) {
      FExt prev = seq[t], this_ = seq[t + 1];
      const auto &lu = chunks_lu_combo[t];
      const auto &lut = chunks_lut_combo[t];
      const auto &ms = chunks_mults[t];
      auto prodExcept = [&](const std::vector<FExt> &xs, int skip) {
        FExt p = FE1();
        for (int i = 0; i < (int)xs.size(); i++)
          if (i != skip) p = p * (alpha - xs[i]);
        return p;
      };
      FExt lu_prod = prodExcept(lu, -1), lut_prod = prodExcept(lut, -1);
      FExt lu_sum_prod = FE0(), lut_sum_prod = FE0();
      for (int i = 0; i < (int)lu.size(); i++) lu_sum_prod = lu_sum_prod + prodExcept(lu, i);
      for (int i = 0; i < (int)std::min(ms.size(), lut.size()); i++) lut_sum_prod = lut_sum_prod + ms[i] * prodExcept(lut, i);
      FExt eq_ldc_trans = selector(LS_TransLdc) * (lu_prod * (this_ - prev) + lu_sum_prod);
      FExt eq_sum_trans = selector(LS_TransSre) * (lut_prod * (this_ - prev) - lut_sum_prod);
      eqs_sldc.push_back(eq_sum_trans);
      eqs_sldc.push_back(eq_ldc_trans);
    }
    final_.push_back(eq_last_sldc);
    final_.push_back(eq_ini_sum);
    final_.push_back(eq_ini_re);
    final_.insert(final_.end(), eq_finals_re.begin(), eq_finals_re.end());
    final_.push_back(eq_re_trans);
    final_.insert(final_.end(), eqs_sldc.begin(), eqs_sldc.end());
  }
  return final_;
}

// ---- Plonk/Vanishing.hs --------------------------------------------------------------------
// combineWithPowersOfAlpha :54-56
inline FExt combineWithPowersOfAlpha(F alpha, const std::vector<FExt> &xs) {
  FExt acc = FE0();
  for (size_t k = xs.size(); k-- > 0;) acc = xs[k] + scaleExt(alpha, acc);
  return acc;
}

struct ConstraintDebug {
  std::vector<std::vector<FExt>> unfiltered;  // per gate
  std::vector<FExt> filters;
};

// evalAllPlonkConstraints :60-111
inline std::vector<FExt> evalAllPlonkConstraints(const CommonCircuitData &c, const ProofWithPublicInputs &pw,
                                                 const ProofChallenges &ch, ConstraintDebug *dbg = nullptr) {
  const OpeningSet &o = pw.proof.openings;
  SelectorConfig selcfg = getSelectorConfig(c);
  ConstantColumns cc = splitConstantColumns(selcfg, o.constants);
  u64 nn = (u64)c.nrows();
  int maxdeg = c.quotient_degree_factor;
  Digest pi_hash = sponge(pw.public_inputs);
  // gate constraints :87-94
  EvaluationVars ev;
  ev.local_selectors = cc.gateSelectors;
  ev.local_lkp_sels = cc.lookupSelectors;
  ev.local_constants = cc.gateConstants;
  ev.local_wires = o.wires;
  for (int i = 0; i < 4; i++) ev.public_inputs_hash.push_back(pi_hash.e[i]);
  std::vector<FExt> sel_values = evalGateSelectors(c, cc.gateSelectors);
  std::vector<FExt> gates;  // combineFilteredGateConstraints = foldl1 (longZipWith 0 0 (+))  :124-125
  if (c.gates.empty()) throw std::runtime_error("foldl1: empty list");
  for (size_t k = 0; k < c.gates.size(); k++) {
    std::vector<FExt> unf = gateConstraints(c.gates[k], ev);
    if (dbg) { dbg->unfiltered.push_back(unf); dbg->filters.push_back(sel_values[k]); }
    if (gates.size() < unf.size()) gates.resize(unf.size(), FE0());
    for (size_t i = 0; i < unf.size(); i++) gates[i] = gates[i] + unf[i] * sel_values[k];
  }
  // permutation constraints :96-111
  std::vector<FExt> zs1;
  for (auto &z : o.plonk_zs) zs1.push_back(evalLagrange0(nn, ch.plonk_zeta) * (z - FE1()));
  auto pp_chunks = partitionK(c.num_partial_products, o.partial_products);
  size_t nrounds = std::min({o.plonk_zs.size(), o.plonk_zs_next.size(), std::min(ch.plonk_betas.size(), ch.plonk_gammas.size()), pp_chunks.size()});
  std::vector<FExt> pp_checks;
  for (size_t rd = 0; rd < nrounds; rd++) {
    F beta = ch.plonk_betas[rd], gamma = ch.plonk_gammas[rd];
    std::vector<FExt> numer_all, denom_all;
    size_t nk = std::min(c.k_is.size(), o.wires.size());
    for (size_t i = 0; i < nk; i++) numer_all.push_back(o.wires[i] + scaleExt(beta * c.k_is[i], ch.plonk_zeta) + fromBase(gamma));
    size_t ns = std::min(o.plonk_sigmas.size(), o.wires.size());
    for (size_t i = 0; i < ns; i++) denom_all.push_back(o.wires[i] + scaleExt(beta, o.plonk_sigmas[i]) + fromBase(gamma));
    auto numers = partitionK(maxdeg, numer_all), denoms = partitionK(maxdeg, denom_all);
    std::vector<FExt> current;
    current.push_back(o.plonk_zs[rd]);
    current.insert(current.end(), pp_chunks[rd].begin(), pp_chunks[rd].end());
    current.push_back(o.plonk_zs_next[rd]);
    size_t np = std::min({current.size() - 1, numers.size(), denoms.size()});
    for (size_t t = 0; t < np; t++) {
      FExt pn = FE1(), pd = FE1();
      for (auto &x : numers[t]) pn = pn * x;
      for (auto &x : denoms[t]) pd = pd * x;
      pp_checks.push_back(current[t] * pn - current[t + 1] * pd);
    }
  }
  std::vector<FExt> lookup_checks;
  if (!c.luts.empty()) lookup_checks = evalLookupEquations(c, cc.lookupSelectors, o, ch);
  std::vector<FExt> finals = zs1;
  finals.insert(finals.end(), pp_checks.begin(), pp_checks.end());
  finals.insert(finals.end(), lookup_checks.begin(), lookup_checks.end());
  finals.insert(finals.end(), gates.begin(), gates.end());
  return finals;
}

// evalCombinedPlonkConstraints :48-51
inline std::vector<FExt> evalCombinedPlonkConstraints(const CommonCircuitData &c, const ProofWithPublicInputs &pw,
                                                      const ProofChallenges &ch) {
  std::vector<FExt> constraints = evalAllPlonkConstraints(c, pw, ch);
  std::vector<FExt> out;
  for (auto &alpha : ch.plonk_alphas) out.push_back(combineWithPowersOfAlpha(alpha, constraints));
  return out;
}

// checkCombinedPlonkEquations', Plonk/Verifier.hs:35-52
inline std::vector<bool> checkCombinedPlonkEquations_(const CommonCircuitData &c, const ProofWithPublicInputs &pw,
                                                      const ProofChallenges &ch, std::vector<FExt> *combined_out = nullptr) {
  int maxdeg = c.quotient_degree_factor;
  FExt zeta_n = powExtU(ch.plonk_zeta, (u64)c.nrows());
  std::vector<FExt> combined = evalCombinedPlonkConstraints(c, pw, ch);
  if (combined_out) *combined_out = combined;
  auto chunks = partitionK(maxdeg, pw.proof.openings.quotient_polys);
  if (chunks.size() != combined.size()) throw std::runtime_error("safeZipWith: different input lengths");
  std::vector<bool> ok;
  for (size_t i = 0; i < chunks.size(); i++) {
    FExt q = FE0();
    for (size_t k = chunks[i].size(); k-- > 0;) q = chunks[i][k] + zeta_n * q;
    ok.push_back(q * (zeta_n - FE1()) == combined[i]);
  }
  return ok;
}

// ---- Plonk/FRI.hs --------------------------------------------------------------------------
struct VerifyError : std::runtime_error {
  int code, query, detail;
  VerifyError(int c, int q, int d, const char *m) : std::runtime_error(m), code(c), query(q), detail(d) {}
};

// checkProofOfWork :212-216
inline bool checkProofOfWork(const FriConfig &fc, const FriChallenges &ch) {
  int b = fc.proof_of_work_bits;
  if (b == 0) return true;
  u64 lo_mask = b >= 64 ? ~(u64)0 : (((u64)1 << b) - 1);
  u64 mask = lo_mask << (64 - b);
  return (ch.fri_pow_response.v & mask) == 0;
}

// precomputeReducedOpenings :128-134
struct PrecomputedReducedOpenings { FExt sum_this_row, sum_next_row; };
inline PrecomputedReducedOpenings precomputeReducedOpenings(const FExt &alpha, const FriOpenings &fo) {
  return PrecomputedReducedOpenings{reduceWithPowers(alpha, fo.batch_this), reduceWithPowers(alpha, fo.batch_next)};
}

// combineInitial :151-207
inline FExt combineInitial(const CommonCircuitData &c, const ProofChallenges &ch, const PrecomputedReducedOpenings &pre,
                           const std::array<std::vector<F>, 4> &oracles, int query_idx) {
  int r = c.num_challenges;
  int npp = divCeil(c.num_routed_wires, c.quotient_degree_factor);
  if (r * (npp + c.num_lookup_polys) != (int)oracles[2].size()) throw std::runtime_error("combineInitial: sanity check failed");
  const std::vector<F> &pp_lookup = oracles[2];
  std::vector<F> oracle_pp(pp_lookup.begin(), pp_lookup.begin() + r * npp), oracle_lookup(pp_lookup.begin() + r * npp, pp_lookup.end());
  std::vector<FExt> firstBatch, secondBatch;
  for (auto &x : oracles[0]) firstBatch.push_back(fromBase(x));
  for (auto &x : oracles[1]) firstBatch.push_back(fromBase(x));
  for (auto &x : oracle_pp) firstBatch.push_back(fromBase(x));
  for (auto &x : oracles[3]) firstBatch.push_back(fromBase(x));
  for (auto &x : oracle_lookup) firstBatch.push_back(fromBase(x));
  for (int i = 0; i < r && i < (int)oracle_pp.size(); i++) secondBatch.push_back(fromBase(oracle_pp[i]));
  for (auto &x : oracle_lookup) secondBatch.push_back(fromBase(x));
  const FExt &alpha = ch.fri_challenges.fri_alpha;
  FExt g0 = reduceWithPowers(alpha, firstBatch), g1 = reduceWithPowers(alpha, secondBatch);
  int logn_small = c.degree_bits, logn_lde = c.lde_bits();
  F omega = subgroupGenerator(logn_small), eta = subgroupGenerator(logn_lde);
  int rev_idx = reverseBitsInt(logn_lde, query_idx);
  FExt point_x = fromBase(F(MUL_GEN) * powi(eta, rev_idx));
  FExt loc0 = ch.plonk_zeta, loc1 = fromBase(omega) * ch.plonk_zeta;
  FExt one = (g0 - pre.sum_this_row) / (point_x - loc0);
  FExt two = (g1 - pre.sum_next_row) / (point_x - loc1);
  return powExtI(alpha, (long long)secondBatch.size()) * one + two;
}

struct Coset { int size_log2; F offset; std::vector<FExt> values; };
// prepareCoset :248-259
inline Coset prepareCoset(F shift, int bigLog2, int idx, const std::vector<FExt> &values) {
  int arity = 0;
  while (((size_t)1 << arity) < values.size()) arity++;
  if (((size_t)1 << arity) != values.size()) throw std::runtime_error("safeLog2: input is not a power of two");
  F eta = subgroupGenerator(bigLog2);
  int start = reverseBitsInt(bigLog2, (idx >> arity) << arity);
  return Coset{arity, shift * powi(eta, start), reverseIndexBitsList(values)};
}
// foldCosetWith :263-279 — literally as written (one inversion per (j,k) via `pow_ x (-k)`)
inline FExt foldCosetWith(const FExt &beta, const Coset &coset) {
  int arity = 1 << coset.size_log2;
  F omega = subgroupGenerator(coset.size_log2);
  F invArity = F(1) / F((u64)arity);
  std::vector<FExt> ys;
  for (int k = 0; k < arity; k++) {
    FExt s = FE0();
    for (int j = 0; j < arity; j++) {
      F x_omega_j = coset.offset * powi(omega, j);
      s = s + scaleExt(powi(x_omega_j, -k), coset.values[j]);
    }
    ys.push_back(s);
  }
  FExt acc = FE0(), bp = FE1();  // sum $ zipWith (*) (powersOf beta) ys
  for (int k = 0; k < arity; k++) { acc = acc + bp * ys[k]; bp = beta * bp; }
  return scaleExt(invArity, acc);
}

// Trace of one query round for differential tests
struct QueryTrace {
  int status = 0;  // P2V_ST_* code of this round (0 = round_ok True)
  int detail = 0;
  FExt combined_eval, final_eval, folded;
};

// checkFRIProof :358-407.  Returns the per-proof status word (FRI part) in reference order:
// pow_ok first, then the query rounds in order with `and` short-circuit.
inline uint32_t checkFRIProofStatus(const CommonCircuitData &c, const VerifierOnlyCircuitData &vk, const Proof &proof,
                                    const ProofChallenges &ch, std::vector<QueryTrace> *trace = nullptr) {
  const FriChallenges &fc = ch.fri_challenges;
  const FriProof &fp = proof.opening_proof;
  size_t ncap = (size_t)1 << c.fri_config.cap_height;
  bool pow_ok = checkProofOfWork(c.fri_config, fc);
  uint32_t result = pow_ok ? (uint32_t)P2V_ST_ACCEPT : (uint32_t)P2V_ST_FALSE_POW;
  bool decided = !pow_ok;
  if (fc.fri_query_indices.size() != fp.query_round_proofs.size()) {
    if (!decided) throw std::runtime_error("safeZipWith: different input lengths");
  }
  FriOpenings fo = toFriOpenings(proof.openings);
  PrecomputedReducedOpenings pre = precomputeReducedOpenings(fc.fri_alpha, fo);
  int logn_lde = c.lde_bits();
  const std::vector<int> &arities = c.fri_config.step_arity_bits;
  const MerkleCap *caps[4] = {&vk.constants_sigmas_cap, &proof.wires_cap, &proof.plonk_zs_partial_products_cap, &proof.quotient_polys_cap};
  auto widths = oracleWidths(c);
  size_t nq = std::min(fc.fri_query_indices.size(), fp.query_round_proofs.size());
  for (size_t q = 0; q < nq; q++) {
    QueryTrace tr;
    int idx = fc.fri_query_indices[q];
    const FriQueryRound &round = fp.query_round_proofs[q];
    try {
      // toMerkleOracles / validateMerkleCapLength :79-97
      for (int o = 0; o < 4; o++)
        if (caps[o]->roots.size() != ncap) throw VerifyError(19, (int)q, o, "validateMerkleCapLength: cap has wrong size");
      // checkInitialTreeProofs :105-117
      const auto &eps = round.initial_trees_proof.evals_proofs;
      if (eps.size() != 4) throw VerifyError(20, (int)q, 0, "checkInitialTreeProofs: expecting 4 Merkle proofs for the 4 oracles");
      int bad = 0;
      for (int o = 0; o < 4; o++)
        if (!checkMerkleProof(*caps[o], idx, eps[o].first, eps[o].second)) bad |= 1 << o;
      if (bad) throw VerifyError(P2V_ST_ERR_INIT_MERKLE, (int)q, bad, "checkInitialTreeProofs: at least one Merkle proof failed");
      std::array<std::vector<F>, 4> oracles;
      for (int o = 0; o < 4; o++) {
        if ((int)eps[o].first.size() != widths[o]) throw VerifyError(21, (int)q, o, "buildListOracle: list size do not match the expected");
        oracles[o] = eps[o].first;
      }
      FExt combined_eval = combineInitial(c, ch, pre, oracles, idx);
      tr.combined_eval = combined_eval;
      // folding :306-323, 387-402
      if (arities.size() != fc.fri_betas.size() || arities.size() != fp.commit_phase_merkle_caps.size() || arities.size() != round.steps.size())
        throw VerifyError(22, (int)q, 0, "safeZipWith4: different input lengths");
      F shift(MUL_GEN);
      int size_log2 = logn_lde, qidx = idx;
      FExt eval = combined_eval;
      for (size_t s = 0; s < arities.size(); s++) {
        int arityLog2 = arities[s];
        int arity = 1 << arityLog2;
        const FriQueryStep &step = round.steps[s];
        int newIdx = qidx >> arityLog2;
        bool proofCheckOK = checkMerkleProof(fp.commit_phase_merkle_caps[s], newIdx, flattenExt(step.evals), step.merkle_proof);
        if (!proofCheckOK) throw VerifyError(P2V_ST_ERR_STEP_MERKLE, (int)q, (int)s, "folding step Merkle proof does not check out");
        if ((size_t)(qidx % arity) >= step.evals.size()) throw std::runtime_error("(!!): index too large");
        bool evalCheckOK = step.evals[qidx % arity] == eval;
        if (!evalCheckOK) throw VerifyError(P2V_ST_ERR_STEP_EVAL, (int)q, (int)s, "folding step evaluation does not match the opening");
        if ((size_t)arity != step.evals.size()) throw VerifyError(23, (int)q, (int)s, "folding step: reduction strategy incompatibility");
        Coset coset = prepareCoset(shift, size_log2, qidx, step.evals);
        FExt newEval = foldCosetWith(fc.fri_betas[s], coset);
        shift = powi(shift, arity);
        size_log2 -= arityLog2;
        qidx = newIdx;
        eval = newEval;
      }
      tr.folded = eval;
      // folding_query_loc :288-291, evalPolynomialAt :325-327
      F x_final = shift * powi(subgroupGenerator(size_log2), reverseBitsInt(size_log2, qidx));
      FExt fpe = FE0(), xp = FE1(), xf = fromBase(x_final);
      for (auto &co : fp.final_poly) { fpe = fpe + co * xp; xp = xf * xp; }
      tr.final_eval = fpe;
      bool round_ok = fpe == eval;
      if (!round_ok) { tr.status = P2V_ST_FALSE_FINAL; }
    } catch (const VerifyError &e) {
      tr.status = e.code;
      tr.detail = e.detail;
    }
    if (trace) trace->push_back(tr);
    if (!decided && tr.status != 0) {
      decided = true;
      result = (uint32_t)tr.status | ((uint32_t)q << 8) | ((uint32_t)tr.detail << 16);
    }
  }
  return result;
}

struct VerifyTrace {
  ProofChallenges challenges;
  std::vector<FExt> combined;
  std::vector<bool> eqs_ok;
  uint32_t fri_status = 0;
  std::vector<QueryTrace> queries;
};

// verifyProof, Plonk/Verifier.hs:56-65: all_ok = eqs_ok && fri_ok (eqs first; FRI only forced if eqs hold)
inline uint32_t verifyProofStatus(const CommonCircuitData &c, const VerifierOnlyCircuitData &vk,
                                  const ProofWithPublicInputs &pw, VerifyTrace *trace = nullptr) {
  ProofChallenges ch = proofChallenges(c, vk, pw);
  std::vector<FExt> combined;
  std::vector<bool> oks = checkCombinedPlonkEquations_(c, pw, ch, &combined);
  uint32_t mask = 0;
  for (size_t i = 0; i < oks.size(); i++)
    if (!oks[i]) mask |= 1u << i;
  uint32_t fri_status = 0;
  std::vector<QueryTrace> qt;
  // The trace wants FRI intermediates even for proofs whose equations fail; the verdict does not.
  if (trace || mask == 0) fri_status = checkFRIProofStatus(c, vk, pw.proof, ch, trace ? &qt : nullptr);
  if (trace) {
    trace->challenges = ch;
    trace->combined = combined;
    trace->eqs_ok = oks;
    trace->fri_status = fri_status;
    trace->queries = qt;
  }
  if (mask) return (uint32_t)P2V_ST_FALSE_EQS | (mask << 16);
  return fri_status;
}

}  // namespace orc
