// ORACLE / FIXTURE GENERATOR — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// The reference ships no fixtures (its json/ directory is git-ignored, .gitignore:4,6; testmain
// reads ../json/<prefix>_{common,vkey,proof}.json, src/testmain.hs:25-33) and no prover.  This is
// the mini-prover of SURVEY.md App. F (extended to real circuits): it produces ACCEPTING Plonky2 proofs of a
// chosen circuit SHAPE (gate list, widths, FRI parameters) in the reference's JSON wire format
// (src/Types.hs aeson instances), so that `testmain` could consume them unchanged.
//
// Real presets (real5, real7): every row carries an ACTIVE gate on an honest witness, Noop rows are wired to the
// active rows by copy constraints (non-identity sigma, real grand product Z and partial products), and the
// quotient polynomial is the real C(X)/Z_H(X) — see `witnessRow` and the p.real branches of `prove`.
// Trivial presets: every row is a NoopGate, the wiring permutation is the identity.
//   - selector column of Noop's group == index(Noop), other selector columns == UNUSED (2^32-1),
//     lookup selectors == 0  =>  every gate filter (Gate/Selector.hs:83-89) and lookup equation
//     vanishes identically;
//   - sigma_i(x) = k_i * x  =>  Z == 1, partial products == 1, quotient == 0
//     (Plonk/Vanishing.hs:96-111, Plonk/Verifier.hs:44-47);
//   - wires (and lookup columns) are uniformly random polynomials of degree < N.
// The FRI part is derived from first principles (commentary/FRI.md:94-134), NOT from the verifier
// restatement: LDE by NTT on the coset g*<eta>, trees over bit-reversed rows, combined polynomial
// from its definition, folding by Lagrange-interpolating each coset and evaluating at beta, final
// polynomial by inverse DFT with a low-degree assertion, proof-of-work grinding.
// The transcript uses oracle/challenger.hpp (shared with the verifier restatement; the Python
// twin oracle/pyref.py re-derives the challenges independently).
//
//   p2v_prover <preset> <out_prefix> [--seed K] [--corrupt-layer S] [--bad-final] [--bad-witness ROW] [--bad-copy] [--bad-lookup]
//
// Presets: s12 (standard recursion shape), mid5, small6, fixed4, arity5, lookup6 (all-Noop rows, quotient == 0);
//          real5, real7, real12 (ACTIVE gates of all 14 standard kinds on honest witnesses, real quotient C/Z_H;
//          real12 = the standard recursion configuration, 2^12 rows),
//          reallu6 (real circuit with an honest lookup argument).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include "plonk.hpp"

using namespace orc;

// ---- deterministic RNG ------------------------------------------------------------------------
struct SplitMix {
  u64 s;
  explicit SplitMix(u64 seed) : s(seed) {}
  u64 next() {
    u64 z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  F felt() { return F(next() % P); }
};

// ---- radix-2 NTT (prover only) ------------------------------------------------------------------
static void ntt(std::vector<F> &a, int logn) {  // in-place, natural order in/out, forward: A_i = sum a_k w^{ik}
  size_t n = (size_t)1 << logn;
  for (size_t i = 0; i < n; i++) {
    size_t j = reverseBits(logn, i);
    if (i < j) std::swap(a[i], a[j]);
  }
  for (int s = 1; s <= logn; s++) {
    size_t m = (size_t)1 << s;
    F wm = subgroupGenerator(s);
    for (size_t k = 0; k < n; k += m) {
      F w(1);
      for (size_t j = 0; j < m / 2; j++) {
        F t = w * a[k + j + m / 2], u = a[k + j];
        a[k + j] = u + t;
        a[k + j + m / 2] = u - t;
        w = w * wm;
      }
    }
  }
}
// evaluations of the polynomial with coefficients `coef` (deg < 2^logn) on g*eta^i, i < 2^loglde
static std::vector<F> ldeCoset(const std::vector<F> &coef, int loglde) {
  size_t m = (size_t)1 << loglde;
  std::vector<F> a(m, F(0));
  F gp(1), g(MUL_GEN);
  for (size_t k = 0; k < coef.size(); k++) { a[k] = coef[k] * gp; gp = gp * g; }
  ntt(a, loglde);
  return a;
}
static F evalPolyBase(const std::vector<F> &coef, F x) {
  F acc(0);
  for (size_t k = coef.size(); k-- > 0;) acc = acc * x + coef[k];
  return acc;
}
static FExt evalPolyAtExt(const std::vector<F> &coef, const FExt &x) {
  FExt acc = FE0();
  for (size_t k = coef.size(); k-- > 0;) acc = acc * x + fromBase(coef[k]);
  return acc;
}

// ---- Merkle trees over bit-reversed rows (commentary/FRI.md:16,27-29) ----------------------------
struct Tree {
  std::vector<std::vector<F>> leaves;        // leaf j (already in tree order)
  std::vector<std::vector<Digest>> levels;   // levels[0] = leaf digests ... last = cap
  int cap_height;
  MerkleCap cap() const { MerkleCap c; c.roots = levels.back(); return c; }
  MerkleProof open(int idx) const {
    MerkleProof p;
    for (size_t l = 0; l + 1 < levels.size(); l++) p.siblings.push_back(levels[l][(idx >> l) ^ 1]);
    return p;
  }
};
static Tree buildTree(std::vector<std::vector<F>> leaves, int cap_height, int threads) {
  Tree t;
  t.cap_height = cap_height;
  size_t n = leaves.size();
  std::vector<Digest> cur(n);
  std::vector<std::thread> th;
  for (int k = 0; k < threads; k++)
    th.emplace_back([&, k]() {
      for (size_t i = k; i < n; i += threads) cur[i] = sponge(leaves[i]);
    });
  for (auto &x : th) x.join();
  t.leaves = std::move(leaves);
  t.levels.push_back(cur);
  while (t.levels.back().size() > ((size_t)1 << cap_height)) {
    const auto &lv = t.levels.back();
    std::vector<Digest> nx(lv.size() / 2);
    for (size_t i = 0; i < nx.size(); i++) nx[i] = compress(lv[2 * i], lv[2 * i + 1]);
    t.levels.push_back(nx);
  }
  return t;
}

// ---- circuit description ---------------------------------------------------------------------------
struct GateSpec { std::string text; Gate gate; int group; };
struct Preset {
  std::string name;
  int degree_bits, rate_bits, cap_height, pow_bits, num_queries;
  bool fixed_strategy;
  std::vector<int> fixed_arities;
  int arity_bits, final_poly_bits;
  int num_wires, num_routed, num_challenges, qdf, num_public_inputs;
  std::vector<GateSpec> gates;
  std::vector<Range> groups;
  int num_lookup_polys = 0;
  std::vector<std::vector<std::pair<u64, u64>>> luts;
  bool real = false;  // rows with ACTIVE gates on honest witnesses + a real quotient polynomial
};

static const char *PHANTOM = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";

static std::vector<F> calcBarycentricWeights(int bits) {  // Gate/Custom/CosetInterp.hs:44-47
  std::vector<F> loc = enumerateSubgroup(bits), w;
  for (size_t i = 0; i < loc.size(); i++) {
    F prod(1);
    for (size_t j = 0; j < loc.size(); j++)
      if (j != i) prod = prod * (loc[i] - loc[j]);
    w.push_back(inv(prod));
  }
  return w;
}

static GateSpec mk(int kind, int p0, int p1, int p2, const std::string &text) {
  GateSpec g;
  g.gate.kind = kind; g.gate.p0 = p0; g.gate.p1 = p1; g.gate.p2 = p2;
  g.text = text;
  g.group = 0;
  return g;
}
static std::string S(long long x) { return std::to_string(x); }
static GateSpec gNoop() { return mk(P2V_GATE_NOOP, 0, 0, 0, "NoopGate"); }
static GateSpec gConst(int n) { return mk(P2V_GATE_CONSTANT, n, 0, 0, "ConstantGate { num_consts: " + S(n) + " }"); }
static GateSpec gPI() { return mk(P2V_GATE_PUBLIC_INPUT, 0, 0, 0, "PublicInputGate"); }
static GateSpec gBaseSum(int l, int b) { return mk(P2V_GATE_BASE_SUM, l, b, 0, "BaseSumGate { num_limbs: " + S(l) + " } + Base: " + S(b)); }
static GateSpec gRedExt(int n) { return mk(P2V_GATE_REDUCING_EXT, n, 0, 0, "ReducingExtensionGate { num_coeffs: " + S(n) + " }"); }
static GateSpec gRed(int n) { return mk(P2V_GATE_REDUCING, n, 0, 0, "ReducingGate { num_coeffs: " + S(n) + " }"); }
static GateSpec gArithExt(int n) { return mk(P2V_GATE_ARITHMETIC_EXT, n, 0, 0, "ArithmeticExtensionGate { num_ops: " + S(n) + " }"); }
static GateSpec gArith(int n) { return mk(P2V_GATE_ARITHMETIC, n, 0, 0, "ArithmeticGate { num_ops: " + S(n) + " }"); }
static GateSpec gMulExt(int n) { return mk(P2V_GATE_MUL_EXT, n, 0, 0, "MulExtensionGate { num_ops: " + S(n) + " }"); }
static GateSpec gExp(int n) { return mk(P2V_GATE_EXPONENTIATION, n, 0, 0, "ExponentiationGate { num_power_bits: " + S(n) + " }"); }
static GateSpec gRA(int b, int c, int e) {
  return mk(P2V_GATE_RANDOM_ACCESS, b, c, e,
            "RandomAccessGate { bits: " + S(b) + ", num_copies: " + S(c) + ", num_extra_constants: " + S(e) + ", _phantom: " + PHANTOM + " }<D=2>");
}
static GateSpec gCoset(int bits, int degree) {
  GateSpec g = mk(P2V_GATE_COSET_INTERP, bits, degree, 0, "");
  g.gate.weights = calcBarycentricWeights(bits);
  std::string w;
  for (size_t i = 0; i < g.gate.weights.size(); i++) w += (i ? ", " : "") + std::to_string(g.gate.weights[i].v);
  g.text = "CosetInterpolationGate { subgroup_bits: " + S(bits) + ", degree: " + S(degree) + ", barycentric_weights: [" + w + "], _phantom: " + PHANTOM + " }<D=2>";
  return g;
}
static GateSpec gPoseidon() { return mk(P2V_GATE_POSEIDON, 12, 0, 0, std::string("PoseidonGate(") + PHANTOM + ")<WIDTH=12>"); }
static GateSpec gPoseidonMds() { return mk(P2V_GATE_POSEIDON_MDS, 12, 0, 0, std::string("PoseidonMdsGate(") + PHANTOM + ")<WIDTH=12>"); }
static std::string lutHashText(int seed) {
  std::string s = "[";
  for (int i = 0; i < 32; i++) s += (i ? ", " : "") + std::to_string((seed * 37 + i * 11) & 255);
  return s + "]";
}
static GateSpec gLookup(int slots) { return mk(P2V_GATE_LOOKUP, slots, 0, 0, "LookupGate { num_slots: " + S(slots) + ", lut_hash: " + lutHashText(1) + " }"); }
static GateSpec gLookupTable(int slots, int last) {
  return mk(P2V_GATE_LOOKUP_TABLE, slots, last, 0, "LookupTableGate { num_slots: " + S(slots) + ", lut_hash: " + lutHashText(1) + ", last_lut_row: " + S(last) + " }");
}

static void assignGroups(Preset &p, const std::vector<int> &sizes) {
  int pos = 0;
  for (size_t g = 0; g < sizes.size(); g++) {
    p.groups.push_back(Range{pos, pos + sizes[g]});
    for (int k = 0; k < sizes[g]; k++) p.gates[pos + k].group = (int)g;
    pos += sizes[g];
  }
  if (pos != (int)p.gates.size()) { fprintf(stderr, "assignGroups: sizes do not cover the gates\n"); exit(2); }
}

static Preset makePreset(const std::string &name) {
  Preset p;
  p.name = name;
  p.fixed_strategy = false;
  p.num_challenges = 2;
  p.qdf = 8;
  if (name == "s12" || name == "mid5") {
    // standard_recursion_config (commentary/FRI.md:46-52, commentary/Layout.md:13-23)
    p.degree_bits = name == "s12" ? 12 : 5;
    p.rate_bits = 3; p.cap_height = name == "s12" ? 4 : 2; p.pow_bits = name == "s12" ? 16 : 6;
    p.num_queries = name == "s12" ? 28 : 6;
    p.arity_bits = name == "s12" ? 4 : 2; p.final_poly_bits = name == "s12" ? 5 : 2;
    p.num_wires = 135; p.num_routed = 80; p.num_public_inputs = 4;
    p.gates = {gNoop(), gConst(2), gPI(), gBaseSum(63, 2), gRedExt(32), gRed(43), gArithExt(10), gArith(20), gMulExt(13),
               gExp(66), gRA(4, 4, 2), gCoset(4, 6), gPoseidon(), gPoseidonMds()};
    assignGroups(p, {6, 5, 3});
  } else if (name == "real5" || name == "real7" || name == "real12") {
    // Same gate set as the standard recursion shape, but the rows carry ACTIVE gates on honestly generated
    // witnesses and the quotient polynomial is the real C(X)/Z_H(X).  Selector groups obey Plonky2's degree rule
    // (group size + gate degree <= quotient_degree_factor + 1): {deg<=2} {deg<=4} {RandomAccess, CosetInterp} {Poseidon, PoseidonMds}.
    p.real = true;
    p.degree_bits = name == "real5" ? 5 : name == "real7" ? 7 : 12;
    p.rate_bits = 3; p.cap_height = 2; p.pow_bits = 6; p.num_queries = 6;
    p.arity_bits = 2; p.final_poly_bits = 2;
    if (name == "real12") {  // the standard recursion configuration itself (like s12), with real rows
      p.cap_height = 4; p.pow_bits = 16; p.num_queries = 28; p.arity_bits = 4; p.final_poly_bits = 5;
    }
    p.num_wires = 135; p.num_routed = 80; p.num_public_inputs = 5;
    p.gates = {gNoop(), gConst(2), gPI(), gBaseSum(63, 2), gRedExt(32), gRed(43), gArithExt(10), gArith(20), gMulExt(13),
               gExp(66), gRA(4, 4, 2), gCoset(4, 6), gPoseidon(), gPoseidonMds()};
    assignGroups(p, {6, 4, 2, 2});
  } else if (name == "small6") {
    p.degree_bits = 6; p.rate_bits = 3; p.cap_height = 1; p.pow_bits = 8; p.num_queries = 5;
    p.arity_bits = 2; p.final_poly_bits = 3;
    p.num_wires = 20; p.num_routed = 16; p.num_public_inputs = 3;
    p.gates = {gNoop(), gConst(2), gArith(4), gPI(), gMulExt(3), gArithExt(2)};
    assignGroups(p, {3, 3});
  } else if (name == "reallu6") {
    // real circuit WITH a lookup argument: LookupGate rows look entries of a 19-entry table up, LookupTableGate rows hold
    // the table with the multiplicities, honest RE / partial-sum polynomials per challenge (Plonk/Lookups.hs:45-132)
    p.real = true;
    p.degree_bits = 6; p.rate_bits = 3; p.cap_height = 2; p.pow_bits = 8; p.num_queries = 5;
    p.arity_bits = 3; p.final_poly_bits = 2;
    p.num_wires = 40; p.num_routed = 24; p.num_public_inputs = 2;
    p.gates = {gNoop(), gConst(2), gLookup(12), gLookupTable(8, 3), gArith(6), gPI()};
    assignGroups(p, {4, 2});
    p.num_lookup_polys = 3;
    std::vector<std::pair<u64, u64>> t0;
    for (u64 i = 0; i < 19; i++) t0.push_back({i, (i * i + 3) & 0xffff});
    p.luts = {t0};
  } else if (name == "arity5") {
    // one folding step of arity 32 followed by one of arity 2 (Fixed [5,1]): exercises wide cosets (a > 4)
    p.degree_bits = 7; p.rate_bits = 1; p.cap_height = 1; p.pow_bits = 5; p.num_queries = 4;
    p.fixed_strategy = true; p.fixed_arities = {5, 1};
    p.num_challenges = 2; p.qdf = 2;
    p.num_wires = 10; p.num_routed = 4; p.num_public_inputs = 1;
    p.gates = {gNoop(), gConst(2), gPI()};
    assignGroups(p, {3});
  } else if (name == "fixed4") {
    p.degree_bits = 4; p.rate_bits = 2; p.cap_height = 0; p.pow_bits = 4; p.num_queries = 4;
    p.fixed_strategy = true; p.fixed_arities = {2, 1};
    p.num_challenges = 3; p.qdf = 4;
    p.num_wires = 12; p.num_routed = 8; p.num_public_inputs = 0;
    p.gates = {gConst(2), gNoop(), gArith(3)};
    assignGroups(p, {3});
  } else if (name == "lookup6") {
    p.degree_bits = 6; p.rate_bits = 3; p.cap_height = 2; p.pow_bits = 8; p.num_queries = 5;
    p.arity_bits = 3; p.final_poly_bits = 2;
    p.num_wires = 40; p.num_routed = 24; p.num_public_inputs = 2;
    p.gates = {gNoop(), gConst(2), gLookup(12), gLookupTable(8, 3), gArith(6), gPI()};
    assignGroups(p, {4, 2});
    p.num_lookup_polys = 3;
    std::vector<std::pair<u64, u64>> t0, t1;
    for (u64 i = 0; i < 19; i++) t0.push_back({i, (i * i + 3) & 0xffff});
    for (u64 i = 0; i < 8; i++) t1.push_back({100 + i, 7 * i + 1});
    p.luts = {t0, t1};
  } else if (name.rfind("rand", 0) == 0 || name.rfind("rreal", 0) == 0) {
    // Random VALID shapes for the property tests (tests/test_random_shapes.py, tests/test_gpu_random_shapes.py): every
    // parameter of CommonCircuitData that the verifier's layout, transcript, gate programs or FRI schedule depend on is
    // drawn from the preset's number.  rand<K>: all-Noop rows (any gate set and any quotient_degree_factor are valid: the
    // filters vanish identically); rreal<K>: the real circuit of real5 (active gates, copy constraints, real quotient,
    // rate_bits 3 = log2 quotient_degree_factor) under random FRI parameters.
    const bool real = name[1] == 'r';
    const u64 k = strtoull(name.c_str() + (real ? 5 : 4), nullptr, 10);
    SplitMix g(0xC0FFEEULL ^ (k * 0x9E3779B97F4A7C15ULL) ^ (real ? 0x5EA1ULL : 0));
    auto pick = [&](int lo, int hi) { return lo + (int)(g.next() % (u64)(hi - lo + 1)); };
    p.real = real;
    p.degree_bits = real ? pick(5, 8) : pick(3, 9);
    p.rate_bits = real ? 3 : pick(1, 4);
    p.qdf = real ? 8 : 1 << pick(1, 3);
    p.num_challenges = real ? 2 : pick(1, 3);
    p.pow_bits = pick(0, 10);
    p.num_queries = pick(1, 9);
    // reduction strategy: ConstantArityBits (a, f) with f chosen so that the folding ends exactly at 2^f coefficients, or Fixed
    int total = 0;
    if (pick(0, 2) == 0) {
      p.fixed_strategy = true;
      for (int left = p.degree_bits; left > 0 && (int)p.fixed_arities.size() < 4 && pick(0, 3) != 0;) {
        int a = pick(1, std::min(5, left));
        p.fixed_arities.push_back(a);
        left -= a;
        total += a;
      }
    } else {
      p.arity_bits = pick(1, 4);
      int steps = pick(0, p.degree_bits / p.arity_bits);
      total = steps * p.arity_bits;
      p.final_poly_bits = p.degree_bits - total;
    }
    p.cap_height = pick(0, std::min(4, p.degree_bits + p.rate_bits - total));  // every commit-phase tree is at least as high as its cap
    if (real) {
      p.num_wires = 135; p.num_routed = 80; p.num_public_inputs = pick(0, 9);
      p.gates = {gNoop(), gConst(2), gPI(), gBaseSum(63, 2), gRedExt(32), gRed(43), gArithExt(10), gArith(20), gMulExt(13),
                 gExp(66), gRA(4, 4, 2), gCoset(4, 6), gPoseidon(), gPoseidonMds()};
      assignGroups(p, {6, 4, 2, 2});
    } else {
      const bool wide = pick(0, 1) == 1;
      p.num_wires = wide ? 135 : pick(8, 60);
      p.num_routed = wide ? 80 : pick(2, p.num_wires);
      p.num_public_inputs = pick(0, 9);
      const int W = p.num_wires;
      std::vector<GateSpec> pool;
      pool.push_back(gConst(pick(1, 2)));
      if (W >= 4) pool.push_back(gPI());
      pool.push_back(gArith(pick(1, std::min(20, W / 4))));
      pool.push_back(gArithExt(pick(1, std::min(10, W / 8))));
      pool.push_back(gMulExt(pick(1, std::min(13, W / 6))));
      pool.push_back(gBaseSum(pick(1, std::min(63, W - 1)), pick(2, 4)));
      if (W >= 9) pool.push_back(gRed(pick(1, std::min(43, (W - 6) / 3))));
      if (W >= 10) pool.push_back(gRedExt(pick(1, std::min(32, (W - 6) / 4))));
      if (W >= 4) pool.push_back(gExp(pick(1, std::min(66, (W - 2) / 2))));
      if (wide) {
        int bits = pick(1, 4);
        pool.push_back(gRA(bits, pick(1, 4), pick(0, 2)));
        int cb = pick(2, 4);
        pool.push_back(gCoset(cb, pick(2, 7)));
        pool.push_back(gPoseidon());
        pool.push_back(gPoseidonMds());
      }
      // a random subset in random order, the NoopGate somewhere among them
      std::vector<GateSpec> chosen;
      for (auto &gs : pool)
        if (pick(0, 2) != 0) chosen.push_back(gs);
      chosen.insert(chosen.begin() + pick(0, (int)chosen.size()), gNoop());
      for (size_t i = chosen.size(); i > 1; i--) std::swap(chosen[i - 1], chosen[(size_t)pick(0, (int)i - 1)]);
      p.gates = chosen;
      std::vector<int> sizes;
      for (int left = (int)chosen.size(); left > 0;) {
        int sz = pick(1, left);
        sizes.push_back(sz);
        left -= sz;
      }
      assignGroups(p, sizes);
    }
  } else {
    fprintf(stderr, "unknown preset %s\n", name.c_str());
    exit(2);
  }
  return p;
}

static std::vector<int> expandStrategy(const Preset &p) {  // Plonk/FRI.hs:337-354
  std::vector<int> out;
  if (p.fixed_strategy) return p.fixed_arities;
  for (int logn = p.degree_bits; logn > p.final_poly_bits; logn -= p.arity_bits) out.push_back(p.arity_bits);
  return out;
}

static CommonCircuitData toCommon(const Preset &p) {
  CommonCircuitData c;
  c.num_wires = p.num_wires; c.num_routed_wires = p.num_routed; c.config_num_constants = 2; c.num_challenges = p.num_challenges;
  c.fri_config.rate_bits = p.rate_bits; c.fri_config.cap_height = p.cap_height; c.fri_config.proof_of_work_bits = p.pow_bits;
  c.fri_config.num_query_rounds = p.num_queries; c.fri_config.step_arity_bits = expandStrategy(p);
  c.degree_bits = p.degree_bits;
  for (auto &g : p.gates) { c.gates.push_back(g.gate); c.selector_indices.push_back(g.group); }
  c.selector_groups = p.groups;
  c.quotient_degree_factor = p.qdf;
  c.num_lookup_selectors = p.luts.empty() ? 0 : 4 + (int)p.luts.size();
  c.num_constants = (int)p.groups.size() + c.num_lookup_selectors + 2;
  c.num_public_inputs = p.num_public_inputs;
  F k(1);
  for (int i = 0; i < p.num_routed; i++) { c.k_is.push_back(k); k = k * F(MUL_GEN); }
  c.num_partial_products = divCeil(p.num_routed, p.qdf) - 1;
  c.num_lookup_polys = p.num_lookup_polys;
  for (auto &t : p.luts) {
    std::vector<std::pair<F, F>> tt;
    for (auto &e : t) tt.push_back({F(e.first), F(e.second)});
    c.luts.push_back(tt);
  }
  return c;
}

// ---- honest witness rows for ACTIVE gates (real circuits) --------------------------------------------------------
// Each generator fills the wires a gate constrains from what the gate MEANS; unused wires stay random.
static FExt EX(const std::vector<F> &w, int i) { return FExt(w[i], w[i + 1]); }
static void putE(std::vector<F> &w, int i, const FExt &v) { w[i] = v.r; w[i + 1] = v.i; }

static void witnessRow(const Gate &g, std::vector<F> &w, F c0, F c1, const Digest &pih, SplitMix &rng) {
  switch (g.kind) {
    case P2V_GATE_ARITHMETIC:
      for (int i = 0; i < g.p0; i++) { int j = 4 * i; w[j + 3] = c0 * w[j] * w[j + 1] + c1 * w[j + 2]; }
      break;
    case P2V_GATE_ARITHMETIC_EXT:
      for (int i = 0; i < g.p0; i++) { int j = 8 * i; putE(w, j + 6, scaleExt(c0, EX(w, j) * EX(w, j + 2)) + scaleExt(c1, EX(w, j + 4))); }
      break;
    case P2V_GATE_MUL_EXT:
      for (int i = 0; i < g.p0; i++) { int j = 6 * i; putE(w, j + 4, scaleExt(c0, EX(w, j) * EX(w, j + 2))); }
      break;
    case P2V_GATE_BASE_SUM: {
      F acc(0), pw(1);
      for (int i = 0; i < g.p0; i++) {
        u64 limb = rng.next() % (u64)g.p1;
        w[1 + i] = F(limb);
        acc = acc + F(limb) * pw;
        pw = pw * F((u64)g.p1);
      }
      w[0] = acc;
      break;
    }
    case P2V_GATE_CONSTANT:
      if (g.p0 > 0) w[0] = c0;
      if (g.p0 > 1) w[1] = c1;
      break;
    case P2V_GATE_PUBLIC_INPUT:
      for (int i = 0; i < 4; i++) w[i] = pih.e[i];
      break;
    case P2V_GATE_EXPONENTIATION: {
      int n = g.p0;
      F base = w[0], acc(1);
      std::vector<int> bits(n);
      for (int i = 0; i < n; i++) { bits[i] = (int)(rng.next() & 1); w[1 + i] = F((u64)bits[i]); }
      for (int i = 0; i < n; i++) {
        int cur = bits[n - 1 - i];
        acc = (i ? acc * acc : F(1)) * (cur ? base : F(1));
        w[n + 2 + i] = acc;
      }
      w[n + 1] = acc;
      break;
    }
    case P2V_GATE_REDUCING: case P2V_GATE_REDUCING_EXT: {
      int n = g.p0;
      bool ext = g.kind == P2V_GATE_REDUCING_EXT;
      FExt alpha = EX(w, 2), acc = EX(w, 4);
      for (int i = 0; i < n; i++) {
        FExt coeff = ext ? EX(w, 6 + 2 * i) : fromBase(w[6 + i]);
        acc = acc * alpha + coeff;
        putE(w, i < n - 1 ? 6 + (ext ? 2 * n : n) + 2 * i : 0, acc);
      }
      break;
    }
    case P2V_GATE_RANDOM_ACCESS: {
      int nb = g.p0, copies = g.p1, extra = g.p2, width = 2 + (1 << nb), start = width * copies + extra;
      for (int k = 0; k < copies; k++) {
        int idx = (int)(rng.next() % (u64)(1 << nb));
        w[k * width] = F((u64)idx);
        w[k * width + 1] = w[k * width + 2 + idx];
        for (int j = 0; j < nb; j++) w[start + k * nb + j] = F((u64)((idx >> j) & 1));
      }
      if (extra > 0) w[copies * width] = c0;
      if (extra > 1) w[copies * width + 1] = c1;
      break;
    }
    case P2V_GATE_COSET_INTERP: {
      int npts = 1 << g.p0, degree = g.p1, nint = (npts - 2) / (degree - 1), base = 1 + 2 * (npts + 2);
      std::vector<F> dom = enumerateSubgroup(g.p0);
      F shift = w[0];
      FExt x0 = EX(w, base + 4 * nint);
      putE(w, 1 + 2 * npts, scaleExt(shift, x0));
      FExt ev = FE0(), pr = FE1();
      int idx = 0, ck = 0;
      while (idx < npts) {
        int len = ck == 0 ? degree : degree - 1;
        for (int k = 0; k < len && idx < npts; k++, idx++) {
          FExt val = scaleExt(g.weights[idx], EX(w, 1 + 2 * idx));
          FExt term = x0 - fromBase(dom[idx]);
          FExt ne = term * ev + val * pr;
          pr = term * pr;
          ev = ne;
        }
        if (idx < npts) { putE(w, base + 2 * ck, ev); putE(w, base + 2 * (nint + ck), pr); }
        ck++;
      }
      putE(w, 1 + 2 * npts + 2, ev);
      break;
    }
    case P2V_GATE_POSEIDON_MDS:
      for (int i = 0; i < 12; i++) {
        FExt acc = FE0();
        for (int j = 0; j < 12; j++) acc = acc + scaleExt(mdsMatrixCoeff(i, j), EX(w, 2 * j));
        putE(w, 2 * (i + 12), acc);
      }
      break;
    case P2V_GATE_POSEIDON: {
      F swap((u64)(rng.next() & 1));
      w[24] = swap;
      F st[12];
      for (int i = 0; i < 4; i++) {
        F delta = swap * (w[i + 4] - w[i]);
        w[25 + i] = delta;
        st[i] = w[i] + delta;
        st[i + 4] = w[i + 4] - delta;
        st[i + 8] = w[i + 8];
      }
      State in;
      for (int i = 0; i < 12; i++) in[i] = st[i];
      auto mds = [&](F *s) { F o[12]; for (int i = 0; i < 12; i++) { F a(0); for (int j = 0; j < 12; j++) a = a + mdsMatrixCoeff(i, j) * s[j]; o[i] = a; } for (int i = 0; i < 12; i++) s[i] = o[i]; };
      for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = st[i] + F(ALL_ROUND_CONSTANTS[r][i]);
        if (r) for (int i = 0; i < 12; i++) w[29 + 12 * (r - 1) + i] = st[i];
        for (int i = 0; i < 12; i++) st[i] = sbox1(st[i]);
        mds(st);
      }
      for (int i = 0; i < 12; i++) st[i] = st[i] + F(FAST_PARTIAL_FIRST_RC[i]);
      {
        F t[12];
        t[0] = st[0];
        for (int i = 0; i < 11; i++) { F a(0); for (int j = 0; j < 11; j++) a = a + partialMdsMatrixCoeff(i, j) * st[j + 1]; t[i + 1] = a; }
        for (int i = 0; i < 12; i++) st[i] = t[i];
      }
      for (int r = 0; r < 22; r++) {
        w[65 + r] = st[0];
        F y = sbox1(st[0]);
        if (r < 21) y = y + F(FAST_PARTIAL_RCS[r]);
        F d = y * mdsMatrixCoeff(0, 0);
        for (int i = 0; i < 11; i++) d = d + st[i + 1] * F(FAST_PARTIAL_W_HATS[r][i]);
        for (int i = 0; i < 11; i++) st[i + 1] = st[i + 1] + y * F(FAST_PARTIAL_VS[r][i]);
        st[0] = d;
      }
      for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) st[i] = st[i] + F(ALL_ROUND_CONSTANTS[26 + r][i]);
        for (int i = 0; i < 12; i++) w[87 + 12 * r + i] = st[i];
        for (int i = 0; i < 12; i++) st[i] = sbox1(st[i]);
        mds(st);
      }
      for (int i = 0; i < 12; i++) w[12 + i] = st[i];
      State out = permutation(in);
      for (int i = 0; i < 12; i++)
        if (out[i] != st[i]) { fprintf(stderr, "PoseidonGate witness != permutation\n"); exit(5); }
      break;
    }
    default: break;  // Noop, Lookup*
  }
}

// inverse NTT on the subgroup of size 2^logn (natural order): values -> coefficients
static std::vector<F> intt(std::vector<F> v, int logn) {
  size_t n = (size_t)1 << logn;
  ntt(v, logn);
  std::vector<F> out(n);
  F ninv = inv(F((u64)n));
  for (size_t k = 0; k < n; k++) out[k] = v[(n - k) % n] * ninv;
  return out;
}

// ---- the prover ---------------------------------------------------------------------------------------
struct ProverOut { VerifierOnlyCircuitData vk; ProofWithPublicInputs pw; };

static ProverOut prove(const Preset &p, const CommonCircuitData &c, u64 seed, int corrupt_layer, bool bad_final, int threads, int bad_witness_row = -1, bool bad_copy = false, bool bad_lookup = false) {
  SplitMix rng(seed);
  int n = c.degree_bits, N = 1 << n, loglde = c.lde_bits(), M = 1 << loglde;
  int r = c.num_challenges;
  int ncap_h = c.fri_config.cap_height;
  int noop_index = -1;
  for (size_t k = 0; k < c.gates.size(); k++)
    if (c.gates[k].kind == P2V_GATE_NOOP) noop_index = (int)k;
  if (noop_index < 0) { fprintf(stderr, "preset needs a NoopGate\n"); exit(2); }
  // --- column polynomials (coefficient form, degree < N) ---
  std::vector<std::vector<F>> const_cols, wire_cols, pp_cols, quot_cols;
  std::vector<std::vector<F>> sigma_vals, real_wvals;  // real circuits: sigma and wire columns as values over H
  std::vector<int> gate_of_row;
  struct { int slots_lu = 0, slots_lut = 0, lu_start = 0, n_lu_rows = 0, lut_first = 0, n_lut_rows = 0, end_row = 0; } lkp;
  ProverOut out;
  Digest pih;
  if (p.real) {
    // rows carry ACTIVE gates (round-robin over the gate list), honest witnesses, per-row gate constants
    SplitMix rpi(seed ^ 0x5EEDF00DULL);
    for (int i = 0; i < c.num_public_inputs; i++) out.pw.public_inputs.push_back(rpi.felt());
    pih = sponge(out.pw.public_inputs);
    int G = (int)c.gates.size(), ngrp = (int)c.selector_groups.size(), nls = c.num_lookup_selectors;
    std::vector<std::vector<F>> cvals(ngrp + nls + 2, std::vector<F>(N, F(0))), wvals(c.num_wires, std::vector<F>(N));
    // lookup argument (Plonk/Lookups.hs:45-132): LookupGate rows, then the LookupTableGate rows holding LUT 0 in
    // REVERSE block order (running evaluation and partial sums accumulate from the end row upwards), then the end row
    bool has_lut = !c.luts.empty();
    int lu_gate = -1, lut_gate = -1;
    for (int k = 0; k < G; k++) {
      if (c.gates[k].kind == P2V_GATE_LOOKUP) lu_gate = k;
      if (c.gates[k].kind == P2V_GATE_LOOKUP_TABLE) lut_gate = k;
    }
    if (has_lut) {
      if (lu_gate < 0 || lut_gate < 0) { fprintf(stderr, "real lookup preset needs LookupGate and LookupTableGate\n"); exit(2); }
      lkp.slots_lu = c.num_routed_wires / 2;
      lkp.slots_lut = c.num_routed_wires / 3;
      lkp.lu_start = 10; lkp.n_lu_rows = 2;
      lkp.lut_first = lkp.lu_start + lkp.n_lu_rows;
      lkp.n_lut_rows = divCeil((int)c.luts[0].size(), lkp.slots_lut);
      lkp.end_row = lkp.lut_first + lkp.n_lut_rows;
      if (lkp.end_row >= N) { fprintf(stderr, "lookup rows do not fit\n"); exit(2); }
    }
    std::vector<u64> mult(has_lut ? c.luts[0].size() : 0, 0);
    gate_of_row.assign(N, 0);
    for (int row = 0; row < N; row++) {
      int gi = row % G;
      enum { PLAIN, LU, LUT } role = PLAIN;
      if (has_lut) {
        if (row >= lkp.lu_start && row < lkp.lut_first) { gi = lu_gate; role = LU; }
        else if (row >= lkp.lut_first && row < lkp.end_row) { gi = lut_gate; role = LUT; }
        else if (row == lkp.end_row || gi == lu_gate || gi == lut_gate) gi = noop_index;
      }
      gate_of_row[row] = gi;
      for (int g = 0; g < ngrp; g++) cvals[g][row] = g == c.selector_indices[gi] ? F::fromInt(gi) : F((u64)0xFFFFFFFFULL);
      if (has_lut) {
        if (role == LUT) cvals[ngrp + LS_TransSre][row] = F(1);
        if (role == LU) cvals[ngrp + LS_TransLdc][row] = F(1);
        if (row == lkp.end_row) cvals[ngrp + LS_InitSre][row] = F(1);
        if (row == lkp.lu_start) cvals[ngrp + LS_LastLdc][row] = F(1);
        if (row == lkp.lut_first) cvals[ngrp + LS_StartEnd + 0][row] = F(1);
      }
      F c0 = rng.felt(), c1 = rng.felt();
      cvals[ngrp + nls][row] = c0; cvals[ngrp + nls + 1][row] = c1;
      std::vector<F> w(c.num_wires);
      for (auto &x : w) x = rng.felt();
      if (role == LU) {            // every slot looks up a random entry of LUT 0
        for (int i = 0; i < lkp.slots_lu; i++) {
          size_t t = (size_t)(rng.next() % (u64)c.luts[0].size());
          if (bad_lookup && row == lkp.lu_start && i == 0) { w[0] = c.luts[0][t].first; w[1] = c.luts[0][t].second + F(1); continue; }  // not in the table
          w[2 * i] = c.luts[0][t].first; w[2 * i + 1] = c.luts[0][t].second;
          mult[t]++;
        }
      } else if (role == LUT) {    // block q of the table sits in row lut_first + (n_lut_rows - 1 - q); padding = entry 0, multiplicity 0
        int q = lkp.n_lut_rows - 1 - (row - lkp.lut_first);
        for (int i = 0; i < lkp.slots_lut; i++) {
          size_t t = (size_t)q * lkp.slots_lut + i;
          bool pad = t >= c.luts[0].size();
          const auto &e = c.luts[0][pad ? 0 : t];
          w[3 * i] = e.first; w[3 * i + 1] = e.second; w[3 * i + 2] = pad ? F(0) : F(mult[t]);
        }
      } else if (row != bad_witness_row) witnessRow(c.gates[gi], w, c0, c1, pih, rng);  // --bad-witness: leave that row's gate unsatisfied
      for (int i = 0; i < c.num_wires; i++) wvals[i][row] = w[i];
    }
    // copy constraints: the routed wires of the Noop rows (unconstrained by any gate) are wired to routed cells of
    // the active rows and to each other, giving 2- and 3-cycles of the wiring permutation sigma
    // (cell (row,col) <-> field element k_col * omega^row, Plonk/Vanishing.hs:96-111).
    int R = c.num_routed_wires;
    std::vector<int> sigma((size_t)N * R);
    for (size_t i = 0; i < sigma.size(); i++) sigma[i] = (int)i;
    auto cell = [&](int row, int col) { return row * R + col; };
    auto joinCycle = [&](int a, int b) { sigma[a] = sigma[b]; sigma[b] = a; };  // a: singleton, joins b's cycle
    {
      std::vector<char> used((size_t)N * R, 0);
      std::vector<int> noop_rows;
      for (int row = 0; row < N; row++) if (gate_of_row[row] == noop_index) noop_rows.push_back(row);
      for (size_t k = 0; k < noop_rows.size(); k++)
        for (int col = 0; col < R; col++) {
          int a = cell(noop_rows[k], col), b;
          if (k > 0 && (col & 1)) b = cell(noop_rows[k - 1], (col * 7 + 3) % R);  // extend an existing cycle
          else
            do { b = cell((int)(rng.next() % (u64)N), (int)(rng.next() % (u64)R)); } while (used[b] || gate_of_row[b / R] == noop_index);
          used[b] = 1;
          wvals[col][noop_rows[k]] = wvals[b % R][b / R];
          joinCycle(a, b);
        }
    }
    if (bad_copy) wvals[0][noop_index] = wvals[0][noop_index] + F(1);  // --bad-copy: one wired cell no longer equals its cycle
    sigma_vals.assign(R, std::vector<F>(N));
    {
      std::vector<F> wp(N);
      F w1 = subgroupGenerator(n), x(1);
      for (int i = 0; i < N; i++) { wp[i] = x; x = x * w1; }
      for (int row = 0; row < N; row++)
        for (int col = 0; col < R; col++) { int t = sigma[cell(row, col)]; sigma_vals[col][row] = c.k_is[t % R] * wp[t / R]; }
    }
    real_wvals = wvals;
    for (auto &col : cvals) const_cols.push_back(intt(col, n));
    for (auto &col : wvals) wire_cols.push_back(intt(col, n));
  } else {
    for (size_t g = 0; g < c.selector_groups.size(); g++) {
      F v = (int)g == c.selector_indices[noop_index] ? F::fromInt(noop_index) : F((u64)0xFFFFFFFFULL);
      const_cols.push_back({v});
    }
    for (int i = 0; i < c.num_lookup_selectors; i++) const_cols.push_back({F(0)});
    for (int i = 0; i < 2; i++) const_cols.push_back({rng.felt()});      // gate constants
  }
  for (int i = 0; i < c.num_routed_wires; i++) {
    if (p.real) const_cols.push_back(intt(sigma_vals[i], n));
    else const_cols.push_back({F(0), c.k_is[i]});  // identity wiring: sigma_i(x) = k_i x
  }
  if (!p.real)
    for (int i = 0; i < c.num_wires; i++) {
      std::vector<F> co(N);
      for (auto &x : co) x = rng.felt();
      wire_cols.push_back(co);
    }
  for (int i = 0; i < r * (1 + c.num_partial_products); i++) pp_cols.push_back({F(1)});  // Z and partial products == 1
  for (int i = 0; i < r * c.num_lookup_polys; i++) {  // lookup columns: arbitrary low-degree
    std::vector<F> co(N);
    for (auto &x : co) x = rng.felt();
    pp_cols.push_back(co);
  }
  for (int i = 0; i < r * c.quotient_degree_factor; i++) quot_cols.push_back({F(0)});  // quotient == 0 unless p.real (below)
  std::vector<std::vector<std::vector<F>> *> mats = {&const_cols, &wire_cols, &pp_cols, &quot_cols};

  // --- LDE + commit: rows at x_i = g*eta^i, leaf j = row rev(j) ---
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::vector<std::vector<F>>> lde(4);  // [oracle][col][i]
  std::vector<Tree> trees(4);
  auto commitOracle = [&](int o) {
    lde[o].assign(mats[o]->size(), {});
    std::vector<std::thread> th;
    for (int k = 0; k < threads; k++)
      th.emplace_back([&, k, o]() {
        for (size_t col = k; col < mats[o]->size(); col += threads) lde[o][col] = ldeCoset((*mats[o])[col], loglde);
      });
    for (auto &x : th) x.join();
    std::vector<std::vector<F>> leaves(M);
    for (int j = 0; j < M; j++) {
      int i = reverseBitsInt(loglde, j);
      leaves[j].resize(lde[o].size());
      for (size_t col = 0; col < lde[o].size(); col++) leaves[j][col] = lde[o][col][i];
    }
    trees[o] = buildTree(std::move(leaves), ncap_h, threads);
  };
  for (int o = 0; o < 2; o++) commitOracle(o);

  out.vk.constants_sigmas_cap = trees[0].cap();
  {  // circuit digest: any 4 field elements (the reference only absorbs it, Challenge/Verifier.hs:73)
    std::vector<F> d;
    for (auto &g : p.gates) for (char ch : g.text) d.push_back(F((u64)(unsigned char)ch));
    for (auto &dg : out.vk.constants_sigmas_cap.roots) for (int i = 0; i < 4; i++) d.push_back(dg.e[i]);
    out.vk.circuit_digest = sponge(d);
  }
  Proof &proof = out.pw.proof;
  if (!p.real)
    for (int i = 0; i < c.num_public_inputs; i++) out.pw.public_inputs.push_back(rng.felt());
  proof.wires_cap = trees[1].cap();

  // --- transcript up to the alphas (Challenge/Verifier.hs:73-88) ---
  Duplex dx(zeroState());
  dx.absorb(out.vk.circuit_digest);
  dx.absorb(sponge(out.pw.public_inputs));
  dx.absorb(proof.wires_cap);
  ProofChallenges pch;
  pch.plonk_betas = dx.squeezeN(r);
  pch.plonk_gammas = dx.squeezeN(r);
  if (c.num_lookup_polys > 0) {
    std::vector<F> all = pch.plonk_betas, deltas = dx.squeezeN(2 * r);
    all.insert(all.end(), pch.plonk_gammas.begin(), pch.plonk_gammas.end());
    all.insert(all.end(), deltas.begin(), deltas.end());
    pch.plonk_deltas = mkLookupDeltaList(all);
  }
  if (p.real) {
    // grand product Z and its partial products per challenge round (Plonk/Vanishing.hs:96-111):
    // current = [Z, pp_0 .. pp_{m-1}, Z(omega x)],  current[t+1] = current[t] * numer_t / denom_t  with the routed
    // wires taken in chunks of quotient_degree_factor.
    int R = c.num_routed_wires, qd = c.quotient_degree_factor, npp = c.num_partial_products;
    std::vector<std::vector<F>> zvals(r, std::vector<F>(N)), ppvals((size_t)r * npp, std::vector<F>(N));
    F w1 = subgroupGenerator(n);
    for (int j = 0; j < r; j++) {
      F beta = pch.plonk_betas[j], gamma = pch.plonk_gammas[j], z(1), x(1);
      for (int row = 0; row < N; row++) {
        zvals[j][row] = z;
        F cur = z;
        for (int t = 0; t * qd < R; t++) {
          F num(1), den(1);
          for (int col = t * qd; col < std::min(R, (t + 1) * qd); col++) {
            num = num * (real_wvals[col][row] + beta * c.k_is[col] * x + gamma);
            den = den * (real_wvals[col][row] + beta * sigma_vals[col][row] + gamma);
          }
          cur = cur * num * inv(den);
          if (t < npp) ppvals[(size_t)j * npp + t][row] = cur;
        }
        z = cur;
        x = x * w1;
      }
      if (z != F(1) && !bad_copy) { fprintf(stderr, "[prover] grand product does not close: sigma is not value-preserving\n"); exit(7); }
    }
    pp_cols.clear();
    for (auto &col : zvals) pp_cols.push_back(intt(col, n));
    for (auto &col : ppvals) pp_cols.push_back(intt(col, n));
    if (!c.luts.empty()) {
      // lookup polynomials per challenge round: RE (running evaluation of the table) and the partial sums SLDC_t of the
      // log-derivative argument, both accumulating from the end row (where they are 0) towards lower row indices:
      //   table row:  RE(row) = Horner_delta(RE(row+1); inp + B out over the slots),  SLDC_t = prev + sum m/(alpha - (inp + A out))
      //   lookup row: SLDC_t = prev - sum 1/(alpha - (inp + A out)),   prev of chunk 0 = last SLDC of row+1.
      int nsl = c.num_lookup_polys - 1, lu_deg = c.quotient_degree_factor - 1, lut_deg = divCeil(lkp.slots_lut, nsl);
      for (int j = 0; j < r; j++) {
        const LookupDelta &ld = pch.plonk_deltas[j];
        std::vector<std::vector<F>> cols(c.num_lookup_polys, std::vector<F>(N, F(0)));
        F re(0), prev(0);
        for (int row = lkp.end_row - 1; row >= lkp.lu_start; row--) {
          bool table = row >= lkp.lut_first;
          int slots = table ? lkp.slots_lut : lkp.slots_lu, deg = table ? lut_deg : lu_deg, stride = table ? 3 : 2;
          if (table) {
            for (int i = 0; i < slots; i++) re = ld.lookup_delta * re + (real_wvals[3 * i][row] + ld.lookup_B * real_wvals[3 * i + 1][row]);
            cols[0][row] = re;
          }
          for (int t = 0; t < nsl; t++) {
            F acc = prev;
            for (int i = t * deg; i < std::min(slots, (t + 1) * deg); i++) {
              F combo = real_wvals[stride * i][row] + ld.lookup_A * real_wvals[stride * i + 1][row];
              F term = inv(ld.lookup_alpha - combo);
              acc = table ? acc + real_wvals[3 * i + 2][row] * term : acc - term;
            }
            cols[1 + t][row] = acc;
            prev = acc;
          }
        }
        if (prev != F(0) && !bad_lookup) { fprintf(stderr, "[prover] lookup sums do not cancel\n"); exit(8); }
        for (auto &col : cols) pp_cols.push_back(intt(col, n));
      }
    }
  }
  commitOracle(2);
  proof.plonk_zs_partial_products_cap = trees[2].cap();
  dx.absorb(proof.plonk_zs_partial_products_cap);
  pch.plonk_alphas = dx.squeezeN(r);

  if (p.real) {
    // --- the real quotient: t_j = C_j / Z_H on the LDE coset (deg C_j < 9N by the selector-group degree rule,
    //     so deg t_j < 8N = M and its M coset values determine it), split into qdf chunks of N coefficients
    //     (Plonk/Verifier.hs:44-47 reassembles sum_k zeta^{kN} chunk_k).  C_j(x) is evaluated with the SAME
    //     function the verifier uses at zeta (evalCombinedPlonkConstraints), fed with the column values at x.
    if (M != N * c.quotient_degree_factor) { fprintf(stderr, "real preset needs 2^rate_bits == quotient_degree_factor\n"); exit(2); }
    std::vector<std::vector<F>> tvals(r, std::vector<F>(M));
    F eta = subgroupGenerator(loglde);
    std::vector<F> xs(M);
    { F x(MUL_GEN); for (int i = 0; i < M; i++) { xs[i] = x; x = x * eta; } }
    std::vector<std::thread> th;
    for (int k = 0; k < threads; k++)
      th.emplace_back([&, k]() {
        for (int i = k; i < M; i += threads) {
          ProofWithPublicInputs fake;
          fake.public_inputs = out.pw.public_inputs;
          OpeningSet &fo = fake.proof.openings;
          for (int t = 0; t < c.num_constants; t++) fo.constants.push_back(fromBase(lde[0][t][i]));
          for (size_t t = c.num_constants; t < lde[0].size(); t++) fo.plonk_sigmas.push_back(fromBase(lde[0][t][i]));
          for (auto &col : lde[1]) fo.wires.push_back(fromBase(col[i]));
          int inext = (i + (M >> n)) % M;  // omega * x_i = x_{i + M/N}
          for (int t = 0; t < r; t++) { fo.plonk_zs.push_back(fromBase(lde[2][t][i])); fo.plonk_zs_next.push_back(fromBase(lde[2][t][inext])); }
          for (int t = 0; t < r * c.num_partial_products; t++) fo.partial_products.push_back(fromBase(lde[2][r + t][i]));
          for (int t = 0; t < r * c.num_lookup_polys; t++) {
            size_t col = (size_t)r * (1 + c.num_partial_products) + t;
            fo.lookup_zs.push_back(fromBase(lde[2][col][i]));
            fo.lookup_zs_next.push_back(fromBase(lde[2][col][inext]));
          }
          ProofChallenges ch = pch;
          ch.plonk_zeta = fromBase(xs[i]);
          std::vector<FExt> cj = evalCombinedPlonkConstraints(c, fake, ch);
          F zh = inv(powu(xs[i], (u64)N) - F(1));
          for (int j = 0; j < r; j++) {
            if (cj[j].i != F(0)) { fprintf(stderr, "[prover] constraint value left the base field\n"); exit(6); }
            tvals[j][i] = cj[j].r * zh;
          }
        }
      });
    for (auto &x : th) x.join();
    quot_cols.clear();
    F ginv = inv(F(MUL_GEN));
    for (int j = 0; j < r; j++) {
      std::vector<F> co = intt(tvals[j], loglde);
      F gp(1);
      for (auto &x : co) { x = x * gp; gp = gp * ginv; }
      for (int k = 0; k < c.quotient_degree_factor; k++) quot_cols.push_back(std::vector<F>(co.begin() + (size_t)k * N, co.begin() + (size_t)(k + 1) * N));
    }
    // sanity: on every row of H the combined constraint must vanish, i.e. t_j is a polynomial (checked through the
    // verifier's identity at zeta by the caller's self-check; here: C_j at the first rows of H)
  }
  commitOracle(3);
  proof.quotient_polys_cap = trees[3].cap();
  fprintf(stderr, "[prover] LDE + 4 trees: %.1fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  dx.absorb(proof.quotient_polys_cap);
  FExt zeta = dx.squeezeExt();
  FExt omega_zeta = fromBase(subgroupGenerator(n)) * zeta;

  // --- openings ---
  OpeningSet &o = proof.openings;
  auto openAll = [&](const std::vector<std::vector<F>> &cols, size_t from, size_t to, const FExt &x, std::vector<FExt> &dst) {
    for (size_t i = from; i < to; i++) dst.push_back(evalPolyAtExt(cols[i], x));
  };
  openAll(const_cols, 0, c.num_constants, zeta, o.constants);
  openAll(const_cols, c.num_constants, const_cols.size(), zeta, o.plonk_sigmas);
  openAll(wire_cols, 0, wire_cols.size(), zeta, o.wires);
  openAll(pp_cols, 0, r, zeta, o.plonk_zs);
  openAll(pp_cols, 0, r, omega_zeta, o.plonk_zs_next);
  openAll(pp_cols, r, r * (1 + c.num_partial_products), zeta, o.partial_products);
  openAll(quot_cols, 0, quot_cols.size(), zeta, o.quotient_polys);
  openAll(pp_cols, r * (1 + c.num_partial_products), pp_cols.size(), zeta, o.lookup_zs);
  openAll(pp_cols, r * (1 + c.num_partial_products), pp_cols.size(), omega_zeta, o.lookup_zs_next);

  // --- FRI: combined polynomial on the LDE domain (commentary/FRI.md:94-114) ---
  FriOpenings fo = toFriOpenings(o);
  dx.absorb(fo.batch_this);
  dx.absorb(fo.batch_next);
  FExt alpha = dx.squeezeExt();
  // column order of the two batches (Plonk/FRI.hs:170-185)
  struct ColRef { int oracle; int col; };
  std::vector<ColRef> first, second;
  int npp_total = r * (1 + c.num_partial_products);
  for (size_t i = 0; i < const_cols.size(); i++) first.push_back({0, (int)i});
  for (size_t i = 0; i < wire_cols.size(); i++) first.push_back({1, (int)i});
  for (int i = 0; i < npp_total; i++) first.push_back({2, i});
  for (size_t i = 0; i < quot_cols.size(); i++) first.push_back({3, (int)i});
  for (size_t i = npp_total; i < pp_cols.size(); i++) first.push_back({2, (int)i});
  for (int i = 0; i < r; i++) second.push_back({2, i});
  for (size_t i = npp_total; i < pp_cols.size(); i++) second.push_back({2, (int)i});
  if (first.size() != fo.batch_this.size() || second.size() != fo.batch_next.size()) { fprintf(stderr, "batch size mismatch\n"); exit(3); }
  std::vector<FExt> apow(first.size() + 1);
  apow[0] = FE1();
  for (size_t k = 1; k < apow.size(); k++) apow[k] = apow[k - 1] * alpha;
  FExt y0 = FE0(), y1 = FE0();
  for (size_t k = 0; k < first.size(); k++) y0 = y0 + apow[k] * fo.batch_this[k];
  for (size_t k = 0; k < second.size(); k++) y1 = y1 + apow[k] * fo.batch_next[k];
  std::vector<FExt> codeword(M);  // natural order i
  {
    F eta = subgroupGenerator(loglde);
    std::vector<F> xs(M);
    F x(MUL_GEN);
    for (int i = 0; i < M; i++) { xs[i] = x; x = x * eta; }
    std::vector<std::thread> th;
    for (int k = 0; k < threads; k++)
      th.emplace_back([&, k]() {
        for (int i = k; i < M; i += threads) {
          FExt g0 = FE0(), g1 = FE0();
          for (size_t t = 0; t < first.size(); t++) g0 = g0 + scaleExt(lde[first[t].oracle][first[t].col][i], apow[t]);
          for (size_t t = 0; t < second.size(); t++) g1 = g1 + scaleExt(lde[second[t].oracle][second[t].col][i], apow[t]);
          FExt px = fromBase(xs[i]);
          codeword[i] = apow[second.size()] * ((g0 - y0) / (px - zeta)) + (g1 - y1) / (px - omega_zeta);
        }
      });
    for (auto &t : th) t.join();
  }
  // bit-reversed order
  std::vector<FExt> cur(M);
  for (int j = 0; j < M; j++) cur[j] = codeword[reverseBitsInt(loglde, j)];

  // --- commit phase (commentary/FRI.md:118-134) ---
  FriProof &fp = proof.opening_proof;
  std::vector<Tree> step_trees;
  std::vector<std::vector<FExt>> layers;  // the codeword each step tree commits to (bit-reversed order)
  F shift(MUL_GEN);
  int bits = loglde;
  std::vector<int> arities = c.fri_config.step_arity_bits;
  for (size_t s = 0; s < arities.size(); s++) {
    int a = arities[s], A = 1 << a;
    if (corrupt_layer == (int)s)
      for (auto &v : cur) v = v + FE1();  // commit to a wrong layer: Merkle passes, eval check fails
    layers.push_back(cur);
    size_t ncos = cur.size() / A;
    std::vector<std::vector<F>> leaves(ncos);
    for (size_t j = 0; j < ncos; j++)
      for (int t = 0; t < A; t++) { leaves[j].push_back(cur[j * A + t].r); leaves[j].push_back(cur[j * A + t].i); }
    int tree_cap = std::min(ncap_h, bits - a);
    step_trees.push_back(buildTree(std::move(leaves), tree_cap, threads));
    fp.commit_phase_merkle_caps.push_back(step_trees.back().cap());
    dx.absorb(fp.commit_phase_merkle_caps.back());
    FExt beta = dx.squeezeExt();
    // fold: Lagrange-interpolate each coset {x_base * w^k} and evaluate at beta
    F eta = subgroupGenerator(bits), w = subgroupGenerator(a);
    std::vector<FExt> nxt(ncos);
    for (size_t j = 0; j < ncos; j++) {
      F x_base = shift * powu(eta, reverseBits(bits - a, j));
      // entry t of the coset (bit-reversed position) sits at x_base * w^{rev_a(t)}
      std::vector<F> pts(A);
      for (int t = 0; t < A; t++) pts[t] = x_base * powu(w, reverseBits(a, t));
      FExt acc = FE0();
      for (int t = 0; t < A; t++) {
        FExt num = FE1();
        F den(1);
        for (int u = 0; u < A; u++)
          if (u != t) { num = num * (beta - fromBase(pts[u])); den = den * (pts[t] - pts[u]); }
        acc = acc + cur[j * A + t] * scaleExt(inv(den), num);
      }
      nxt[j] = acc;
    }
    cur = nxt;
    shift = powu(shift, A);
    bits -= a;
  }
  // --- final polynomial: inverse DFT on the coset shift*<eta_f>, must have degree < final_len ---
  {
    int total = 0;
    for (int a : arities) total += a;
    size_t final_len = (size_t)1 << (n - total);
    size_t m = cur.size();
    F eta = subgroupGenerator(bits);
    std::vector<FExt> nat(m);
    for (size_t j = 0; j < m; j++) nat[reverseBits(bits, j)] = cur[j];
    F minv = inv(F((u64)m));
    std::vector<FExt> coeffs;
    for (size_t k = 0; k < m; k++) {
      FExt acc = FE0();
      for (size_t i = 0; i < m; i++) {
        F xi = shift * powu(eta, i);
        acc = acc + scaleExt(powu(inv(xi), k), nat[i]);
      }
      coeffs.push_back(scaleExt(minv, acc));
    }
    for (size_t k = final_len; k < m; k++)
      if (coeffs[k] != FE0() && corrupt_layer < 0) { fprintf(stderr, "[prover] folded codeword is NOT low-degree (k=%zu)\n", k); exit(4); }
    coeffs.resize(final_len);
    if (bad_final) coeffs[0] = coeffs[0] + FE1();
    fp.final_poly = coeffs;
  }
  dx.absorb(fp.final_poly);
  // --- grinding (Challenge/FRI.hs:89-90, Plonk/FRI.hs:212-216) ---
  {
    u64 w = 0;
    for (;; w++) {
      Duplex trial = dx;
      trial.absorb(F(w));
      FriChallenges tmp;
      tmp.fri_pow_response = trial.squeezeFelt();
      if (checkProofOfWork(c.fri_config, tmp)) { dx = trial; break; }
    }
    fp.pow_witness = F(w);
    fprintf(stderr, "[prover] pow_witness = %llu\n", (unsigned long long)w);
  }
  // --- queries ---
  for (int q = 0; q < c.fri_config.num_query_rounds; q++) {
    int idx = (int)(dx.squeezeFelt().v % (u64)M);
    FriQueryRound qr;
    for (int orc_i = 0; orc_i < 4; orc_i++) qr.initial_trees_proof.evals_proofs.push_back({trees[orc_i].leaves[idx], trees[orc_i].open(idx)});
    int qi = idx;
    for (size_t s = 0; s < arities.size(); s++) {
      int a = arities[s], A = 1 << a;
      FriQueryStep st;
      int cos = qi >> a;
      for (int t = 0; t < A; t++) st.evals.push_back(layers[s][(size_t)cos * A + t]);
      st.merkle_proof = step_trees[s].open(cos);
      qr.steps.push_back(st);
      qi = cos;
    }
    fp.query_round_proofs.push_back(qr);
  }
  return out;
}

// ---- JSON writers (schema: SURVEY.md App. A) -----------------------------------------------------------
static std::string jF(F x) { return std::to_string(x.v); }
static std::string jE(const FExt &x) { return "[" + jF(x.r) + "," + jF(x.i) + "]"; }
static std::string jD(const Digest &d) { return "{\"elements\":[" + jF(d.e[0]) + "," + jF(d.e[1]) + "," + jF(d.e[2]) + "," + jF(d.e[3]) + "]}"; }
template <class T, class Fn> static std::string jList(const std::vector<T> &v, Fn f) {
  std::string s = "[";
  for (size_t i = 0; i < v.size(); i++) s += (i ? "," : "") + f(v[i]);
  return s + "]";
}
static std::string jCap(const MerkleCap &c) { return jList(c.roots, jD); }
static std::string jPath(const MerkleProof &p) { return "{\"siblings\":" + jList(p.siblings, jD) + "}"; }
static std::string jEsc(const std::string &s) {
  std::string o = "\"";
  for (char ch : s) { if (ch == '"' || ch == '\\') o += '\\'; o += ch; }
  return o + "\"";
}
static std::string jBool(bool b) { return b ? "true" : "false"; }

static std::string commonJson(const Preset &p, const CommonCircuitData &c) {
  std::ostringstream o;
  std::string strat;
  if (p.fixed_strategy) strat = "{\"Fixed\":" + jList(p.fixed_arities, [](int a) { return std::to_string(a); }) + "}";
  else strat = "{\"ConstantArityBits\":[" + S(p.arity_bits) + "," + S(p.final_poly_bits) + "]}";
  std::string fri_config = "{\"rate_bits\":" + S(p.rate_bits) + ",\"cap_height\":" + S(p.cap_height) + ",\"proof_of_work_bits\":" + S(p.pow_bits) +
                           ",\"reduction_strategy\":" + strat + ",\"num_query_rounds\":" + S(p.num_queries) + "}";
  o << "{\"config\":{\"num_wires\":" << c.num_wires << ",\"num_routed_wires\":" << c.num_routed_wires << ",\"num_constants\":2"
    << ",\"use_base_arithmetic_gate\":true,\"security_bits\":100,\"num_challenges\":" << c.num_challenges
    << ",\"zero_knowledge\":false,\"randomize_unused_wires\":false,\"max_quotient_degree_factor\":" << c.quotient_degree_factor
    << ",\"fri_config\":" << fri_config << "},";
  o << "\"fri_params\":{\"config\":" << fri_config << ",\"hiding\":false,\"degree_bits\":" << c.degree_bits
    << ",\"reduction_arity_bits\":" << jList(c.fri_config.step_arity_bits, [](int a) { return std::to_string(a); }) << "},";
  o << "\"gates\":" << jList(p.gates, [](const GateSpec &g) { return jEsc(g.text); }) << ",";
  o << "\"selectors_info\":{\"selector_indices\":" << jList(c.selector_indices, [](int a) { return std::to_string(a); })
    << ",\"groups\":" << jList(c.selector_groups, [](const Range &r) { return "{\"start\":" + S(r.start) + ",\"end\":" + S(r.end) + "}"; }) << "},";
  o << "\"quotient_degree_factor\":" << c.quotient_degree_factor << ",\"num_gate_constraints\":";
  int maxc = 0;
  {
    EvaluationVars ev;
    ev.local_constants.assign(2, FE0());
    ev.local_wires.assign(c.num_wires, FE0());
    ev.public_inputs_hash.assign(4, F(0));
    for (auto &g : c.gates) maxc = std::max(maxc, (int)gateConstraints(g, ev).size());
  }
  o << maxc << ",\"num_constants\":" << c.num_constants << ",\"num_public_inputs\":" << c.num_public_inputs
    << ",\"k_is\":" << jList(c.k_is, jF) << ",\"num_partial_products\":" << c.num_partial_products
    << ",\"num_lookup_polys\":" << c.num_lookup_polys << ",\"num_lookup_selectors\":" << c.num_lookup_selectors << ",\"luts\":"
    << jList(p.luts, [](const std::vector<std::pair<u64, u64>> &t) {
         return jList(t, [](const std::pair<u64, u64> &e) { return "[" + std::to_string(e.first) + "," + std::to_string(e.second) + "]"; });
       })
    << "}";
  return o.str();
}
static std::string vkeyJson(const VerifierOnlyCircuitData &v) {
  return "{\"constants_sigmas_cap\":" + jCap(v.constants_sigmas_cap) + ",\"circuit_digest\":" + jD(v.circuit_digest) + "}";
}
static std::string proofJson(const ProofWithPublicInputs &pw) {
  const Proof &p = pw.proof;
  const OpeningSet &op = p.openings;
  std::ostringstream o;
  o << "{\"proof\":{\"wires_cap\":" << jCap(p.wires_cap) << ",\"plonk_zs_partial_products_cap\":" << jCap(p.plonk_zs_partial_products_cap)
    << ",\"quotient_polys_cap\":" << jCap(p.quotient_polys_cap) << ",\"openings\":{\"constants\":" << jList(op.constants, jE)
    << ",\"plonk_sigmas\":" << jList(op.plonk_sigmas, jE) << ",\"wires\":" << jList(op.wires, jE) << ",\"plonk_zs\":" << jList(op.plonk_zs, jE)
    << ",\"plonk_zs_next\":" << jList(op.plonk_zs_next, jE) << ",\"partial_products\":" << jList(op.partial_products, jE)
    << ",\"quotient_polys\":" << jList(op.quotient_polys, jE) << ",\"lookup_zs\":" << jList(op.lookup_zs, jE)
    << ",\"lookup_zs_next\":" << jList(op.lookup_zs_next, jE) << "},\"opening_proof\":{\"commit_phase_merkle_caps\":"
    << jList(p.opening_proof.commit_phase_merkle_caps, jCap) << ",\"query_round_proofs\":"
    << jList(p.opening_proof.query_round_proofs,
             [](const FriQueryRound &qr) {
               std::string s = "{\"initial_trees_proof\":{\"evals_proofs\":" +
                               jList(qr.initial_trees_proof.evals_proofs,
                                     [](const std::pair<std::vector<F>, MerkleProof> &ep) { return "[" + jList(ep.first, jF) + "," + jPath(ep.second) + "]"; }) +
                               "},\"steps\":" +
                               jList(qr.steps, [](const FriQueryStep &st) { return "{\"evals\":" + jList(st.evals, jE) + ",\"merkle_proof\":" + jPath(st.merkle_proof) + "}"; }) + "}";
               return s;
             })
    << ",\"final_poly\":{\"coeffs\":" << jList(p.opening_proof.final_poly, jE) << "},\"pow_witness\":" << jF(p.opening_proof.pow_witness)
    << "}},\"public_inputs\":" << jList(pw.public_inputs, jF) << "}";
  return o.str();
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: p2v_prover <preset> <out_prefix> [--seed K] [--corrupt-layer S] [--bad-final] [--bad-witness ROW] [--bad-copy] [--bad-lookup] [--threads T]\n");
    return 2;
  }
  std::string preset = argv[1], prefix = argv[2];
  u64 seed = 1;
  int corrupt_layer = -1, bad_witness_row = -1, threads = (int)std::thread::hardware_concurrency();
  bool bad_final = false, bad_copy = false, bad_lookup = false;
  for (int i = 3; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
    else if (a == "--corrupt-layer" && i + 1 < argc) corrupt_layer = atoi(argv[++i]);
    else if (a == "--bad-final") bad_final = true;
    else if (a == "--bad-copy") bad_copy = true;
    else if (a == "--bad-lookup") bad_lookup = true;
    else if (a == "--bad-witness" && i + 1 < argc) bad_witness_row = atoi(argv[++i]);
    else if (a == "--threads" && i + 1 < argc) threads = atoi(argv[++i]);
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (threads < 1) threads = 1;
  activePermutation() = permutationBulk;  // bit-identical to `permutation` (tests/test_oracle.py)
  Preset p = makePreset(preset);
  CommonCircuitData c = toCommon(p);
  ProverOut out = prove(p, c, seed, corrupt_layer, bad_final, threads, bad_witness_row, bad_copy, bad_lookup);
  // self-check with the verifier restatement (dense-MDS permutation)
  activePermutation() = permutation;
  permCounter() = 0;
  VerifyTrace tr;
  uint32_t st = verifyProofStatus(c, out.vk, out.pw, &tr);
  fprintf(stderr, "[prover] oracle verdict: status=0x%x (%s), permutations=%llu\n", st, st == 0 ? "ACCEPT" : "REJECT", permCounter());
  std::ofstream(prefix + "_common.json") << commonJson(p, c);
  std::ofstream(prefix + "_vkey.json") << vkeyJson(out.vk);
  std::ofstream(prefix + "_proof.json") << proofJson(out.pw);
  printf("%u\n", st);
  return 0;
}
