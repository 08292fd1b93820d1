// Host side of the boundary: the reference's JSON wire format and gate-string grammar.
//   p2v_parse_gate    <- recognizeGate / gateP, src/Gate/Parser.hs:107-240 (Parsec semantics kept:
//                        alternatives tried in the same order, `withEOF` only where the reference has it)
//   p2v_parse_common  <- FromJSON CommonCircuitData, src/Types.hs:47-173
//   p2v_parse_vkey    <- FromJSON VerifierOnlyCircuitData, src/Types.hs:236-240
//   p2v_parse_proof   <- FromJSON ProofWithPublicInputs, src/Types.hs:176-279
//   p2v_shape_layout  <- the flat blob order (include/p2v.h) + oracleWidths, src/Plonk/FRI.hs:56-65
// No GPU needed here.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../../include/p2v.h"
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include "json.hpp"

using namespace p2vhost;

extern thread_local std::string p2v_tls_error;
static const bool g_no_fast_parse = getenv("P2V_NO_FAST_PARSE") != nullptr;  // force the tape decoder (tests)
static int fail(int code, const std::string &msg) {
  p2v_tls_error = msg;
  return code;
}

// ---- Parsec-like cursor ------------------------------------------------------------------------
namespace {
struct Cur {
  const char *s;
  size_t n, pos = 0;
  bool eof() const { return pos == n; }
  bool str(const char *lit) {  // `string lit`
    size_t l = strlen(lit);
    if (n - pos < l || memcmp(s + pos, lit, l) != 0) return false;
    pos += l;
    return true;
  }
  bool chr(char c) {
    if (pos < n && s[pos] == c) { pos++; return true; }
    return false;
  }
  void spaces() {  // Parsec `spaces` = skipMany space (isSpace)
    while (pos < n && (s[pos] == ' ' || s[pos] == '\t' || s[pos] == '\n' || s[pos] == '\r' || s[pos] == '\f' || s[pos] == '\v')) pos++;
  }
  bool digits(std::string &out) {  // many1 digit
    size_t st = pos;
    while (pos < n && s[pos] >= '0' && s[pos] <= '9') pos++;
    if (pos == st) return false;
    out.assign(s + st, pos - st);
    return true;
  }
  bool intP(long long &v) {
    std::string d;
    if (!digits(d)) return false;
    v = 0;
    for (char c : d) {
      if (v > (long long)9e17) return false;
      v = v * 10 + (c - '0');
    }
    return true;
  }
  bool fieldP(uint64_t &v) {  // mkGoldilocks <$> integerP
    std::string d;
    if (!digits(d)) return false;
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    unsigned __int128 acc = 0;
    for (char c : d) acc = (acc * 10 + (unsigned)(c - '0')) % P;
    v = (uint64_t)acc;
    return true;
  }
  bool commaP() {
    if (!chr(',')) return false;
    spaces();
    return true;
  }
  // keyValueP key p = string key; spaces; char ':'; spaces; p; spaces
  bool keyInt(const char *key, long long &v) {
    if (!str(key)) return false;
    spaces();
    if (!chr(':')) return false;
    spaces();
    if (!intP(v)) return false;
    spaces();
    return true;
  }
  template <class Fn> bool keyList(const char *key, Fn elem) {  // keyValueP key (listP elem)
    if (!str(key)) return false;
    spaces();
    if (!chr(':')) return false;
    spaces();
    if (!chr('[')) return false;
    spaces();
    // sepBy elem commaP
    size_t save = pos;
    if (elem(*this)) {
      for (;;) {
        size_t s2 = pos;
        if (!commaP()) { pos = s2; break; }
        if (!elem(*this)) return false;  // sepBy fails if the separator consumed input and the element fails
      }
    } else {
      if (pos != save) return false;
    }
    if (!chr(']')) return false;
    spaces();
    spaces();
    return true;
  }
  // rustStructP name p = string name; spaces; '{'; spaces; p; spaces; '}'; spaces
  bool structOpen(const char *name) {
    if (!str(name)) return false;
    spaces();
    if (!chr('{')) return false;
    spaces();
    return true;
  }
  bool structClose() {
    spaces();
    if (!chr('}')) return false;
    spaces();
    return true;
  }
};

const char *PHANTOM_FIELD = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";

// A gate string may carry any integer (the reference reads them into 64-bit Ints); the shape record holds ints.  Values beyond
// INT_MAX are saturated, never truncated, so that validateShape refuses them instead of seeing a small number.
static int satInt(long long v) { return v > 0x7FFFFFFFLL ? 0x7FFFFFFF : v < -0x7FFFFFFFLL ? -0x7FFFFFFF : (int)v; }

int numConstraints(const p2v_gate &g) {  // lengths of the committed lists, SURVEY.md App. H (64-bit arithmetic, saturated)
  const long long p0 = g.p0, p1 = g.p1, p2 = g.p2;
  switch (g.kind) {
    case P2V_GATE_ARITHMETIC: return satInt(p0);
    case P2V_GATE_ARITHMETIC_EXT: return satInt(2 * p0);
    case P2V_GATE_BASE_SUM: return satInt(1 + p0);
    case P2V_GATE_COSET_INTERP: {
      if (p1 < 2 || p0 < 0 || p0 > 30) return 0;  // refused by validateShape / checkShapeSupported
      long long n_points = 1LL << p0, d = p1;
      long long n_int = (n_points - 2) / (d - 1);
      long long chunks = 1 + (n_points > d ? (n_points - d + d - 2) / (d - 1) : 0);
      long long wchunks = 1 + (g.weights_len > d ? (g.weights_len - d + d - 2) / (d - 1) : 0);
      long long nstuff = std::min(std::min(chunks, wchunks), n_int + 1);
      return satInt(2 + 4 * (nstuff - 1) + 2);
    }
    case P2V_GATE_CONSTANT: return satInt(p0);
    case P2V_GATE_EXPONENTIATION: return satInt(p0 + 1);
    case P2V_GATE_MUL_EXT: return satInt(2 * p0);
    case P2V_GATE_PUBLIC_INPUT: return 4;
    case P2V_GATE_POSEIDON: return 123;
    case P2V_GATE_POSEIDON_MDS: return 24;
    case P2V_GATE_RANDOM_ACCESS: return satInt(p1 * (p0 + 2) + p2);
    case P2V_GATE_REDUCING: case P2V_GATE_REDUCING_EXT: return satInt(2 * p0);
    default: return 0;
  }
}

// One alternative of gateP; returns true on success (like `try p`, the caller resets the cursor).
bool parseAlt(int alt, Cur &c, p2v_gate &g, uint64_t *weights) {
  long long a = 0, b = 0, d = 0;
  switch (alt) {
    case 0:  // arithmeticGateP (withEOF)
      if (!c.structOpen("ArithmeticGate") || !c.keyInt("num_ops", a) || !c.structClose() || !c.eof()) return false;
      g.kind = P2V_GATE_ARITHMETIC; g.p0 = satInt(a); return true;
    case 1:  // arithmeticExtensionGateP (withEOF)
      if (!c.structOpen("ArithmeticExtensionGate") || !c.keyInt("num_ops", a) || !c.structClose() || !c.eof()) return false;
      g.kind = P2V_GATE_ARITHMETIC_EXT; g.p0 = satInt(a); return true;
    case 2:  // baseSumGateP (withEOF): "BaseSumGate { num_limbs: 63 } + Base: 2"
      if (!c.structOpen("BaseSumGate") || !c.keyInt("num_limbs", a) || !c.structClose()) return false;
      if (!c.chr('+')) return false;
      c.spaces();
      if (!c.keyInt("Base", b) || !c.eof()) return false;
      g.kind = P2V_GATE_BASE_SUM; g.p0 = satInt(a); g.p1 = satInt(b); return true;
    case 3: {  // cosetInterpolationGateP (withEOF)
      if (!c.structOpen("CosetInterpolationGate")) return false;
      if (!c.keyInt("subgroup_bits", a) || !c.commaP() || !c.keyInt("degree", b) || !c.commaP()) return false;
      int nw = 0;
      bool overflow = false;
      auto elem = [&](Cur &cc) {
        uint64_t v;
        if (!cc.fieldP(v)) return false;
        if (nw < P2V_MAX_WEIGHTS) weights[nw] = v; else overflow = true;
        nw++;
        return true;
      };
      if (!c.keyList("barycentric_weights", elem)) return false;
      if (!c.commaP() || !c.str(PHANTOM_FIELD)) return false;
      c.spaces();
      if (!c.structClose() || !c.str("<D=2>") || !c.eof()) return false;
      if (overflow) throw JsonError("CosetInterpolationGate: more than P2V_MAX_WEIGHTS barycentric weights");
      g.kind = P2V_GATE_COSET_INTERP; g.p0 = satInt(a); g.p1 = satInt(b); g.weights_len = nw; return true;
    }
    case 4:  // constantGateP (no EOF check in the reference)
      if (!c.structOpen("ConstantGate") || !c.keyInt("num_consts", a) || !c.structClose()) return false;
      g.kind = P2V_GATE_CONSTANT; g.p0 = satInt(a); return true;
    case 5:
      if (!c.structOpen("ExponentiationGate") || !c.keyInt("num_power_bits", a) || !c.structClose()) return false;
      g.kind = P2V_GATE_EXPONENTIATION; g.p0 = satInt(a); return true;
    case 6: {  // lookupGateP
      if (!c.structOpen("LookupGate") || !c.keyInt("num_slots", a) || !c.commaP()) return false;
      auto byte = [&](Cur &cc) { long long v; return cc.intP(v); };
      if (!c.keyList("lut_hash", byte) || !c.structClose()) return false;
      g.kind = P2V_GATE_LOOKUP; g.p0 = satInt(a); return true;
    }
    case 7: {  // lookupTableGateP
      if (!c.structOpen("LookupTableGate") || !c.keyInt("num_slots", a) || !c.commaP()) return false;
      auto byte = [&](Cur &cc) { long long v; return cc.intP(v); };
      if (!c.keyList("lut_hash", byte) || !c.commaP() || !c.keyInt("last_lut_row", b) || !c.structClose()) return false;
      g.kind = P2V_GATE_LOOKUP_TABLE; g.p0 = satInt(a); g.p1 = satInt(b); return true;
    }
    case 8:
      if (!c.structOpen("MulExtensionGate") || !c.keyInt("num_ops", a) || !c.structClose()) return false;
      g.kind = P2V_GATE_MUL_EXT; g.p0 = satInt(a); return true;
    case 9:  // noopGateP = string "NoopGate"
      if (!c.str("NoopGate")) return false;
      g.kind = P2V_GATE_NOOP; return true;
    case 10:
      if (!c.str("PublicInputGate")) return false;
      g.kind = P2V_GATE_PUBLIC_INPUT; return true;
    case 11:  // poseidonGateP (eof)
      if (!c.str("PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=") || !c.intP(a) || !c.str(">") || !c.eof()) return false;
      g.kind = P2V_GATE_POSEIDON; g.p0 = satInt(a); return true;
    case 12:
      if (!c.str("PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=") || !c.intP(a) || !c.str(">") || !c.eof()) return false;
      g.kind = P2V_GATE_POSEIDON_MDS; g.p0 = satInt(a); return true;
    case 13:  // randomAccessGateP
      if (!c.structOpen("RandomAccessGate")) return false;
      if (!c.keyInt("bits", a) || !c.commaP() || !c.keyInt("num_copies", b) || !c.commaP() || !c.keyInt("num_extra_constants", d)) return false;
      if (!c.commaP() || !c.str(PHANTOM_FIELD)) return false;
      c.spaces();
      if (!c.structClose() || !c.str("<D=2>")) return false;
      g.kind = P2V_GATE_RANDOM_ACCESS; g.p0 = satInt(a); g.p1 = satInt(b); g.p2 = satInt(d); return true;
    case 14:  // reducingGateP: the struct, then `optional (string "<D=2>")`
      if (!c.structOpen("ReducingGate") || !c.keyInt("num_coeffs", a)) return false;
      {  // `optional $ string "<D=2>"` sits INSIDE rustStructP's user parser in the reference (:230-234)
        size_t save = c.pos;
        if (!c.str("<D=2>")) { if (c.pos != save) return false; }
      }
      if (!c.structClose()) return false;
      g.kind = P2V_GATE_REDUCING; g.p0 = satInt(a); return true;
    case 15:
      if (!c.structOpen("ReducingExtensionGate") || !c.keyInt("num_coeffs", a)) return false;
      {
        size_t save = c.pos;
        if (!c.str("<D=2>")) { if (c.pos != save) return false; }
      }
      if (!c.structClose()) return false;
      g.kind = P2V_GATE_REDUCING_EXT; g.p0 = satInt(a); return true;
  }
  return false;
}

void parseGateString(const char *s, size_t n, p2v_gate &g, uint64_t *weights) {
  for (int alt = 0; alt < 16; alt++) {
    Cur c{s, n};
    p2v_gate tmp;
    memset(&tmp, 0, sizeof tmp);
    if (parseAlt(alt, c, tmp, weights)) {
      g = tmp;
      g.num_constraints = numConstraints(g);
      return;
    }
  }
  memset(&g, 0, sizeof g);
  g.kind = P2V_GATE_UNKNOWN;  // UnknownGate <$> many anyToken
}

// every integer of a shape that enters an offset, a loop bound or a shift, range-checked BEFORE it is used (a shape may
// come from p2v_parse_common or straight from a caller)
int validateShape(const p2v_shape &s) {
  auto in = [](long long v, long long lo, long long hi) { return v >= lo && v <= hi; };
  if (!in(s.num_wires, 1, 1 << 16) || !in(s.num_routed_wires, 1, P2V_MAX_ROUTED) || s.num_routed_wires > s.num_wires)
    return fail(P2V_E_UNSUPPORTED, "num_wires / num_routed_wires out of range");
  if (!in(s.num_gate_constants, 0, 1 << 12) || !in(s.num_constants, 0, 1 << 12)) return fail(P2V_E_UNSUPPORTED, "number of constants out of range");
  if (!in(s.num_challenges, 1, 8)) return fail(P2V_E_UNSUPPORTED, "num_challenges out of range");
  if (!in(s.degree_bits, 0, 31) || !in(s.rate_bits, 0, 31) || s.degree_bits + s.rate_bits > 31) return fail(P2V_E_UNSUPPORTED, "FRI parameters out of range");
  if (!in(s.cap_height, 0, 24) || s.cap_height > s.degree_bits + s.rate_bits) return fail(P2V_E_UNSUPPORTED, "cap_height out of range");
  if (!in(s.pow_bits, 0, 64) || !in(s.num_queries, 1, 1 << 12)) return fail(P2V_E_UNSUPPORTED, "proof_of_work_bits / num_query_rounds out of range");
  if (!in(s.num_steps, 0, P2V_MAX_STEPS)) return fail(P2V_E_UNSUPPORTED, "too many FRI reduction steps");
  int total = 0;
  for (int i = 0; i < s.num_steps; i++) {
    if (!in(s.step_arity_bits[i], 1, 8)) return fail(P2V_E_UNSUPPORTED, "FRI arity bits out of range");
    total += s.step_arity_bits[i];
  }
  if (total > s.degree_bits || s.final_poly_len != (1 << (s.degree_bits - total))) return fail(P2V_E_UNSUPPORTED, "reduction strategy and final polynomial length disagree");
  if (!in(s.quotient_degree_factor, 1, 64) || !in(s.num_public_inputs, 0, 1 << 20) || !in(s.num_partial_products, 0, 1 << 12))
    return fail(P2V_E_UNSUPPORTED, "circuit parameters out of range");
  if (!in(s.num_lookup_polys, 0, 1 << 10) || !in(s.num_lookup_selectors, 0, 4 + P2V_MAX_LUTS) || !in(s.num_luts, 0, P2V_MAX_LUTS))
    return fail(P2V_E_UNSUPPORTED, "lookup parameters out of range");
  // Plonk/Lookups.hs:57: lookup_deg = quotient_degree_factor - 1 chunks the looking columns; 0 would divide by zero
  if (s.num_luts > 0 && s.quotient_degree_factor < 2) return fail(P2V_E_UNSUPPORTED, "lookup tables need quotient_degree_factor >= 2");
  if (!in(s.num_gates, 0, P2V_MAX_GATES) || !in(s.num_groups, 0, P2V_MAX_GROUPS) || !in(s.num_weights, 0, P2V_MAX_WEIGHTS))
    return fail(P2V_E_UNSUPPORTED, "gate list out of range");
  for (int g = 0; g < s.num_groups; g++)
    if (!in(s.group_start[g], 0, s.num_gates) || !in(s.group_end[g], s.group_start[g], s.num_gates)) return fail(P2V_E_SHAPE, "selector group range out of bounds");
  for (int k = 0; k < s.num_gates; k++) {
    const p2v_gate &g = s.gates[k];
    if (!in(g.group, 0, s.num_groups - 1)) return fail(P2V_E_SHAPE, "selector index out of range ((!!) in Gate/Selector.hs:85)");
    if (!in(g.p0, 0, 1 << 16) || !in(g.p1, 0, 1 << 30) || !in(g.p2, 0, 1 << 16)) return fail(P2V_E_UNSUPPORTED, "gate parameter out of range");
    if (g.weights_len < 0 || g.weights_off < 0 || g.weights_off + g.weights_len > s.num_weights) return fail(P2V_E_SHAPE, "gate weights out of bounds");
  }
  for (int l = 0; l < s.num_luts; l++)
    if (s.lut_off[l] < 0 || s.lut_off[l + 1] < s.lut_off[l]) return fail(P2V_E_SHAPE, "lookup table offsets are not increasing");
  return P2V_OK;
}

int layoutImpl(const p2v_shape &s, p2v_layout &L) {
  memset(&L, 0, sizeof L);
  int rc = validateShape(s);
  if (rc != P2V_OK) return rc;
  {  // sizes in 64 bits first: the 32-bit offsets below are only computed when everything fits
    long long capw = 4LL << s.cap_height, r = s.num_challenges;
    long long proof = 3 * capw + 2LL * (s.num_constants + s.num_routed_wires + s.num_wires + 2 * r + r * s.num_partial_products +
                                         r * s.quotient_degree_factor + 2 * r * s.num_lookup_polys) +
                      s.num_steps * capw + 2LL * s.final_poly_len + 1 + s.num_public_inputs;
    long long plen = s.degree_bits + s.rate_bits - s.cap_height;
    long long query = (long long)s.num_constants + s.num_routed_wires + s.num_wires + r * (1 + s.num_partial_products + s.num_lookup_polys) +
                      r * s.quotient_degree_factor + 16 * plen;
    for (int st = 0; st < s.num_steps; st++) query += (2LL << s.step_arity_bits[st]) + 4 * plen;
    if (proof + (long long)s.num_queries * query > 0x7fffffffLL) return fail(P2V_E_UNSUPPORTED, "proof blob too large");
  }
  int r = s.num_challenges;
  int cap = 4 << s.cap_height;
  L.cap_words = cap;
  int pos = 0;
  L.off_wires_cap = pos; pos += cap;
  L.off_zs_pp_cap = pos; pos += cap;
  L.off_quotient_cap = pos; pos += cap;
  auto open = [&](int count, int32_t &off, int32_t &n) { off = pos; n = count; pos += 2 * count; };
  open(s.num_constants, L.off_open_constants, L.n_open_constants);
  open(s.num_routed_wires, L.off_open_sigmas, L.n_open_sigmas);
  open(s.num_wires, L.off_open_wires, L.n_open_wires);
  open(r, L.off_open_zs, L.n_open_zs);
  open(r, L.off_open_zs_next, L.n_open_zs_next);
  open(r * s.num_partial_products, L.off_open_pp, L.n_open_pp);
  open(r * s.quotient_degree_factor, L.off_open_quotient, L.n_open_quotient);
  open(r * s.num_lookup_polys, L.off_open_lookup_zs, L.n_open_lookup_zs);
  open(r * s.num_lookup_polys, L.off_open_lookup_zs_next, L.n_open_lookup_zs_next);
  L.off_commit_caps = pos; pos += s.num_steps * cap;
  L.off_final_poly = pos; pos += 2 * s.final_poly_len;
  L.off_pow_witness = pos; pos += 1;
  L.off_public_inputs = pos; pos += s.num_public_inputs;
  L.proof_words = pos;
  L.oracle_width[0] = s.num_constants + s.num_routed_wires;
  L.oracle_width[1] = s.num_wires;
  L.oracle_width[2] = r * (1 + s.num_partial_products + s.num_lookup_polys);
  L.oracle_width[3] = r * s.quotient_degree_factor;
  L.init_path_len = s.degree_bits + s.rate_bits - s.cap_height;
  if (L.init_path_len < 0) return fail(P2V_E_UNSUPPORTED, "cap_height exceeds the LDE tree height");
  int q = 0;
  for (int o = 0; o < 4; o++) {
    L.q_off_leaf[o] = q; q += L.oracle_width[o];
    L.q_off_sibs[o] = q; q += 4 * L.init_path_len;
  }
  int bits = s.degree_bits + s.rate_bits;
  for (int st = 0; st < s.num_steps; st++) {
    int a = s.step_arity_bits[st];
    L.q_off_step_evals[st] = q; q += 2 << a;
    bits -= a;
    int plen = bits - s.cap_height;
    if (plen < 0) return fail(P2V_E_UNSUPPORTED, "FRI step tree is smaller than the Merkle cap");
    L.step_path_len[st] = plen;
    L.q_off_step_sibs[st] = q; q += 4 * plen;
  }
  L.query_words = q;
  long long total = (long long)L.proof_words + (long long)s.num_queries * q;
  if (total > 0x7fffffffLL) return fail(P2V_E_UNSUPPORTED, "proof blob too large");
  L.blob_words = (int)total;
  L.vkey_words = cap + 4;
  return P2V_OK;
}

void putDigest(const JValue &v, uint64_t *&w) {
  const auto &el = v.at("elements").list();
  if (el.size() != 4) throw JsonError("Digest: expecting 4 elements");  // listToDigest [a,b,c,d]
  for (int i = 0; i < 4; i++) *w++ = el[i].felt();
}
struct ShapeErr : std::runtime_error { using std::runtime_error::runtime_error; };
void putCap(const JValue &v, int ncap, uint64_t *&w, const char *what) {
  const auto &ds = v.list();
  if ((int)ds.size() != ncap) throw ShapeErr(std::string(what) + ": cap has wrong size (validateMerkleCapLength)");
  for (auto &d : ds) putDigest(d, w);
}
void putExts(const JValue &v, int count, uint64_t *&w, const char *what) {
  const auto &xs = v.list();
  if ((int)xs.size() != count) throw ShapeErr(std::string(what) + ": expected " + std::to_string(count) + " entries, got " + std::to_string(xs.size()));
  for (auto &x : xs) {
    const auto &pr = x.list();
    if (pr.size() != 2) throw JsonError("FExt: expecting [a,b]");
    *w++ = pr[0].felt();
    *w++ = pr[1].felt();
  }
}
void putPath(const JValue &v, int len, uint64_t *&w, const char *what) {
  const auto &sib = v.at("siblings").list();
  if ((int)sib.size() != len) throw ShapeErr(std::string(what) + ": Merkle path has " + std::to_string(sib.size()) + " siblings, shape expects " + std::to_string(len));
  for (auto &d : sib) putDigest(d, w);
}

// ---- fast path for p2v_parse_proof -----------------------------------------------------------------------------
// serde_json writes `ProofWithPublicInputs` with the keys in declaration order and plain non-negative integers, so
// the common case needs no tree at all: one forward scan that checks the expected punctuation and keys and converts
// every number straight into its slot of the blob (eight digits at a time).  ANY deviation — other key order, a
// negative / fractional / exponent token, a count that does not match the shape, malformed text — makes it return
// false and the tape decoder below takes over (and produces the error, if there is one).  aeson accepts any key
// order, so the slow path stays the definition; tests/test_host.py checks both against each other.
struct FastScan {
  const char *p, *end;
  bool ok = true;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool ch(char c) {
    ws();
    if (p < end && *p == c) { p++; return true; }
    return ok = false;
  }
  bool peek(char c) { ws(); return p < end && *p == c; }
  bool key(const char *k) {  // "k":
    ws();
    size_t n = strlen(k);
    if ((size_t)(end - p) < n + 3 || *p != '"' || memcmp(p + 1, k, n) != 0 || p[n + 1] != '"') return ok = false;
    p += n + 2;
    return ch(':');
  }
  bool felt(uint64_t &v) {
    ws();
    const char *t = p;
    while (end - p >= 8 && eightDigits(p)) p += 8;
    while (p < end && (unsigned)(*p - '0') < 10) p++;
    size_t nd = (size_t)(p - t);
    if (nd == 0 || nd > 38) return ok = false;
    if (nd > 1 && t[0] == '0') return ok = false;  // leading zeros are not JSON (aeson rejects them): let the tape reader decide
    if (p < end && (*p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) return ok = false;
    const uint64_t P = 0xFFFFFFFF00000001ULL;
    size_t n1 = nd > 19 ? nd - 19 : 0, i = 0;
    uint64_t hi = 0, lo = 0;
    for (; i < n1; i++) hi = hi * 10 + (unsigned)(t[i] - '0');
    for (; i + 8 <= nd; i += 8) lo = lo * 100000000ULL + parseEightDigits(t + i);
    for (; i < nd; i++) lo = lo * 10 + (unsigned)(t[i] - '0');
    if (n1 == 0) v = lo >= P ? lo - P : lo;
    else {
      static const uint64_t TEN19 = 10000000000000000000ULL;
      v = JValue::reduce128((unsigned __int128)hi * TEN19 + lo);
    }
    return true;
  }
  // [a, b, ...] of exactly `count` field elements
  bool felts(int count, uint64_t *&w) {
    if (!ch('[')) return false;
    for (int i = 0; i < count; i++) {
      if (i && !ch(',')) return false;
      if (!felt(*w++)) return false;
    }
    return ch(']');
  }
  bool digest(uint64_t *&w) { return ch('{') && key("elements") && felts(4, w) && ch('}'); }
  bool digests(int count, uint64_t *&w) {
    if (!ch('[')) return false;
    for (int i = 0; i < count; i++) {
      if (i && !ch(',')) return false;
      if (!digest(w)) return false;
    }
    return ch(']');
  }
  bool exts(int count, uint64_t *&w) {
    if (!ch('[')) return false;
    for (int i = 0; i < count; i++) {
      if (i && !ch(',')) return false;
      if (!felts(2, w)) return false;
    }
    return ch(']');
  }
  bool path(int len, uint64_t *&w) { return ch('{') && key("siblings") && digests(len, w) && ch('}'); }
};

static bool fastParseProof(const char *json, size_t len, const p2v_shape &s, const p2v_layout &L, uint64_t *out) {
  FastScan f{json, json + len};
  int ncap = 1 << s.cap_height;
  uint64_t *w = out;
  if (!(f.ch('{') && f.key("proof") && f.ch('{'))) return false;
  if (!(f.key("wires_cap") && f.digests(ncap, w) && f.ch(','))) return false;
  if (!(f.key("plonk_zs_partial_products_cap") && f.digests(ncap, w) && f.ch(','))) return false;
  if (!(f.key("quotient_polys_cap") && f.digests(ncap, w) && f.ch(','))) return false;
  if (!(f.key("openings") && f.ch('{'))) return false;
  const char *names[9] = {"constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys", "lookup_zs", "lookup_zs_next"};
  const int counts[9] = {L.n_open_constants, L.n_open_sigmas, L.n_open_wires, L.n_open_zs, L.n_open_zs_next, L.n_open_pp, L.n_open_quotient, L.n_open_lookup_zs, L.n_open_lookup_zs_next};
  for (int i = 0; i < 9; i++) {
    if (i && !f.ch(',')) return false;
    if (!(f.key(names[i]) && f.exts(counts[i], w))) return false;
  }
  if (!(f.ch('}') && f.ch(',') && f.key("opening_proof") && f.ch('{'))) return false;
  if (!(f.key("commit_phase_merkle_caps") && f.ch('['))) return false;
  for (int i = 0; i < s.num_steps; i++) {
    if (i && !f.ch(',')) return false;
    if (!f.digests(ncap, w)) return false;
  }
  if (!(f.ch(']') && f.ch(','))) return false;
  if (w - out != L.off_final_poly) return false;
  // query rounds come before final_poly / pow_witness / public_inputs in the text, after them in the blob
  uint64_t *q = out + L.proof_words;
  if (!(f.key("query_round_proofs") && f.ch('['))) return false;
  for (int r = 0; r < s.num_queries; r++) {
    if (r && !f.ch(',')) return false;
    uint64_t *q0 = q;
    if (!(f.ch('{') && f.key("initial_trees_proof") && f.ch('{') && f.key("evals_proofs") && f.ch('['))) return false;
    for (int o = 0; o < 4; o++) {
      if (o && !f.ch(',')) return false;
      if (!(f.ch('[') && f.felts(L.oracle_width[o], q) && f.ch(',') && f.path(L.init_path_len, q) && f.ch(']'))) return false;
    }
    if (!(f.ch(']') && f.ch('}') && f.ch(',') && f.key("steps") && f.ch('['))) return false;
    for (int st = 0; st < s.num_steps; st++) {
      if (st && !f.ch(',')) return false;
      if (!(f.ch('{') && f.key("evals") && f.exts(1 << s.step_arity_bits[st], q) && f.ch(',') && f.key("merkle_proof") &&
            f.path(L.step_path_len[st], q) && f.ch('}')))
        return false;
    }
    if (!(f.ch(']') && f.ch('}'))) return false;
    if (q - q0 != L.query_words) return false;
  }
  if (!(f.ch(']') && f.ch(','))) return false;
  if (!(f.key("final_poly") && f.ch('{') && f.key("coeffs") && f.exts(s.final_poly_len, w) && f.ch('}') && f.ch(','))) return false;
  if (!(f.key("pow_witness") && f.felt(*w++) && f.ch('}') && f.ch('}') && f.ch(','))) return false;
  if (!(f.key("public_inputs") && f.felts(s.num_public_inputs, w) && f.ch('}'))) return false;
  f.ws();
  return f.p == f.end && w - out == L.proof_words;
}
}  // namespace

extern "C" {

int p2v_parse_gate(const char *str, size_t len, p2v_gate *out, uint64_t *weights) {
  if (!str || !out || !weights) return fail(P2V_E_INVALID, "p2v_parse_gate: NULL argument");
  try {
    parseGateString(str, len, *out, weights);
  } catch (const std::exception &e) {
    return fail(P2V_E_PARSE, e.what());
  }
  return P2V_OK;
}

// what the kernels implement, beyond the ranges of validateShape (was in verify_api.cu; host-only so that a circuit description can
// be vetted, and fuzzed, without a GPU)
static int checkShapeSupported(const p2v_shape &s) {
  if (s.num_challenges > P2V_MAX_CHALLENGES)
    return fail(P2V_E_UNSUPPORTED, "num_challenges > 4 is not supported");
  for (int k = 0; k < s.num_gates; k++) {
    const p2v_gate &g = s.gates[k];
    switch (g.kind) {
      case P2V_GATE_UNKNOWN:
        return fail(P2V_E_UNSUPPORTED, "gateConstraints: unknown gate (Gate/Constraints.hs:108)");
      case P2V_GATE_POSEIDON:
      case P2V_GATE_POSEIDON_MDS:
        if (g.p0 != 12) return fail(P2V_E_UNSUPPORTED, "gateConstraints/PoseidonGate: unsupported width (Gate/Constraints.hs:93,97)");
        break;
      case P2V_GATE_COSET_INTERP:
        if (g.p0 < 1 || g.p0 > 5 || g.p1 < 2 || g.weights_len != (1 << g.p0))
          return fail(P2V_E_UNSUPPORTED, "CosetInterpolationGate: need 2^subgroup_bits weights, degree >= 2, subgroup_bits <= 5");
        break;
      case P2V_GATE_RANDOM_ACCESS:
        if (g.p0 < 0 || g.p0 > 7) return fail(P2V_E_UNSUPPORTED, "RandomAccessGate: bits out of range");
        break;
      case P2V_GATE_BASE_SUM:
        // the limb range check is a product over the base (Gate/Constraints.hs: prod (limb - k), k < B): B per limb and proof
        if (g.p1 < 0 || g.p1 > 256) return fail(P2V_E_UNSUPPORTED, "BaseSumGate: base above 256");
        break;
      default: break;
    }
  }
  // every wire / constant index a gate touches must exist (Array `!` raises in the reference)
  // (64-bit: the parameters are bounded by validateShape, their products are not — num_copies = 2^30 would wrap an int)
  auto need = [&](int k) -> long long {
    const p2v_gate &g = s.gates[k];
    const long long p0 = g.p0, p1 = g.p1, p2 = g.p2;
    switch (g.kind) {
      case P2V_GATE_ARITHMETIC: return 4 * p0;
      case P2V_GATE_ARITHMETIC_EXT: return 8 * p0;
      case P2V_GATE_MUL_EXT: return 6 * p0;
      case P2V_GATE_BASE_SUM: return 1 + p0;
      case P2V_GATE_CONSTANT: return p0;
      case P2V_GATE_PUBLIC_INPUT: return 4;
      case P2V_GATE_EXPONENTIATION: return 2 * p0 + 2;
      case P2V_GATE_POSEIDON: return 135;
      case P2V_GATE_POSEIDON_MDS: return 48;
      case P2V_GATE_RANDOM_ACCESS: return ((2 + (1LL << p0)) * p1 + p2 + p0 * p1);
      case P2V_GATE_REDUCING: return p0 ? 3 * p0 + 4 : 0;
      case P2V_GATE_REDUCING_EXT: return p0 ? 4 * p0 + 4 : 0;
      case P2V_GATE_COSET_INTERP: {
        long long np = 1LL << p0, ni = (np - 2) / (p1 - 1);
        return 1 + 2 * (np + 2) + 4 * ni + 2;
      }
      default: return 0;
    }
  };
  for (int k = 0; k < s.num_gates; k++) {
    if (need(k) > s.num_wires) return fail(P2V_E_UNSUPPORTED, "a gate reads a wire beyond num_wires (array index error in the reference)");
    const p2v_gate &g = s.gates[k];
    int nconst = g.kind == P2V_GATE_CONSTANT ? g.p0 : g.kind == P2V_GATE_RANDOM_ACCESS ? g.p2 : (g.kind == P2V_GATE_ARITHMETIC || g.kind == P2V_GATE_ARITHMETIC_EXT) ? 2 : g.kind == P2V_GATE_MUL_EXT ? 1 : 0;
    if (g.p0 > 0 && nconst > s.num_gate_constants) return fail(P2V_E_UNSUPPORTED, "a gate reads a constant beyond config.num_constants");
  }
  if (s.num_luts > 0 && s.num_lookup_polys < 2) return fail(P2V_E_UNSUPPORTED, "lookup tables need at least 2 lookup polynomials");
  if (s.num_luts > 0 && (s.num_routed_wires / 3 < 1 || 3 * (s.num_routed_wires / 3) > s.num_wires))
    return fail(P2V_E_UNSUPPORTED, "lookup slots exceed the wires");
  return P2V_OK;
}


int p2v_shape_check(const p2v_shape *shape) {
  if (!shape) return fail(P2V_E_INVALID, "p2v_shape_check: NULL argument");
  int rc = validateShape(*shape);
  if (rc) return rc;
  return checkShapeSupported(*shape);
}

int p2v_shape_layout(const p2v_shape *shape, p2v_layout *out) {
  if (!shape || !out) return fail(P2V_E_INVALID, "p2v_shape_layout: NULL argument");
  return layoutImpl(*shape, *out);
}

int p2v_challenges_words(const p2v_shape *s) {
  if (!s) return 0;
  int r = s->num_challenges;
  return 3 * r + (s->num_lookup_polys > 0 ? 4 * r : 0) + 2 + 2 + 2 * s->num_steps + 1 + s->num_queries;
}

void p2v_shape_free(p2v_shape *shape) {
  if (shape && shape->lut_pairs) {
    delete[] shape->lut_pairs;
    shape->lut_pairs = nullptr;
  }
}

int p2v_parse_common(const char *json, size_t len, p2v_shape *out) {
  if (!json || !out) return fail(P2V_E_INVALID, "p2v_parse_common: NULL argument");
  memset(out, 0, sizeof *out);
  uint64_t *luts = nullptr;
  try {
    JsonDoc doc(json, len);
    JValue root = doc.root();
    const JValue &cfg = root.at("config");
    p2v_shape &s = *out;
    // long long -> int32 with a range check (an out-of-range value must not wrap into a plausible one)
    auto i32 = [](const JValue &v, const char *what) {
      long long x = v.integer();
      if (x < -(1LL << 30) || x > (1LL << 30)) throw JsonError(std::string(what) + ": integer out of range");
      return (int)x;
    };
    s.num_wires = i32(cfg.at("num_wires"), "num_wires");
    s.num_routed_wires = i32(cfg.at("num_routed_wires"), "num_routed_wires");
    s.num_gate_constants = i32(cfg.at("num_constants"), "config.num_constants");
    s.num_challenges = i32(cfg.at("num_challenges"), "num_challenges");
    // the remaining CircuitConfig fields are required by the generic aeson instance (Types.hs:87)
    cfg.at("use_base_arithmetic_gate").boolean();
    cfg.at("security_bits").integer();
    bool zk = cfg.at("zero_knowledge").boolean();
    cfg.at("randomize_unused_wires").boolean();
    cfg.at("max_quotient_degree_factor").integer();
    const JValue &fc = cfg.at("fri_config");
    const JValue &fp = root.at("fri_params");
    const JValue &fc2 = fp.at("config");
    auto friInts = [&](const JValue &c, int v[4]) {
      v[0] = i32(c.at("rate_bits"), "rate_bits");
      v[1] = i32(c.at("cap_height"), "cap_height");
      v[2] = i32(c.at("proof_of_work_bits"), "proof_of_work_bits");
      v[3] = i32(c.at("num_query_rounds"), "num_query_rounds");
    };
    int a[4], b[4];
    friInts(fc, a);
    friInts(fc2, b);
    // The reference reads some FRI parameters from config.fri_config and others from
    // fri_params.config (Challenge/FRI.hs:68-69 vs Plonk/FRI.hs:91-92,190-191); a real Plonky2
    // circuit has them equal, and a fixed-shape batch needs them equal.
    for (int i = 0; i < 4; i++)
      if (a[i] != b[i]) return fail(P2V_E_UNSUPPORTED, "config.fri_config and fri_params.config differ");
    s.rate_bits = a[0]; s.cap_height = a[1]; s.pow_bits = a[2]; s.num_queries = a[3];
    bool hiding = fp.at("hiding").boolean();
    if (zk || hiding) return fail(P2V_E_UNSUPPORTED, "zero-knowledge / hiding (salted leaves) is not supported (reference README.md:34)");
    s.degree_bits = i32(fp.at("degree_bits"), "degree_bits");
    fp.at("reduction_arity_bits").list();
    if (s.degree_bits < 0 || s.degree_bits + s.rate_bits > 31 || s.rate_bits < 0 || s.cap_height < 0 || s.pow_bits < 0 || s.pow_bits > 64)
      return fail(P2V_E_UNSUPPORTED, "FRI parameters out of range");
    // expandReductionStrategy, Plonk/FRI.hs:337-354 (NOT fri_params.reduction_arity_bits)
    const JValue &strat = fc.at("reduction_strategy");
    if (!strat.isObject() || strat.objSize() != 1) throw JsonError("FriReductionStrategy: expecting a singleton object");
    const std::string key = strat.objKey(0);
    const JValue sval = strat.objVal(0);
    s.num_steps = 0;
    int total = 0;
    auto addStep = [&](int ar) {
      if (s.num_steps >= P2V_MAX_STEPS) throw JsonError("too many FRI reduction steps");
      if (ar < 1 || ar > 8) throw JsonError("FRI arity bits out of range");
      s.step_arity_bits[s.num_steps++] = ar;
      total += ar;
    };
    if (key == "ConstantArityBits") {
      const auto &ab = sval.list();
      if (ab.size() != 2) throw JsonError("ConstantArityBits: expecting [arity_bits, final_poly_bits]");
      int arity = i32(ab[0], "arity_bits"), final_bits = i32(ab[1], "final_poly_bits");
      if (arity < 1) throw JsonError("ConstantArityBits: arity_bits < 1 does not terminate");
      for (int logn = s.degree_bits; logn > final_bits; logn -= arity) addStep(arity);
    } else if (key == "Fixed") {
      for (auto &x : sval.list()) addStep(i32(x, "Fixed arity"));
    } else if (key == "MinSize") {
      return fail(P2V_E_UNSUPPORTED, "reduction strategy not implemented (Plonk/FRI.hs:342)");
    } else {
      throw JsonError("FromJSON/FriReductionStrategy: unrecognized FRI reduction strategy: `" + key + "`");
    }
    if (total > s.degree_bits) return fail(P2V_E_UNSUPPORTED, "reduction strategy folds below degree 1");
    s.final_poly_len = 1 << (s.degree_bits - total);
    s.quotient_degree_factor = i32(root.at("quotient_degree_factor"), "quotient_degree_factor");
    root.at("num_gate_constraints").integer();
    s.num_constants = i32(root.at("num_constants"), "num_constants");
    s.num_public_inputs = i32(root.at("num_public_inputs"), "num_public_inputs");
    s.num_partial_products = i32(root.at("num_partial_products"), "num_partial_products");
    s.num_lookup_polys = i32(root.at("num_lookup_polys"), "num_lookup_polys");
    s.num_lookup_selectors = i32(root.at("num_lookup_selectors"), "num_lookup_selectors");
    const auto &kis = root.at("k_is").list();
    if ((int)kis.size() > P2V_MAX_ROUTED || s.num_routed_wires > P2V_MAX_ROUTED) return fail(P2V_E_UNSUPPORTED, "more than P2V_MAX_ROUTED routed wires");
    // Vanishing.hs:107 zips k_is with the wires: extra k_is are ignored, missing ones shorten the product
    if ((int)kis.size() != s.num_routed_wires) return fail(P2V_E_UNSUPPORTED, "k_is length differs from num_routed_wires");
    for (size_t i = 0; i < kis.size(); i++) s.k_is[i] = kis[i].felt();
    // gates + selectors
    const auto &gates = root.at("gates").list();
    const JValue &sel = root.at("selectors_info");
    const auto &sidx = sel.at("selector_indices").list();
    const auto &groups = sel.at("groups").list();
    if (gates.size() > P2V_MAX_GATES) return fail(P2V_E_UNSUPPORTED, "more than P2V_MAX_GATES gates");
    if (groups.size() > P2V_MAX_GROUPS) return fail(P2V_E_UNSUPPORTED, "more than P2V_MAX_GROUPS selector groups");
    if (sidx.size() != gates.size()) return fail(P2V_E_SHAPE, "selector_indices and gates have different lengths");
    s.num_gates = (int)gates.size();
    s.num_groups = (int)groups.size();
    for (size_t g = 0; g < groups.size(); g++) {
      s.group_start[g] = i32(groups[g].at("start"), "group start");
      s.group_end[g] = i32(groups[g].at("end"), "group end");
    }
    for (size_t k = 0; k < gates.size(); k++) {
      uint64_t w[P2V_MAX_WEIGHTS];
      p2v_gate g;
      const std::string &txt = gates[k].str();
      parseGateString(txt.data(), txt.size(), g, w);
      g.group = i32(sidx[k], "selector index");
      if (g.group < 0 || g.group >= s.num_groups) return fail(P2V_E_SHAPE, "selector index out of range ((!!) in Gate/Selector.hs:85)");
      if (g.weights_len) {
        if (s.num_weights + g.weights_len > P2V_MAX_WEIGHTS) return fail(P2V_E_UNSUPPORTED, "too many barycentric weights");
        g.weights_off = s.num_weights;
        for (int i = 0; i < g.weights_len; i++) s.weights[s.num_weights++] = w[i];
      }
      s.gates[k] = g;
    }
    // luts
    const auto &lutl = root.at("luts").list();
    if (lutl.size() > P2V_MAX_LUTS) return fail(P2V_E_UNSUPPORTED, "more than P2V_MAX_LUTS lookup tables");
    s.num_luts = (int)lutl.size();
    size_t total_pairs = 0;
    for (auto &t : lutl) total_pairs += t.list().size();
    if (total_pairs) luts = new uint64_t[2 * total_pairs];
    size_t pos = 0;
    for (size_t l = 0; l < lutl.size(); l++) {
      s.lut_off[l] = (int)pos;
      for (auto &e : lutl[l].list()) {
        const auto &pr = e.list();
        if (pr.size() != 2) throw JsonError("LookupTable: expecting [inp,out] pairs");
        // Word64 pairs, then toF (Types.hs:28-32)
        luts[2 * pos] = pr[0].felt();
        luts[2 * pos + 1] = pr[1].felt();
        pos++;
      }
    }
    s.lut_off[lutl.size()] = (int)pos;
    s.lut_pairs = luts;
    luts = nullptr;
    // getSelectorConfig, Gate/Selector.hs:31-47 (raised once per proof in the reference)
    int expected = s.num_luts == 0 ? 0 : 4 + s.num_luts;
    if (s.num_lookup_selectors != expected) {
      p2v_shape_free(out);
      return fail(P2V_E_SHAPE, "getSelectorConfig: fatal: num_lookup_selectors /= (4 + #nluts)");
    }
    if (s.num_constants != s.num_groups + s.num_lookup_selectors + s.num_gate_constants) {
      p2v_shape_free(out);
      return fail(P2V_E_SHAPE, "getSelectorConfig: fatal: constant columns tally does not add up!");
    }
    {
      int vrc = validateShape(s);
      if (vrc != P2V_OK) {
        p2v_shape_free(out);
        return vrc;
      }
    }
  } catch (const std::exception &e) {
    delete[] luts;
    p2v_shape_free(out);
    return fail(P2V_E_PARSE, std::string("p2v_parse_common: ") + e.what());
  }
  return P2V_OK;
}

int p2v_parse_vkey(const char *json, size_t len, const p2v_shape *shape, uint64_t *out) {
  if (!json || !shape || !out) return fail(P2V_E_INVALID, "p2v_parse_vkey: NULL argument");
  try {
    JsonDoc doc(json, len);
    JValue root = doc.root();
    uint64_t *w = out;
    putCap(root.at("constants_sigmas_cap"), 1 << shape->cap_height, w, "constants_sigmas_cap");
    putDigest(root.at("circuit_digest"), w);
  } catch (const ShapeErr &e) {
    return fail(P2V_E_SHAPE, std::string("p2v_parse_vkey: ") + e.what());
  } catch (const std::exception &e) {
    return fail(P2V_E_PARSE, std::string("p2v_parse_vkey: ") + e.what());
  }
  return P2V_OK;
}

int p2v_parse_proof(const char *json, size_t len, const p2v_shape *shape, uint64_t *out) {
  if (!json || !shape || !out) return fail(P2V_E_INVALID, "p2v_parse_proof: NULL argument");
  p2v_layout L;
  int rc = layoutImpl(*shape, L);
  if (rc) return rc;
  if (!g_no_fast_parse && fastParseProof(json, len, *shape, L, out)) return P2V_OK;
  try {
    JsonDoc doc(json, len);
    JValue root = doc.root();
    const JValue &proof = root.at("proof");
    int ncap = 1 << shape->cap_height;
    uint64_t *w = out;
    putCap(proof.at("wires_cap"), ncap, w, "wires_cap");
    putCap(proof.at("plonk_zs_partial_products_cap"), ncap, w, "plonk_zs_partial_products_cap");
    putCap(proof.at("quotient_polys_cap"), ncap, w, "quotient_polys_cap");
    const JValue &op = proof.at("openings");
    putExts(op.at("constants"), L.n_open_constants, w, "openings.constants");
    putExts(op.at("plonk_sigmas"), L.n_open_sigmas, w, "openings.plonk_sigmas");
    putExts(op.at("wires"), L.n_open_wires, w, "openings.wires");
    putExts(op.at("plonk_zs"), L.n_open_zs, w, "openings.plonk_zs");
    putExts(op.at("plonk_zs_next"), L.n_open_zs_next, w, "openings.plonk_zs_next");
    putExts(op.at("partial_products"), L.n_open_pp, w, "openings.partial_products");
    putExts(op.at("quotient_polys"), L.n_open_quotient, w, "openings.quotient_polys");
    putExts(op.at("lookup_zs"), L.n_open_lookup_zs, w, "openings.lookup_zs");
    putExts(op.at("lookup_zs_next"), L.n_open_lookup_zs_next, w, "openings.lookup_zs_next");
    const JValue &fri = proof.at("opening_proof");
    const auto &ccaps = fri.at("commit_phase_merkle_caps").list();
    if ((int)ccaps.size() != shape->num_steps) throw ShapeErr("commit_phase_merkle_caps: wrong number of caps (safeZipWith4)");
    for (auto &c : ccaps) putCap(c, ncap, w, "commit_phase_merkle_caps");
    putExts(fri.at("final_poly").at("coeffs"), shape->final_poly_len, w, "final_poly");
    *w++ = fri.at("pow_witness").felt();
    const auto &pis = root.at("public_inputs").list();
    if ((int)pis.size() != shape->num_public_inputs) throw ShapeErr("public_inputs: wrong length");
    for (auto &x : pis) *w++ = x.felt();
    if (w - out != L.proof_words) throw ShapeErr("internal: proof part size mismatch");
    const auto &rounds = fri.at("query_round_proofs").list();
    if ((int)rounds.size() != shape->num_queries) throw ShapeErr("query_round_proofs: wrong number of rounds (safeZipWith)");
    for (auto &rd : rounds) {
      uint64_t *q0 = w;
      const auto &eps = rd.at("initial_trees_proof").at("evals_proofs").list();
      if (eps.size() != 4) throw ShapeErr("checkInitialTreeProofs: expecting 4 Merkle proofs for the 4 oracles");
      for (int o = 0; o < 4; o++) {
        const auto &pr = eps[o].list();
        if (pr.size() != 2) throw JsonError("evals_proofs: expecting [leaf, proof] pairs");
        const auto &leaf = pr[0].list();
        if ((int)leaf.size() != L.oracle_width[o]) throw ShapeErr("buildListOracle: list size do not match the expected");
        for (auto &x : leaf) *w++ = x.felt();
        putPath(pr[1], L.init_path_len, w, "initial tree proof");
      }
      const auto &steps = rd.at("steps").list();
      if ((int)steps.size() != shape->num_steps) throw ShapeErr("steps: wrong number of folding steps (safeZipWith4)");
      for (int st = 0; st < shape->num_steps; st++) {
        putExts(steps[st].at("evals"), 1 << shape->step_arity_bits[st], w, "step evals (reduction strategy incompatibility)");
        putPath(steps[st].at("merkle_proof"), L.step_path_len[st], w, "step proof");
      }
      if (w - q0 != L.query_words) throw ShapeErr("internal: query part size mismatch");
    }
  } catch (const ShapeErr &e) {
    return fail(P2V_E_SHAPE, std::string("p2v_parse_proof: ") + e.what());
  } catch (const std::exception &e) {
    return fail(P2V_E_PARSE, std::string("p2v_parse_proof: ") + e.what());
  }
  return P2V_OK;
}

int p2v_parse_proofs(const char *const *jsons, const size_t *lens, size_t n, const p2v_shape *shape, uint64_t *blobs,
                     int threads, int32_t *rcs) {
  if ((n && (!jsons || !lens || !blobs)) || !shape) return fail(P2V_E_INVALID, "p2v_parse_proofs: NULL argument");
  p2v_layout L;
  int rc = layoutImpl(*shape, L);
  if (rc) return rc;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  threads = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, n));
  std::atomic<size_t> next_item{0};
  std::mutex mu;
  size_t first_bad = n;
  int first_rc = P2V_OK;
  std::string first_msg;
  auto work = [&]() {
    for (;;) {
      size_t i = next_item.fetch_add(1);
      if (i >= n) return;
      uint64_t *out = blobs + i * (size_t)L.blob_words;
      int r = jsons[i] ? p2v_parse_proof(jsons[i], lens[i], shape, out) : fail(P2V_E_INVALID, "p2v_parse_proofs: NULL text");
      if (rcs) rcs[i] = r;
      if (r != P2V_OK) {
        std::fill(out, out + L.blob_words, (uint64_t)0);
        std::lock_guard<std::mutex> lk(mu);
        if (i < first_bad) { first_bad = i; first_rc = r; first_msg = "proof " + std::to_string(i) + ": " + p2v_tls_error; }
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(work);
  work();
  for (auto &t : pool) t.join();
  if (first_rc != P2V_OK) return fail(first_rc, first_msg);
  return P2V_OK;
}

}  // extern "C"
