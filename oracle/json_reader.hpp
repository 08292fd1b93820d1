// ORACLE — TEST INFRASTRUCTURE ONLY (see gl.hpp header).
//
// The oracle's OWN reader for the reference's JSON wire format, so that a GPU-vs-oracle comparison does not share the
// product's parser (plonky2-verifier_b200/csrc/host/parse.cpp): a small DOM parser (numbers kept as digit strings and
// reduced mod p like `mkGoldilocks`, Algebra/Goldilocks.hs:132) and the FromJSON instances of src/Types.hs:
//   CommonCircuitData :70 (fieldLabelModifier = drop 8), CircuitConfig :87 (drop 7), SelectorsInfo :104-108,
//   FriConfig :125 (drop 4), FriReductionStrategy :133-143, FriParams :172 (drop 4), LookupTable :36,
//   VerifierOnlyCircuitData :236-240, ProofWithPublicInputs / Proof / OpeningSet / FriProof / FriQueryRound /
//   FriInitialTreeProof / FriQueryStep / MerkleProof / MerkleCap :176-279, Digest (Hash/Digest.hs: {"elements": [..]}),
//   FExt as a 2-element list (Algebra/GoldilocksExt.hs), and the Rust debug strings of the gates (Gate/Parser.hs:107-240).
// Written from the Haskell, not from the product's parser; tests compare the two on every fixture.
#pragma once
#include <cctype>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "types.hpp"

namespace orc {
namespace js {

struct Value {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  std::string s;  // Num: the token, Str: the unescaped text
  std::vector<Value> a;
  std::vector<std::pair<std::string, Value>> o;
  const Value &at(const std::string &key) const {
    for (auto &kv : o)
      if (kv.first == key) return kv.second;
    throw std::runtime_error("json: key not found: " + key);
  }
  const Value *find(const std::string &key) const {
    for (auto &kv : o)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct Parser {
  const char *p, *end;
  Parser(const char *s, size_t n) : p(s), end(s + n) {}
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("json: ") + what); }
  Value value() {
    ws();
    if (p >= end) fail("unexpected end");
    Value v;
    char c = *p;
    if (c == '{') {
      v.kind = Value::Obj;
      p++;
      ws();
      if (p < end && *p == '}') { p++; return v; }
      for (;;) {
        ws();
        Value k = value();
        if (k.kind != Value::Str) fail("object key is not a string");
        ws();
        if (p >= end || *p != ':') fail("expected ':'");
        p++;
        v.o.emplace_back(k.s, value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; return v; }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      v.kind = Value::Arr;
      p++;
      ws();
      if (p < end && *p == ']') { p++; return v; }
      for (;;) {
        v.a.push_back(value());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; return v; }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') {
      v.kind = Value::Str;
      p++;
      while (p < end && *p != '"') {
        if (*p == '\\') {
          p++;
          if (p >= end) fail("bad escape");
          switch (*p) {
            case 'n': v.s += '\n'; break;
            case 't': v.s += '\t'; break;
            case 'r': v.s += '\r'; break;
            case 'b': v.s += '\b'; break;
            case 'f': v.s += '\f'; break;
            case 'u': {
              if (end - p < 5) fail("bad \\u escape");
              unsigned cp = (unsigned)std::stoul(std::string(p + 1, p + 5), nullptr, 16);
              if (cp < 0x80) v.s += (char)cp;
              else if (cp < 0x800) { v.s += (char)(0xC0 | (cp >> 6)); v.s += (char)(0x80 | (cp & 0x3F)); }
              else { v.s += (char)(0xE0 | (cp >> 12)); v.s += (char)(0x80 | ((cp >> 6) & 0x3F)); v.s += (char)(0x80 | (cp & 0x3F)); }
              p += 4;
              break;
            }
            default: v.s += *p;
          }
          p++;
        } else v.s += *p++;
      }
      if (p >= end) fail("unterminated string");
      p++;
      return v;
    }
    if (c == 't' && end - p >= 4 && std::string(p, p + 4) == "true") { v.kind = Value::Bool; v.b = true; p += 4; return v; }
    if (c == 'f' && end - p >= 5 && std::string(p, p + 5) == "false") { v.kind = Value::Bool; v.b = false; p += 5; return v; }
    if (c == 'n' && end - p >= 4 && std::string(p, p + 4) == "null") { p += 4; return v; }
    if (c == '-' || std::isdigit((unsigned char)c)) {
      v.kind = Value::Num;
      const char *q = p;
      if (*p == '-') p++;
      while (p < end && (std::isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) p++;
      v.s.assign(q, p);
      return v;
    }
    fail("unexpected character");
  }
};

inline Value parse(const char *s, size_t n) {
  Parser ps(s, n);
  Value v = ps.value();
  ps.ws();
  if (ps.p != ps.end) throw std::runtime_error("json: trailing characters");
  return v;
}

// decimal digit string -> value mod p (mkGoldilocks on an Integer, Algebra/Goldilocks.hs:132): Horner in the field
inline F feltOfDigits(const std::string &t) {
  if (t.empty()) throw std::runtime_error("json: empty number");
  bool neg = t[0] == '-';
  F acc(0), ten(10);
  for (size_t i = neg ? 1 : 0; i < t.size(); i++) {
    if (!std::isdigit((unsigned char)t[i])) throw std::runtime_error("json: not an integer: " + t);
    acc = acc * ten + F((uint64_t)(t[i] - '0'));
  }
  return neg ? (F(0) - acc) : acc;
}
inline F felt(const Value &v) {
  if (v.kind != Value::Num) throw std::runtime_error("json: expected a number");
  return feltOfDigits(v.s);
}
inline long long integer(const Value &v) {
  if (v.kind != Value::Num) throw std::runtime_error("json: expected an integer");
  return std::stoll(v.s);
}
inline const std::vector<Value> &arr(const Value &v) {
  if (v.kind != Value::Arr) throw std::runtime_error("json: expected an array");
  return v.a;
}
inline FExt ext(const Value &v) {  // [real, imag]
  auto &a = arr(v);
  if (a.size() != 2) throw std::runtime_error("json: extension element is not a pair");
  return FExt(felt(a[0]), felt(a[1]));
}
inline std::vector<FExt> exts(const Value &v) { std::vector<FExt> r; for (auto &x : arr(v)) r.push_back(ext(x)); return r; }
inline std::vector<F> felts(const Value &v) { std::vector<F> r; for (auto &x : arr(v)) r.push_back(felt(x)); return r; }
inline Digest digest(const Value &v) {  // Hash/Digest.hs: {"elements":[a,b,c,d]}
  auto &a = arr(v.at("elements"));
  if (a.size() != 4) throw std::runtime_error("json: digest does not have 4 elements");
  Digest d;
  for (int i = 0; i < 4; i++) d.e[i] = felt(a[i]);
  return d;
}
inline MerkleCap cap(const Value &v) { MerkleCap c; for (auto &x : arr(v)) c.roots.push_back(digest(x)); return c; }
inline MerkleProof merkleProof(const Value &v) { MerkleProof m; for (auto &x : arr(v.at("siblings"))) m.siblings.push_back(digest(x)); return m; }

// ---- gate strings (Gate/Parser.hs:107-240) ----------------------------------------------------------------------
struct GateScan {
  const std::string &s;
  size_t i = 0;
  explicit GateScan(const std::string &t) : s(t) {}
  bool lit(const char *t) {
    size_t n = strlen(t);
    if (s.compare(i, n, t) == 0) { i += n; return true; }
    return false;
  }
  void spaces() { while (i < s.size() && std::isspace((unsigned char)s[i])) i++; }
  bool digits(std::string &out) {
    size_t j = i;
    while (j < s.size() && std::isdigit((unsigned char)s[j])) j++;
    if (j == i) return false;
    out = s.substr(i, j - i);
    i = j;
    return true;
  }
  bool keyInt(const char *key, long long &out) {  // keyValueP key intP
    std::string d;
    if (!lit(key)) return false;
    spaces();
    if (!lit(":")) return false;
    spaces();
    if (!digits(d)) return false;
    out = std::stoll(d);
    spaces();
    return true;
  }
  bool comma() { if (!lit(",")) return false; spaces(); return true; }
  bool eof() const { return i == s.size(); }
};

// rustStructP name body = string name; spaces; '{'; spaces; body; spaces; '}' (Gate/Parser.hs:78-86)
inline Gate recognizeGate(const std::string &str) {
  const char *PH = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";
  auto open = [](GateScan &g, const char *name) {
    if (!g.lit(name)) return false;
    g.spaces();
    if (!g.lit("{")) return false;
    g.spaces();
    return true;
  };
  auto close = [](GateScan &g) { g.spaces(); if (!g.lit("}")) return false; g.spaces(); return true; };
  auto oneInt = [&](const char *name, const char *key, int kind, bool need_eof, Gate &out) {
    GateScan g(str);
    long long x;
    if (!open(g, name) || !g.keyInt(key, x) || !close(g)) return false;
    if (need_eof && !g.eof()) return false;
    out.kind = kind; out.p0 = (int)x;
    return true;
  };
  auto listOf = [](GateScan &g, std::vector<std::string> &out) {  // listP
    if (!g.lit("[")) return false;
    g.spaces();
    std::string d;
    if (g.digits(d)) {
      out.push_back(d);
      for (;;) {
        size_t save = g.i;
        if (!g.comma()) { g.i = save; break; }
        if (!g.digits(d)) return false;
        out.push_back(d);
      }
    }
    if (!g.lit("]")) return false;
    g.spaces();
    return true;
  };
  Gate out;
  // the order of the alternatives is gateP's (first match wins)
  if (oneInt("ArithmeticGate", "num_ops", P2V_GATE_ARITHMETIC, true, out)) return out;
  if (oneInt("ArithmeticExtensionGate", "num_ops", P2V_GATE_ARITHMETIC_EXT, true, out)) return out;
  {  // "BaseSumGate { num_limbs: 63 } + Base: 2"
    GateScan g(str);
    long long limbs, base;
    if (open(g, "BaseSumGate") && g.keyInt("num_limbs", limbs) && close(g)) {
      g.spaces();
      if (g.lit("+")) {
        g.spaces();
        if (g.keyInt("Base", base) && g.eof()) { out.kind = P2V_GATE_BASE_SUM; out.p0 = (int)limbs; out.p1 = (int)base; return out; }
      }
    }
  }
  {  // CosetInterpolationGate { subgroup_bits, degree, barycentric_weights: [..], _phantom: .. }<D=2>
    GateScan g(str);
    long long sb, deg;
    std::vector<std::string> ws;
    if (open(g, "CosetInterpolationGate") && g.keyInt("subgroup_bits", sb) && g.comma() && g.keyInt("degree", deg) && g.comma() &&
        g.lit("barycentric_weights") && (g.spaces(), g.lit(":")) && (g.spaces(), listOf(g, ws)) && g.comma() && g.lit(PH) && close(g) && g.lit("<D=2>") &&
        g.eof()) {
      out.kind = P2V_GATE_COSET_INTERP; out.p0 = (int)sb; out.p1 = (int)deg;
      for (auto &w : ws) out.weights.push_back(feltOfDigits(w));
      return out;
    }
  }
  if (oneInt("ConstantGate", "num_consts", P2V_GATE_CONSTANT, false, out)) return out;
  if (oneInt("ExponentiationGate", "num_power_bits", P2V_GATE_EXPONENTIATION, false, out)) return out;
  {  // LookupGate { num_slots, lut_hash: [bytes] }
    GateScan g(str);
    long long slots;
    std::vector<std::string> h;
    if (open(g, "LookupGate") && g.keyInt("num_slots", slots) && g.comma() && g.lit("lut_hash") && (g.spaces(), g.lit(":")) && (g.spaces(), listOf(g, h)) && close(g)) {
      out.kind = P2V_GATE_LOOKUP; out.p0 = (int)slots;
      return out;
    }
  }
  {
    GateScan g(str);
    long long slots, last;
    std::vector<std::string> h;
    if (open(g, "LookupTableGate") && g.keyInt("num_slots", slots) && g.comma() && g.lit("lut_hash") && (g.spaces(), g.lit(":")) && (g.spaces(), listOf(g, h)) && g.comma() &&
        g.keyInt("last_lut_row", last) && close(g)) {
      out.kind = P2V_GATE_LOOKUP_TABLE; out.p0 = (int)slots; out.p1 = (int)last;
      return out;
    }
  }
  if (oneInt("MulExtensionGate", "num_ops", P2V_GATE_MUL_EXT, false, out)) return out;
  if (str.compare(0, 8, "NoopGate") == 0) { out.kind = P2V_GATE_NOOP; return out; }
  if (str.compare(0, 15, "PublicInputGate") == 0) { out.kind = P2V_GATE_PUBLIC_INPUT; return out; }
  for (int mds = 0; mds < 2; mds++) {  // "PoseidonGate(PhantomData<...>)<WIDTH=12>"
    GateScan g(str);
    std::string d;
    std::string head = std::string(mds ? "PoseidonMdsGate" : "PoseidonGate") + "(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=";
    if (g.lit(head.c_str()) && g.digits(d) && g.lit(">") && g.eof()) {
      out.kind = mds ? P2V_GATE_POSEIDON_MDS : P2V_GATE_POSEIDON; out.p0 = std::stoi(d);
      return out;
    }
  }
  {  // RandomAccessGate { bits, num_copies, num_extra_constants, _phantom: .. }<D=2>
    GateScan g(str);
    long long b, c, e;
    if (open(g, "RandomAccessGate") && g.keyInt("bits", b) && g.comma() && g.keyInt("num_copies", c) && g.comma() && g.keyInt("num_extra_constants", e) &&
        g.comma() && g.lit(PH) && close(g) && g.lit("<D=2>")) {
      out.kind = P2V_GATE_RANDOM_ACCESS; out.p0 = (int)b; out.p1 = (int)c; out.p2 = (int)e;
      return out;
    }
  }
  if (oneInt("ReducingGate", "num_coeffs", P2V_GATE_REDUCING, false, out)) return out;
  if (oneInt("ReducingExtensionGate", "num_coeffs", P2V_GATE_REDUCING_EXT, false, out)) return out;
  out.kind = P2V_GATE_UNKNOWN;
  return out;
}

// ---- Types.hs records ---------------------------------------------------------------------------------------------
inline void friConfig(const Value &v, int degree_bits, FriConfig &fc) {
  fc.rate_bits = (int)integer(v.at("rate_bits"));
  fc.cap_height = (int)integer(v.at("cap_height"));
  fc.proof_of_work_bits = (int)integer(v.at("proof_of_work_bits"));
  fc.num_query_rounds = (int)integer(v.at("num_query_rounds"));
  const Value &st = v.at("reduction_strategy");
  if (st.kind != Value::Obj || st.o.size() != 1) throw std::runtime_error("FromJSON/FriReductionStrategy: expecting a singleton object");
  const std::string &key = st.o[0].first;
  const Value &val = st.o[0].second;
  fc.step_arity_bits.clear();
  // expandReductionStrategy, Plonk/FRI.hs:337-354
  if (key == "ConstantArityBits") {
    auto &ab = arr(val);
    if (ab.size() != 2) throw std::runtime_error("ConstantArityBits: expecting [arity_bits, final_poly_bits]");
    int arity = (int)integer(ab[0]), final_bits = (int)integer(ab[1]);
    if (arity <= 0) throw std::runtime_error("ConstantArityBits: arity_bits must be positive");
    for (int logn = degree_bits; logn > final_bits; logn -= arity) fc.step_arity_bits.push_back(arity);
  } else if (key == "Fixed") {
    for (auto &x : arr(val)) fc.step_arity_bits.push_back((int)integer(x));
  } else {
    throw std::runtime_error("reduction strategy not implemented");
  }
}

inline CommonCircuitData commonFromJson(const char *text, size_t len) {
  Value v = parse(text, len);
  CommonCircuitData c;
  const Value &cfg = v.at("config");
  c.num_wires = (int)integer(cfg.at("num_wires"));
  c.num_routed_wires = (int)integer(cfg.at("num_routed_wires"));
  c.config_num_constants = (int)integer(cfg.at("num_constants"));
  c.num_challenges = (int)integer(cfg.at("num_challenges"));
  const Value &fp = v.at("fri_params");
  c.degree_bits = (int)integer(fp.at("degree_bits"));
  // the verifier reads config_fri_config circuit_config (Plonk/FRI.hs:367, Challenge/FRI.hs:67-68)
  friConfig(cfg.at("fri_config"), c.degree_bits, c.fri_config);
  for (auto &g : arr(v.at("gates"))) {
    if (g.kind != Value::Str) throw std::runtime_error("json: gate is not a string");
    c.gates.push_back(recognizeGate(g.s));
  }
  const Value &sel = v.at("selectors_info");
  for (auto &x : arr(sel.at("selector_indices"))) c.selector_indices.push_back((int)integer(x));
  for (auto &x : arr(sel.at("groups"))) c.selector_groups.push_back(Range{(int)integer(x.at("start")), (int)integer(x.at("end"))});
  c.quotient_degree_factor = (int)integer(v.at("quotient_degree_factor"));
  c.num_constants = (int)integer(v.at("num_constants"));
  c.num_public_inputs = (int)integer(v.at("num_public_inputs"));
  c.k_is = felts(v.at("k_is"));
  c.num_partial_products = (int)integer(v.at("num_partial_products"));
  c.num_lookup_polys = (int)integer(v.at("num_lookup_polys"));
  c.num_lookup_selectors = (int)integer(v.at("num_lookup_selectors"));
  for (auto &t : arr(v.at("luts"))) {
    std::vector<std::pair<F, F>> lut;
    for (auto &pr : arr(t)) {
      auto &ab = arr(pr);
      if (ab.size() != 2) throw std::runtime_error("json: lookup table entry is not a pair");
      lut.push_back({felt(ab[0]), felt(ab[1])});
    }
    c.luts.push_back(lut);
  }
  return c;
}

inline VerifierOnlyCircuitData vkeyFromJson(const char *text, size_t len) {
  Value v = parse(text, len);
  VerifierOnlyCircuitData vk;
  vk.constants_sigmas_cap = cap(v.at("constants_sigmas_cap"));
  vk.circuit_digest = digest(v.at("circuit_digest"));
  return vk;
}

inline ProofWithPublicInputs proofFromJson(const char *text, size_t len) {
  Value v = parse(text, len);
  ProofWithPublicInputs pw;
  const Value &pv = v.at("proof");
  Proof &p = pw.proof;
  p.wires_cap = cap(pv.at("wires_cap"));
  p.plonk_zs_partial_products_cap = cap(pv.at("plonk_zs_partial_products_cap"));
  p.quotient_polys_cap = cap(pv.at("quotient_polys_cap"));
  const Value &ov = pv.at("openings");
  OpeningSet &o = p.openings;
  o.constants = exts(ov.at("constants"));
  o.plonk_sigmas = exts(ov.at("plonk_sigmas"));
  o.wires = exts(ov.at("wires"));
  o.plonk_zs = exts(ov.at("plonk_zs"));
  o.plonk_zs_next = exts(ov.at("plonk_zs_next"));
  o.partial_products = exts(ov.at("partial_products"));
  o.quotient_polys = exts(ov.at("quotient_polys"));
  o.lookup_zs = exts(ov.at("lookup_zs"));
  o.lookup_zs_next = exts(ov.at("lookup_zs_next"));
  const Value &fv = pv.at("opening_proof");
  FriProof &fp = p.opening_proof;
  for (auto &c : arr(fv.at("commit_phase_merkle_caps"))) fp.commit_phase_merkle_caps.push_back(cap(c));
  for (auto &q : arr(fv.at("query_round_proofs"))) {
    FriQueryRound qr;
    for (auto &ep : arr(q.at("initial_trees_proof").at("evals_proofs"))) {
      auto &pr = arr(ep);  // (leaf values, MerkleProof) as a 2-element list
      if (pr.size() != 2) throw std::runtime_error("json: evals_proofs entry is not a pair");
      qr.initial_trees_proof.evals_proofs.push_back({felts(pr[0]), merkleProof(pr[1])});
    }
    for (auto &s : arr(q.at("steps"))) {
      FriQueryStep st;
      st.evals = exts(s.at("evals"));
      st.merkle_proof = merkleProof(s.at("merkle_proof"));
      qr.steps.push_back(st);
    }
    fp.query_round_proofs.push_back(qr);
  }
  fp.final_poly = exts(fv.at("final_poly").at("coeffs"));
  fp.pow_witness = felt(fv.at("pow_witness"));
  pw.public_inputs = felts(v.at("public_inputs"));
  return pw;
}

}  // namespace js
}  // namespace orc
