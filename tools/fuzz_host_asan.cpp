// ASAN/UBSAN fuzz of the host parsers (csrc/host/parse.cpp, json.hpp): mutated common / vkey / proof / gate texts.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -Iinclude tools/fuzz_host_asan.cpp \
//       plonky2-verifier_b200/csrc/host/parse.cpp -o /tmp/fuzz_host -lpthread && /tmp/fuzz_host tests/golden 20000
// Mutations: byte edits, deletions, insertions, truncation, raw bytes, and (3 of 8) whole number tokens replaced by boundary
// values (0, 2^k +- 1, 2^31, 2^32, p, 2^64 - 1, 2^128, negative, fractional, leading zeros).  A mutated `common` that still parses is
// laid out and the honest proof is decoded against it.  Any sanitizer report aborts.  tests/test_host.py runs a short session.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include "p2v.h"
static std::string slurp(const std::string &p) { std::ifstream f(p); std::stringstream s; s << f.rdbuf(); return s.str(); }
static std::string mutate(const std::string &t, std::mt19937_64 &r) {
  std::string b = t;
  static const char al[] = "0123456789[]{},:\" -e.+Eabtrufalsn\\/\n\t";
  int k = (int[]){1, 1, 2, 5, 12}[r() % 5];
  for (int i = 0; i < k && b.size() > 1; i++) {
    size_t pos = r() % b.size();
    switch (r() % 8 < 3 ? 4 : r() % 6) {
      case 0: b[pos] = al[r() % (sizeof(al) - 1)]; break;
      case 1: b.erase(pos, 1 + r() % 8); break;
      case 2: { std::string ins; for (int j = 0, n = 1 + r() % 6; j < n; j++) ins += al[r() % (sizeof(al) - 1)]; b.insert(pos, ins); break; }
      case 3: b.resize(pos); break;
      case 4: { // swap in a huge / odd number for a whole number token
        for (int tries = 0; tries < 64 && !isdigit((unsigned char)b[pos]); tries++) pos = r() % b.size();
        while (pos > 0 && isdigit((unsigned char)b[pos - 1])) pos--;
        static const char *nums[] = {"18446744073709551615", "18446744069414584321", "340282366920938463463374607431768211456", "-1", "1e5", "0.5", "00", "4294967296", "2147483648", "99999999999999999999999999999999999999999999"};
        static const char *small[] = {"0", "1", "2", "3", "7", "8", "9", "15", "16", "17", "31", "32", "33", "63", "64", "65", "255", "256", "1023", "4096", "65535", "65536", "1048576", "2147483647"};
        if (r() % 2) { size_t e2 = pos; while (e2 < b.size() && isdigit((unsigned char)b[e2])) e2++; b.replace(pos, e2 - pos, small[r() % 24]); break; }
        size_t e = pos; while (e < b.size() && isdigit((unsigned char)b[e])) e++;
        b.replace(pos, e - pos, nums[r() % 10]); break; }
      case 5: b[pos] = (char)(r() & 0xFF); break;
    }
  }
  return b;
}
thread_local std::string p2v_tls_error;
int main(int argc, char **argv) {
  std::string dir = argv[1];
  int rounds = argc > 2 ? atoi(argv[2]) : 2000;
  const char *names[] = {"small6", "fixed4", "lookup6", "real5", "reallu6", "arity5", "mid5"};
  std::mt19937_64 r(12345);
  long ok[4] = {}, bad[4] = {};
  for (const char *nm : names) {
    std::string common = slurp(dir + "/" + nm + "_common.json"), vkey = slurp(dir + "/" + nm + "_vkey.json"), proof = slurp(dir + "/" + nm + "_proof.json");
    p2v_shape sh;
    if (p2v_parse_common(common.data(), common.size(), &sh)) { printf("%s: common does not parse\n", nm); return 1; }
    p2v_layout lay;
    p2v_shape_layout(&sh, &lay);
    std::vector<uint64_t> blob(lay.blob_words), vk(lay.vkey_words);
    for (int i = 0; i < rounds; i++) {
      std::string m = mutate(proof, r);
      (p2v_parse_proof(m.data(), m.size(), &sh, blob.data()) ? bad : ok)[0]++;
      if (i % 4 == 0) {
        m = mutate(vkey, r);
        (p2v_parse_vkey(m.data(), m.size(), &sh, vk.data()) ? bad : ok)[1]++;
        m = mutate(common, r);
        p2v_shape s2;
        memset(&s2, 0, sizeof s2);
        if (p2v_parse_common(m.data(), m.size(), &s2) == 0) {
          ok[2]++;
          p2v_layout l2;
          if (p2v_shape_layout(&s2, &l2) == 0 && l2.blob_words < (1 << 22)) {  // a shape that parses must be usable: decode the proof against it
            std::vector<uint64_t> b2(l2.blob_words);
            p2v_parse_proof(proof.data(), proof.size(), &s2, b2.data());
          }
          p2v_shape_free(&s2);
        } else bad[2]++;
      }
    }
    p2v_shape_free(&sh);
  }
  const char *gates[] = {"NoopGate", "ConstantGate { num_consts: 2 }", "PublicInputGate", "BaseSumGate { num_limbs: 63 } + Base: 2",
                         "ArithmeticGate { num_ops: 20 }", "RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
                         "CosetInterpolationGate { subgroup_bits: 4, degree: 6, barycentric_weights: [17293822565076172801, 256, 1048576, 4294967296], _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>",
                         "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", "LookupGate {num_slots: 40, lut_hash: [1,2,3]}", "ExponentiationGate { num_power_bits: 67, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>"};
  for (int i = 0; i < rounds * 20; i++) {
    std::string m = mutate(gates[r() % 10], r);
    p2v_gate g;
    std::vector<uint64_t> w(P2V_MAX_WEIGHTS);
    (p2v_parse_gate(m.data(), m.size(), &g, w.data()) ? bad : ok)[3]++;
  }
  printf("proof ok/bad %ld/%ld  vkey %ld/%ld  common %ld/%ld  gate %ld/%ld\n", ok[0], bad[0], ok[1], bad[1], ok[2], bad[2], ok[3], bad[3]);
  return 0;
}
