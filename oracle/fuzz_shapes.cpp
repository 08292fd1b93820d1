// TEST INFRASTRUCTURE.  Hostile circuit descriptions: mutated `*_common.json` texts that the product's host parser accepts AND
// p2v_shape_check passes are handed to the CPU oracle (compiled with ASAN/UBSAN and _GLIBCXX_ASSERTIONS) together with a proof blob
// of the shape's size.  The oracle indexes its arrays where the reference indexes its lists, so a sanitizer report here means the
// shape check lets through a circuit on which the reference raises an index error — and on which a kernel would read out of bounds.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined -D_GLIBCXX_ASSERTIONS -Iinclude -Ioracle \
//       oracle/fuzz_shapes.cpp plonky2-verifier_b200/csrc/host/parse.cpp -o /tmp/fuzz_shapes -lpthread && /tmp/fuzz_shapes tests/golden 3000
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include "oracle_capi.cpp"
thread_local std::string p2v_tls_error;
static std::string slurp(const std::string &p) { std::ifstream f(p); std::stringstream s; s << f.rdbuf(); return s.str(); }
static std::string g_current;
static std::atomic<long> g_tick{0};
static std::string mutate(const std::string &t, std::mt19937_64 &r) {
  std::string b = t;
  static const char *vals[] = {"0", "1", "2", "3", "4", "5", "7", "8", "9", "12", "15", "16", "17", "20", "28", "31", "32", "33", "63", "64", "65", "80", "135", "255", "256", "1023", "4096", "65535", "65536"};
  int k = 1 + (int)(r() % 3);
  for (int i = 0; i < k; i++) {
    // number tokens, chosen uniformly; field elements (k_is, barycentric weights: 13+ digits) are data, not structure — rarely touched
    std::vector<std::pair<size_t, size_t>> toks;
    for (size_t p = 0; p < b.size();) {
      if (isdigit((unsigned char)b[p])) {
        size_t e = p;
        while (e < b.size() && isdigit((unsigned char)b[e])) e++;
        if (e - p <= 12 || r() % 64 == 0) toks.push_back({p, e - p});
        p = e;
      } else p++;
    }
    if (toks.empty()) break;
    auto tk = toks[r() % toks.size()];
    b.replace(tk.first, tk.second, vals[r() % 29]);
  }
  if (r() % 8 == 0) {  // drop or duplicate a gate string / list element
    size_t q = b.find("\"gates\"");
    if (q != std::string::npos) {
      size_t a = b.find('"', q + 8), z = a == std::string::npos ? a : b.find('"', a + 1);
      if (z != std::string::npos && z + 2 < b.size()) { if (r() % 2) b.erase(a, z - a + 2); else b.insert(a, b.substr(a, z - a + 2)); }
    }
  }
  return b;
}
int main(int argc, char **argv) {
  std::string dir = argv[1];
  int rounds = argc > 2 ? atoi(argv[2]) : 1000;
  std::thread([] {  // watchdog: a case that runs for more than 30 s is a finding too (a loop bound the check should have refused)
    long last = -1; int same = 0;
    for (;;) {
      std::this_thread::sleep_for(std::chrono::seconds(5));
      long t = g_tick.load();
      same = t == last ? same + 1 : 0; last = t;
      if (same >= 6) { fprintf(stderr, "TIMEOUT on:\n%s\n", g_current.c_str()); _Exit(3); }
    }
  }).detach();
  const char *names[] = {"small6", "fixed4", "lookup6", "real5", "reallu6", "arity5"};
  std::mt19937_64 r(99);
  long ee = 0, accepted = 0, refused = 0, unparsed = 0, ran = 0, proof_fits = 0;
  for (int i = 0; i < rounds; i++) {
    const char *nm = names[r() % 6];
    std::string common = slurp(dir + "/" + nm + "_common.json"), proof = slurp(dir + "/" + nm + "_proof.json"), vkey = slurp(dir + "/" + nm + "_vkey.json");
    std::string m = mutate(common, r);
    g_current = m;
    g_tick++;
    p2v_shape sh;
    memset(&sh, 0, sizeof sh);
    if (p2v_parse_common(m.data(), m.size(), &sh)) { unparsed++; continue; }
    if (p2v_shape_check(&sh)) { refused++; p2v_shape_free(&sh); continue; }
    accepted++;
    p2v_layout lay;
    if (p2v_shape_layout(&sh, &lay) || lay.blob_words > (1 << 19)) { p2v_shape_free(&sh); continue; }
    std::vector<uint64_t> blob(lay.blob_words), vk(lay.vkey_words);
    if (p2v_parse_proof(proof.data(), proof.size(), &sh, blob.data()) == 0) proof_fits++;
    else for (auto &w : blob) w = r() % 0xFFFFFFFF00000001ULL;  // a blob of the right size with arbitrary field elements
    if (p2v_parse_vkey(vkey.data(), vkey.size(), &sh, vk.data())) for (auto &w : vk) w = r() % 0xFFFFFFFF00000001ULL;
    int Q = sh.num_queries, rr = sh.num_challenges;
    std::vector<uint64_t> ch(4096), comb(2 * rr), folded(2 * Q);
    std::vector<uint32_t> qs(Q);
    uint8_t eq; uint32_t st, fst; unsigned long long pc;
    ch.resize((size_t)p2v_challenges_words(&sh) + 8);
    orc_verify_batch(&sh, vk.data(), blob.data(), 1, 1, 1, ch.data(), comb.data(), &eq, &st, &fst, qs.data(), folded.data(), &pc);
    ran++;
    if (st == 0xEE) {
      ee++;
      if (ee <= 5) {  // show what was changed
        size_t d = 0; while (d < m.size() && d < common.size() && m[d] == common[d]) d++;
        fprintf(stderr, "oracle raised (0xEE) on %s, first change at byte %zu: ...%s...\n", nm, d, m.substr(d > 60 ? d - 60 : 0, 140).c_str());
      }
    }
    p2v_shape_free(&sh);
  }
  printf("mutated commons: %ld unparsed, %ld refused by p2v_shape_check, %ld accepted (%ld verified by the oracle, honest proof still fits %ld), oracle exceptions %ld\n", unparsed, refused, accepted, ran, proof_fits, ee);
  return 0;
}
