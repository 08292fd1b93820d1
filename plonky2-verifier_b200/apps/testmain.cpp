// p2v_testmain — the reference's driver (src/testmain.hs:24-63) on the GPU, above the C ABI of include/p2v.h.
//
//   p2v_testmain <dir> <prefix>                       reads <dir>/<prefix>_{common,vkey,proof}.json like testmain.hs:31-33
//   p2v_testmain <common.json> <vkey.json> <proof.json>
//
// Prints what the Haskell prints, line for line and with its Show instances (`MkDigest a b c d`, Hash/Digest.hs:36-38;
// `(re + X*im)`, Algebra/GoldilocksExt.hs:37-38; Haskell list / Bool syntax):
//   public inputs hash (`sponge (public_inputs proof_data)`, testmain.hs:40-41), the nine opening counts (:43-52),
//   `evalCombinedPlonkConstraints` (:58), `checkCombinedPlonkEquations'` (:59), `verifyProof` (:63).
// A proof that runs into one of the reference's `error` sites ends with "testmain: <message>" — what GHC prints on stderr
// before it aborts (Plonk/FRI.hs:108, :310, :311).
// Every number comes from libp2v's kernels (p2v_hash_leaves, p2v_verify_intermediates); there is no CPU arithmetic here and
// no CPU fallback: without a B200 the program stops with the library's P2V_E_NOGPU message and exit status 3.
// tests/test_gpu_testmain.py diffs its output against tests/golden/<name>.testmain.txt for every bundled fixture.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "p2v.h"

static bool slurp(const std::string &path, std::string &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char buf[1 << 16];
  size_t k;
  out.clear();
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
  fclose(f);
  return true;
}

static int die(int rc, const char *what, const p2v_ctx *ctx = nullptr) {
  fprintf(stderr, "p2v_testmain: %s: %s (code %d)\n", what, p2v_last_error(ctx), rc);
  return rc == P2V_E_NOGPU ? 3 : 2;
}

int main(int argc, char **argv) {
  std::string pc, pv, pp;
  if (argc == 3) {
    std::string base = std::string(argv[1]) + "/" + argv[2];
    pc = base + "_common.json"; pv = base + "_vkey.json"; pp = base + "_proof.json";
  } else if (argc == 4) {
    pc = argv[1]; pv = argv[2]; pp = argv[3];
  } else {
    fprintf(stderr, "usage: %s <dir> <prefix>  |  %s <common.json> <vkey.json> <proof.json>\n", argv[0], argv[0]);
    return 64;
  }
  std::string tc, tv, tp;
  if (!slurp(pc, tc) || !slurp(pv, tv) || !slurp(pp, tp)) {
    fprintf(stderr, "p2v_testmain: cannot read %s / %s / %s\n", pc.c_str(), pv.c_str(), pp.c_str());
    return 66;
  }

  // `decode text_common`, `decode text_vkey`, `decode text_proof` (testmain.hs:35-37)
  p2v_shape shape;
  p2v_layout lay;
  int rc = p2v_parse_common(tc.data(), tc.size(), &shape);
  if (rc) return die(rc, "common");
  if ((rc = p2v_shape_layout(&shape, &lay))) return die(rc, "layout");
  std::vector<uint64_t> vkey(lay.vkey_words), blob(lay.blob_words);
  if ((rc = p2v_parse_vkey(tv.data(), tv.size(), &shape, vkey.data()))) return die(rc, "vkey");
  if ((rc = p2v_parse_proof(tp.data(), tp.size(), &shape, blob.data()))) return die(rc, "proof");

  p2v_ctx *ctx = nullptr;
  if ((rc = p2v_ctx_create(0, &ctx))) return die(rc, "p2v_ctx_create");
  p2v_circuit *cir = nullptr;
  if ((rc = p2v_circuit_create(ctx, &shape, vkey.data(), &cir))) return die(rc, "p2v_circuit_create", ctx);

  // sponge (public_inputs proof_data): one "leaf" of num_public_inputs words, SoA [w][1]
  uint64_t pih[4] = {0, 0, 0, 0};
  if ((rc = p2v_hash_leaves(ctx, blob.data() + lay.off_public_inputs, (uint32_t)shape.num_public_inputs, 1, pih)))
    return die(rc, "p2v_hash_leaves", ctx);
  printf("public inputs hash = MkDigest %" PRIu64 " %" PRIu64 " %" PRIu64 " %" PRIu64 "\n", pih[0], pih[1], pih[2], pih[3]);

  const struct { const char *name; int n; } counts[9] = {
      {"# opening_constants", lay.n_open_constants},       {"# opening_plonk_sigmas", lay.n_open_sigmas},
      {"# opening_wires", lay.n_open_wires},               {"# opening_plonk_zs", lay.n_open_zs},
      {"# opening_plonk_zs_next", lay.n_open_zs_next},     {"# opening_partial_products", lay.n_open_pp},
      {"# opening_quotient_polys", lay.n_open_quotient},   {"# opening_lookup_zs", lay.n_open_lookup_zs},
      {"# opening_lookup_zs_next", lay.n_open_lookup_zs_next}};
  for (const auto &c : counts) printf("%-26s = %d\n", c.name, c.n);

  const int r = shape.num_challenges;
  std::vector<uint64_t> combined(2 * (size_t)r);
  uint8_t eqmask = 0;
  uint32_t status = 0;
  p2v_intermediates io;
  memset(&io, 0, sizeof io);
  io.combined = combined.data();
  io.eq_ok_mask = &eqmask;
  io.status = &status;
  if ((rc = p2v_verify_intermediates(ctx, cir, blob.data(), 1, &io))) return die(rc, "p2v_verify_intermediates", ctx);

  // print $ evalCombinedPlonkConstraints ...   (a list of FExt; SoA [2r][1]: plane 2i = real part, 2i+1 = X part of round i)
  printf("[");
  for (int i = 0; i < r; i++) printf("%s(%" PRIu64 " + X*%" PRIu64 ")", i ? "," : "", combined[2 * i], combined[2 * i + 1]);
  printf("]\n");
  // print $ checkCombinedPlonkEquations' ...   (a list of Bool)
  printf("[");
  for (int i = 0; i < r; i++) printf("%s%s", i ? "," : "", ((eqmask >> i) & 1) ? "True" : "False");
  printf("]\n");
  // putStrLn $ "proof verification result = " ++ show (verifyProof vkey proof_data)
  const char *verdict;
  char other[64];
  switch (status & 0xFF) {
    case P2V_ST_ACCEPT: verdict = "True"; break;
    case P2V_ST_FALSE_EQS:
    case P2V_ST_FALSE_POW:
    case P2V_ST_FALSE_FINAL: verdict = "False"; break;
    case P2V_ST_ERR_INIT_MERKLE: verdict = "testmain: checkInitialTreeProofs: at least one Merkle proof failed"; break;
    case P2V_ST_ERR_STEP_MERKLE: verdict = "testmain: folding step Merkle proof does not check out"; break;
    case P2V_ST_ERR_STEP_EVAL: verdict = "testmain: folding step evaluation does not match the opening"; break;
    default: snprintf(other, sizeof other, "testmain: error site %u", status & 0xFF); verdict = other;
  }
  printf("proof verification result = %s\n", verdict);

  p2v_circuit_destroy(cir);
  p2v_ctx_destroy(ctx);
  p2v_shape_free(&shape);
  return 0;
}
