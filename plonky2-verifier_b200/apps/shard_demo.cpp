// p2v_shard_demo — BASELINE config 5 from a host that has neither Python nor torch: one PROCESS per GPU, the communicator
// bootstrapped with p2v_nccl_unique_id / p2v_nccl_init (the 128-byte id travels through a file), the SAME batch cut with
// p2v_shard_bounds, every rank verifying its slice with p2v_verify_batch_sharded (`verifyProof`, Plonk/Verifier.hs:56-65,
// mapped over the slice + one all-gather of the accept bitmap).  Every rank then verifies the WHOLE batch alone and
// requires the gathered bitmap to be identical, first with the NCCL gather, then with the peer-store gather
// (p2v_peer_enable).  This is the call sequence a `foreign import ccall` host or any C host uses.
//
//   p2v_shard_demo <common.json> <vkey.json> <proof.json> <world> [n_total]
//
// Uses include/p2v.h only.  Exit status 0 and a line "SHARD_DEMO_OK ..." from rank 0 when every rank agrees.
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "p2v.h"

static const uint64_t GL_P = 0xFFFFFFFF00000001ULL;

static bool slurp(const char *path, std::string &out) {
  FILE *f = fopen(path, "rb");
  if (!f) return false;
  char buf[1 << 16];
  size_t k;
  out.clear();
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
  fclose(f);
  return true;
}

#define CHECK(call, ctx)                                                                                   \
  do {                                                                                                     \
    int rc__ = (call);                                                                                     \
    if (rc__ != P2V_OK) {                                                                                  \
      fprintf(stderr, "[rank %d] %s: %s (code %d)\n", rank, #call, p2v_last_error(ctx), rc__);             \
      return 2;                                                                                            \
    }                                                                                                      \
  } while (0)

static int run_rank(int rank, int world, const p2v_shape &shape, const p2v_layout &lay, const std::vector<uint64_t> &vkey,
                    const std::vector<uint64_t> &batch, size_t n_total, const std::string &id_path) {
  p2v_ctx *ctx = nullptr;
  CHECK(p2v_ctx_create(rank, &ctx), nullptr);
  p2v_circuit *cir = nullptr;
  CHECK(p2v_circuit_create(ctx, &shape, vkey.data(), &cir), ctx);

  // bootstrap: rank 0 makes the id and publishes it (write + rename = atomic), the others wait for the file
  unsigned char id[P2V_NCCL_UNIQUE_ID_BYTES];
  if (rank == 0) {
    CHECK(p2v_nccl_unique_id(id), nullptr);
    std::string tmp = id_path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(id, 1, sizeof id, f) != sizeof id) return 2;
    fclose(f);
    if (rename(tmp.c_str(), id_path.c_str()) != 0) return 2;
  } else {
    FILE *f = nullptr;
    for (int tries = 0; tries < 3000 && !(f = fopen(id_path.c_str(), "rb")); tries++) usleep(10000);
    if (!f || fread(id, 1, sizeof id, f) != sizeof id) {
      fprintf(stderr, "[rank %d] no NCCL id at %s\n", rank, id_path.c_str());
      return 2;
    }
    fclose(f);
  }
  CHECK(p2v_nccl_init(ctx, id, rank, world), ctx);

  // this rank alone, the whole batch: the result every gathered bitmap must equal
  const size_t words = (n_total + 31) / 32;
  std::vector<uint32_t> bits1(words, 0), status1(n_total, 0);
  CHECK(p2v_verify_batch(ctx, cir, batch.data(), n_total, bits1.data(), status1.data()), ctx);

  size_t start = 0, stop = 0;
  CHECK(p2v_shard_bounds(n_total, rank, world, &start, &stop), ctx);
  const size_t words_full = p2v_shard_slice_len(n_total, world) / 32 * (size_t)world;
  const uint64_t *local = batch.data() + start * (size_t)lay.blob_words;
  size_t accepted = 0;
  for (int mode = 0; mode < 2; mode++) {  // 0: ncclAllGather, 1: direct stores into the peers' buffers
    if (mode == 1) CHECK(p2v_peer_enable(ctx), ctx);
    for (int rep = 0; rep < 3; rep++) {
      std::vector<uint32_t> full(words_full, 0xDEADBEEFu), st(stop - start + 1, 0xFFFFFFFFu);
      CHECK(p2v_verify_batch_sharded(ctx, cir, local, n_total, rank, world, full.data(), st.data()), ctx);
      if (memcmp(full.data(), bits1.data(), words * 4) != 0) {
        fprintf(stderr, "[rank %d] gathered bitmap differs from the single-GPU run (mode %d)\n", rank, mode);
        return 1;
      }
      for (size_t i = start; i < stop; i++)
        if (st[i - start] != status1[i]) {
          fprintf(stderr, "[rank %d] status of proof %zu differs (mode %d)\n", rank, i, mode);
          return 1;
        }
    }
  }
  for (size_t i = 0; i < n_total; i++) accepted += (bits1[i / 32] >> (i % 32)) & 1;
  int r2 = -1, w2 = -1, ver = 0;
  CHECK(p2v_nccl_info(ctx, &r2, &w2, &ver), ctx);
  if (r2 != rank || w2 != world) return 1;
  CHECK(p2v_peer_disable(ctx), ctx);
  CHECK(p2v_nccl_finalize(ctx), ctx);
  if (rank == 0)
    printf("SHARD_DEMO_OK world=%d n=%zu accepted=%zu nccl=%d launches=%" PRIu64 "\n", world, n_total, accepted, ver, p2v_ctx_launch_count(ctx));
  p2v_circuit_destroy(cir);
  p2v_ctx_destroy(ctx);
  return 0;
}

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <common.json> <vkey.json> <proof.json> <world> [n_total]\n", argv[0]);
    return 64;
  }
  const int world = atoi(argv[4]);
  const size_t n_total = argc > 5 ? (size_t)atoll(argv[5]) : 1000;
  if (world < 1 || world > 16 || n_total < 1) return 64;
  std::string tc, tv, tp;
  if (!slurp(argv[1], tc) || !slurp(argv[2], tv) || !slurp(argv[3], tp)) {
    fprintf(stderr, "cannot read the JSON files\n");
    return 66;
  }
  // host-side decoding needs no GPU and happens BEFORE the fork (no CUDA state is inherited by the children)
  p2v_shape shape;
  p2v_layout lay;
  int rank = -1;
  CHECK(p2v_parse_common(tc.data(), tc.size(), &shape), nullptr);
  CHECK(p2v_shape_layout(&shape, &lay), nullptr);
  std::vector<uint64_t> vkey(lay.vkey_words), blob(lay.blob_words);
  CHECK(p2v_parse_vkey(tv.data(), tv.size(), &shape, vkey.data()), nullptr);
  CHECK(p2v_parse_proof(tp.data(), tp.size(), &shape, blob.data()), nullptr);
  // the batch: copies of the fixture; copy i is left alone when i % 4 == 0, otherwise ONE word is changed, walking through
  // the whole blob (openings, caps, PoW witness, leaves, siblings, coset evaluations ... of every query)
  std::vector<uint64_t> batch(n_total * (size_t)lay.blob_words);
  for (size_t i = 0; i < n_total; i++) {
    uint64_t *b = batch.data() + i * (size_t)lay.blob_words;
    memcpy(b, blob.data(), (size_t)lay.blob_words * 8);
    if (i % 4) {
      size_t w = (i * 2654435761ULL) % (size_t)lay.blob_words;
      b[w] = b[w] % GL_P + 1 == GL_P ? 0 : b[w] % GL_P + 1;
    }
  }
  char id_path[64];
  snprintf(id_path, sizeof id_path, "/tmp/p2v_shard_demo_%d.id", (int)getpid());
  unlink(id_path);
  std::vector<pid_t> kids;
  for (int r = 0; r < world; r++) {
    pid_t pid = fork();
    if (pid < 0) return 71;
    if (pid == 0) {
      int rc = run_rank(r, world, shape, lay, vkey, batch, n_total, id_path);
      fflush(stdout);
      _exit(rc);
    }
    kids.push_back(pid);
  }
  int bad = 0;
  for (pid_t pid : kids) {
    int st = 0;
    waitpid(pid, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
  }
  unlink(id_path);
  p2v_shape_free(&shape);
  if (bad) fprintf(stderr, "p2v_shard_demo: %d of %d ranks failed\n", bad, world);
  return bad ? 1 : 0;
}
