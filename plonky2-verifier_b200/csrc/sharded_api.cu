// C-ABI entry points for the multi-GPU form of the path (BASELINE config 5): every proof is
// independent, so a batch is cut into contiguous slices, one per rank (one process per GPU); each rank runs the whole
// verifier on its slice and the ONLY exchange is one all-gather of the packed accept bitmap.
//
//   p2v_shard_*               the slicing rule (multiples of 32 proofs, so bitmap words never straddle two ranks)
//   p2v_nccl_unique_id/init   bootstrap of a communicator for callers that have none (a Haskell or C host, bench.py)
//   p2v_nccl_attach           use a communicator the caller already owns
//   p2v_verify_batch_sharded  verifyProof on the slice + ncclAllGather on the context's stream (in place)
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a torch process that is the copy torch already loaded,
// elsewhere the system library; libp2v.so itself loads on machines without NCCL and the single-GPU API is unaffected.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <memory>
#include <mutex>
#include <vector>
#include "ctx.hpp"

namespace {

// the part of nccl.h this file needs (NCCL 2.x ABI: ncclUniqueId is 128 opaque bytes passed by value)
struct NcclUniqueId {
  char internal[128];
};
typedef void *NcclComm;
enum { NCCL_UINT32 = 3 };
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*CommCount)(NcclComm, int *) = nullptr;
  int (*CommUserRank)(NcclComm, int *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int *) = nullptr;
  std::string why;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      api.why = std::string("libnccl.so.2 could not be loaded: ") + dlerror();
      return;
    }
    auto sym = [&](const char *s) {
      void *p = dlsym(api.handle, s);
      if (!p && api.why.empty()) api.why = std::string("libnccl lacks ") + s;
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.CommCount = (decltype(api.CommCount))sym("ncclCommCount");
    api.CommUserRank = (decltype(api.CommUserRank))sym("ncclCommUserRank");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  });
  return api;
}

int ncclReady(p2v_ctx *ctx) {
  NcclApi &a = nccl();
  if (!a.why.empty()) return p2v_fail(ctx, P2V_E_UNSUPPORTED, a.why);
  return P2V_OK;
}

#define P2V_NCCL(ctx, call)                                                                                       \
  do {                                                                                                            \
    int r__ = (call);                                                                                             \
    if (r__ != 0) return p2v_fail((ctx), P2V_E_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r__));     \
  } while (0)

// ---- the gather without NCCL: every rank stores its slice straight into every peer's buffer over NVLink ---------------
// ("K7 writes its slice into peer-mapped memory + a system-scope flag").  Opt-in (p2v_peer_enable); the
// NCCL path stays the default.  Each rank owns one PeerBuf (cudaMalloc, exported with cudaIpcGetMemHandle, the handles
// travel once through the communicator); a call with epoch e uses half e & 1 of the buffers:
//   k_peer_publish  (after the verdict kernels of all lanes have been joined): the W words of this rank's slice go to
//                   bits[e&1][rank*W ..] of EVERY rank (remote stores, 4 B each), then __threadfence_system() and a
//                   release store of e into flags[e&1][rank] of every rank;
//   k_peer_wait     one lane per peer spins (acquire loads, bounded by a timeout) until flags[e&1][r] == e for all r.
// Re-use of a half is safe without a barrier: a rank publishes e+1 only after its own copy-out of e (stream order), and
// nobody can enter e+2 before every rank has published e+1.
#define P2V_PEER_MAX_WORDS (1u << 18) /* 8.4 M proofs per call */
#define P2V_PEER_MAX_WORLD 16
struct PeerBuf {
  uint32_t bits[2][P2V_PEER_MAX_WORDS];
  uint32_t flags[2][P2V_PEER_MAX_WORLD];
  uint32_t timed_out;
};

struct PeerPtrs {
  PeerBuf *p[P2V_PEER_MAX_WORLD];
};

__global__ void k_peer_publish(const uint32_t *__restrict__ mine, size_t words, PeerPtrs peers, int world, int rank, uint32_t epoch) {
  const int half = (int)(epoch & 1u);
  for (size_t i = threadIdx.x; i < words * (size_t)world; i += blockDim.x) {
    size_t r = i / words, w = i - r * words;
    peers.p[r]->bits[half][(size_t)rank * words + w] = mine[w];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < (unsigned)world) {
    uint32_t *flag = &peers.p[threadIdx.x]->flags[half][rank];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
  }
}

__global__ void k_peer_wait(PeerBuf *local, int world, uint32_t epoch, long long timeout_cycles) {
  const int half = (int)(epoch & 1u);
  if (threadIdx.x < (unsigned)world) {
    const uint32_t *flag = &local->flags[half][threadIdx.x];
    long long t0 = clock64();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v == epoch) break;
      if (clock64() - t0 > timeout_cycles) {
        local->timed_out = epoch;
        break;
      }
      __nanosleep(200);
    }
  }
}

}  // namespace

struct p2v_peer_state {
  PeerBuf *local = nullptr;
  PeerPtrs peers = {};
  int world = 0, rank = 0;
  uint32_t epoch = 0;
};

extern "C" {

size_t p2v_shard_slice_len(size_t n_total, int world) {
  if (world < 1) return 0;
  size_t per = (n_total + (size_t)world - 1) / (size_t)world;
  return (per + 31) / 32 * 32;
}

int p2v_shard_bounds(size_t n_total, int rank, int world, size_t *start, size_t *stop) {
  if (world < 1 || rank < 0 || rank >= world || !start || !stop) return P2V_E_INVALID;
  size_t per = p2v_shard_slice_len(n_total, world);
  size_t s = (size_t)rank * per;
  if (s > n_total) s = n_total;
  size_t e = s + per;
  if (e > n_total) e = n_total;
  *start = s;
  *stop = e;
  return P2V_OK;
}

int p2v_nccl_unique_id(void *out128) {
  if (!out128) return p2v_fail(nullptr, P2V_E_INVALID, "p2v_nccl_unique_id: NULL argument");
  int rc = ncclReady(nullptr);
  if (rc) return rc;
  NcclUniqueId id;
  P2V_NCCL(nullptr, nccl().GetUniqueId(&id));
  memcpy(out128, id.internal, sizeof id.internal);
  return P2V_OK;
}

int p2v_peer_disable(p2v_ctx *ctx);
int p2v_nccl_finalize(p2v_ctx *ctx) {
  if (!ctx) return P2V_E_INVALID;
  p2v_peer_disable(ctx);
  if (ctx->nccl_comm && ctx->nccl_owned) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    nccl().CommDestroy((NcclComm)ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->nccl_owned = false;
  ctx->nccl_rank = 0;
  ctx->nccl_world = 1;
  return P2V_OK;
}

int p2v_nccl_init(p2v_ctx *ctx, const void *id128, int rank, int world) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return p2v_fail(ctx, P2V_E_INVALID, "p2v_nccl_init: bad argument");
  int rc = ncclReady(ctx);
  if (rc) return rc;
  p2v_nccl_finalize(ctx);
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id;
  memcpy(id.internal, id128, sizeof id.internal);
  NcclComm comm = nullptr;
  P2V_NCCL(ctx, nccl().CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->nccl_owned = true;
  ctx->nccl_rank = rank;
  ctx->nccl_world = world;
  return P2V_OK;
}

int p2v_nccl_attach(p2v_ctx *ctx, void *nccl_comm) {
  if (!ctx || !nccl_comm) return p2v_fail(ctx, P2V_E_INVALID, "p2v_nccl_attach: NULL argument");
  int rc = ncclReady(ctx);
  if (rc) return rc;
  int world = 0, rank = -1;
  P2V_NCCL(ctx, nccl().CommCount((NcclComm)nccl_comm, &world));
  P2V_NCCL(ctx, nccl().CommUserRank((NcclComm)nccl_comm, &rank));
  p2v_nccl_finalize(ctx);
  ctx->nccl_comm = nccl_comm;
  ctx->nccl_owned = false;
  ctx->nccl_rank = rank;
  ctx->nccl_world = world;
  return P2V_OK;
}

int p2v_nccl_info(p2v_ctx *ctx, int *rank, int *world, int *version) {
  if (!ctx) return P2V_E_INVALID;
  if (rank) *rank = ctx->nccl_comm ? ctx->nccl_rank : 0;
  if (world) *world = ctx->nccl_comm ? ctx->nccl_world : 1;
  if (version) {
    *version = 0;
    if (ncclReady(nullptr) == P2V_OK && nccl().GetVersion) nccl().GetVersion(version);
  }
  return P2V_OK;
}

int p2v_peer_disable(p2v_ctx *ctx) {
  if (!ctx) return P2V_E_INVALID;
  p2v_peer_state *ps = (p2v_peer_state *)ctx->peer;
  if (!ps) return P2V_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < ps->world; r++)
    if (r != ps->rank && ps->peers.p[r]) cudaIpcCloseMemHandle(ps->peers.p[r]);
  if (ps->local) cudaFree(ps->local);
  cudaGetLastError();
  delete ps;
  ctx->peer = nullptr;
  return P2V_OK;
}

// Collective over the context's communicator: afterwards p2v_verify_batch_sharded gathers the bitmap with direct stores
// into the peers' buffers instead of ncclAllGather.  All ranks must live on one node with peer access (NVLink/NVSwitch).
int p2v_peer_enable(p2v_ctx *ctx) {
  if (!ctx) return P2V_E_INVALID;
  if (!ctx->nccl_comm) return p2v_fail(ctx, P2V_E_INVALID, "p2v_peer_enable: no communicator (p2v_nccl_init / p2v_nccl_attach first)");
  const int world = ctx->nccl_world, rank = ctx->nccl_rank;
  if (world > P2V_PEER_MAX_WORLD) return p2v_fail(ctx, P2V_E_UNSUPPORTED, "p2v_peer_enable: more than 16 ranks");
  p2v_peer_disable(ctx);
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  std::unique_ptr<p2v_peer_state> ps(new p2v_peer_state());
  ps->world = world;
  ps->rank = rank;
  P2V_CUDA(ctx, cudaMalloc(&ps->local, sizeof(PeerBuf)));
  P2V_CUDA(ctx, cudaMemsetAsync(ps->local, 0, sizeof(PeerBuf), ctx->stream));
  // exchange the IPC handles through the communicator (64 bytes per rank)
  cudaIpcMemHandle_t mine;
  cudaError_t e = cudaIpcGetMemHandle(&mine, ps->local);
  char *d_handles = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_handles, sizeof(mine) * (size_t)world);
  if (e != cudaSuccess) {
    cudaFree(ps->local);
    cudaGetLastError();
    return p2v_fail(ctx, P2V_E_CUDA, std::string("p2v_peer_enable: ") + cudaGetErrorString(e));
  }
  std::vector<cudaIpcMemHandle_t> all(world);
  int rc = P2V_OK;
  do {
    if (cudaMemcpyAsync(d_handles + sizeof(mine) * rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = P2V_E_CUDA; break; }
    if (nccl().AllGather(d_handles + sizeof(mine) * rank, d_handles, sizeof(mine), /*ncclUint8*/ 1, (NcclComm)ctx->nccl_comm, ctx->stream) != 0) { rc = P2V_E_CUDA; break; }
    if (cudaMemcpyAsync(all.data(), d_handles, sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { rc = P2V_E_CUDA; break; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = P2V_E_CUDA; break; }
    for (int r = 0; r < world && rc == P2V_OK; r++) {
      if (r == rank) { ps->peers.p[r] = ps->local; continue; }
      void *ptr = nullptr;
      e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { rc = P2V_E_CUDA; p2v_fail(ctx, rc, std::string("p2v_peer_enable: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); break; }
      ps->peers.p[r] = (PeerBuf *)ptr;
    }
  } while (0);
  cudaFree(d_handles);
  if (rc != P2V_OK) {
    for (int r = 0; r < world; r++)
      if (r != rank && ps->peers.p[r]) cudaIpcCloseMemHandle(ps->peers.p[r]);
    cudaFree(ps->local);
    cudaGetLastError();
    if (ctx->err.empty()) p2v_fail(ctx, rc, "p2v_peer_enable: handle exchange failed");
    return rc;
  }
  ctx->peer = ps.release();
  return P2V_OK;
}

// verifyProof (Plonk/Verifier.hs:56-65) over a batch of n_total proofs of which this rank holds the slice
// [start, stop) = p2v_shard_bounds(n_total, rank, world):  blobs_local is AoS [stop - start][blob_words].
//   accept_bits_full: world * slice_len / 32 words (host or device); after the call the first ceil(n_total/32) words
//                     are the accept bitmap of the WHOLE batch, identical on every rank (padding bits are 0);
//   status_local:     [stop - start] status words of this rank's proofs (may be NULL).
// world == 1 needs no communicator and is exactly p2v_verify_batch.
int p2v_verify_batch_sharded(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs_local, size_t n_total, int rank, int world,
                             uint32_t *accept_bits_full, uint32_t *status_local) {
  if (!ctx || !c || !accept_bits_full) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch_sharded: NULL argument");
  size_t start = 0, stop = 0;
  if (p2v_shard_bounds(n_total, rank, world, &start, &stop) != P2V_OK) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch_sharded: bad rank/world");
  const size_t n_local = stop - start;
  if (n_local && !blobs_local) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch_sharded: blobs_local is NULL");
  if (world > 1) {
    if (!ctx->nccl_comm) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch_sharded: no communicator (p2v_nccl_init / p2v_nccl_attach first)");
    if (ctx->nccl_world != world || ctx->nccl_rank != rank)
      return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch_sharded: rank/world differ from the communicator's");
  }
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t words_per_rank = p2v_shard_slice_len(n_total, world) / 32;
  const size_t words_full = words_per_rank * (size_t)world;
  if (words_full == 0) return P2V_OK;
  // the gathered bitmap lives on the device: the caller's buffer if it is device memory, else a pool temporary
  DevOut full;
  int rc;
  if ((rc = full.init(ctx, accept_bits_full, words_full * 4))) return rc;
  uint32_t *mine = full.as<uint32_t>() + (size_t)rank * words_per_rank;
  // padding words of a short (or empty) last slice are zero; K7 writes ceil(n_local/32) words
  P2V_CUDA(ctx, cudaMemsetAsync(mine, 0, words_per_rank * 4, ctx->stream));
  // The call is a COLLECTIVE: a rank whose slice cannot be verified (bad pointer, out of memory, shape mismatch ...) must
  // still take part in the exchange, or every other rank would wait for it forever.  It contributes an all-zero slice
  // (nothing accepted), completes the gather, and only then reports its error.
  int rc_local = P2V_OK;
  std::string err_local;
  if (n_local) {
    rc_local = p2v_verify_batch(ctx, c, blobs_local, n_local, mine, status_local);
    if (rc_local != P2V_OK) {
      err_local = ctx->err;
      if (world == 1) return rc_local;
      cudaGetLastError();
      if (cudaMemsetAsync(mine, 0, words_per_rank * 4, ctx->stream) != cudaSuccess) return rc_local;  // sticky CUDA error: nothing more can run
    }
  }
#define P2V_SHARDED_DONE() \
  do { if (rc_local != P2V_OK) return p2v_fail(ctx, rc_local, err_local + " (this rank contributed an all-zero slice to the gather)"); return P2V_OK; } while (0)
  p2v_peer_state *ps = (p2v_peer_state *)ctx->peer;
  if (world > 1 && ps && ps->world == world && ps->rank == rank && words_full <= P2V_PEER_MAX_WORDS) {
    // direct stores into every peer's buffer + flags (no NCCL kernel on the path)
    const uint32_t epoch = ++ps->epoch;
    P2V_LAUNCH(ctx, k_peer_publish, 1, 256, 0, mine, words_per_rank, ps->peers, world, rank, epoch);
    static const double timeout_s = getenv("P2V_PEER_TIMEOUT_S") ? atof(getenv("P2V_PEER_TIMEOUT_S")) : 30.0;
    P2V_LAUNCH(ctx, k_peer_wait, 1, 32, 0, ps->local, world, epoch, (long long)(timeout_s * 1.9e9));
    P2V_CUDA(ctx, cudaMemcpyAsync(full.as<uint32_t>(), ps->local->bits[epoch & 1u], words_full * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    if (full.host) {
      // a host caller gets the time-out as an error code; a device caller must watch its own stream
      uint32_t timed_out = 0;
      if ((rc = full.finish())) return rc;
      P2V_CUDA(ctx, cudaMemcpyAsync(&timed_out, &ps->local->timed_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
      P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (timed_out) return p2v_fail(ctx, P2V_E_CUDA, "p2v_verify_batch_sharded: a peer did not publish its slice in time");
    }
    P2V_SHARDED_DONE();
  }
  if (world > 1) {
    // in place: rank r's words already sit at recvbuff + r * count
    P2V_NCCL(ctx, nccl().AllGather(mine, full.as<uint32_t>(), words_per_rank, NCCL_UINT32, (NcclComm)ctx->nccl_comm, ctx->stream));
  }
  if ((rc = full.finish())) return rc;
  if (full.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  P2V_SHARDED_DONE();
#undef P2V_SHARDED_DONE
}

}  // extern "C"
