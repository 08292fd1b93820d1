// C-ABI entry points: context + L2 hash layer (K1-K3) + measurement helpers.  See include/p2v.h.
#include "ctx.hpp"
#include "hash_kernels.cuh"

thread_local std::string p2v_tls_error;

// streams, events and the private pool of a new context; on failure the caller destroys the partly built context
static int ctxInit(p2v_ctx *ctx, int device) {
  P2V_CUDA(nullptr, cudaDeviceGetStreamPriorityRange(&ctx->prio_lo, &ctx->prio_hi));
  P2V_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  P2V_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 1; i < P2V_MAX_DEPTH; i++) {
    P2V_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->lane_stream[i], cudaStreamNonBlocking));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->lane_join[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < P2V_MAX_DEPTH; i++) {
    P2V_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->side_stream[i], cudaStreamNonBlocking));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->staged_ev[i], cudaEventDisableTiming));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->transcript_ev[i], cudaEventDisableTiming));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->k0_done[i], cudaEventDisableTiming));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->lane_done[i], cudaEventDisableTiming));
  }
  P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
  P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->pp_filled, cudaEventDisableTiming));
  for (auto &ev : ctx->ev) P2V_CUDA(nullptr, cudaEventCreate(&ev));
  for (int i = 0; i < P2V_MAX_DEPTH + 1; i++) {
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->stage_filled[i], cudaEventDisableTiming));
    P2V_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming));
  }
  cudaMemPoolProps props = {};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = device;
  P2V_CUDA(nullptr, cudaMemPoolCreate(&ctx->pool, &props));
  uint64_t keep = UINT64_MAX;
  P2V_CUDA(nullptr, cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  return P2V_OK;
}

extern "C" {

int p2v_abi_version(void) { return P2V_ABI_VERSION; }

const char *p2v_last_error(const p2v_ctx *ctx) { return ctx ? ctx->err.c_str() : p2v_tls_error.c_str(); }

int p2v_ctx_create(int device, p2v_ctx **out) {
  if (!out) return p2v_fail(nullptr, P2V_E_INVALID, "p2v_ctx_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return p2v_fail(nullptr, P2V_E_NOGPU,
                    "p2v_ctx_create: no CUDA device available; libp2v has no CPU fallback");
  }
  if (device < 0 || device >= count) return p2v_fail(nullptr, P2V_E_INVALID, "p2v_ctx_create: bad device index");
  cudaDeviceProp prop;
  P2V_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return p2v_fail(nullptr, P2V_E_NOGPU, "p2v_ctx_create: kernels are built for sm_100a only (Blackwell B200)");
  P2V_CUDA(nullptr, cudaSetDevice(device));
  p2v_ctx *ctx = new p2v_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  int rc = ctxInit(ctx, device);
  if (rc != P2V_OK) {
    std::string why = p2v_tls_error;  // p2v_ctx_destroy tolerates a partly built context
    p2v_ctx_destroy(ctx);
    return p2v_fail(nullptr, rc, why);
  }
  *out = ctx;
  return P2V_OK;
}

void p2v_ctx_destroy(p2v_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  for (int i = 1; i < P2V_MAX_DEPTH; i++)
    if (ctx->lane_stream[i]) cudaStreamSynchronize(ctx->lane_stream[i]);
  for (int i = 0; i < P2V_MAX_DEPTH; i++) {
    if (ctx->side_stream[i]) { cudaStreamSynchronize(ctx->side_stream[i]); cudaStreamDestroy(ctx->side_stream[i]); }
    if (ctx->staged_ev[i]) cudaEventDestroy(ctx->staged_ev[i]);
    if (ctx->transcript_ev[i]) cudaEventDestroy(ctx->transcript_ev[i]);
    if (ctx->k0_done[i]) cudaEventDestroy(ctx->k0_done[i]);
    if (ctx->lane_done[i]) cudaEventDestroy(ctx->lane_done[i]);
  }
  if (ctx->pp_filled) cudaEventDestroy(ctx->pp_filled);
  if (ctx->pp_stage) cudaFree(ctx->pp_stage);
  p2v_nccl_finalize(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  for (int i = 1; i < P2V_MAX_DEPTH; i++) {
    if (ctx->lane_ws[i]) cudaFree(ctx->lane_ws[i]);
    if (ctx->lane_join[i]) cudaEventDestroy(ctx->lane_join[i]);
    if (ctx->lane_stream[i]) cudaStreamDestroy(ctx->lane_stream[i]);
  }
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  for (auto &b : ctx->stage_buf)
    if (b) cudaFree(b);
  for (auto &ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < P2V_MAX_DEPTH + 1; i++) {
    if (ctx->stage_filled[i]) cudaEventDestroy(ctx->stage_filled[i]);
    if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]);
  }
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaGetLastError();
  delete ctx;
}

void *p2v_ctx_stream(p2v_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int p2v_ctx_sync(p2v_ctx *ctx) {
  if (!ctx) return P2V_E_INVALID;
  P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

uint64_t p2v_ctx_launch_count(const p2v_ctx *ctx) { return ctx ? ctx->launches : 0; }

int p2v_ctx_set_pipeline(p2v_ctx *ctx, int depth) {
  if (!ctx || depth < 1 || depth > P2V_MAX_DEPTH) return p2v_fail(ctx, P2V_E_INVALID, "p2v_ctx_set_pipeline: depth must be 1.." + std::to_string(P2V_MAX_DEPTH));
  ctx->pipeline = depth;
  return P2V_OK;
}

int p2v_ctx_set_chunk(p2v_ctx *ctx, size_t proofs_per_chunk) {
  if (!ctx) return P2V_E_INVALID;
  ctx->chunk = proofs_per_chunk;
  return P2V_OK;
}

int p2v_host_alloc(size_t bytes, void **out) {
  if (!out) return P2V_E_INVALID;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return p2v_fail(nullptr, P2V_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  }
  return P2V_OK;
}
void p2v_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- K1 -----------------------------------------------------------------------------------
int p2v_poseidon_permute(p2v_ctx *ctx, const uint64_t *in, uint64_t *out, size_t n) {
  if (!ctx || !in || !out) return p2v_fail(ctx, P2V_E_INVALID, "p2v_poseidon_permute: NULL argument");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  DevIn din;
  DevOut dout;
  int rc;
  if ((rc = din.init(ctx, in, n * 12 * sizeof(u64)))) return rc;
  if ((rc = dout.init(ctx, out, n * 12 * sizeof(u64)))) return rc;
  P2V_LAUNCH(ctx, k_poseidon_permute, p2v_grid_for(ctx, n, 256, 16), 256, 0, din.as<u64>(), dout.as<u64>(), n);
  if ((rc = dout.finish())) return rc;
  if (dout.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

// ---- K2 -----------------------------------------------------------------------------------
int p2v_hash_leaves(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, size_t n, uint64_t *digests) {
  if (!ctx || !digests || (!leaves && w)) return p2v_fail(ctx, P2V_E_INVALID, "p2v_hash_leaves: NULL argument");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  DevIn din;
  DevOut dout;
  int rc;
  if ((rc = din.init(ctx, leaves, n * (size_t)w * sizeof(u64)))) return rc;
  if ((rc = dout.init(ctx, digests, n * 4 * sizeof(u64)))) return rc;
  P2V_LAUNCH(ctx, k_hash_leaves, p2v_grid_for(ctx, n, 256, 16), 256, 0, din.as<u64>(), w, n, dout.as<u64>());
  if ((rc = dout.finish())) return rc;
  if (dout.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

int p2v_compress(p2v_ctx *ctx, const uint64_t *left, const uint64_t *right, uint64_t *out, size_t n) {
  if (!ctx || !left || !right || !out) return p2v_fail(ctx, P2V_E_INVALID, "p2v_compress: NULL argument");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  DevIn dl, dr;
  DevOut dout;
  int rc;
  if ((rc = dl.init(ctx, left, n * 4 * sizeof(u64)))) return rc;
  if ((rc = dr.init(ctx, right, n * 4 * sizeof(u64)))) return rc;
  if ((rc = dout.init(ctx, out, n * 4 * sizeof(u64)))) return rc;
  P2V_LAUNCH(ctx, k_compress, p2v_grid_for(ctx, n, 256, 16), 256, 0, dl.as<u64>(), dr.as<u64>(), n, (size_t)1,
             dout.as<u64>(), n, n);
  if ((rc = dout.finish())) return rc;
  if (dout.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

// ---- K3 -----------------------------------------------------------------------------------
int p2v_merkle_verify(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, const uint32_t *idx, const uint64_t *siblings,
                      uint32_t path_len, const uint64_t *cap, uint32_t cap_height, size_t n, uint32_t *ok_bits,
                      uint64_t *roots_out) {
  if (!ctx || !idx || !cap || !ok_bits || (!leaves && w) || (!siblings && path_len))
    return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_verify: NULL argument");
  if (cap_height > 20 || path_len > 32) return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_verify: bad heights");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  DevIn dl, di, ds, dc;
  DevOut dok, droots;
  int rc;
  if ((rc = dl.init(ctx, leaves, n * (size_t)w * sizeof(u64)))) return rc;
  if ((rc = di.init(ctx, idx, n * sizeof(u32)))) return rc;
  if ((rc = ds.init(ctx, siblings, n * (size_t)path_len * 4 * sizeof(u64)))) return rc;
  if ((rc = dc.init(ctx, cap, ((size_t)4 << cap_height) * sizeof(u64)))) return rc;
  if ((rc = dok.init(ctx, ok_bits, (n + 31) / 32 * sizeof(u32)))) return rc;
  if ((rc = droots.init(ctx, roots_out, n * 4 * sizeof(u64)))) return rc;
  P2V_LAUNCH(ctx, k_merkle_verify, p2v_grid_for(ctx, n, 256, 16), 256, 0, dl.as<u64>(), w, di.as<u32>(), ds.as<u64>(),
             path_len, dc.as<u64>(), cap_height, n, dok.as<u32>(), droots.as<u64>());
  if ((rc = dok.finish())) return rc;
  if ((rc = droots.finish())) return rc;
  if (dok.host || droots.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

int p2v_merkle_build(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, uint32_t log_n, uint32_t cap_height,
                     uint64_t *digests_out) {
  if (!ctx || !digests_out || (!leaves && w)) return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_build: NULL argument");
  if (log_n > 30 || cap_height > log_n) return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_build: bad heights");
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t n = (size_t)1 << log_n;
  size_t total = 4 * (((size_t)2 << log_n) - ((size_t)1 << cap_height));
  DevIn dl;
  DevOut dout;
  int rc;
  if ((rc = dl.init(ctx, leaves, n * (size_t)w * sizeof(u64)))) return rc;
  if ((rc = dout.init(ctx, digests_out, total * sizeof(u64)))) return rc;
  u64 *lvl = dout.as<u64>();
  P2V_LAUNCH(ctx, k_hash_leaves, p2v_grid_for(ctx, n, 256, 16), 256, 0, dl.as<u64>(), w, n, lvl);
  for (uint32_t l = 0; l < log_n - cap_height; l++) {
    size_t cur_n = n >> l, nxt_n = cur_n >> 1;
    u64 *nxt = lvl + 4 * cur_n;
    // children 2t (left) and 2t+1 (right) of level l -> node t of level l+1
    P2V_LAUNCH(ctx, k_compress, p2v_grid_for(ctx, nxt_n, 256, 16), 256, 0, lvl, lvl + 1, cur_n, (size_t)2, nxt, nxt_n,
               nxt_n);
    lvl = nxt;
  }
  if ((rc = dout.finish())) return rc;
  if (dout.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

int p2v_merkle_open(p2v_ctx *ctx, const uint64_t *leaves, uint32_t w, uint32_t log_n, uint32_t cap_height,
                    const uint64_t *digests, const uint32_t *idx, size_t n, uint64_t *leaves_out,
                    uint64_t *siblings_out, uint64_t *cap_out) {
  if (log_n > 30 || cap_height > log_n) return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_open: bad heights");
  // buffers of zero size may be NULL: no leaves (w == 0), no siblings (the cap is the leaf level)
  if (!ctx || !digests || !idx || !cap_out || ((!leaves || !leaves_out) && w) || (!siblings_out && log_n > cap_height))
    return p2v_fail(ctx, P2V_E_INVALID, "p2v_merkle_open: NULL argument");
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t n_leaves = (size_t)1 << log_n;
  size_t total = 4 * (((size_t)2 << log_n) - ((size_t)1 << cap_height));
  uint32_t path_len = log_n - cap_height;
  DevIn dl, dd, di;
  DevOut dlo, dso, dco;
  int rc;
  if ((rc = dl.init(ctx, leaves, n_leaves * (size_t)w * sizeof(u64)))) return rc;
  if ((rc = dd.init(ctx, digests, total * sizeof(u64)))) return rc;
  if ((rc = di.init(ctx, idx, n * sizeof(u32)))) return rc;
  if ((rc = dlo.init(ctx, leaves_out, n * (size_t)w * sizeof(u64)))) return rc;
  if ((rc = dso.init(ctx, siblings_out, n * (size_t)path_len * 4 * sizeof(u64)))) return rc;
  if ((rc = dco.init(ctx, cap_out, ((size_t)4 << cap_height) * sizeof(u64)))) return rc;
  if (n) {
    P2V_LAUNCH(ctx, k_merkle_open, p2v_grid_for(ctx, n, 256, 16), 256, 0, dl.as<u64>(), w, log_n, cap_height,
               dd.as<u64>(), di.as<u32>(), n, dlo.as<u64>(), dso.as<u64>());
  }
  uint32_t ncap = 1u << cap_height;
  const u64 *cap_level = dd.as<u64>() + (total - 4 * (size_t)ncap);
  P2V_LAUNCH(ctx, k_cap_transpose, (ncap * 4 + 255) / 256, 256, 0, cap_level, ncap, dco.as<u64>());
  if ((rc = dlo.finish())) return rc;
  if ((rc = dso.finish())) return rc;
  if ((rc = dco.finish())) return rc;
  if (dlo.host || dso.host || dco.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

// ---- test hook: device field arithmetic --------------------------------------------------------
int p2v_debug_field_op(p2v_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
  if (!ctx || !a || !out || op < 0 || op > 24) return p2v_fail(ctx, P2V_E_INVALID, "p2v_debug_field_op: bad argument");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t words = op < 16 ? n : 2 * n;
  DevIn da, db;
  DevOut dout;
  int rc;
  if ((rc = da.init(ctx, a, words * 8))) return rc;
  if ((rc = db.init(ctx, b, words * 8))) return rc;
  if ((rc = dout.init(ctx, out, words * 8))) return rc;
  P2V_LAUNCH(ctx, k_field_op, p2v_grid_for(ctx, n, 128, 8), 128, 0, op, da.as<u64>(), db.as<u64>(), dout.as<u64>(), n);
  if ((rc = dout.finish())) return rc;
  if (dout.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

// ---- measurement helper -------------------------------------------------------------------
int p2v_int_pipe_peak(p2v_ctx *ctx, int mode, double *ops_per_s) {
  if (!ctx || !ops_per_s || mode < 0 || mode > 23) return p2v_fail(ctx, P2V_E_INVALID, "p2v_int_pipe_peak: bad argument");
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  u64 *d = nullptr;
  P2V_CUDA(ctx, cudaMalloc(&d, 8));
  const uint32_t iters = 4096;
  int grid = ctx->sm_count * 8, block = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    P2V_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    switch (mode) {
      case 0: P2V_LAUNCH(ctx, k_int_pipe<0>, grid, block, 0, d, iters, 12345u); break;
      case 1: P2V_LAUNCH(ctx, k_int_pipe<1>, grid, block, 0, d, iters, 12345u); break;
      case 2: P2V_LAUNCH(ctx, k_int_pipe<2>, grid, block, 0, d, iters, 12345u); break;
      case 3: P2V_LAUNCH(ctx, k_int_pipe<3>, grid, block, 0, d, iters, 12345u); break;
      case 4: P2V_LAUNCH(ctx, k_int_pipe<4>, grid, block, 0, d, iters, 12345u); break;
      case 5: P2V_LAUNCH(ctx, k_int_pipe<5>, grid, block, 0, d, iters, 12345u); break;
      case 6: P2V_LAUNCH(ctx, k_int_pipe<6>, grid, block, 0, d, iters, 12345u); break;
      case 7: P2V_LAUNCH(ctx, k_int_pipe<7>, grid, block, 0, d, iters, 12345u); break;
      case 8: P2V_LAUNCH(ctx, k_int_pipe<8>, grid, block, 0, d, iters, 12345u); break;
      case 9: P2V_LAUNCH(ctx, k_int_pipe<9>, grid, block, 0, d, iters, 12345u); break;
      case 10: P2V_LAUNCH(ctx, k_int_pipe<10>, grid, block, 0, d, iters, 12345u); break;
      case 11: P2V_LAUNCH(ctx, k_int_pipe<11>, grid, block, 0, d, iters, 12345u); break;
      case 12: P2V_LAUNCH(ctx, k_int_pipe<12>, grid, block, 0, d, iters, 12345u); break;
      case 13: P2V_LAUNCH(ctx, k_int_pipe<13>, grid, block, 0, d, iters, 12345u); break;
      case 14: P2V_LAUNCH(ctx, k_int_pipe<14>, grid, block, 0, d, iters, 12345u); break;
      case 15: P2V_LAUNCH(ctx, k_int_pipe<15>, grid, block, 0, d, iters, 12345u); break;
      case 16: P2V_LAUNCH(ctx, k_rf_probe<16>, grid, block, 0, d, iters, 12345u); break;
      case 17: P2V_LAUNCH(ctx, k_rf_probe<17>, grid, block, 0, d, iters, 12345u); break;
      case 18: P2V_LAUNCH(ctx, k_rf_probe<18>, grid, block, 0, d, iters, 12345u); break;
      case 19: P2V_LAUNCH(ctx, k_rf_probe<19>, grid, block, 0, d, iters, 12345u); break;
      case 20: P2V_LAUNCH(ctx, k_rf_probe<20>, grid, block, 0, d, iters, 12345u); break;
      case 21: P2V_LAUNCH(ctx, k_rf_probe<21>, grid, block, 0, d, iters, 12345u); break;
      case 22: P2V_LAUNCH(ctx, k_rf_probe<22>, grid, block, 0, d, iters, 12345u); break;
      default: P2V_LAUNCH(ctx, k_rf_probe<23>, grid, block, 0, d, iters, 12345u); break;
    }
    P2V_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    P2V_CUDA(ctx, cudaEventSynchronize(ctx->ev[1]));
    float ms = 0;
    P2V_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaFree(d);
  double per_thread = (double)iters * 64.0;  // groups per thread
  *ops_per_s = per_thread * (double)grid * block / (best * 1e-3);
  return P2V_OK;
}

}  // extern "C"
