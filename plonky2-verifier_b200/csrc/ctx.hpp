// Context, error plumbing and host/device pointer staging shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <map>
#include <vector>
#include "../../include/p2v.h"

#define P2V_MAX_DEPTH 4

struct p2v_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // chunk pipeline: lane 0 is `stream` + `ws`; lanes 1.. have their own stream and workspace, so the latency-bound
  // per-proof kernels (K0, K4, K5) of the next chunks run under the Merkle kernel of the current one
  cudaStream_t lane_stream[P2V_MAX_DEPTH] = {};  // [0] unused (= stream)
  void *lane_ws[P2V_MAX_DEPTH] = {};             // [0] unused (= ws)
  size_t lane_ws_bytes[P2V_MAX_DEPTH] = {};
  cudaEvent_t lane_join[P2V_MAX_DEPTH] = {};
  // per lane: K4/K5 (transcript, constraints) run on a side stream next to the leaf phase of the Merkle kernel
  cudaStream_t side_stream[P2V_MAX_DEPTH] = {};
  cudaEvent_t staged_ev[P2V_MAX_DEPTH] = {}, transcript_ev[P2V_MAX_DEPTH] = {};
  int prio_hi = 0, prio_lo = 0;    // cudaDeviceGetStreamPriorityRange: greatest (numerically lowest) and least priority
  int pipeline = 4;                // 1 = strictly serial chunks (per-section timings valid), 2..P2V_MAX_DEPTH = overlapped
  cudaEvent_t fork_ev = nullptr;
  std::string err;
  uint64_t launches = 0;
  size_t chunk = 0;  // proofs per pass; 0 = default
  std::map<std::string, float> last_ms;
  // verify workspace (lazily grown)
  void *ws = nullptr;
  size_t ws_bytes = 0;
  void *stage_buf[P2V_MAX_DEPTH + 1] = {};  // ring of AoS chunk buffers for host input (depth + 1 in use)
  size_t stage_bytes = 0;
  int stage_count = 0;
  cudaEvent_t stage_filled[P2V_MAX_DEPTH + 1] = {}, stage_free[P2V_MAX_DEPTH + 1] = {};
  // host input, transcript-first schedule: the per-proof parts of a whole window of chunks are copied ahead of the query parts
  void *pp_stage = nullptr;  // [window][proof_words]
  size_t pp_stage_bytes = 0;
  cudaEvent_t pp_filled = nullptr, k0_done[P2V_MAX_DEPTH] = {}, lane_done[P2V_MAX_DEPTH] = {};
  cudaEvent_t ev[8] = {};
  // Private stream-ordered pool for the staged copies of host inputs/outputs (DevIn/DevOut).  Its release
  // threshold is unlimited: with the default pool (threshold 0) every synchronisation hands the freed blocks
  // back to the driver and the next call re-maps them, which cost 30-160 ms of host time per call at random
  // (measured with P2V_TRACE: identical GPU timelines, end-to-end throughput between 2.0e5 and 3.4e5 proofs/s).
  cudaMemPool_t pool = nullptr;
  // multi-GPU (sharded_api.cu): communicator for the accept-bitmap all-gather
  void *nccl_comm = nullptr;   // ncclComm_t
  bool nccl_owned = false;     // created by p2v_nccl_init (destroyed with the context) or attached by the caller
  int nccl_rank = 0, nccl_world = 1;
  void *peer = nullptr;        // p2v_peer_state (sharded_api.cu): peer-mapped gather buffers, opt-in
};

extern thread_local std::string p2v_tls_error;

static inline int p2v_fail(p2v_ctx *ctx, int code, const std::string &msg) {
  p2v_tls_error = msg;
  if (ctx) ctx->err = msg;
  return code;
}

#define P2V_CUDA(ctx, call)                                                                     \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess)                                                                     \
      return p2v_fail((ctx), P2V_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define P2V_LAUNCH_ON(ctx, strm, kernel, grid, block, smem, ...)                \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                   \
    (ctx)->launches++;                                                          \
    P2V_CUDA((ctx), cudaGetLastError());                                        \
  } while (0)
#define P2V_LAUNCH(ctx, kernel, grid, block, smem, ...) P2V_LAUNCH_ON(ctx, (ctx)->stream, kernel, grid, block, smem, __VA_ARGS__)

// Launch with an execution priority (cudaLaunchAttributePriority; numerically LOWER = scheduled first when SM resources
// free up — running blocks are never pre-empted).  The chunk pipeline uses it to let the short latency-bound kernels of a
// chunk (K0, K4, K5) and the closing kernels of the OLDEST chunk overtake the bulk Merkle blocks of younger chunks.
template <typename... KArgs, typename... Args>
static inline cudaError_t p2v_launch_prio(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t strm, int prio, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = strm;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = prio;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define P2V_LAUNCH_PRIO(ctx, strm, prio, kernel, grid, block, smem, ...)                                   \
  do {                                                                                                     \
    P2V_CUDA((ctx), p2v_launch_prio(kernel, dim3(grid), dim3(block), (smem), (strm), (prio), __VA_ARGS__)); \
    (ctx)->launches++;                                                                                     \
    P2V_CUDA((ctx), cudaGetLastError());                                                                   \
  } while (0)

static inline bool p2v_is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Input that may live on the host: staged to the device through the context's stream.
struct DevIn {
  p2v_ctx *ctx;
  const void *dev = nullptr;
  void *tmp = nullptr;
  int init(p2v_ctx *c, const void *p, size_t bytes) {
    ctx = c;
    if (!p || bytes == 0) { dev = p; return 0; }
    if (p2v_is_device_ptr(p)) { dev = p; return 0; }
    P2V_CUDA(ctx, cudaMallocFromPoolAsync(&tmp, bytes, ctx->pool, ctx->stream));
    P2V_CUDA(ctx, cudaMemcpyAsync(tmp, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    dev = tmp;
    return 0;
  }
  ~DevIn() { if (tmp) cudaFreeAsync(tmp, ctx->stream); }
  template <class T> const T *as() const { return (const T *)dev; }
};

// Output that may live on the host: computed on the device, copied back by finish().
struct DevOut {
  p2v_ctx *ctx;
  void *host = nullptr;
  void *dev = nullptr;
  void *tmp = nullptr;
  size_t bytes = 0;
  int init(p2v_ctx *c, void *p, size_t nbytes) {
    ctx = c;
    bytes = nbytes;
    if (!p || nbytes == 0) { dev = nullptr; return 0; }
    if (p2v_is_device_ptr(p)) { dev = p; return 0; }
    host = p;
    P2V_CUDA(ctx, cudaMallocFromPoolAsync(&tmp, nbytes, ctx->pool, ctx->stream));
    dev = tmp;
    return 0;
  }
  // enqueue the copy back; the caller synchronises the stream afterwards
  int finish() {
    if (tmp && host) P2V_CUDA(ctx, cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
  }
  ~DevOut() { if (tmp) cudaFreeAsync(tmp, ctx->stream); }
  template <class T> T *as() const { return (T *)dev; }
};

static inline int p2v_grid_for(p2v_ctx *ctx, size_t n, int block, int blocks_per_sm) {
  size_t need = (n + block - 1) / block;
  size_t cap = (size_t)ctx->sm_count * blocks_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}
