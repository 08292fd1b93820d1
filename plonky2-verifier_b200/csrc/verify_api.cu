// C-ABI entry points for circuits and proof batches: p2v_circuit_create, p2v_challenges,
// p2v_constraints, p2v_fri, p2v_verify_batch, p2v_synth_batch.  See include/p2v.h.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include "constraints.cuh"
#include "ctx.hpp"
#include "host/field.hpp"

struct p2v_circuit {
  p2v_ctx *ctx = nullptr;
  p2v_shape shape;
  DevCircuit dev;
  u64 *d_blob = nullptr;  // vkey | k_is | weights | lut_pairs | tab
};

namespace {

using namespace p2vhost;

void buildTranscript(DevCircuit &d) {
  const p2v_layout &L = d.L;
  int n = 0;
  auto add = [&](int kind, int off, int count, int dst) {
    if (n >= P2V_MAX_TOPS) { n++; return; }  // counted, reported by the caller (cannot happen within P2V_MAX_STEPS)
    d.ops[n].kind = kind; d.ops[n].off = off; d.ops[n].count = count; d.ops[n].dst = dst;
    n++;
  };
  int r = d.r;
  // sponge public_inputs (Challenge/Verifier.hs:67)
  add(TOP_ABSORB_PROOF, L.off_public_inputs, d.num_pi, 0);
  add(TOP_SPONGE_FINISH, 0, 0, 0);
  // :73-79
  add(TOP_ABSORB_VKEY, L.cap_words, 4, 0);
  add(TOP_ABSORB_PIH, 0, 4, 0);
  add(TOP_ABSORB_PROOF, L.off_wires_cap, L.cap_words, 0);
  add(TOP_SQUEEZE, 0, r, d.ch_betas);
  add(TOP_SQUEEZE, 0, r, d.ch_gammas);
  if (d.num_lookup_polys > 0) add(TOP_SQUEEZE, 0, 2 * r, d.ch_deltas + 2 * r);  // :82-86
  add(TOP_ABSORB_PROOF, L.off_zs_pp_cap, L.cap_words, 0);                      // :88-89
  add(TOP_SQUEEZE, 0, r, d.ch_alphas);
  add(TOP_ABSORB_PROOF, L.off_quotient_cap, L.cap_words, 0);                   // :91-92
  add(TOP_SQUEEZE, 0, 2, d.ch_zeta);
  // friChallenges (Challenge/FRI.hs:73): batch_this = constants ++ sigmas ++ wires ++ zs ++ pp ++ quotient ++ lookup_zs
  add(TOP_ABSORB_PROOF, L.off_open_constants, 2 * (L.n_open_constants + L.n_open_sigmas + L.n_open_wires + L.n_open_zs), 0);
  add(TOP_ABSORB_PROOF, L.off_open_pp, 2 * (L.n_open_pp + L.n_open_quotient + L.n_open_lookup_zs), 0);
  add(TOP_ABSORB_PROOF, L.off_open_zs_next, 2 * L.n_open_zs_next, 0);
  add(TOP_ABSORB_PROOF, L.off_open_lookup_zs_next, 2 * L.n_open_lookup_zs_next, 0);
  add(TOP_SQUEEZE, 0, 2, d.ch_fri_alpha);                                      // :76
  for (int s = 0; s < d.nsteps; s++) {                                         // :79-81
    add(TOP_ABSORB_PROOF, L.off_commit_caps + s * L.cap_words, L.cap_words, 0);
    add(TOP_SQUEEZE, 0, 2, d.ch_fri_betas + 2 * s);
  }
  add(TOP_ABSORB_PROOF, L.off_final_poly, 2 * d.final_len, 0);                 // :83
  add(TOP_ABSORB_PROOF, L.off_pow_witness, 1, 0);                              // :89
  add(TOP_SQUEEZE, 0, 1, d.ch_pow);                                            // :90
  add(TOP_SQUEEZE, 0, d.Q, d.ch_idx);                                          // :93-97
  d.nops = n;
}

struct WsAlloc {
  Workspace ws;
  size_t bytes;
};

// carve the per-chunk planes out of one allocation
size_t carve(const DevCircuit &d, size_t m, char *base, Workspace *ws, bool want_folded, bool want_roots = false) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return base ? base + o : (char *)nullptr;
  };
  u64 *pp = (u64 *)take((size_t)d.L.proof_words * m * 8);
  u64 *ch = (u64 *)take((size_t)d.ch_words * m * 8);
  u64 *pih = (u64 *)take(4 * m * 8);
  u64 *pre = (u64 *)take(4 * m * 8);
  u64 *apow = (u64 *)take((size_t)2 * d.n_first * m * 8);
  u64 *comb = (u64 *)take((size_t)2 * d.r * m * 8);
  u32 *qstat = (u32 *)take((size_t)d.Q * m * 4);
  uint8_t *tree_ok = (uint8_t *)take((size_t)(4 + d.nsteps) * d.Q * m);
  u64 *leafdig = (u64 *)take((size_t)4 * (4 + d.nsteps) * d.Q * m * 8);
  u64 *folded = want_folded ? (u64 *)take((size_t)2 * d.Q * m * 8) : nullptr;
  uint8_t *eq = (uint8_t *)take(m);
  u64 *roots = want_roots ? (u64 *)take((size_t)4 * (4 + d.nsteps) * d.Q * m * 8) : nullptr;
  if (ws) {
    ws->roots = roots; ws->ch_in = nullptr; ws->ch_in_n = 0; ws->ch_in_off = 0;
    ws->pp = pp; ws->aos = nullptr; ws->ch = ch; ws->pih = pih; ws->pre = pre; ws->apow = apow; ws->comb = comb; ws->qstat = qstat;
    ws->folded = folded; ws->eqmask = eq; ws->tree_ok = tree_ok; ws->leafdig = leafdig;
  }
  return off;
}

int ensureWorkspace(p2v_ctx *ctx, int which, size_t bytes) {
  void *&ws = which == 0 ? ctx->ws : ctx->lane_ws[which];
  size_t &have = which == 0 ? ctx->ws_bytes : ctx->lane_ws_bytes[which];
  if (have >= bytes) return P2V_OK;
  if (ws) {
    P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 1; i < P2V_MAX_DEPTH; i++) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->lane_stream[i]));
    for (int i = 0; i < P2V_MAX_DEPTH; i++) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->side_stream[i]));
    cudaFree(ws);
    ws = nullptr;
    have = 0;
  }
  cudaError_t e = cudaMalloc(&ws, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return p2v_fail(ctx, P2V_E_NOMEM, "workspace allocation of " + std::to_string(bytes >> 20) + " MiB failed: " + cudaGetErrorString(e) +
                                          " (lower it with p2v_ctx_set_chunk)");
  }
  have = bytes;
  return P2V_OK;
}

// Host input: a ring of `count` device buffers receives the chunks as they are (AoS); the kernels read the query parts in
// place, so a buffer is free again when the LAST kernel of its chunk has finished (ctx->stage_free[b]).
int ensureStage(p2v_ctx *ctx, size_t bytes, int count) {
  if (ctx->stage_bytes >= bytes && ctx->stage_count >= count) return P2V_OK;
  P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  P2V_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  for (int i = 1; i < P2V_MAX_DEPTH; i++) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->lane_stream[i]));
  for (auto &b : ctx->stage_buf) {
    if (b) cudaFree(b);
    b = nullptr;
  }
  bytes = std::max(bytes, ctx->stage_bytes);
  count = std::max(count, ctx->stage_count);
  ctx->stage_bytes = 0;
  ctx->stage_count = 0;
  for (int i = 0; i < count; i++) {
    cudaError_t e = cudaMalloc(&ctx->stage_buf[i], bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return p2v_fail(ctx, P2V_E_NOMEM, std::string("staging buffer allocation failed: ") + cudaGetErrorString(e) + " (lower the chunk with p2v_ctx_set_chunk)");
    }
  }
  ctx->stage_bytes = bytes;
  ctx->stage_count = count;
  return P2V_OK;
}

// Host input, transcript-first schedule: one buffer for the per-proof parts of a whole window of chunks
int ensurePpStage(p2v_ctx *ctx, size_t bytes) {
  if (ctx->pp_stage_bytes >= bytes) return P2V_OK;
  P2V_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  for (int i = 0; i < P2V_MAX_DEPTH; i++) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->side_stream[i]));
  if (ctx->pp_stage) cudaFree(ctx->pp_stage);
  ctx->pp_stage = nullptr;
  ctx->pp_stage_bytes = 0;
  cudaError_t e = cudaMalloc(&ctx->pp_stage, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return p2v_fail(ctx, P2V_E_NOMEM, std::string("staging buffer allocation failed: ") + cudaGetErrorString(e));
  }
  ctx->pp_stage_bytes = bytes;
  return P2V_OK;
}

enum { RUN_CHALLENGES = 1, RUN_CONSTRAINTS = 2, RUN_FRI = 4 };

struct Outputs {
  u64 *challenges = nullptr;   // SoA [ch_words][n]
  u64 *combined = nullptr;     // SoA [2r][n]
  uint8_t *eqmask = nullptr;   // [n]
  u32 *status = nullptr;       // [n]
  u32 *accept_bits = nullptr;  // ceil(n/32)
  u32 *qstatus = nullptr;      // [n][Q]
  u64 *folded = nullptr;       // SoA [2][n*Q]
  u64 *roots = nullptr;        // SoA [(4+nsteps)*4][n*Q]
  const u64 *challenges_in = nullptr;  // SoA [ch_words][n]: test hook, replaces the transcript's challenges
  int verdict_mode = 0;        // bit0 eqs, bit1 fri
};

// device-side transposes for the debug outputs
__global__ void k_copy_planes(const u64 *__restrict__ src, size_t m, int planes, u64 *__restrict__ dst, size_t n_total, size_t c0) {
  size_t total = (size_t)planes * m;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t pl = i / m, t = i - pl * m;
    dst[pl * n_total + c0 + t] = src[i];
  }
}
__global__ void k_copy_qstat(const u32 *__restrict__ qstat, size_t m, int Q, u32 *__restrict__ dst, size_t c0) {
  size_t total = (size_t)Q * m;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t q = i / m, t = i - q * m;
    dst[(c0 + t) * Q + q] = qstat[i];
  }
}
__global__ void k_copy_folded(const u64 *__restrict__ folded, size_t m, int Q, u64 *__restrict__ dst, size_t n_total, size_t c0) {
  // src [2][Q*m] with t = q*m + p ; dst [2][n_total*Q] with index p*Q + q
  size_t total = (size_t)2 * Q * m;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t half = i / ((size_t)Q * m), rem = i - half * (size_t)Q * m;
    size_t q = rem / m, t = rem - q * m;
    dst[half * n_total * Q + (c0 + t) * Q + q] = folded[i];
  }
}
__global__ void k_copy_roots(const u64 *__restrict__ roots, size_t m, int Q, int ntrees, u64 *__restrict__ dst, size_t n_total, size_t c0) {
  // src [4][ntrees*Q*m] with t = tr*Q*m + q*m + p ; dst [(tr*4+i)][n_total*Q] with index p*Q + q
  size_t per = (size_t)ntrees * Q * m, total = 4 * per;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
    size_t i = k / per, t = k - i * per;
    size_t tr = t / ((size_t)Q * m), rem = t - tr * (size_t)Q * m;
    size_t q = rem / m, p = rem - q * m;
    dst[(tr * 4 + i) * n_total * Q + (c0 + p) * Q + q] = roots[k];
  }
}
__global__ void k_copy_bytes(const uint8_t *__restrict__ src, size_t m, uint8_t *__restrict__ dst) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) dst[i] = src[i];
}
__global__ void k_lookup_delta_copies(DevCircuit c, Workspace ws, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
    for (int i = 0; i < c.r; i++) {
      // mkLookupDeltaList (betas ++ gammas ++ deltas), Challenge/Verifier.hs:82-86
      ws.ch[(size_t)(c.ch_deltas + i) * n + p] = ws.ch[(size_t)(c.ch_betas + i) * n + p];
      ws.ch[(size_t)(c.ch_deltas + c.r + i) * n + p] = ws.ch[(size_t)(c.ch_gammas + i) * n + p];
    }
}

// P2V_TRACE=1 in the environment: per-chunk timeline of the pipelined path on stderr (tuning aid; adds events only)
struct TracePoint { const char *what; int chunk; size_t m; cudaEvent_t ev; };
#ifndef P2V_RAMP_NUM
#define P2V_RAMP_NUM 9
#define P2V_RAMP_DEN 8
#endif
#ifndef P2V_RAMP_START_DIV
#define P2V_RAMP_START_DIV 8
#endif

// One unit of work of a batch call: n proofs of ONE circuit.  p2v_verify_batch & co. run a single job; p2v_verify_groups
// runs one job per circuit, and the chunks of all jobs share the lanes of the pipeline (a kernel launch is shape-homogeneous
// because the circuit is a kernel parameter, a batch call need not be).
struct Job {
  const p2v_circuit *cir = nullptr;
  const uint64_t *blobs = nullptr;
  size_t n = 0;
  Outputs out;
  int rc = P2V_OK;       // per-job result: a job that cannot run does not stop the others
  std::string err;
};
struct JobState {  // everything runJobs keeps per job while the chunks are in flight
  bool src_dev = false;
  size_t chunk = 0, blob_words = 0;
  size_t window = 0;  // host input, transcript-first schedule: proofs whose per-proof parts are copied ahead in one piece (0 = off)
  DevOut o_ch, o_comb, o_eq, o_status, o_bits, o_qs, o_folded, o_roots;
  DevIn i_ch;
};
struct Chunk { int job; size_t c0, m; size_t win_n = 0; };  // win_n > 0: first chunk of a window of win_n proofs (transcript-first schedule)

int runJobs(p2v_ctx *ctx, std::vector<Job> &jobs, int what) {
  auto host_t0 = std::chrono::steady_clock::now();
  auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
  if (!ctx) return p2v_fail(ctx, P2V_E_INVALID, "NULL argument");
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  // ---- per-job validation and chunking -----------------------------------------------------------------------------
  // Serial mode: as few chunks as memory allows.  Pipelined mode (default): chunks go round-robin over `pipeline` lanes
  // (stream + workspace each).  Device-resident input: ~3 GiB chunks.  Host input: ~0.5 GiB chunks — a chunk cannot start
  // before it has arrived and PCIe delivers proofs about as fast as the GPU verifies them, so there is never a backlog of
  // big chunks to overlap (tools/chunk_sweep.sh).
  static const bool pp_first_on = !(getenv("P2V_PPFIRST") && atoi(getenv("P2V_PPFIRST")) == 0);  // transcript-first host schedule, see below
  std::vector<JobState> js(jobs.size());
  auto reject = [&](Job &j, int code, const std::string &msg) { j.rc = code; j.err = msg; };
  size_t total_chunks = 0;
  for (size_t g = 0; g < jobs.size(); g++) {
    Job &j = jobs[g];
    JobState &t = js[g];
    if (j.n == 0) continue;
    if (!j.cir || !j.blobs) { reject(j, P2V_E_INVALID, "NULL argument"); continue; }
    if (j.cir->ctx != ctx) { reject(j, P2V_E_INVALID, "circuit belongs to another context"); continue; }
    t.blob_words = (size_t)j.cir->dev.L.blob_words;
    if ((t.blob_words + 31) / 32 > 65535) { reject(j, P2V_E_UNSUPPORTED, "proof blob longer than 2^21 words (K0 tiles words along grid.y)"); continue; }
    t.src_dev = p2v_is_device_ptr(j.blobs);
    size_t chunk = ctx->chunk;
    if (chunk == 0) {
      // host input: 0.25 GiB with the transcript-first schedule (measured at 10^5 proofs, tools/e2e_chunk_tail.sh: 1056-proof chunks
      // 373 k proofs/s, 2112: 425 k, 3200: 422 k, 4256: 419 k, 6400: 412 k), 0.5 GiB without it
      size_t budget = ctx->pipeline > 1 ? (t.src_dev ? ((size_t)3 << 30) : ((size_t)(pp_first_on ? 256 : 512) << 20)) : (t.src_dev ? ((size_t)16 << 30) : ((size_t)2 << 30));
      chunk = budget / (t.blob_words * 8);
      if (chunk < 1024) chunk = 1024;
    }
    chunk = (chunk + 31) / 32 * 32;
    if (chunk > j.n) chunk = (j.n + 31) / 32 * 32;
    t.chunk = chunk;
    total_chunks += (j.n + chunk - 1) / chunk;
  }
  auto live = [&](size_t g) { return jobs[g].rc == P2V_OK && jobs[g].n > 0; };
  auto first_error = [&]() {
    for (auto &j : jobs)
      if (j.rc != P2V_OK) return p2v_fail(ctx, j.rc, j.err);
    return (int)P2V_OK;
  };
  if (total_chunks == 0) return first_error();
  const int depth = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->pipeline, total_chunks));
  // ---- workspaces: one per lane, sized for the largest job ---------------------------------------------------------------
  size_t ws_bytes = 0, stage_bytes = 0;
  for (size_t g = 0; g < jobs.size(); g++) {
    if (!live(g)) continue;
    const DevCircuit &d = jobs[g].cir->dev;
    ws_bytes = std::max(ws_bytes, carve(d, js[g].chunk, nullptr, nullptr, jobs[g].out.folded != nullptr, jobs[g].out.roots != nullptr));
    if (!js[g].src_dev) stage_bytes = std::max(stage_bytes, js[g].chunk * js[g].blob_words * 8);
  }
  int rc;
  for (int k = 0; k < depth; k++)
    if ((rc = ensureWorkspace(ctx, k, ws_bytes))) return rc;
  const int nstage_bufs = depth + 1;  // one per lane in flight + one being filled
  if (stage_bytes && (rc = ensureStage(ctx, stage_bytes, nstage_bufs))) return rc;
  // Host input on the pipeline: transcript-first schedule (P2V_PPFIRST=0 switches it off).  The per-proof parts (5.5% of a
  // blob at the standard shape) of a whole WINDOW of chunks are copied first with one strided copy, the query parts follow
  // chunk by chunk (strided as well, so every byte still crosses PCIe once).  K0/K4/K5 of a chunk — a latency-bound chain of
  // 114 dependent permutations per proof, ~10 ms next to the Merkle blocks of other chunks — then run long before the chunk's
  // query parts arrive, and what is left after the LAST copy is the Merkle work of one short chunk instead of that chain.
  const size_t win_bytes = getenv("P2V_PPFIRST_BYTES") ? (size_t)atoll(getenv("P2V_PPFIRST_BYTES")) : ((size_t)1 << 30);  // test hook: small windows
  size_t pp_bytes = 0;
  if (pp_first_on && depth >= 2)
    for (size_t g = 0; g < jobs.size(); g++) {
      if (!live(g) || js[g].src_dev) continue;
      const size_t row = (size_t)jobs[g].cir->dev.L.proof_words * 8;
      size_t w = std::max<size_t>(1, win_bytes / row / js[g].chunk) * js[g].chunk;  // ~1 GiB of per-proof parts, whole chunks
      js[g].window = std::min(w, (jobs[g].n + 31) / 32 * 32);
      pp_bytes = std::max(pp_bytes, js[g].window * row);
    }
  if (pp_bytes && (rc = ensurePpStage(ctx, pp_bytes))) return rc;
  // ---- outputs that may live on the host (temporaries are allocated in stream order on the primary stream) ---------------
  for (size_t g = 0; g < jobs.size(); g++) {
    if (!live(g)) continue;
    const DevCircuit &d = jobs[g].cir->dev;
    const Outputs &out = jobs[g].out;
    JobState &t = js[g];
    const size_t n = jobs[g].n;
    if ((rc = t.i_ch.init(ctx, out.challenges_in, (size_t)d.ch_words * n * 8))) return rc;
    if ((rc = t.o_roots.init(ctx, out.roots, (size_t)4 * (4 + d.nsteps) * n * d.Q * 8))) return rc;
    if ((rc = t.o_ch.init(ctx, out.challenges, (size_t)d.ch_words * n * 8))) return rc;
    if ((rc = t.o_comb.init(ctx, out.combined, (size_t)2 * d.r * n * 8))) return rc;
    if ((rc = t.o_eq.init(ctx, out.eqmask, n))) return rc;
    if ((rc = t.o_status.init(ctx, out.status, n * 4))) return rc;
    if ((rc = t.o_bits.init(ctx, out.accept_bits, (n + 31) / 32 * 4))) return rc;
    if ((rc = t.o_qs.init(ctx, out.qstatus, n * d.Q * 4))) return rc;
    if ((rc = t.o_folded.init(ctx, out.folded, (size_t)2 * n * d.Q * 8))) return rc;
  }
  // Declared after the DevOut temporaries (js), so it is destroyed BEFORE them: any early return below (a failed launch or
  // CUDA call in the chunk loop) first waits for the lanes that were already forked — their kernels may still be writing
  // the temporaries that the DevOut destructors hand back to the pool on the primary stream.
  struct LaneGuard {
    p2v_ctx *ctx;
    bool armed = false;
    ~LaneGuard() {
      if (!armed) return;
      for (int i = 0; i < P2V_MAX_DEPTH; i++) cudaStreamSynchronize(ctx->side_stream[i]);
      for (int i = 1; i < P2V_MAX_DEPTH; i++) cudaStreamSynchronize(ctx->lane_stream[i]);
      cudaStreamSynchronize(ctx->copy_stream);
      cudaStreamSynchronize(ctx->stream);
      cudaGetLastError();
    }
  } lane_guard{ctx};
  cudaStream_t streams[P2V_MAX_DEPTH];
  streams[0] = ctx->stream;
  for (int i = 1; i < P2V_MAX_DEPTH; i++) streams[i] = ctx->lane_stream[i];
  if (depth >= 2) {
    // fork: the other streams start after everything already queued on the primary one
    P2V_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));
    lane_guard.armed = true;
    for (int i = 1; i < depth; i++) P2V_CUDA(ctx, cudaStreamWaitEvent(streams[i], ctx->fork_ev, 0));
  }
  if (stage_bytes) lane_guard.armed = true;  // the copy stream runs ahead of the primary one as well
  const bool timed = depth == 1;
  static const int force_split = getenv("P2V_SPLIT") ? atoi(getenv("P2V_SPLIT")) : -1;
  static const bool trace_on = getenv("P2V_TRACE") != nullptr;
  // Launch priorities (P2V_PRIO: 0 = none (default), 1 = K0/K4/K5 first, 2 = also the closing kernels of a chunk before younger
  // chunks' leaf blocks).  Measured at 10^5 proofs end to end: +1.3% with the chunk-by-chunk schedule (404 k -> 409 k proofs/s:
  // K4's chain runs 11 ms instead of 22 ms next to the Merkle blocks), nothing with the transcript-first schedule, where K4 is
  // off the critical path anyway; halving the last chunks (tail taper) changed nothing in either schedule and was removed.
  static const int prio_mode = getenv("P2V_PRIO") ? atoi(getenv("P2V_PRIO")) : 0;
  const int prio_head = prio_mode >= 1 ? ctx->prio_hi : 0;
  const int prio_close = prio_mode >= 2 ? std::min(ctx->prio_lo, ctx->prio_hi + 1) : 0;
  const int prio_bulk = prio_mode >= 1 ? ctx->prio_lo : 0;
  std::vector<TracePoint> trace;
  auto mark = [&](const char *what_, int chunk_i, size_t m_i, cudaStream_t s) {
    if (!trace_on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    trace.push_back({what_, chunk_i, m_i, e});
  };
  double host_setup = host_ms();
  mark("begin", -1, 0, ctx->stream);
  // ---- chunk schedule ----------------------------------------------------------------------------------------------------
  // Per job: equal chunks, except for host input with a user-set chunk >= 8192 proofs, which ramps up from chunk/8 by x9/8
  // per chunk — the copy of chunk k+1 must not take longer than the kernels of chunk k (measured with 2 GiB chunks,
  // tools/ramp_sweep.sh: x2 292k, x1.5 296k, x1.25 300k, x1.125 319k proofs/s).  Round 2 tried geometric ramps up to 4 k / 8 k /
  // 16 k / 24 k-proof chunks mirrored at the end, and shorter chunks at the ends of the default schedule: 250 / 258 / 270 / 272 ms
  // against 250 ms for constant 0.5 GiB chunks (the copies run back to back either way; what is left
  // is the work in flight when the last copy ends, about three chunks whatever their size).  Jobs follow each other in the
  // order given; their chunks share the lanes.
  std::vector<Chunk> sched;
  for (size_t g = 0; g < jobs.size(); g++) {
    if (!live(g)) continue;
    const size_t n = jobs[g].n, chunk = js[g].chunk;
    const bool src_dev = js[g].src_dev;
    std::vector<size_t> sizes;
    {
      bool ramped = !src_dev && depth >= 2 && chunk >= 8 * 1024;
      size_t step = ramped ? chunk / P2V_RAMP_START_DIV / 32 * 32 : chunk;
      for (size_t done = 0; done < n;) {
        size_t m = std::min(step, n - done);
        sizes.push_back(m);
        done += m;
        if (ramped) step = std::min(chunk, (step * P2V_RAMP_NUM / P2V_RAMP_DEN + 31) / 32 * 32);
      }
    }
    size_t c0 = 0, win_left = 0;
    for (size_t i = 0; i < sizes.size(); i++) {
      const size_t m = sizes[i];
      Chunk ck{(int)g, c0, m};
      if (js[g].window) {
        if (win_left == 0) {
          // a new window starts here: as many whole chunks as fit (at least this one)
          size_t wn = m;
          for (size_t j = i + 1; j < sizes.size() && wn + sizes[j] <= js[g].window; j++) wn += sizes[j];
          ck.win_n = wn;
          win_left = wn;
        }
        win_left -= m;
      }
      sched.push_back(ck);
      c0 += m;
    }
  }
  int nstaged = 0;  // host-input chunks issued so far: they alternate between the two staging buffers
  bool k0_recorded[P2V_MAX_DEPTH] = {}, lane_recorded[P2V_MAX_DEPTH] = {};
  std::vector<size_t> win_base(jobs.size(), 0);
  for (int k = 0; k < (int)sched.size(); k++) {
    const Chunk &ck = sched[k];
    const Job &job = jobs[ck.job];
    JobState &t = js[ck.job];
    const DevCircuit &d = job.cir->dev;
    const Outputs &out = job.out;
    const size_t m = ck.m, c0 = ck.c0, n = job.n, blob_words = t.blob_words;
    const bool src_dev = t.src_dev;
    const u64 *src = job.blobs + c0 * blob_words;
    int lane = k % depth;       // stream + workspace of this chunk
    cudaStream_t st = streams[lane];
    Workspace ws;
    carve(d, m, (char *)(lane == 0 ? ctx->ws : ctx->lane_ws[lane]), &ws, out.folded != nullptr, out.roots != nullptr);
    ws.ch_in = t.i_ch.as<u64>(); ws.ch_in_n = n; ws.ch_in_off = c0;
    int b = -1;
    const bool pp_first = !src_dev && t.window > 0;
    const size_t pw = (size_t)d.L.proof_words;
    if (pp_first && ck.win_n) {
      // the per-proof parts of the whole window, ahead of its query parts; the buffer is reused, so every K0 of the previous
      // window (the ones that can still be pending are the last one of each lane) must have read it
      for (int i = 0; i < depth; i++)
        if (k0_recorded[i]) P2V_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->k0_done[i], 0));
      mark("pp_copy_start", k, ck.win_n, ctx->copy_stream);
      P2V_CUDA(ctx, cudaMemcpy2DAsync(ctx->pp_stage, pw * 8, src, blob_words * 8, pw * 8, ck.win_n, cudaMemcpyHostToDevice, ctx->copy_stream));
      P2V_CUDA(ctx, cudaEventRecord(ctx->pp_filled, ctx->copy_stream));
      mark("pp_copy_end", k, ck.win_n, ctx->copy_stream);
      win_base[ck.job] = c0;
    }
    ws.aos_pitch = blob_words;
    ws.aos_qoff = d.L.proof_words;
    if (!src_dev) {
      // staging ring: the H2D copy of the next chunks overlaps the kernels of the current ones; buffer b is free again
      // once the last kernel that reads it (the chunk that used it depth + 1 staged chunks ago) has finished
      b = nstaged++ % nstage_bufs;
      P2V_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[b], 0));
      mark("copy_start", k, m, ctx->copy_stream);
      if (pp_first) {
        const size_t qw = blob_words - pw;  // the query parts only, packed
        P2V_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_buf[b], qw * 8, src + pw, blob_words * 8, qw * 8, m, cudaMemcpyHostToDevice, ctx->copy_stream));
        ws.aos_pitch = qw;
        ws.aos_qoff = 0;
      } else {
        P2V_CUDA(ctx, cudaMemcpyAsync(ctx->stage_buf[b], src, m * blob_words * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
      }
      P2V_CUDA(ctx, cudaEventRecord(ctx->stage_filled[b], ctx->copy_stream));
      mark("copy_end", k, m, ctx->copy_stream);
      P2V_CUDA(ctx, cudaStreamWaitEvent(st, ctx->stage_filled[b], 0));
      src = (const u64 *)ctx->stage_buf[b];
    }
    ws.aos = src;
    // K0
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
    const bool split = (what & RUN_FRI) && (force_split >= 0 ? force_split == 1 : !src_dev);
    cudaStream_t side = ((split || pp_first) && depth >= 2) ? ctx->side_stream[lane] : st;
    {
      // only the per-proof part is transposed; the query parts are read in place
      dim3 grid((unsigned)((m + 31) / 32), (unsigned)((d.L.proof_words + 31) / 32));
      if (pp_first) {
        // K0 on the side stream, from the window's buffer: it waits for the window's copy and for the previous chunk of this
        // lane (whose kernels still read the lane's planes), not for this chunk's query parts
        if (lane_recorded[lane]) P2V_CUDA(ctx, cudaStreamWaitEvent(side, ctx->lane_done[lane], 0));
        P2V_CUDA(ctx, cudaStreamWaitEvent(side, ctx->pp_filled, 0));
        const u64 *ppsrc = (const u64 *)ctx->pp_stage + (c0 - win_base[ck.job]) * pw;
        P2V_LAUNCH_PRIO(ctx, side, prio_head, k_stage_transpose, grid, 256, 0, ppsrc, m, d.L.proof_words, d.L.proof_words, d.L.query_words, d.Q, ws.pp,
                        (u64 *)nullptr, d.L.proof_words);
        P2V_CUDA(ctx, cudaEventRecord(ctx->k0_done[lane], side));
        k0_recorded[lane] = true;
      } else {
        P2V_LAUNCH_PRIO(ctx, st, prio_head, k_stage_transpose, grid, 256, 0, src, m, (int)blob_words, d.L.proof_words, d.L.query_words, d.Q, ws.pp,
                        (u64 *)nullptr, d.L.proof_words);
      }
    }
    mark("k0_end", k, m, pp_first ? side : st);
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
    // K4 + K5 on the lane's side stream when the Merkle kernel runs in two phases: the leaf phase needs nothing from the
    // transcript, so the per-proof chains of K4/K5 (latency-bound, ~5 ms per chunk whatever its size) hide beside it.
    // Measured at 10^5 proofs (bench.py): host input, 0.5 GiB chunks of 4.2 k proofs, where the ~5 ms chain of K4 is half of a
    // chunk's Merkle time: 379 k -> 397 k proofs/s end to end with the split; device-resident input, 25 k-proof chunks: the GPU
    // is saturated either way (the step is already within 2% of the sum of all kernels' work) and the second launch costs 1%
    // (436 k -> 431 k) — so the split follows the input.  P2V_SPLIT=0/1 forces it (A/B aid).
    if (side != st && !pp_first) {
      P2V_CUDA(ctx, cudaEventRecord(ctx->staged_ev[lane], st));
      P2V_CUDA(ctx, cudaStreamWaitEvent(side, ctx->staged_ev[lane], 0));
    }
    P2V_LAUNCH_PRIO(ctx, side, prio_head, k_challenges, (unsigned)((m + 127) / 128), 128, 0, d, ws, m);
    if (d.num_lookup_polys > 0) P2V_LAUNCH_PRIO(ctx, side, prio_head, k_lookup_delta_copies, p2v_grid_for(ctx, m, 256, 8), 256, 0, d, ws, m);
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
    // K5
    if (what & RUN_CONSTRAINTS) P2V_LAUNCH_PRIO(ctx, side, prio_head, k_constraints, (unsigned)((m + 127) / 128), 128, 0, d, ws, m);
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
    if (side != st) P2V_CUDA(ctx, cudaEventRecord(ctx->transcript_ev[lane], side));
    mark("k5_end", k, m, side);
    // K6
    if (what & RUN_FRI) {
      size_t items = m * (size_t)d.Q * (4 + d.nsteps);
      // serial mode: persistent grid of 256-thread blocks; pipelined mode: one 128-thread block per 128 openings, so
      // that blocks of all lanes' kernels interleave on the SMs as resources free up
      const bool pipe = depth >= 2;
      unsigned grid = pipe ? (unsigned)((items + P2V_MERKLE_BLOCK_PIPE - 1) / P2V_MERKLE_BLOCK_PIPE)
                           : (unsigned)p2v_grid_for(ctx, items, P2V_MERKLE_BLOCK, P2V_MERKLE_MINBLOCKS);
#define P2V_MERKLE_LAUNCH(PHASE, PRIO)                                                                                                                            \
  do {                                                                                                                                                          \
    if (pipe) P2V_LAUNCH_PRIO(ctx, st, PRIO, (k_fri_merkle<P2V_MERKLE_BLOCK_PIPE, P2V_MERKLE_MINBLOCKS_PIPE, PHASE>), grid, P2V_MERKLE_BLOCK_PIPE, 0, d, ws, m); \
    else P2V_LAUNCH_ON(ctx, st, (k_fri_merkle<P2V_MERKLE_BLOCK, P2V_MERKLE_MINBLOCKS, PHASE>), grid, P2V_MERKLE_BLOCK, 0, d, ws, m);                             \
  } while (0)
      if (split) {
        P2V_MERKLE_LAUNCH(MERKLE_LEAF, prio_bulk);
        if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
        if (side != st) P2V_CUDA(ctx, cudaStreamWaitEvent(st, ctx->transcript_ev[lane], 0));
        P2V_MERKLE_LAUNCH(MERKLE_PATH, prio_close);
      } else {
        if (side != st) P2V_CUDA(ctx, cudaStreamWaitEvent(st, ctx->transcript_ev[lane], 0));
        P2V_MERKLE_LAUNCH(MERKLE_ALL, prio_bulk);
        if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
      }
#undef P2V_MERKLE_LAUNCH
      if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
      P2V_LAUNCH_PRIO(ctx, st, prio_close, k_fri_query, p2v_grid_for(ctx, m * d.Q, 256, 4), 256, 0, d, ws, m);
    } else if (side != st) {
      P2V_CUDA(ctx, cudaStreamWaitEvent(st, ctx->transcript_ev[lane], 0));
    }
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
    // K7
    if (out.verdict_mode) {
      u32 *stp = t.o_status.as<u32>() ? t.o_status.as<u32>() + c0 : nullptr;
      u32 *bits = t.o_bits.as<u32>() ? t.o_bits.as<u32>() + c0 / 32 : nullptr;
      P2V_LAUNCH_PRIO(ctx, st, prio_close, k_verdict, p2v_grid_for(ctx, m, 256, 8), 256, 0, d, ws, m, out.verdict_mode, stp, bits);
    }
    if (timed) P2V_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    if (b >= 0) P2V_CUDA(ctx, cudaEventRecord(ctx->stage_free[b], st));  // K6a/K6b were the last readers of the chunk's blobs
    mark("chunk_end", k, m, st);
    // optional intermediate outputs
    if (t.o_ch.dev) P2V_LAUNCH_ON(ctx, st, k_copy_planes, p2v_grid_for(ctx, m * d.ch_words, 256, 8), 256, 0, ws.ch, m, d.ch_words, t.o_ch.as<u64>(), n, c0);
    if (t.o_comb.dev) P2V_LAUNCH_ON(ctx, st, k_copy_planes, p2v_grid_for(ctx, m * 2 * d.r, 256, 8), 256, 0, ws.comb, m, 2 * d.r, t.o_comb.as<u64>(), n, c0);
    if (t.o_eq.dev) P2V_LAUNCH_ON(ctx, st, k_copy_bytes, p2v_grid_for(ctx, m, 256, 8), 256, 0, ws.eqmask, m, t.o_eq.as<uint8_t>() + c0);
    if (t.o_qs.dev) P2V_LAUNCH_ON(ctx, st, k_copy_qstat, p2v_grid_for(ctx, m * d.Q, 256, 8), 256, 0, ws.qstat, m, d.Q, t.o_qs.as<u32>(), c0);
    if (t.o_roots.dev) P2V_LAUNCH_ON(ctx, st, k_copy_roots, p2v_grid_for(ctx, m * d.Q * 4 * (4 + d.nsteps), 256, 8), 256, 0, ws.roots, m, d.Q, 4 + d.nsteps, t.o_roots.as<u64>(), n, c0);
    if (t.o_folded.dev) P2V_LAUNCH_ON(ctx, st, k_copy_folded, p2v_grid_for(ctx, m * d.Q, 256, 8), 256, 0, ws.folded, m, d.Q, t.o_folded.as<u64>(), n, c0);
    if (pp_bytes) {
      P2V_CUDA(ctx, cudaEventRecord(ctx->lane_done[lane], st));
      lane_recorded[lane] = true;
    }
  }
  if (depth >= 2) {
    // join: whoever orders work after us on the primary stream (D2H below, the caller's events, NCCL) sees all lanes
    for (int i = 1; i < depth; i++) {
      P2V_CUDA(ctx, cudaEventRecord(ctx->lane_join[i], streams[i]));
      P2V_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->lane_join[i], 0));
    }
  }
  lane_guard.armed = false;  // every lane is joined into the primary stream from here on
  double host_issued = host_ms();
  bool any_host = false;
  for (size_t g = 0; g < jobs.size(); g++) {
    if (!live(g)) continue;
    JobState &t = js[g];
    for (DevOut *o : {&t.o_ch, &t.o_comb, &t.o_eq, &t.o_status, &t.o_bits, &t.o_qs, &t.o_folded, &t.o_roots}) {
      if ((rc = o->finish())) return rc;
      any_host = any_host || (o->host != nullptr);
    }
  }
  if (any_host) {
    P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  if (trace_on) {
    double host_synced = host_ms();
    P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    fprintf(stderr, "[p2v trace] host: setup %.3f ms, all chunks issued %.3f ms, outputs synced %.3f ms (%zu jobs, %zu chunks)\n", host_setup,
            host_issued, host_synced, jobs.size(), sched.size());
    for (auto &tp : trace) {
      float ms = 0;
      cudaEventElapsedTime(&ms, trace[0].ev, tp.ev);
      fprintf(stderr, "[p2v trace] %8.3f ms  chunk %3d (%6zu proofs)  %s\n", ms, tp.chunk, tp.m, tp.what);
    }
    for (auto &tp : trace) cudaEventDestroy(tp.ev);
  }
  ctx->last_ms.clear();
  if (timed) ctx->last_ms["_pending"] = 1.0f;
  return first_error();
}

int runBatch(p2v_ctx *ctx, const p2v_circuit *cir, const uint64_t *blobs, size_t n, int what, Outputs out) {
  if (!ctx || !cir || !blobs) return p2v_fail(ctx, P2V_E_INVALID, "NULL argument");
  if (n == 0) return P2V_OK;
  std::vector<Job> jobs(1);
  jobs[0].cir = cir;
  jobs[0].blobs = blobs;
  jobs[0].n = n;
  jobs[0].out = out;
  return runJobs(ctx, jobs, what);
}

}  // namespace

// resolve the per-section timings of the last chunk lazily (needs a finished stream)
static int resolveTimings(p2v_ctx *ctx) {
  if (ctx->last_ms.count("_pending") == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const char *names[5] = {"stage", "challenges", "constraints", "fri", "verdict"};
  ctx->last_ms.clear();
  for (int i = 0; i < 5; i++) {
    float ms = 0;
    P2V_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]));
    ctx->last_ms[names[i]] = ms;
  }
  {  // "fri" = k_fri_merkle + k_fri_query; the Merkle part separately (only valid if FRI ran)
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[6]) == cudaSuccess) ctx->last_ms["fri_merkle"] = ms;
    else cudaGetLastError();
    if (cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[7]) == cudaSuccess) ctx->last_ms["fri_merkle_leaf"] = ms;
    else cudaGetLastError();
  }
  return P2V_OK;
}

extern "C" {

int p2v_ctx_last_ms(p2v_ctx *ctx, const char *section, float *ms) {
  if (!ctx || !section || !ms) return P2V_E_INVALID;
  int rc = resolveTimings(ctx);
  if (rc) return rc;
  auto it = ctx->last_ms.find(section);
  if (it == ctx->last_ms.end()) return p2v_fail(ctx, P2V_E_INVALID, std::string("no timing for section ") + section);
  *ms = it->second;
  return P2V_OK;
}

int p2v_circuit_create(p2v_ctx *ctx, const p2v_shape *shape, const uint64_t *vkey, p2v_circuit **out) {
  if (!ctx || !shape || !vkey || !out) return p2v_fail(ctx, P2V_E_INVALID, "p2v_circuit_create: NULL argument");
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = p2v_shape_check(shape);  // host-side (csrc/host/parse.cpp): ranges + what the kernels implement
  if (rc) return p2v_fail(ctx, rc, p2v_tls_error);
  p2v_layout L;
  if ((rc = p2v_shape_layout(shape, &L))) {
    ctx->err = p2v_tls_error;
    return rc;
  }
  std::unique_ptr<p2v_circuit> c(new p2v_circuit());
  c->ctx = ctx;
  c->shape = *shape;
  c->shape.lut_pairs = nullptr;
  DevCircuit &d = c->dev;
  memset(&d, 0, sizeof d);
  d.L = L;
  d.num_wires = shape->num_wires; d.num_routed = shape->num_routed_wires; d.num_gate_constants = shape->num_gate_constants;
  d.r = shape->num_challenges;
  d.degree_bits = shape->degree_bits; d.rate_bits = shape->rate_bits; d.lde_bits = shape->degree_bits + shape->rate_bits;
  d.cap_height = shape->cap_height; d.pow_bits = shape->pow_bits; d.Q = shape->num_queries; d.nsteps = shape->num_steps;
  d.cum_bits[0] = 0;
  for (int s = 0; s < shape->num_steps; s++) {
    d.arity_bits[s] = shape->step_arity_bits[s];
    d.cum_bits[s + 1] = d.cum_bits[s] + shape->step_arity_bits[s];
    d.inv_arity[s] = hinv((uint64_t)1 << shape->step_arity_bits[s]);
    d.inv_omega[s] = hinv(hroot(shape->step_arity_bits[s]));
  }
  d.final_len = shape->final_poly_len; d.qdf = shape->quotient_degree_factor; d.num_constants = shape->num_constants;
  d.num_pi = shape->num_public_inputs; d.num_pp = shape->num_partial_products; d.num_lookup_polys = shape->num_lookup_polys;
  d.num_lookup_sel = shape->num_lookup_selectors; d.num_gates = shape->num_gates; d.num_groups = shape->num_groups;
  d.num_luts = shape->num_luts;
  memcpy(d.group_start, shape->group_start, sizeof d.group_start);
  memcpy(d.group_end, shape->group_end, sizeof d.group_end);
  memcpy(d.lut_off, shape->lut_off, sizeof d.lut_off);
  memcpy(d.gates, shape->gates, sizeof d.gates);
  int r = d.r;
  d.ch_betas = 0; d.ch_gammas = r; d.ch_alphas = 2 * r; d.ch_deltas = 3 * r;
  d.ch_zeta = 3 * r + (shape->num_lookup_polys > 0 ? 4 * r : 0);
  d.ch_fri_alpha = d.ch_zeta + 2; d.ch_fri_betas = d.ch_fri_alpha + 2; d.ch_pow = d.ch_fri_betas + 2 * d.nsteps;
  d.ch_idx = d.ch_pow + 1; d.ch_words = d.ch_idx + d.Q;
  d.omega = hroot(d.degree_bits);
  d.n_first = L.oracle_width[0] + L.oracle_width[1] + L.oracle_width[2] + L.oracle_width[3];
  // device tables
  size_t n_lut_words = 2 * (size_t)(shape->num_luts ? shape->lut_off[shape->num_luts] : 0);
  size_t words = (size_t)L.vkey_words + P2V_MAX_ROUTED + P2V_MAX_WEIGHTS + n_lut_words + TAB_WORDS;
  std::vector<u64> h(words, 0);
  size_t o_vkey = 0, o_kis = o_vkey + L.vkey_words, o_w = o_kis + P2V_MAX_ROUTED, o_lut = o_w + P2V_MAX_WEIGHTS, o_tab = o_lut + n_lut_words;
  memcpy(&h[o_vkey], vkey, (size_t)L.vkey_words * 8);
  memcpy(&h[o_kis], shape->k_is, sizeof shape->k_is);
  memcpy(&h[o_w], shape->weights, sizeof shape->weights);
  if (n_lut_words) {
    if (!shape->lut_pairs) return p2v_fail(ctx, P2V_E_INVALID, "shape has lookup tables but lut_pairs is NULL");
    memcpy(&h[o_lut], shape->lut_pairs, n_lut_words * 8);
  }
  {
    uint64_t eta = hroot(d.lde_bits), ieta = hinv(eta), g = HGL_MUL_GEN, ig = hinv(g);
    for (int k = 0; k < 32; k++) {
      h[o_tab + TAB_ETA + k] = eta; h[o_tab + TAB_INV_ETA + k] = ieta; h[o_tab + TAB_G + k] = g; h[o_tab + TAB_INV_G + k] = ig;
      eta = hmul(eta, eta); ieta = hmul(ieta, ieta); g = hmul(g, g); ig = hmul(ig, ig);
    }
    for (int a = 1; a <= 8; a++) {  // inverse roots of unity of every possible folding arity
      uint64_t iw = hinv(hroot(a)), x = 1;
      for (int e = 0; e < (1 << a); e++) { h[o_tab + TAB_INVW + (1 << a) + e] = x; x = hmul(x, iw); }
    }
  }
  P2V_CUDA(ctx, cudaMalloc(&c->d_blob, words * 8));
  P2V_CUDA(ctx, cudaMemcpy(c->d_blob, h.data(), words * 8, cudaMemcpyHostToDevice));
  d.vkey = c->d_blob + o_vkey; d.k_is = c->d_blob + o_kis; d.weights = c->d_blob + o_w; d.lut_pairs = c->d_blob + o_lut; d.tab = c->d_blob + o_tab;
  buildTranscript(d);
  if (d.nops > P2V_MAX_TOPS) {
    cudaFree(c->d_blob);
    return p2v_fail(ctx, P2V_E_UNSUPPORTED, "transcript has more than P2V_MAX_TOPS operations");
  }
  *out = c.release();
  return P2V_OK;
}

void p2v_circuit_destroy(p2v_circuit *c) {
  if (!c) return;
  if (c->d_blob) {
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    cudaFree(c->d_blob);
  }
  delete c;
}

int p2v_challenges(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint64_t *challenges_out) {
  Outputs o;
  o.challenges = challenges_out;
  return runBatch(ctx, c, blobs, n, RUN_CHALLENGES, o);
}

int p2v_constraints(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint64_t *combined_out, uint8_t *eq_ok_mask) {
  Outputs o;
  o.combined = combined_out;
  o.eqmask = eq_ok_mask;
  return runBatch(ctx, c, blobs, n, RUN_CHALLENGES | RUN_CONSTRAINTS, o);
}

int p2v_fri(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint32_t *status, uint32_t *query_status, uint64_t *folded_out) {
  Outputs o;
  o.status = status;
  o.qstatus = query_status;
  o.folded = folded_out;
  o.verdict_mode = 2;
  return runBatch(ctx, c, blobs, n, RUN_CHALLENGES | RUN_FRI, o);
}

int p2v_verify_intermediates(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, const p2v_intermediates *io) {
  if (!io) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_intermediates: io is NULL");
  Outputs o;
  o.challenges_in = io->challenges_in;
  o.challenges = io->challenges;
  o.combined = io->combined;
  o.eqmask = io->eq_ok_mask;
  o.status = io->status;
  o.accept_bits = io->accept_bits;
  o.qstatus = io->query_status;
  o.folded = io->folded;
  o.roots = io->roots;
  o.verdict_mode = 3;
  return runBatch(ctx, c, blobs, n, RUN_CHALLENGES | RUN_CONSTRAINTS | RUN_FRI, o);
}

int p2v_verify_batch(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint32_t *accept_bits, uint32_t *status) {
  if (!accept_bits) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_batch: accept_bits is NULL");
  Outputs o;
  o.status = status;
  o.accept_bits = accept_bits;
  o.verdict_mode = 3;
  return runBatch(ctx, c, blobs, n, RUN_CHALLENGES | RUN_CONSTRAINTS | RUN_FRI, o);
}

int p2v_stage(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *blobs, size_t n, uint64_t *planes_out) {
  if (!ctx || !c || !blobs || !planes_out) return p2v_fail(ctx, P2V_E_INVALID, "p2v_stage: NULL argument");
  if (c->ctx != ctx) return p2v_fail(ctx, P2V_E_INVALID, "circuit belongs to another context");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  const DevCircuit &d = c->dev;
  const size_t bw = (size_t)d.L.blob_words;
  if ((bw + 31) / 32 > 65535) return p2v_fail(ctx, P2V_E_UNSUPPORTED, "p2v_stage: blob longer than 2^21 words");
  DevIn in;
  DevOut out;
  int rc;
  if ((rc = in.init(ctx, blobs, n * bw * 8))) return rc;
  if ((rc = out.init(ctx, planes_out, n * bw * 8))) return rc;
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((bw + 31) / 32));
  u64 *pp = out.as<u64>();
  P2V_LAUNCH(ctx, k_stage_transpose, grid, 256, 0, in.as<u64>(), n, (int)bw, d.L.proof_words, d.L.query_words, d.Q, pp, pp + (size_t)d.L.proof_words * n, (int)bw);
  if ((rc = out.finish())) return rc;
  if (out.host) P2V_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return P2V_OK;
}

int p2v_verify_groups(p2v_ctx *ctx, size_t n_groups, const p2v_circuit *const *circuits, const uint64_t *const *blobs,
                      const size_t *counts, uint32_t *const *accept_bits, uint32_t *const *status, int32_t *rcs) {
  if (!ctx || (n_groups && (!circuits || !blobs || !counts || !accept_bits))) return p2v_fail(ctx, P2V_E_INVALID, "p2v_verify_groups: NULL argument");
  std::vector<Job> jobs(n_groups);
  for (size_t g = 0; g < n_groups; g++) {
    jobs[g].cir = circuits[g];
    jobs[g].blobs = blobs[g];
    jobs[g].n = counts[g];
    jobs[g].out.status = status ? status[g] : nullptr;
    jobs[g].out.accept_bits = accept_bits[g];
    jobs[g].out.verdict_mode = 3;
    if (counts[g] && !accept_bits[g]) { jobs[g].rc = P2V_E_INVALID; jobs[g].err = "p2v_verify_groups: accept_bits of a non-empty group is NULL"; }
  }
  int rc = runJobs(ctx, jobs, RUN_CHALLENGES | RUN_CONSTRAINTS | RUN_FRI);
  if (rcs)
    for (size_t g = 0; g < n_groups; g++) rcs[g] = jobs[g].rc;
  return rc;
}

int p2v_synth_batch(p2v_ctx *ctx, const p2v_circuit *c, const uint64_t *template_blob, size_t n, const int32_t *tamper_word,
                    const uint64_t *tamper_delta, uint64_t *blobs_out) {
  if (!ctx || !c || !template_blob || !blobs_out) return p2v_fail(ctx, P2V_E_INVALID, "p2v_synth_batch: NULL argument");
  if (!p2v_is_device_ptr(blobs_out)) return p2v_fail(ctx, P2V_E_INVALID, "p2v_synth_batch: blobs_out must be device memory");
  if ((tamper_word == nullptr) != (tamper_delta == nullptr)) return p2v_fail(ctx, P2V_E_INVALID, "p2v_synth_batch: tamper arrays must come together");
  if (n == 0) return P2V_OK;
  P2V_CUDA(ctx, cudaSetDevice(ctx->device));
  int bw = c->dev.L.blob_words;
  DevIn t, tw, td;
  int rc;
  if ((rc = t.init(ctx, template_blob, (size_t)bw * 8))) return rc;
  if ((rc = tw.init(ctx, tamper_word, n * 4))) return rc;
  if ((rc = td.init(ctx, tamper_delta, n * 8))) return rc;
  P2V_LAUNCH(ctx, k_synth, p2v_grid_for(ctx, n * (size_t)bw, 256, 8), 256, 0, t.as<u64>(), bw, n, tw.as<int32_t>(), td.as<u64>(), blobs_out);
  return P2V_OK;
}

}  // extern "C"
