// K0 (stage/transpose), K4 (Fiat-Shamir challenger), K6 (FRI query rounds), K7 (verdict).
//
// Data layout in HBM (DESIGN.md "data layout"): a chunk of n proofs arrives as AoS blobs
// [n][blob_words] (include/p2v.h p2v_layout).  K0 transposes the PER-PROOF part (5.5% of a blob at S12) into word planes
//   pp[w][n]                 w < proof_words
// so that the one-thread-per-proof kernels (K4, K5, the tails of K6a/K6b) read plane[w][proof]: a warp touches 256
// contiguous bytes per word.  The PER-QUERY parts stay where they are (round 2): a thread (q, proof) of K6a/K6b consumes
// its OWN contiguous words — 8 per sponge block, 4 per sibling, the whole leaf in combineInitial — so every 32-byte sector
// it touches is used completely whether or not the warp's addresses are adjacent, and the 12 GB transposed copy of round 1
// (written once, read once, 1.0x the batch in workspace) bought nothing.  p2v_stage still produces the full plane layout.  All kernels are shape-generic: the shape lives in a DevCircuit
// passed by value (kernel parameter space, read through the constant cache).
#pragma once
#include "poseidon.cuh"
#include "../../include/p2v.h"

#define P2V_MAX_TOPS 64

// One operation of the transcript (Challenge/Verifier.hs:73-94, Challenge/FRI.hs:73-97)
enum { TOP_ABSORB_PROOF = 0, TOP_ABSORB_VKEY = 1, TOP_ABSORB_PIH = 2, TOP_SPONGE_FINISH = 3, TOP_SQUEEZE = 4 };
struct TOp {
  int32_t kind, off, count, dst;
};

struct DevCircuit {
  p2v_layout L;
  int32_t num_wires, num_routed, num_gate_constants, r;
  int32_t degree_bits, rate_bits, lde_bits, cap_height, pow_bits, Q, nsteps;
  int32_t arity_bits[P2V_MAX_STEPS], cum_bits[P2V_MAX_STEPS + 1];
  int32_t final_len, qdf, num_constants, num_pi, num_pp, num_lookup_polys, num_lookup_sel;
  int32_t num_gates, num_groups, num_luts;
  int32_t group_start[P2V_MAX_GROUPS], group_end[P2V_MAX_GROUPS];
  int32_t lut_off[P2V_MAX_LUTS + 1];
  p2v_gate gates[P2V_MAX_GATES];
  int32_t nops;
  TOp ops[P2V_MAX_TOPS];
  // offsets of the challenge planes (include/p2v.h "Challenges of one proof")
  int32_t ch_betas, ch_gammas, ch_alphas, ch_deltas, ch_zeta, ch_fri_alpha, ch_fri_betas, ch_pow, ch_idx, ch_words;
  int32_t n_first;       // length of the first FRI batch = sum of the four oracle widths (powers of alpha kept per proof)
  // device-resident tables
  const u64 *vkey;       // cap [2^cap_height][4] ++ circuit_digest[4]
  const u64 *k_is;       // [num_routed]
  const u64 *weights;    // barycentric weights
  const u64 *lut_pairs;  // (inp,out) pairs
  const u64 *tab;        // [4][32]: eta^(2^k), eta^-(2^k), g^(2^k), g^-(2^k)   (eta = LDE generator, g = mulGen),
                         // then TAB_INVW: entry 2^a + e = subgroupGenerator(a)^-e for a in 1..8, e < 2^a
  u64 omega;             // subgroupGenerator(degree_bits)
  u64 inv_arity[P2V_MAX_STEPS];  // 1/2^arity_bits
  u64 inv_omega[P2V_MAX_STEPS];  // subgroupGenerator(arity_bits)^-1
};
#define TAB_ETA 0
#define TAB_INV_ETA 32
#define TAB_G 64
#define TAB_INV_G 96
#define TAB_INVW 128
#define TAB_WORDS (128 + 512)

// Per-chunk workspace planes
struct Workspace {
  u64 *pp;     // [proof_words][n]
  const u64 *aos;  // the chunk's query parts as they arrived: query part q of proof p starts at aos[p * aos_pitch + aos_qoff + q * query_words]
  size_t aos_pitch;  // words per row: blob_words (whole blobs, device-resident input) or Q * query_words (host input: query parts only)
  int aos_qoff;      // proof_words or 0
  u64 *ch;     // [ch_words][n]   challenges (canonical)
  u64 *pih;    // [4][n]          sponge(public_inputs)
  u64 *pre;    // [4][n]          precomputed reduced openings Y0, Y1 (Plonk/FRI.hs:128-134)
  u64 *apow;   // [2*n_first][n]  alpha_fri^k, k < n_first (K4 -> K6b: combineInitial as a plain dot product)
  u64 *comb;   // [2r][n]         combined constraints
  u32 *qstat;  // [Q][n]          per-query status
  uint8_t *tree_ok;  // [4+nsteps][Q][n]  Merkle opening verdicts (K6a -> K6b)
  u64 *leafdig;      // [4][(4+nsteps)*Q*n]  leaf digests (K6a leaf phase -> K6a path phase)
  u64 *folded; // [2][Q*n]        final folded evaluation per query (debug/parity output)
  uint8_t *eqmask;  // [n]        bit j: round j of the quotient identity holds
  u64 *roots;  // [4][(4+nsteps)*Q*n]  recomputed Merkle roots of every opening (debug/parity output, else NULL)
  const u64 *ch_in;  // test hook: SoA [ch_words][ch_in_n] challenges that replace the transcript's (else NULL)
  size_t ch_in_n, ch_in_off;  // plane length of ch_in and this chunk's first proof in it
};

__device__ __forceinline__ u32 bitrev(u32 x, int bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// base^e for a table of base^(2^k): product over the set bits of e
__device__ __forceinline__ u64 pow_from_table(const u64 *__restrict__ tab, u32 e) {
  u64 acc = 1;
  int k = 0;
  while (e) {
    if (e & 1u) acc = gl_mul(acc, __ldg(tab + k));
    e >>= 1;
    k++;
  }
  return acc;
}

// sum of 64x64-bit products kept as an exact 192-bit integer; 2^128 = 2^96 * 2^32 = -2^32 (mod p)
struct DotAcc {
  u64 lo = 0, hi = 0;
  u32 top = 0;
  __device__ __forceinline__ void mac(u64 a, u64 x) {
    unsigned __int128 m = (unsigned __int128)a * x;
    asm("add.cc.u64 %0,%0,%3;\n\taddc.cc.u64 %1,%1,%4;\n\taddc.u32 %2,%2,0;" : "+l"(lo), "+l"(hi), "+r"(top) : "l"((u64)m), "l"((u64)(m >> 64)));
  }
  __device__ __forceinline__ u64 reduce() const { return gl_sub(gl_reduce128(lo, hi), (u64)top << 32); }
};

// ---- K0: AoS blobs -> SoA planes -----------------------------------------------------------------
// 32x32 tiles through shared memory: reads are coalesced along the blob (256 B per proof row),
// writes are coalesced along the proof index (256 B per word plane).
// `words` = how many leading words of every blob are transposed: proof_words inside the verifier (qp unused), blob_words for
// p2v_stage.
__global__ void __launch_bounds__(256) k_stage_transpose(const u64 *__restrict__ blobs, size_t n, int blob_words, int proof_words,
                                                         int query_words, int Q, u64 *__restrict__ pp, u64 *__restrict__ qp, int words) {
  __shared__ u64 tile[32][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  // proof tiles run along grid.x (up to 2^31-1 blocks): grid.y is limited to 65535, i.e. ~2.09 M proofs per chunk
  size_t p0 = (size_t)blockIdx.x * 32;
  int w0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    size_t p = p0 + ty + 8 * k;
    int w = w0 + tx;
    if (p < n && w < words) tile[ty + 8 * k][tx] = blobs[p * (size_t)blob_words + w];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int w = w0 + ty + 8 * k;
    size_t p = p0 + tx;
    if (p < n && w < words) {
      u64 v = tile[tx][ty + 8 * k];
      if (w < proof_words) {
        pp[(size_t)w * n + p] = v;
      } else {
        int rel = w - proof_words;
        int q = rel / query_words, wq = rel - q * query_words;
        qp[((size_t)wq * Q + q) * n + p] = v;
      }
    }
  }
}

// Replicate a template blob n times with one tampered word per copy (synthetic batches).
__global__ void k_synth(const u64 *__restrict__ tmpl, int blob_words, size_t n, const int32_t *__restrict__ tamper_word,
                        const u64 *__restrict__ tamper_delta, u64 *__restrict__ out) {
  size_t total = n * (size_t)blob_words;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    size_t p = i / blob_words;
    int w = (int)(i - p * blob_words);
    u64 v = tmpl[w];
    if (tamper_word && tamper_word[p] == w) v = gl_canon(gl_add(gl_canon(v), gl_canon(tamper_delta[p])));
    out[i] = v;
  }
}

// ---- K4: proofChallenges (Challenge/Verifier.hs:58-103, Challenge/FRI.hs:65-104) ---------------
// One thread per proof; the duplex sponge of Challenge/Pure.hs:27-69 as an explicit state machine
// with ONE permutation call site.  Pending inputs / produced outputs live in a tiny local array
// (dynamic index), the Poseidon state stays in registers.
__global__ void __launch_bounds__(128) k_challenges(const __grid_constant__ DevCircuit c, Workspace ws, size_t n) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 *__restrict__ pp = ws.pp;
  u64 s[12];
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  u64 buf[8];          // Absorbing: pending inputs; Squeezing: state[0..7] after the last permutation
  int nbuf = 0;        // Absorbing: number of pending inputs
  int nout = 0;        // Squeezing: outputs left; next output is buf[nout-1]  (reverse . take 8, Pure.hs:41-43)
  bool absorbing = true;
  int pc = 0, sub = 0;  // program counter, position inside the current op
  for (;;) {
    bool need_perm = false;
    // run the transcript until a permutation is required
    while (pc < c.nops && !need_perm) {
      TOp op = c.ops[pc];
      if (op.kind == TOP_SQUEEZE) {
        if (sub >= op.count) { pc++; sub = 0; continue; }
        if (absorbing || nout == 0) { need_perm = true; break; }  // Pure.hs:60-69
        u64 y = gl_canon(buf[nout - 1]);
        nout--;
        int dst = op.dst + sub;
        // query indices: canonical value mod 2^lde_bits (Challenge/FRI.hs:93-97)
        if (dst >= c.ch_idx) y &= ((1ull << c.lde_bits) - 1);
        ws.ch[(size_t)dst * n + p] = y;
        sub++;
      } else if (op.kind == TOP_SPONGE_FINISH) {
        // sponge(public_inputs), Hash/Sponge.hs:26-31: flush the last partial block, keep the digest
        if (absorbing && nbuf > 0) { need_perm = true; break; }
        // state after the last permutation is in buf[0..7] (or zero if there were no inputs)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          u64 d = (absorbing ? 0 : gl_canon(buf[i]));
          ws.pih[(size_t)i * n + p] = d;
        }
        // runDuplex action zeroState (Challenge/Verifier.hs:59)
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = 0;
        absorbing = true;
        nbuf = 0;
        nout = 0;
        pc++;
        sub = 0;
      } else {
        if (sub >= op.count) { pc++; sub = 0; continue; }
        // absorbFelt, Pure.hs:50-58
        if (!absorbing) { absorbing = true; nbuf = 0; }
        if (nbuf == 8) { need_perm = true; break; }
        u64 x;
        if (op.kind == TOP_ABSORB_PROOF) x = pp[(size_t)(op.off + sub) * n + p];
        else if (op.kind == TOP_ABSORB_VKEY) x = __ldg(c.vkey + op.off + sub);
        else x = ws.pih[(size_t)sub * n + p];
        buf[nbuf++] = x;
        sub++;
      }
    }
    if (!need_perm) break;
    if (absorbing) {
      // duplex inp old = permutation (overwrite inp old), Pure.hs:35-39
#pragma unroll
      for (int i = 0; i < 8; i++)
        if (i < nbuf) s[i] = buf[i];
    }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 8; i++) buf[i] = s[i];
    // a permutation triggered by an absorb overflow keeps absorbing; one triggered by a squeeze
    // (or by the sponge finish) switches to Squeezing with 8 fresh outputs
    TOp op = c.ops[pc];
    if (op.kind == TOP_SQUEEZE || op.kind == TOP_SPONGE_FINISH) {
      absorbing = false;
      nout = 8;
    } else {
      nbuf = 0;
    }
  }
  if (ws.ch_in) {
    // test hook (p2v_verify_intermediates): verify against GIVEN challenges, so that branches no honest transcript
    // reaches (zeta = 1 in evalLagrange0, x = zeta in combineInitial) can be compared with the oracle
#pragma unroll 1
    for (int w = 0; w < c.ch_words; w++) {
      u64 v = gl_canon(ws.ch_in[(size_t)w * ws.ch_in_n + ws.ch_in_off + p]);
      if (w >= c.ch_idx) v &= ((1ull << c.lde_bits) - 1);
      ws.ch[(size_t)w * n + p] = v;
    }
  }
  // precomputeReducedOpenings, Plonk/FRI.hs:128-134: Y0 = sum alpha^i batch_this_i, Y1 over batch_next.
  // toFriOpenings order (Challenge/FRI.hs:46-61): constants, sigmas, wires, zs, partial_products,
  // quotient, lookup_zs | zs_next, lookup_zs_next.  Horner from the last element (Goldilocks.hs:180-183).
  {
    gl2 alpha = gl2_make(ws.ch[(size_t)c.ch_fri_alpha * n + p], ws.ch[(size_t)(c.ch_fri_alpha + 1) * n + p]);
    const p2v_layout &L = c.L;
    int seg_off[7] = {L.off_open_constants, L.off_open_sigmas, L.off_open_wires, L.off_open_zs, L.off_open_pp, L.off_open_quotient, L.off_open_lookup_zs};
    int seg_n[7] = {L.n_open_constants, L.n_open_sigmas, L.n_open_wires, L.n_open_zs, L.n_open_pp, L.n_open_quotient, L.n_open_lookup_zs};
    gl2 y0 = gl2_make(0, 0);
#pragma unroll 1
    for (int sg = 6; sg >= 0; sg--) {
#pragma unroll 1
      for (int i = seg_n[sg] - 1; i >= 0; i--) {
        gl2 x = gl2_make(pp[(size_t)(seg_off[sg] + 2 * i) * n + p], pp[(size_t)(seg_off[sg] + 2 * i + 1) * n + p]);
        y0 = gl2_add(x, gl2_mul(alpha, y0));
      }
    }
    int seg2_off[2] = {L.off_open_zs_next, L.off_open_lookup_zs_next};
    int seg2_n[2] = {L.n_open_zs_next, L.n_open_lookup_zs_next};
    gl2 y1 = gl2_make(0, 0);
#pragma unroll 1
    for (int sg = 1; sg >= 0; sg--) {
#pragma unroll 1
      for (int i = seg2_n[sg] - 1; i >= 0; i--) {
        gl2 x = gl2_make(pp[(size_t)(seg2_off[sg] + 2 * i) * n + p], pp[(size_t)(seg2_off[sg] + 2 * i + 1) * n + p]);
        y1 = gl2_add(x, gl2_mul(alpha, y1));
      }
    }
    {  // alpha^k for K6b: G = sum alpha^k col_k costs 2 multiplications per column there instead of Horner's 5
      gl2 ap = gl2_make(1, 0);
#pragma unroll 1
      for (int k = 0; k < c.n_first; k++) {
        ws.apow[(size_t)(2 * k) * n + p] = ap.a;
        ws.apow[(size_t)(2 * k + 1) * n + p] = ap.b;
        ap = gl2_mul(ap, alpha);
      }
    }
    ws.pre[0 * n + p] = gl_canon(y0.a);
    ws.pre[1 * n + p] = gl_canon(y0.b);
    ws.pre[2 * n + p] = gl_canon(y1.a);
    ws.pre[3 * n + p] = gl_canon(y1.b);
  }
}

// ---- K6a: Merkle openings of the FRI query rounds (checkInitialTreeProofs Plonk/FRI.hs:105-117,
// the proofCheckOK of foldingStep :316) -----------------------------------------------------------
// Thread t = (tree, q, proof), tree-major so that a warp works on one tree: trees 0..3 are the
// initial oracles (leaf = row of the oracle, path to the cap), trees 4.. are the commit-phase
// trees (leaf = flattened coset evals).  ~97% of all permutations of a verification run here,
// through a single permutation call site.  Occupancy: the pipes, not latency, are the limit, and ptxas needs
// registers more than the SM needs warps — measured on 5x10^4 S12 proofs (fraction of the integer-pipe roofline):
// 4 blocks of 256 per SM (64 registers, 112 B of spills, 8 warps/scheduler) 79.5%, 3 blocks (80 registers) 82.2%,
// 2 blocks (115 registers, no spills, 4 warps/scheduler) 83.9%.
#ifndef P2V_MERKLE_MINBLOCKS
#define P2V_MERKLE_MINBLOCKS 3
#endif
#ifndef P2V_MERKLE_BLOCK
#define P2V_MERKLE_BLOCK 256
#endif
// Two instantiations: BLOCK = P2V_MERKLE_BLOCK for a kernel that has the GPU to itself (serial mode), and
// P2V_MERKLE_BLOCK_PIPE for the chunk pipeline, where blocks of several lanes' kernels share the SMs: 128-thread
// blocks interleave better there.  Round 1: 5 per SM (96 registers) +1% over 256 x 2.  Round 2, after the equivalent-constant
// partial rounds (10^5 device-resident proofs): 128 x 5 481 k proofs/s, 192 x 4 485 k, 256 x 3 486 k, 128 x 6 (80 registers) 487 k.
#ifndef P2V_MERKLE_BLOCK_PIPE
#define P2V_MERKLE_BLOCK_PIPE 128
#endif
#ifndef P2V_MERKLE_MINBLOCKS_PIPE
#define P2V_MERKLE_MINBLOCKS_PIPE 6
#endif
// PHASE: the kernel runs in two launches per chunk.  The leaf sponges (43% of the permutations at the standard shape)
// need nothing from the transcript, so PHASE 1 starts right after K0 while K4/K5 of the same chunk — one thread per proof,
// a 114-permutation dependent chain, ~5 ms whatever the chunk size — run beside it on the lane's side stream; PHASE 2
// (the sibling compressions, which need the query indices) follows when both are done.  PHASE 0 = both in one launch.
enum { MERKLE_ALL = 0, MERKLE_LEAF = 1, MERKLE_PATH = 2 };
template <int BLOCK, int MINBLOCKS, int PHASE>
__global__ void __launch_bounds__(BLOCK, MINBLOCKS) k_fri_merkle(const __grid_constant__ DevCircuit c, Workspace ws, size_t n) {
  const int Q = c.Q;
  const size_t per_tree = n * (size_t)Q;
  const size_t total = per_tree * (size_t)(4 + c.nsteps);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const u64 *__restrict__ pp = ws.pp;
  const p2v_layout &L = c.L;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int tr = (int)(t / per_tree);
    size_t rem = t - (size_t)tr * per_tree;  // = q*n + proof
    int q = (int)(rem / n);
    size_t p = rem - (size_t)q * n;
    // query part q of proof p, read in place: word w at qbase[w]
    const u64 *__restrict__ qbase = ws.aos + p * ws.aos_pitch + ws.aos_qoff + (size_t)q * L.query_words;
    u32 index = PHASE == MERKLE_LEAF ? 0u : (u32)ws.ch[(size_t)(c.ch_idx + q) * n + p];  // the leaf phase must not touch K4's output
    int leaf_off, width, sib_off, plen, cap_off;
    if (tr < 4) {
      leaf_off = L.q_off_leaf[tr]; width = L.oracle_width[tr]; sib_off = L.q_off_sibs[tr]; plen = L.init_path_len;
      cap_off = tr == 1 ? L.off_wires_cap : tr == 2 ? L.off_zs_pp_cap : L.off_quotient_cap;
    } else {
      int st = tr - 4;
      leaf_off = L.q_off_step_evals[st]; width = 2 << c.arity_bits[st]; sib_off = L.q_off_step_sibs[st]; plen = L.step_path_len[st];
      index >>= c.cum_bits[st + 1];  // query_index_rev newQueryIdx
      cap_off = L.off_commit_caps + st * L.cap_words;
    }
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    int nblk = (width + 7) >> 3;
    int iters = PHASE == MERKLE_LEAF ? nblk : nblk + plen;
    int it0 = PHASE == MERKLE_PATH ? nblk : 0;
    if (PHASE == MERKLE_PATH) {
#pragma unroll
      for (int i = 0; i < 4; i++) s[i] = ws.leafdig[(size_t)i * total + t];
    }
#pragma unroll 1
    for (int it = it0; it < iters; it++) {
      if (PHASE != MERKLE_PATH && (PHASE == MERKLE_LEAF || it < nblk)) {
        // sponge block, Hash/Sponge.hs:26-31 (overwrite the first k lanes)
        int k = width - it * 8;
        const u64 *src = qbase + (leaf_off + it * 8);
#pragma unroll
        for (int i = 0; i < 8; i++)
          if (i < k) s[i] = src[i];
      } else {
        // compress with the sibling, Hash/Merkle.hs:30-37
        const u64 *src = qbase + (sib_off + (it - nblk) * 4);
        bool even = (index & 1u) == 0;
        index >>= 1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          u64 sib = src[i];
          u64 node = s[i];
          s[i] = even ? node : sib;
          s[4 + i] = even ? sib : node;
          s[8 + i] = 0;
        }
      }
      poseidon_permute(s);
    }
    if (PHASE == MERKLE_LEAF) {
#pragma unroll
      for (int i = 0; i < 4; i++) ws.leafdig[(size_t)i * total + t] = s[i];
      continue;
    }
    // compare with cap[index] (Merkle.hs:39-42); cap 0 is the verifier key, the others come with the proof
    bool ok = index < (1u << c.cap_height);
    u32 ci = ok ? index : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      u64 want = tr == 0 ? __ldg(c.vkey + ci * 4 + i) : pp[(size_t)(cap_off + ci * 4 + i) * n + p];
      ok = ok && (gl_canon(s[i]) == gl_canon(want));
    }
    ws.tree_ok[t] = ok ? 1 : 0;
    if (ws.roots) {
#pragma unroll
      for (int i = 0; i < 4; i++) ws.roots[(size_t)i * total + t] = gl_canon(s[i]);
    }
  }
}

// ---- K6b: the arithmetic of a query round (checkQueryRound, Plonk/FRI.hs:380-407): combineInitial,
// coset folding, final-polynomial check; merges the Merkle verdicts of K6a into the per-query
// status, keeping the FIRST failure in reference order: INIT_MERKLE, then per
// step STEP_MERKLE, STEP_EVAL, finally FALSE_FINAL.  Thread t = (q, proof).
__global__ void __launch_bounds__(256) k_fri_query(const __grid_constant__ DevCircuit c, Workspace ws, size_t n) {
  size_t total = n * (size_t)c.Q;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const u64 *__restrict__ pp = ws.pp;
  const p2v_layout &L = c.L;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int q = (int)(t / n);
    size_t p = t - (size_t)q * n;
    // query part q of proof p, read in place (word w at qbase[w]): a thread streams its own leaf values
    const u64 *__restrict__ qbase = ws.aos + p * ws.aos_pitch + ws.aos_qoff + (size_t)q * L.query_words;
    u32 idx = (u32)ws.ch[(size_t)(c.ch_idx + q) * n + p];
    u32 init_bad = 0, step_bad = 0;
#pragma unroll 1
    for (int tr = 0; tr < 4 + c.nsteps; tr++) {
      if (!ws.tree_ok[(size_t)tr * total + t]) {
        if (tr < 4) init_bad |= 1u << tr;
        else step_bad |= 1u << (tr - 4);
      }
    }

    // ---------------- phase 2: combineInitial (Plonk/FRI.hs:151-207) ----------------
    gl2 alpha = gl2_make(ws.ch[(size_t)c.ch_fri_alpha * n + p], ws.ch[(size_t)(c.ch_fri_alpha + 1) * n + p]);
    gl2 zeta = gl2_make(ws.ch[(size_t)c.ch_zeta * n + p], ws.ch[(size_t)(c.ch_zeta + 1) * n + p]);
    int r = c.r;
    int npp = (c.num_routed + c.qdf - 1) / c.qdf;  // divCeil routed qdf
    int w2 = L.oracle_width[2];
    int n_pp = r * npp < w2 ? r * npp : w2;        // splitAt (r*npp)
    // firstBatch = constants ++ witness ++ oracle_pp ++ quotient ++ oracle_lookup:  G0 = sum_k alpha^k col_k with the
    // per-proof powers of K4; the 128-bit products are accumulated unreduced (192 bits) and reduced once.
    const u64 *__restrict__ apw = ws.apow + p;
    gl2 g0, g1;
    int n_second = 0;
    {
      int seg_off[5] = {L.q_off_leaf[0], L.q_off_leaf[1], L.q_off_leaf[2], L.q_off_leaf[3], L.q_off_leaf[2] + n_pp};
      int seg_n[5] = {L.oracle_width[0], L.oracle_width[1], n_pp, L.oracle_width[3], w2 - n_pp};
      DotAcc re, im;
      int k = 0;
#pragma unroll 1
      for (int sg = 0; sg < 5; sg++) {
#pragma unroll 2
        for (int i = 0; i < seg_n[sg]; i++, k++) {
          u64 x = qbase[seg_off[sg] + i];
          re.mac(apw[(size_t)(2 * k) * n], x);
          im.mac(apw[(size_t)(2 * k + 1) * n], x);
        }
      }
      g0 = gl2_make(re.reduce(), im.reduce());
    }
    // secondBatch = take r oracle_pp ++ oracle_lookup
    {
      int take_r = r < n_pp ? r : n_pp;
      int seg_off[2] = {L.q_off_leaf[2], L.q_off_leaf[2] + n_pp};
      int seg_n[2] = {take_r, w2 - n_pp};
      n_second = seg_n[0] + seg_n[1];
      DotAcc re, im;
      int k = 0;
#pragma unroll 1
      for (int sg = 0; sg < 2; sg++) {
#pragma unroll 1
        for (int i = 0; i < seg_n[sg]; i++, k++) {
          u64 x = qbase[seg_off[sg] + i];
          re.mac(apw[(size_t)(2 * k) * n], x);
          im.mac(apw[(size_t)(2 * k + 1) * n], x);
        }
      }
      g1 = gl2_make(re.reduce(), im.reduce());
    }
    gl2 y0 = gl2_make(ws.pre[0 * n + p], ws.pre[1 * n + p]);
    gl2 y1 = gl2_make(ws.pre[2 * n + p], ws.pre[3 * n + p]);
    // point_x = mulGen * eta^rev(idx)
    u64 point_x = gl_mul(GL_MUL_GEN_C, pow_from_table(c.tab + TAB_ETA, bitrev(idx, c.lde_bits)));
    gl2 loc1 = gl2_scale(c.omega, zeta);
    // the two divisions share one inversion (inv 0 = 0 is kept: a zero denominator takes the separate path)
    gl2 den0 = gl2_make(gl_sub(point_x, zeta.a), gl_neg(zeta.b)), den1 = gl2_make(gl_sub(point_x, loc1.a), gl_neg(loc1.b));
    gl2 inv0, inv1;
    if (gl2_eq(den0, gl2_make(0, 0)) || gl2_eq(den1, gl2_make(0, 0))) {
      inv0 = gl2_inv(den0);
      inv1 = gl2_inv(den1);
    } else {
      gl2 ip = gl2_inv(gl2_mul(den0, den1));
      inv0 = gl2_mul(ip, den1);
      inv1 = gl2_mul(ip, den0);
    }
    gl2 one = gl2_mul(gl2_sub(g0, y0), inv0);
    gl2 two = gl2_mul(gl2_sub(g1, y1), inv1);
    gl2 apow = gl2_make(apw[(size_t)(2 * n_second) * n], apw[(size_t)(2 * n_second + 1) * n]);  // n_second <= w2 < n_first
    gl2 eval = gl2_add(gl2_mul(apow, one), two);

    // ---------------- folding steps (Plonk/FRI.hs:306-323) ----------------
    u32 eval_bad = 0;
    u32 qidx = idx;
#pragma unroll 1
    for (int st = 0; st < c.nsteps; st++) {
      int a = c.arity_bits[st], A = 1 << a;
      int bits = c.lde_bits - c.cum_bits[st];
      const u64 *ev = qbase + L.q_off_step_evals[st];
      // evals !! (idx mod arity) == upstream eval  (:317)
      u32 pos = qidx & (u32)(A - 1);
      gl2 opened = gl2_make(ev[2 * pos], ev[2 * pos + 1]);
      if (!gl2_eq(opened, eval)) eval_bad |= 1u << st;
      // coset offset ofs = shift * eta_big^rev(bigLog2, (idx>>a)<<a); here its inverse, from the
      // inverse tables: shift = g^(2^cum), eta_big = eta^(2^cum)   (prepareCoset :248-259)
      u32 start = bitrev((qidx >> a) << a, bits);
      u64 inv_ofs = gl_mul(__ldg(c.tab + TAB_INV_G + c.cum_bits[st]), pow_from_table(c.tab + TAB_INV_ETA + c.cum_bits[st], start));
      // foldCosetWith (:263-279) is the value at beta of the interpolant P through (ofs w^j, v_j), j < A.  Computed
      // as `a` radix-2 folds (the FRI recursion itself): P(x) = Pe(x^2) + x Po(x^2) gives on each pair of opposite
      // points  g(x_j^2) = [ (v_j + v_{j+A/2}) + (beta / x_j)(v_j - v_{j+A/2}) ] / 2  and  P(beta) = g(beta^2);
      // level l uses beta^(2^l), ofs^(2^l), w^(2^l).  The evals are stored bit-reversed, so every level pairs
      // ADJACENT entries and its output is again bit-reversed: one streaming pass with a stack of `a` partial
      // results.  15 combines of 7 multiplications for A = 16 (the closed form with S(t) = prod (1 + t^(2^i)) took
      // 47 per point); the 1/2 of every level is applied once as 1/A.  Same field element, bit for bit.
      gl2 beta = gl2_make(ws.ch[(size_t)(c.ch_fri_betas + 2 * st) * n + p], ws.ch[(size_t)(c.ch_fri_betas + 2 * st + 1) * n + p]);
      gl2 tl[8], stack[8];
      tl[0] = gl2_scale(inv_ofs, beta);
#pragma unroll 1
      for (int l = 1; l < a; l++) tl[l] = gl2_sqr(tl[l - 1]);
      const u64 *__restrict__ winv = c.tab + TAB_INVW + A;
      gl2 acc = gl2_make(0, 0);
#pragma unroll 1
      for (int i = 0; i < A; i++) {
        gl2 cur = gl2_make(ev[2 * i], ev[2 * i + 1]);
        u32 ix = (u32)i;
        int lvl = 0;
        while (ix & 1u) {
          u32 pair = ix >> 1;  // pair number at this level in storage order; its natural index is the bit reversal
          u32 e = bitrev(pair, a - lvl - 1) << lvl;
          gl2 first = stack[lvl];
          gl2 f = e ? gl2_scale(__ldg(winv + e), tl[lvl]) : tl[lvl];
          cur = gl2_add(gl2_add(first, cur), gl2_mul(f, gl2_sub(first, cur)));
          ix >>= 1;
          lvl++;
        }
        if (lvl < a) stack[lvl] = cur;
        acc = cur;
      }
      eval = gl2_scale(c.inv_arity[st], acc);
      qidx >>= a;
    }
    // ---------------- final polynomial (:288-291, 325-327, 404-407) ----------------
    int cum = c.cum_bits[c.nsteps];
    int fbits = c.lde_bits - cum;
    u64 x_final = gl_mul(__ldg(c.tab + TAB_G + cum), pow_from_table(c.tab + TAB_ETA + cum, bitrev(qidx, fbits)));
    gl2 fpe = gl2_make(0, 0);
#pragma unroll 1
    for (int i = c.final_len - 1; i >= 0; i--) {
      gl2 co = gl2_make(pp[(size_t)(L.off_final_poly + 2 * i) * n + p], pp[(size_t)(L.off_final_poly + 2 * i + 1) * n + p]);
      fpe = gl2_add(gl2_scale(x_final, fpe), co);
    }
    bool final_ok = gl2_eq(fpe, eval);

    // ---------------- status in reference order ----------------
    u32 status = P2V_ST_ACCEPT;
    if (init_bad) status = P2V_ST_ERR_INIT_MERKLE | (init_bad << 16);
    else {
      bool decided = false;
#pragma unroll 1
      for (int st = 0; st < c.nsteps && !decided; st++) {
        if (step_bad & (1u << st)) { status = P2V_ST_ERR_STEP_MERKLE | ((u32)st << 16); decided = true; }
        else if (eval_bad & (1u << st)) { status = P2V_ST_ERR_STEP_EVAL | ((u32)st << 16); decided = true; }
      }
      if (!decided && !final_ok) status = P2V_ST_FALSE_FINAL;
    }
    if (status != P2V_ST_ACCEPT) status |= (u32)q << 8;
    ws.qstat[t] = status;
    if (ws.folded) {
      gl2 ce = gl2_canon(eval);
      ws.folded[t] = ce.a;
      ws.folded[total + t] = ce.b;
    }
  }
}

// ---- K7: verdict (verifyProof, Plonk/Verifier.hs:56-65; checkFRIProof `ok`, Plonk/FRI.hs:370-372) --
// mode bit 0: include the quotient-identity check; bit 1: include the FRI part.
__global__ void __launch_bounds__(256) k_verdict(const __grid_constant__ DevCircuit c, Workspace ws, size_t n, int mode,
                                                 u32 *__restrict__ status_out, u32 *__restrict__ accept_bits) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n_round = (n + 31) / 32 * 32;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
    bool live = t < n;
    size_t p = live ? t : n - 1;
    u32 status = P2V_ST_ACCEPT;
    bool decided = false;
    if (mode & 1) {
      u32 ok = ws.eqmask[p];
      u32 all = (1u << c.r) - 1;
      if ((ok & all) != all) { status = P2V_ST_FALSE_EQS | ((~ok & all) << 16); decided = true; }
    }
    if (!decided && (mode & 2)) {
      // checkProofOfWork, Plonk/FRI.hs:212-216: the top pow_bits of the canonical response are zero
      u64 resp = ws.ch[(size_t)c.ch_pow * n + p];
      bool pow_ok = c.pow_bits == 0 || (resp >> (64 - c.pow_bits)) == 0;
      if (!pow_ok) { status = P2V_ST_FALSE_POW; decided = true; }
      for (int q = 0; q < c.Q && !decided; q++) {
        u32 qs = ws.qstat[(size_t)q * n + p];
        if (qs != P2V_ST_ACCEPT) { status = qs; decided = true; }
      }
    }
    if (live && status_out) status_out[p] = status;
    u32 ballot = __ballot_sync(0xffffffffu, live && status == P2V_ST_ACCEPT);
    if (accept_bits && (threadIdx.x & 31) == 0) accept_bits[t / 32] = ballot;
  }
}
