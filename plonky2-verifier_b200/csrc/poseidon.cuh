// Width-12 Poseidon permutation over Goldilocks, one thread per state, state in registers.
//
// Computes exactly `permutation` of the reference (src/Hash/Poseidon.hs:42-101: 4 full + 22
// partial + 4 full rounds, x^7 s-box, MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20)+diag(8,0..)),
// but in Plonky2's fast-partial-round schedule, the same one the reference's PoseidonGate uses
// (src/Gate/Custom/Poseidon.hs:92-100,125-137; tables src/Hash/Constants.hs:27-113).  The two
// schedules are equal as functions (asserted by tests/test_oracle.py on the CPU and by the GPU
// parity tests against the dense-MDS oracle).
//
// Cost model: 8 full rounds x (12 s-boxes x 4 mulmods + 288 small IMADs)
// + 11x11 pre-matrix + 22 partial rounds x (1 s-box + 2x11 wide MACs).  Round constants are
// folded into the accumulators of the preceding linear layer.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

// Round constants of the 8 full rounds, pre-split for use as IMAD.WIDE addends:
//   c_full_rc[r][i][0] = lo32(rc), [1] = hi32(rc)      (r = 0..3 initial, 4..7 final = rounds 26..29)
// plus the fast-partial tables.
struct PoseidonTables {
  u64 rc[30][12];       // all_ROUND_CONSTANTS
  u64 first_rc[12];     // fast_PARTIAL_FIRST_ROUND_CONSTANT
  u64 partial_rc[22];   // fast_PARTIAL_ROUND_CONSTANTS
  u64 vs[22][11];       // fast_PARTIAL_ROUND_VS
  u64 w_hats[22][11];   // fast_PARTIAL_ROUND_W_HATS
  u64 init_mat[11][11]; // fast_PARTIAL_ROUND_INITIAL_MATRIX, row-major as in the source
};

static __constant__ PoseidonTables c_pt = {
    P2V_ALL_ROUND_CONSTANTS, P2V_FAST_PARTIAL_FIRST_RC, P2V_FAST_PARTIAL_RCS,
    P2V_FAST_PARTIAL_VS,     P2V_FAST_PARTIAL_W_HATS,   P2V_FAST_PARTIAL_INIT_MAT};

#define POSEIDON_MDS_ROW                                              \
  { 17u, 15u, 41u, 16u, 2u, 28u, 13u, 13u, 39u, 18u, 34u, 20u }

// x^7 with 2 squarings + 2 multiplications (sbox1, Hash/Poseidon.hs:79-80)
__device__ __forceinline__ u64 poseidon_sbox(u64 x) {
  u64 x2 = gl_sqr(x);
  u64 x3 = gl_mul(x, x2);
  u64 x4 = gl_sqr(x2);
  return gl_mul(x3, x4);
}

// out_i = add_i + sum_j M[i][j] * s_j  with M[i][j] = circ[(j-i) mod 12] + (i==j ? diag[i] : 0)
// (mdsMatrixCoeff, Hash/Constants.hs:24-25).  The state is split into 32-bit halves; each half is
// accumulated in a u64 (row sum <= 264, so < 2^41: no carries), then recombined with one
// reduction per output.  `addlo/addhi` carry the next round's constants.
__device__ __forceinline__ void poseidon_mds(u64 (&s)[12], const u64 *__restrict__ add) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
  u32 lo[12], hi[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    lo[j] = (u32)s[j];
    hi[j] = (u32)(s[j] >> 32);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) {
    u64 a = add ? add[i] : 0;
    u64 L = (u32)a, H = (u32)(a >> 32);
#pragma unroll
    for (int j = 0; j < 12; j++) {
      u32 c = C[(j - i + 12) % 12] + ((i == 0 && j == 0) ? 8u : 0u);
      L += (u64)lo[j] * c;
      H += (u64)hi[j] * c;
    }
    // value = L + 2^32*H,  H = Hh*2^32 + Hl  ->  L + Hh*(2^32-1) + (Hl << 32)
    u32 Hl = (u32)H, Hh = (u32)(H >> 32);
    u64 A = L + (u64)Hh * 0xFFFFFFFFu;  // < 2^42
    u64 r = A + ((u64)Hl << 32);
    if (r < A) r += GL_EPS;  // wrapped r < 2^42: no second wrap
    s[i] = r;
  }
}

__device__ __forceinline__ void poseidon_full_round(u64 (&s)[12], const u64 *__restrict__ next_rc) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(s[i]);
  poseidon_mds(s, next_rc);
}

// The permutation.  Input: lazy u64 (any values); output: lazy u64 (apply gl_canon before use as data).
__device__ __forceinline__ void poseidon_permute(u64 (&s)[12]) {
  // round 0 constants
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], c_pt.rc[0][i]);
  // initial full rounds 0..3; rounds 0..2 fold in the next round's constants, round 3 folds in
  // FAST_PARTIAL_FIRST_ROUND_CONSTANT.
#pragma unroll 1
  for (int r = 0; r < 4; r++) poseidon_full_round(s, r < 3 ? c_pt.rc[r + 1] : c_pt.first_rc);
  // pre-partial matrix: s[1..] <- INITIAL_MATRIX^T-style product (mdsInitPartial,
  // Gate/Custom/Poseidon.hs:121-125: out_i = sum_j INITIAL_MATRIX[j][i] * s_{j+1})
  {
    // register-rotating loop over the input lane j: acc_i += INITIAL_MATRIX[j][i] * s_{j+1};
    // all register indices stay static so nothing falls into local memory.
    u64 acc[11];
#pragma unroll
    for (int i = 0; i < 11; i++) acc[i] = 0;
#pragma unroll 1
    for (int j = 0; j < 11; j++) {
      u64 x = s[1];
#pragma unroll
      for (int i = 1; i < 11; i++) s[i] = s[i + 1];
      s[11] = x;
#pragma unroll
      for (int i = 0; i < 11; i++) acc[i] = gl_add(acc[i], gl_mul(c_pt.init_mat[j][i], x));
    }
#pragma unroll
    for (int i = 0; i < 11; i++) s[i + 1] = acc[i];
  }
  // 22 partial rounds
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    u64 y = poseidon_sbox(s[0]);
    y = gl_add(y, c_pt.partial_rc[r]);  // entry 21 is 0
    // d = 25*y + sum s_{i+1} * W_HAT[r][i];   s_{i+1} += y * VS[r][i]
    u64 d = gl_mul_small(y, 25u);
#pragma unroll
    for (int i = 0; i < 11; i++) {
      d = gl_add(d, gl_mul(s[i + 1], c_pt.w_hats[r][i]));
      s[i + 1] = gl_add(s[i + 1], gl_mul(y, c_pt.vs[r][i]));
    }
    s[0] = d;
  }
  // final full rounds 26..29
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], c_pt.rc[26][i]);
#pragma unroll 1
  for (int r = 0; r < 4; r++) poseidon_full_round(s, r < 3 ? c_pt.rc[27 + r] : nullptr);
}
