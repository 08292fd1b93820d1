// Width-12 Poseidon permutation over Goldilocks, one thread per state, state in registers.
//
// Computes exactly `permutation` of the reference (src/Hash/Poseidon.hs:42-101): 30 rounds
// (4 full + 22 partial + 4 full), each = add round constants, x^7 s-box (all 12 lanes in a full
// round, lane 0 in a partial round), then the dense MDS layer
//   M[i][j] = circ[(j-i) mod 12] + (i==j ? diag[i] : 0),  circ = (17,15,41,16,2,28,13,13,39,18,34,20), diag = (8,0,...)
// (mdsMatrixCoeff, src/Hash/Constants.hs:21-25; linearDiffusion, Hash/Poseidon.hs:100-101).
//
// Why the DENSE layer in every round and not Plonky2's fast-partial-round tables: measured on B200
// (p2v_int_pipe_peak), IMAD.WIDE.U32 issues every 4 cycles per SM sub-partition, a 32-bit IMAD every 2 (like an
// ALU op), a DFMA every 2 on a pipe of its own.  A 64x64 mulmod needs 4 IMAD.WIDE + ~13 other instructions, so the
// 22 mulmods per round of the "fast" partial form cost far more FMA-pipe time than a dense layer made of
// small-constant multiply-adds that never need a wide multiply.  Several generations of that layer live here, all
// bit-exact (tests/test_gpu_hash.py), selected by POSEIDON_MDS_F64:
//   0  poseidon_mds        22/22/20-bit limbs, 3 x 144 carry-free 32-bit IMADs
//   1  poseidon_mds_f64    32-bit halves as doubles, 2 x 144 exact DFMAs (the FP64 pipe is otherwise idle)
//   2  poseidon_mds_mixed  low 43 bits: 144 DFMAs, high 21 bits: 144 IMADs (two pipes)
//   3  poseidon_mds_crt    the same split after a CRT step on the circulant: 2 x 72 multiply-adds (round-1 default)
//   4  poseidon_mds_crt64  (default) CRT layer entirely on the FP64 pipe, 32/32 split, halved coefficients, mask-free fold
// The fast-partial tables are still used by the PoseidonGate constraint program (constraints.cuh), as in the
// reference.
//
// Round constants are folded into the accumulators of the preceding linear layer.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

#ifndef POSEIDON_SBOX_GROUP
/* s-boxes per iteration of the register-rotating loop of a full round: 3, 4, 6 or 12 (= no loop, no rotation moves).
 * With the 80-register budget of the Merkle kernel the fully unrolled form measured best (12: 82.3% of the roofline,
 * 6: 82.0%, 4: 81.0%); at 64 registers all of them were equal. */
#define POSEIDON_SBOX_GROUP 12
#endif

struct PoseidonTables {
  u64 rc[30][12];        // all_ROUND_CONSTANTS
  u64 first_rc[12];      // fast_PARTIAL_FIRST_ROUND_CONSTANT   } used by the PoseidonGate
  u64 partial_rc[22];    // fast_PARTIAL_ROUND_CONSTANTS        } constraint program only
  u64 vs[22][11];        // fast_PARTIAL_ROUND_VS
  u64 w_hats[22][11];    // fast_PARTIAL_ROUND_W_HATS
  u64 init_mat[11][11];  // fast_PARTIAL_ROUND_INITIAL_MATRIX, row-major as in the source
};

static __constant__ PoseidonTables c_pt = {
    P2V_ALL_ROUND_CONSTANTS, P2V_FAST_PARTIAL_FIRST_RC, P2V_FAST_PARTIAL_RCS,
    P2V_FAST_PARTIAL_VS,     P2V_FAST_PARTIAL_W_HATS,   P2V_FAST_PARTIAL_INIT_MAT};

// Round constants cut into 22/22/20-bit limbs (accumulator seeds of the linear layer); row r holds
// the constants ADDED AFTER the linear layer of round r-1, i.e. rc[r]; row 30 is zero.
struct PoseidonRcLimbs {
  u32 v[31][12][4];
};
constexpr PoseidonRcLimbs poseidon_make_rc_limbs() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  PoseidonRcLimbs t{};
  for (int r = 0; r < 30; r++)
    for (int i = 0; i < 12; i++) {
      t.v[r][i][0] = (u32)(rc[r * 12 + i] & 0x3FFFFFULL);
      t.v[r][i][1] = (u32)((rc[r * 12 + i] >> 22) & 0x3FFFFFULL);
      t.v[r][i][2] = (u32)(rc[r * 12 + i] >> 44);
    }
  return t;
}
static __constant__ PoseidonRcLimbs c_rc3 = poseidon_make_rc_limbs();

#define POSEIDON_MDS_ROW                                              \
  { 17u, 15u, 41u, 16u, 2u, 28u, 13u, 13u, 39u, 18u, 34u, 20u }

// x^7 with 2 squarings + 2 multiplications (sbox1, Hash/Poseidon.hs:79-80)
#ifndef P2V_WHATIF
#define P2V_WHATIF 0 /* analysis builds only (WRONG results, tools/whatif.sh): marginal cost of a component of the permutation */
#endif
__device__ __forceinline__ u64 poseidon_sbox(u64 x) {
#if P2V_WHATIF == 1
  return gl_mul(x, x);  // one mulmod instead of four
#elif P2V_WHATIF == 2
  return x + 1;         // no s-box at all
#endif
  u64 x2 = gl_mul(x, x);
  u64 x3 = gl_mul(x, x2);
  u64 x4 = gl_mul(x2, x2);
  return gl_mul(x3, x4);
}

// s <- rc_next + MDS * s on 22/22/20-bit limbs.  Bounds: limb < 2^22, row sum of the coefficients
// <= 264, seed < 2^22  =>  every accumulator < 2^31.
template <int I, int J>
__device__ __forceinline__ void poseidon_mds_acc(u32 &a0, u32 &a1, u32 &a2, const u32 (&l0)[12], const u32 (&l1)[12], const u32 (&l2)[12]) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
  constexpr u32 c = C[(J - I + 12) % 12] + ((I == 0 && J == 0) ? 8u : 0u);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a0) : "r"(l0[J]), "n"(c));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a1) : "r"(l1[J]), "n"(c));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a2) : "r"(l2[J]), "n"(c));
  if constexpr (J + 1 < 12) poseidon_mds_acc<I, J + 1>(a0, a1, a2, l0, l1, l2);
}
template <int I>
__device__ __forceinline__ void poseidon_mds_row(u64 (&s)[12], const u32 (&l0)[12], const u32 (&l1)[12], const u32 (&l2)[12], int next_round) {
  u32 a0 = c_rc3.v[next_round][I][0], a1 = c_rc3.v[next_round][I][1], a2 = c_rc3.v[next_round][I][2];
  poseidon_mds_acc<I, 0>(a0, a1, a2, l0, l1, l2);
  // V = a0 + a1*2^22 + a2*2^44 (< 2^74) as words w0,w1,w2 (w2 < 2^10), then
  // (w1:w0) + w2*2^64 == (w1 + w2 : w0) - w2 (mod p); a wrap of w1 + w2 adds another 2^64 == 2^32 - 1.
  // The final subtraction cannot borrow out: its high word is >= 1 whenever w2 > 0.
  u32 r0, r1;
  asm("{\n\t.reg .u32 t,u,w2,c;\n\t"
      "shl.b32 t,%3,22;\n\tadd.cc.u32 %0,%2,t;\n\t"
      "shr.u32 t,%3,10;\n\tshl.b32 u,%4,12;\n\taddc.cc.u32 %1,t,u;\n\t"
      "shr.u32 w2,%4,20;\n\taddc.u32 w2,w2,0;\n\t"
      "add.cc.u32 %1,%1,w2;\n\taddc.u32 c,0,0;\n\t"
      "add.u32 w2,w2,c;\n\tadd.u32 %1,%1,c;\n\t"
      "sub.cc.u32 %0,%0,w2;\n\tsubc.u32 %1,%1,0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(a0), "r"(a1), "r"(a2));
  s[I] = ((u64)r1 << 32) | r0;
  if constexpr (I + 1 < 12) poseidon_mds_row<I + 1>(s, l0, l1, l2, next_round);
}
__device__ __forceinline__ void poseidon_mds(u64 (&s)[12], int next_round) {
  u32 l0[12], l1[12], l2[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    u32 xl = (u32)s[j], xh = (u32)(s[j] >> 32);
    l0[j] = xl & 0x3FFFFFu;
    l1[j] = __funnelshift_r(xl, xh, 22) & 0x3FFFFFu;
    l2[j] = xh >> 12;
  }
  poseidon_mds_row<0>(s, l0, l1, l2, next_round);
}

// ---- alternative linear layer on the FP64 pipe ---------------------------------------------------------------
// B200 has a full-rate FP64 pipe (64 DFMA/clk/SM) that integer code leaves idle.  The MDS layer is a sum of
// small-constant products, and a double holds integers < 2^53 exactly, so the layer can run there while the
// FMA pipe (IMAD / IMAD.WIDE of the s-boxes) and the ALU pipe do the rest:
//   lo_j, hi_j = the 32-bit halves of s_j as doubles (exact: u32 -> double by the 2^52 trick)
//   L_i = rc_lo + sum_j M[i][j]*lo_j   (< 2^41),   H_i likewise          288 DFMAs per layer, all exact
//   back to integers with the same trick (mantissa of x + 2^52), then value = L + 2^32*H reduced once:
//   L = Ll + 2^32*Lh, H = Hl + 2^32*Hh (Lh,Hh < 2^10):  w0 = Ll, w1 = Lh + Hl, w2 = Hh + carry,
//   (w1:w0) + w2*2^64 == (w1 + w2 : w0) - w2  (mod p), a wrap of w1 + w2 adds another 2^32 - 1.
struct PoseidonRcF64 {
  double v[31][12][2];
};
constexpr PoseidonRcF64 poseidon_make_rc_f64() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  PoseidonRcF64 t{};
  for (int r = 0; r < 30; r++)
    for (int i = 0; i < 12; i++) {
      t.v[r][i][0] = (double)(rc[r * 12 + i] & 0xFFFFFFFFULL);
      t.v[r][i][1] = (double)(rc[r * 12 + i] >> 32);
    }
  return t;
}
static __constant__ PoseidonRcF64 c_rcd = poseidon_make_rc_f64();

#define P2V_TWO52 4503599627370496.0
__device__ __forceinline__ double poseidon_u32_to_f64(u32 x) { return __hiloint2double(0x43300000, (int)x) - P2V_TWO52; }

template <int I, int J>
__device__ __forceinline__ void poseidon_mds_acc_f64(double &L, double &H, const double (&lo)[12], const double (&hi)[12]) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
  constexpr double c = (double)(C[(J - I + 12) % 12] + ((I == 0 && J == 0) ? 8u : 0u));
  L = fma(lo[J], c, L);
  H = fma(hi[J], c, H);
  if constexpr (J + 1 < 12) poseidon_mds_acc_f64<I, J + 1>(L, H, lo, hi);
}
template <int I>
__device__ __forceinline__ void poseidon_mds_row_f64(u64 (&s)[12], const double (&lo)[12], const double (&hi)[12], int next_round) {
  double L = c_rcd.v[next_round][I][0], H = c_rcd.v[next_round][I][1];
  poseidon_mds_acc_f64<I, 0>(L, H, lo, hi);
  L += P2V_TWO52;
  H += P2V_TWO52;
  u32 Ll = (u32)__double2loint(L), Lh = (u32)__double2hiint(L) & 0xFFFFFu;
  u32 Hl = (u32)__double2loint(H), Hh = (u32)__double2hiint(H) & 0xFFFFFu;
  u32 r0, r1;
  asm("{\n\t.reg .u32 w2,c;\n\t"
      "add.cc.u32 %1,%3,%4;\n\taddc.u32 w2,%5,0;\n\t"       // w1 = Lh + Hl, w2 = Hh + carry
      "add.cc.u32 %1,%1,w2;\n\taddc.u32 c,0,0;\n\t"         // w1 += w2, c = wrap
      "add.u32 w2,w2,c;\n\tadd.u32 %1,%1,c;\n\t"
      "sub.cc.u32 %0,%2,w2;\n\tsubc.u32 %1,%1,0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(Ll), "r"(Lh), "r"(Hl), "r"(Hh));
  s[I] = ((u64)r1 << 32) | r0;
  if constexpr (I + 1 < 12) poseidon_mds_row_f64<I + 1>(s, lo, hi, next_round);
}
__device__ __forceinline__ void poseidon_mds_f64(u64 (&s)[12], int next_round) {
  double lo[12], hi[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    lo[j] = poseidon_u32_to_f64((u32)s[j]);
    hi[j] = poseidon_u32_to_f64((u32)(s[j] >> 32));
  }
  poseidon_mds_row_f64<0>(s, lo, hi, next_round);
}

// ---- mixed linear layer: low 43 bits on the FP64 pipe, high 21 bits on 32-bit IMADs -----------------------
// x = lo43 + 2^43*hi21.  L_i = rc_lo43 + sum_j M[i][j]*lo43_j < 264*2^43 + 2^43 < 2^52 is exact in a double and
// still extractable with the 2^52 trick; H_i = rc_hi21 + sum_j M[i][j]*hi21_j < 2^30 fits a 32-bit IMAD chain.
// Same 288 multiply-adds per layer as the pure FP64 form, but split across two pipes.
//   value = L + 2^43*H:  w0 = L[31:0], w1 = L[51:32] + (H << 11)[31:0], w2 = (H >> 21) + carry  (< 2^10)
struct PoseidonRcMixed {
  double lo[31][12];
  u32 hi[31][12];
};
constexpr PoseidonRcMixed poseidon_make_rc_mixed() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  PoseidonRcMixed t{};
  for (int r = 0; r < 30; r++)
    for (int i = 0; i < 12; i++) {
      t.lo[r][i] = (double)(rc[r * 12 + i] & ((1ULL << 43) - 1));
      t.hi[r][i] = (u32)(rc[r * 12 + i] >> 43);
    }
  return t;
}
static __constant__ PoseidonRcMixed c_rcm = poseidon_make_rc_mixed();

// Accumulation order: inputs 1..11 first, input 0 LAST.  In a partial round only lane 0 comes out of an s-box
// (a ~100-cycle dependent chain on the FMA pipe); with this order the 11/12 of the layer that do not depend on
// it (DFMA + IMAD chains) are independent work the scheduler can interleave with that chain.
template <int I, int J>
__device__ __forceinline__ void poseidon_mds_acc_mixed(double &L, u32 &H, const double (&lo)[12], const u32 (&hi)[12]) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
  constexpr u32 ci = C[(J - I + 12) % 12] + ((I == 0 && J == 0) ? 8u : 0u);
  constexpr double c = (double)ci;
  L = fma(lo[J], c, L);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(H) : "r"(hi[J]), "n"(ci));
  if constexpr (J == 0) return;
  else if constexpr (J + 1 < 12) poseidon_mds_acc_mixed<I, J + 1>(L, H, lo, hi);
  else poseidon_mds_acc_mixed<I, 0>(L, H, lo, hi);
}
template <int I>
__device__ __forceinline__ void poseidon_mds_row_mixed(u64 (&s)[12], const double (&lo)[12], const u32 (&hi)[12], int next_round) {  // s = output
  double L = c_rcm.lo[next_round][I];
  u32 H = c_rcm.hi[next_round][I];
  poseidon_mds_acc_mixed<I, 1>(L, H, lo, hi);
  L += P2V_TWO52;
  u32 Ll = (u32)__double2loint(L), Lh = (u32)__double2hiint(L) & 0xFFFFFu;
  u32 r0, r1;
  asm("{\n\t.reg .u32 t,w2,c;\n\t"
      "shl.b32 t,%4,11;\n\tadd.cc.u32 %1,%3,t;\n\t"          // w1 = Lh + (H << 11)
      "shr.u32 w2,%4,21;\n\taddc.u32 w2,w2,0;\n\t"           // w2 = (H >> 21) + carry
      "add.cc.u32 %1,%1,w2;\n\taddc.u32 c,0,0;\n\t"
      "add.u32 w2,w2,c;\n\tadd.u32 %1,%1,c;\n\t"
      "sub.cc.u32 %0,%2,w2;\n\tsubc.u32 %1,%1,0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(Ll), "r"(Lh), "r"(H));
  s[I] = ((u64)r1 << 32) | r0;
  if constexpr (I + 1 < 12) poseidon_mds_row_mixed<I + 1>(s, lo, hi, next_round);
}
__device__ __forceinline__ void poseidon_mds_mixed(u64 (&s)[12], int next_round) {
  double lo[12];
  u32 hi[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    u32 xl = (u32)s[j], xh = (u32)(s[j] >> 32);
    lo[j] = __hiloint2double((int)(0x43300000u | (xh & 0x7FFu)), (int)xl) - P2V_TWO52;  // lo43 as a double, exact
    hi[j] = xh >> 11;
  }
#if POSEIDON_MDS_SEPARATE_OUT
  u64 o[12];
  poseidon_mds_row_mixed<0>(o, lo, hi, next_round);
#pragma unroll
  for (int j = 0; j < 12; j++) s[j] = o[j];
#else
  poseidon_mds_row_mixed<0>(s, lo, hi, next_round);
#endif
}

// ---- CRT-split mixed linear layer (POSEIDON_MDS_F64 == 3) ---------------------------------------------------
// The MDS matrix is circ(c) + 8*E00, and x^12 - 1 = (x^6 - 1)(x^6 + 1): with X+_j = x_j + x_{j+6},
// X-_j = x_j - x_{j+6} (j < 6),  p_d = c_d + c_{d+6} = (30,28,80,34,36,48),  q_d = c_d - c_{d+6} = (4,2,2,-2,-32,8),
// q_{d+6} = -q_d:
//     S_i = sum_{j<6} p_{(j-i) mod 6} X+_j        D_i = sum_{j<6} q_{(j-i) mod 12} X-_j        (i < 6)
//     y_i = (S_i + D_i)/2 + [i == 0] 8 x_0         y_{i+6} = (S_i - D_i)/2
// 2 x 36 multiply-adds per half instead of 144, plus 12 + 12 additions: 25% fewer instructions in the layer.
// Same two-pipe split as the mixed form (low 43 bits: exact DFMAs, |values| < 2^53; high 21 bits: 32-bit IMADs,
// wrap-around signed arithmetic with a true result in [0, 2^31.1)).  The halving is free: it rides on the
// magic-number add for the low part (fma(T, 0.5, 2^52)) and on the shift amounts of the fold for the high part.
// Round constants enter as seeds: S_i <- rc_i + rc_{i+6}, D_i <- rc_i - rc_{i+6}.
#define POSEIDON_MDS_P {30, 28, 80, 34, 36, 48}
#define POSEIDON_MDS_Q {4, 2, 2, -2, -32, 8}
struct PoseidonRcCrt {
  double s_lo[31][6], d_lo[31][6];
  u32 s_hi[31][6], d_hi[31][6];  // d_hi: two's complement
};
constexpr PoseidonRcCrt poseidon_make_rc_crt() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  PoseidonRcCrt t{};
  for (int r = 0; r < 30; r++)
    for (int i = 0; i < 6; i++) {
      u64 a = rc[r * 12 + i], b = rc[r * 12 + i + 6];
      double al = (double)(a & ((1ULL << 43) - 1)), bl = (double)(b & ((1ULL << 43) - 1));
      t.s_lo[r][i] = al + bl;
      t.d_lo[r][i] = al - bl;
      t.s_hi[r][i] = (u32)(a >> 43) + (u32)(b >> 43);
      t.d_hi[r][i] = (u32)(a >> 43) - (u32)(b >> 43);
    }
  return t;
}
static __constant__ PoseidonRcCrt c_rcc = poseidon_make_rc_crt();

// input pairs 1..5 first, pair 0 (the one lane 0 feeds) LAST — see poseidon_mds_acc_mixed
template <int I, int J>
__device__ __forceinline__ void poseidon_crt_acc(double &S, double &D, u32 &Sh, u32 &Dh, const double (&xp)[6], const double (&xm)[6],
                                                 const u32 (&hp)[6], const u32 (&hm)[6]) {
  constexpr int Pc[6] = POSEIDON_MDS_P;
  constexpr int Qc[6] = POSEIDON_MDS_Q;
  constexpr int d = (J - I + 12) % 12;
  constexpr int pc = Pc[d % 6];
  constexpr int qc = d < 6 ? Qc[d] : -Qc[d - 6];
  S = fma(xp[J], (double)pc, S);
  D = fma(xm[J], (double)qc, D);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(Sh) : "r"(hp[J]), "n"(pc));
  asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(Dh) : "r"(hm[J]), "n"(qc));
  if constexpr (J == 0) return;
  else if constexpr (J + 1 < 6) poseidon_crt_acc<I, J + 1>(S, D, Sh, Dh, xp, xm, hp, hm);
  else poseidon_crt_acc<I, 0>(S, D, Sh, Dh, xp, xm, hp, hm);
}
// value = T/2 + 2^43 * (Th/2):  T < 2^53 even (exact double), Th < 2^31.1 even
__device__ __forceinline__ u64 poseidon_crt_fold(double T, u32 Th) {
  double L = fma(T, 0.5, P2V_TWO52);
  u32 Ll = (u32)__double2loint(L), Lh = (u32)__double2hiint(L) & 0xFFFFFu;
  u32 r0, r1;
  asm("{\n\t.reg .u32 t,w2,c;\n\t"
      "shl.b32 t,%4,10;\n\tadd.cc.u32 %1,%3,t;\n\t"          // w1 = Lh + ((Th/2) << 11)
      "shf.l.clamp.b32 w2,%4,0,10;\n\taddc.u32 w2,w2,0;\n\t" // w2 = ((Th/2) >> 21) + carry  (funnel shift: keeps ptxas from fusing the pair into an IMAD.WIDE by 1024)
      "add.cc.u32 %1,%1,w2;\n\taddc.u32 c,0,0;\n\t"
      "add.u32 w2,w2,c;\n\tadd.u32 %1,%1,c;\n\t"
      "sub.cc.u32 %0,%2,w2;\n\tsubc.u32 %1,%1,0;\n\t}"
      : "=&r"(r0), "=&r"(r1)
      : "r"(Ll), "r"(Lh), "r"(Th));
  return ((u64)r1 << 32) | r0;
}
template <int I>
__device__ __forceinline__ void poseidon_crt_row(u64 (&s)[12], const double (&xp)[6], const double (&xm)[6], const u32 (&hp)[6],
                                                 const u32 (&hm)[6], double lo0, u32 hi0, int next_round) {
  double S = c_rcc.s_lo[next_round][I], D = c_rcc.d_lo[next_round][I];
  u32 Sh = c_rcc.s_hi[next_round][I], Dh = c_rcc.d_hi[next_round][I];
  poseidon_crt_acc<I, 1>(S, D, Sh, Dh, xp, xm, hp, hm);
  double T = S + D, U = S - D;
  u32 Th = Sh + Dh, Uh = Sh - Dh;
  if constexpr (I == 0) {  // + 8 x_0 on the diagonal (doubled like everything else before the halving)
    T = fma(lo0, 16.0, T);
    asm("mad.lo.u32 %0, %1, 16, %0;" : "+r"(Th) : "r"(hi0));
  }
  s[I] = poseidon_crt_fold(T, Th);
  s[I + 6] = poseidon_crt_fold(U, Uh);
  if constexpr (I + 1 < 6) poseidon_crt_row<I + 1>(s, xp, xm, hp, hm, lo0, hi0, next_round);
}
__device__ __forceinline__ void poseidon_mds_crt(u64 (&s)[12], int next_round) {
  double lo[12], xp[6], xm[6];
  u32 hi[12], hp[6], hm[6];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    u32 xl = (u32)s[j], xh = (u32)(s[j] >> 32);
    lo[j] = __hiloint2double((int)(0x43300000u | (xh & 0x7FFu)), (int)xl) - P2V_TWO52;
    hi[j] = xh >> 11;
  }
#pragma unroll
  for (int j = 0; j < 6; j++) {
    xp[j] = lo[j] + lo[j + 6];
    xm[j] = lo[j] - lo[j + 6];
    hp[j] = hi[j] + hi[j + 6];
    hm[j] = hi[j] - hi[j + 6];
  }
  poseidon_crt_row<0>(s, xp, xm, hp, hm, lo[0], hi[0], next_round);
}

// ---- CRT layer entirely on the FP64 pipe (POSEIDON_MDS_F64 == 4) ------------------------------------------------
// Round 2.  ncu of the mixed CRT layer: the FMA pipe is the throttle (IMAD.WIDE holds it 4 cycles, the 72 IMADs of the
// high part and the adds ptxas parks there another 280 cycles per layer) while the FP64 pipe idles at 24%.  This form
// takes the integer multiply-adds out of the layer altogether:
//   * 32/32 split, free on a register pair: b = hiloint2double(0x43300000, word) = 2^52 + word (exact);
//   * the CRT butterfly absorbs the bias:  xm_j = b_j - b_{j+6},  xp_j = (b_j - 2^53) + b_{j+6}  (all exact), 18 DADDs
//     per part instead of 12 conversions + 12 butterfly adds;
//   * p_d and q_d are all even, so the halved coefficients p/2 = (15,14,40,17,18,24), q/2 = (2,1,1,-1,-16,4) give
//     y_i = S'_i + D'_i, y_{i+6} = S'_i - D'_i directly (no halving step, one bit more headroom);
//   * the seeds carry the round constants, the 2^52 "magic" that leaves the integer in the mantissa, and a constant that
//     cancels the exponent bits of the raw high words (below), so the result words come out of the register pairs of
//     T = S' + D' and U = S' - D' without a mask:  s0 + d0 = a, s0 - d0 = b with  a = rc_i + c, b = rc_{i+6} + c (mod p)
//     split into 32-bit parts of equal parity (b is moved by multiples of p and by 2^32 between its parts until they
//     match; both stay non-negative);
//   * fold: raw words (Ll, K + Lh), (Hl, K + Hh), K = 0x43300000:  value = Ll + 2^32 (Lh + Hl) + 2^64 Hh =
//     raw - K (2^32 + 2^64); c = -K (2^32 + 2^64) mod p sits in the seeds, so the fold works on the raw words:
//     Z = RL + Hl + RH (< 2^33), result = (Zlo + c : Ll) - (RH + c) with c = Z >> 32: 5 ALU instructions.
// Bounds: word sums <= 272 (2^32 - 1) + 2^34 < 2^41: exact in doubles, high field of T below 2^20.
#ifndef POSEIDON_CVT_I2F
#define POSEIDON_CVT_I2F 1
#endif
#define POSEIDON_MDS_PH {15, 14, 40, 17, 18, 24}
#define POSEIDON_MDS_QH {2, 1, 1, -1, -16, 4}
#define P2V_F64_K 0x43300000u
struct PoseidonRcCrt64 {
  double sl[31][6], dl[31][6], sh[31][6], dh[31][6];
};
__host__ __device__ constexpr u64 poseidon_addmod_c(u64 a, u64 b) {  // a, b < p
  u64 r = a + b;
  return (r < a || r >= GL_P) ? r - GL_P : r;
}
constexpr PoseidonRcCrt64 poseidon_make_rc_crt64() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  constexpr u64 kadj = ((u64)P2V_F64_K << 33) - (u64)P2V_F64_K;  // K 2^32 + K 2^64 = K 2^33 - K (mod p), < p
  constexpr u64 cadj = GL_P - kadj;
  PoseidonRcCrt64 t{};
  for (int r = 0; r < 31; r++)
    for (int i = 0; i < 6; i++) {
      u64 ra = r < 30 ? rc[r * 12 + i] : 0, rb = r < 30 ? rc[r * 12 + i + 6] : 0;
      if (ra >= GL_P) ra -= GL_P;
      if (rb >= GL_P) rb -= GL_P;
      u64 a = poseidon_addmod_c(ra, cadj), b = poseidon_addmod_c(rb, cadj);
      u64 alo = a & 0xFFFFFFFFULL, ahi = a >> 32, blo = b & 0xFFFFFFFFULL, bhi = b >> 32;
      bool flo = ((alo ^ blo) & 1) != 0, fhi = ((ahi ^ bhi) & 1) != 0;
      // value-preserving moves (mod p): m1 = (+1, +2^32-1) adds p; m2 = (+2^32, -1)
      if (flo && fhi) { blo += 1; bhi += 0xFFFFFFFFULL; }
      else if (flo) { blo += 1 + (1ULL << 32); bhi += 0xFFFFFFFEULL; }
      else if (fhi) {
        if (bhi >= 1) { blo += 1ULL << 32; bhi -= 1; }
        else { blo += 2 + (1ULL << 32); bhi += 2 * 0xFFFFFFFFULL - 1; }
      }
      t.sl[r][i] = P2V_TWO52 + (double)((alo + blo) / 2);
      t.dl[r][i] = (double)(((long long)alo - (long long)blo) / 2);
      t.sh[r][i] = P2V_TWO52 + (double)((ahi + bhi) / 2);
      t.dh[r][i] = (double)(((long long)ahi - (long long)bhi) / 2);
    }
  return t;
}
static __constant__ PoseidonRcCrt64 c_rc64 = poseidon_make_rc_crt64();


// contribution of input pair J to rows I..5 (column-major: a pair can be accumulated as soon as its two s-boxes are done)
template <int J, int I>
__device__ __forceinline__ void poseidon_crt64_col(double (&SL)[6], double (&DL)[6], double (&SH)[6], double (&DH)[6], double xpL,
                                                   double xmL, double xpH, double xmH) {
  constexpr int Pc[6] = POSEIDON_MDS_PH;
  constexpr int Qc[6] = POSEIDON_MDS_QH;
  constexpr int d = (J - I + 12) % 12;
  constexpr double pc = (double)Pc[d % 6];
  constexpr double qc = (double)(d < 6 ? Qc[d] : -Qc[d - 6]);
  SL[I] = fma(xpL, pc, SL[I]);
  DL[I] = fma(xmL, qc, DL[I]);
#if P2V_WHATIF == 4
  if (I == 0) { SH[I] = fma(xpH, pc, SH[I]); DH[I] = fma(xmH, qc, DH[I]); }  // 1/6 of the high-part multiply-adds
#else
  SH[I] = fma(xpH, pc, SH[I]);
  DH[I] = fma(xmH, qc, DH[I]);
#endif
  if constexpr (I + 1 < 6) poseidon_crt64_col<J, I + 1>(SL, DL, SH, DH, xpL, xmL, xpH, xmH);
}
// biased butterfly of the pair (x_J, x_{J+6}) and its accumulation
template <int J>
__device__ __forceinline__ void poseidon_crt64_pair(u64 xj, u64 xk, double (&SL)[6], double (&DL)[6], double (&SH)[6], double (&DH)[6]) {
#if POSEIDON_CVT_I2F
  // conversion instruction (I2F.F64.U32) instead of the register-pair trick: no (word, 0x43300000) pairs to assemble
  double bjl = __uint2double_rn((u32)xj), bjh = __uint2double_rn((u32)(xj >> 32));
  double bkl = __uint2double_rn((u32)xk), bkh = __uint2double_rn((u32)(xk >> 32));
  double xmL = bjl - bkl, xmH = bjh - bkh;
  double xpL = bjl + bkl, xpH = bjh + bkh;
#else
  double bjl = __hiloint2double((int)P2V_F64_K, (int)(u32)xj), bjh = __hiloint2double((int)P2V_F64_K, (int)(u32)(xj >> 32));
  double bkl = __hiloint2double((int)P2V_F64_K, (int)(u32)xk), bkh = __hiloint2double((int)P2V_F64_K, (int)(u32)(xk >> 32));
  double xmL = bjl - bkl, xmH = bjh - bkh;
  double xpL = (bjl - 2.0 * P2V_TWO52) + bkl, xpH = (bjh - 2.0 * P2V_TWO52) + bkh;
#endif
  poseidon_crt64_col<J, 0>(SL, DL, SH, DH, xpL, xmL, xpH, xmH);
}
// value = Ll + 2^32 (RL + Hl) + 2^64 RH  =  Ll - RH + 2^32 Z  with  Z = RL + Hl + RH < 2^33  (RL, RH = K + small, K < 2^31):
// Z = Zlo + 2^32 c2 and 2^64 c2 = (2^32 - 1) c2 (mod p) give  value = (Zlo + c2 : Ll) - (RH + c2),  exact in 64 bits (no wrap of
// Zlo + c2: c2 = 1 leaves Zlo < 2K + 2^21; no borrow out: the high word is >= 1 whenever something is subtracted).  Written on
// 64-bit integers so that ptxas uses the three-input IADD3 with both carry outputs: 5 SASS instructions (the carry-chain form
// in PTX, which has no three-input add.cc, took 7).
__device__ __forceinline__ u64 poseidon_crt64_fold(double TL, double TH) {
  u32 Ll = (u32)__double2loint(TL), RL = (u32)__double2hiint(TL), Hl = (u32)__double2loint(TH), RH = (u32)__double2hiint(TH);
#if P2V_WHATIF == 3
  return ((u64)(RL + Hl + RH) << 32) | Ll;  // no modular fold
#endif
  u64 Z = (u64)RL + (u64)Hl + (u64)RH;
  u32 c2 = (u32)(Z >> 32), Zlo = (u32)Z;
  return ((((u64)(Zlo + c2)) << 32) | Ll) - (u64)(RH + c2);
}
__device__ __forceinline__ void poseidon_crt64_seed(double (&SL)[6], double (&DL)[6], double (&SH)[6], double (&DH)[6], int next_round) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    SL[i] = c_rc64.sl[next_round][i];
    DL[i] = c_rc64.dl[next_round][i];
    SH[i] = c_rc64.sh[next_round][i];
    DH[i] = c_rc64.dh[next_round][i];
  }
}
// (l0, h0) = the two parts of the value of lane 0 that entered the layer (for the +8 on the diagonal)
__device__ __forceinline__ void poseidon_crt64_finish_d(u64 (&s)[12], double l0, double h0, const double (&SL)[6], const double (&DL)[6],
                                                        const double (&SH)[6], const double (&DH)[6]) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double TL = SL[i] + DL[i], UL = SL[i] - DL[i], TH = SH[i] + DH[i], UH = SH[i] - DH[i];
    if (i == 0) {
      TL = fma(l0, 8.0, TL);
      TH = fma(h0, 8.0, TH);
    }
    s[i] = poseidon_crt64_fold(TL, TH);
    s[i + 6] = poseidon_crt64_fold(UL, UH);
  }
}
__device__ __forceinline__ void poseidon_crt64_finish(u64 (&s)[12], u64 x0, const double (&SL)[6], const double (&DL)[6],
                                                      const double (&SH)[6], const double (&DH)[6]) {
#if POSEIDON_CVT_I2F
  double l0 = __uint2double_rn((u32)x0), h0 = __uint2double_rn((u32)(x0 >> 32));
#else
  double l0 = __hiloint2double((int)P2V_F64_K, (int)(u32)x0) - P2V_TWO52, h0 = __hiloint2double((int)P2V_F64_K, (int)(u32)(x0 >> 32)) - P2V_TWO52;
#endif
  poseidon_crt64_finish_d(s, l0, h0, SL, DL, SH, DH);
}
// the whole layer; pair 0 (the one lane 0 of a partial round feeds) is accumulated last
__device__ __forceinline__ void poseidon_mds_crt64(u64 (&s)[12], int next_round) {
  double SL[6], DL[6], SH[6], DH[6];
  poseidon_crt64_seed(SL, DL, SH, DH, next_round);
  poseidon_crt64_pair<1>(s[1], s[7], SL, DL, SH, DH);
  poseidon_crt64_pair<2>(s[2], s[8], SL, DL, SH, DH);
  poseidon_crt64_pair<3>(s[3], s[9], SL, DL, SH, DH);
  poseidon_crt64_pair<4>(s[4], s[10], SL, DL, SH, DH);
  poseidon_crt64_pair<5>(s[5], s[11], SL, DL, SH, DH);
  poseidon_crt64_pair<0>(s[0], s[6], SL, DL, SH, DH);
  poseidon_crt64_finish(s, s[0], SL, DL, SH, DH);
}
// ---- second CRT level on the cyclic half (POSEIDON_CRT_LEVEL2) -------------------------------------------------------
// S'_i = sum_j p'_{(j-i) mod 6} X_j is a 6-point CYCLIC convolution, and x^6 - 1 = (x^3 - 1)(x^3 + 1): with
// A_k = X_k + X_{k+3}, B_k = X_k - X_{k+3},  p'_e + p'_{e+3} = (32,32,64),  p'_e - p'_{e+3} = (-2,-4,16), halved again:
//     SS_i = 16 (A_0 + A_1 + A_2) + 16 A_{(i+2) mod 3}            SD_i = sum_k m_{(k-i) mod 6} B_k,  m = (-1,-2,8,1,2,-8)
//     S'_i = SS_i + SD_i,   S'_{i+3} = SS_i - SD_i                 (i < 3)
// The cyclic part collapses to additions and one scale by 16 that rides on the recombining DFMA: 29 FP64 operations
// instead of 36 per part.  Seeds: SD_i starts from g_i, E_i = A_0 + A_1 + A_2 + A_{(i+2)%3} gets e_i (a multiple of 1/16 next to
// 2^48: exact), with 16 e_i + g_i = 2^52 + sigma_i and 16 e_i - g_i = 2^52 + sigma_{i+3}; that needs sigma_i + sigma_{i+3} even, which
// the table builder arranges with value-preserving moves (multiples of 2p, 2^33 between the parts) on a_{i+3}.
#ifndef POSEIDON_CRT_LEVEL2
#define POSEIDON_CRT_LEVEL2 1
#endif
#ifndef POSEIDON_FUSE_SBOX
#define POSEIDON_FUSE_SBOX 1 /* full rounds hand the last product of x^7 to the layer unreduced (in partial rounds it measured 1% slower) */
#endif
#if POSEIDON_CRT_LEVEL2
// Equivalent round constants (POSEIDON_EQUIV_RC).  In a partial round only lane 0 goes through the s-box, so the constants of lanes
// 1..11 commute with it and can be pushed through the linear layer into the next round's constants:
//     M S(u) = M S(u - d) + M d      for d with d_0 = 0            (S = x^7 on lane 0, identity elsewhere)
// Rounds 5..25 are left with a constant on lane 0 only; what has been pushed along arrives, as a full vector, in the constants of
// round 26 (the first of the closing full rounds).  The state between rounds differs from the textbook one, the result of the
// permutation does not.  The point: 18 of the 24 seeds of a partial-round layer become the same in every round, so they are
// compile-time addresses in the constant bank (operands) instead of 18 indexed uniform loads per round.
struct PoseidonEquivRc {
  u64 v[31][12];
};
__host__ __device__ constexpr u64 poseidon_mulsmall_c(u64 x, u32 c) {  // x < p
  u64 r = 0, t = x;
  for (; c; c >>= 1) {
    if (c & 1) r = poseidon_addmod_c(r, t);
    t = poseidon_addmod_c(t, t);
  }
  return r;
}
__host__ __device__ constexpr PoseidonEquivRc poseidon_make_equiv_rc(bool push) {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  constexpr u32 circ[12] = POSEIDON_MDS_ROW;
  PoseidonEquivRc t{};
  u64 d[12] = {};  // what the previous round pushed forward (lane 0 = 0)
  for (int r = 0; r < 30; r++) {
    u64 v[12] = {};
    for (int i = 0; i < 12; i++) {
      u64 x = rc[r * 12 + i];
      if (x >= GL_P) x -= GL_P;
      u64 md = 0;  // (M d)_i
      for (int j = 1; j < 12; j++) md = poseidon_addmod_c(md, poseidon_mulsmall_c(d[j], circ[(j - i + 12) % 12]));
      v[i] = poseidon_addmod_c(x, md);
    }
    const bool strip = push && r >= 5 && r <= 25;
    for (int i = 0; i < 12; i++) {
      t.v[r][i] = (strip && i > 0) ? 0 : v[i];
      d[i] = (strip && i > 0) ? v[i] : 0;
    }
  }
  return t;
}
#ifndef POSEIDON_EQUIV_RC
#define POSEIDON_EQUIV_RC 1
#endif

// one 16-byte aligned row of 24 seeds per round, in the order they are consumed (ptxas fetches them with 128-bit uniform loads):
// [0..5] D' low, [6..11] D' high, [12..14] e low, [15..17] g low, [18..20] e high, [21..23] g high
struct alignas(16) PoseidonRcCrt64L2 {
  double v[31][24];
};
enum { L2_DL = 0, L2_DH = 6, L2_EL = 12, L2_GL = 15, L2_EH = 18, L2_GH = 21 };
constexpr PoseidonRcCrt64L2 poseidon_make_rc_crt64_l2() {
  constexpr PoseidonEquivRc rce = poseidon_make_equiv_rc(POSEIDON_EQUIV_RC != 0);
  constexpr u64 kadj = ((u64)P2V_F64_K << 33) - (u64)P2V_F64_K;
  constexpr u64 cadj = GL_P - kadj;
  PoseidonRcCrt64L2 t{};
  for (int r = 0; r < 31; r++) {
    u64 alo[6] = {}, ahi[6] = {}, blo[6] = {}, bhi[6] = {};
    for (int i = 0; i < 6; i++) {
      u64 ra = rce.v[r][i], rb = rce.v[r][i + 6];  // canonical; row 30 is zero
      u64 a = poseidon_addmod_c(ra, cadj), b = poseidon_addmod_c(rb, cadj);
      alo[i] = a & 0xFFFFFFFFULL; ahi[i] = a >> 32; blo[i] = b & 0xFFFFFFFFULL; bhi[i] = b >> 32;
#if POSEIDON_FUSE_SBOX
      // raw s-box products enter the layer with a low part in (-2^34, 2^32): lift the low parts by 2^42 and take 2^10 off the
      // high parts (same value); a high part below 2^10 is first moved up by p = 1 + 2^32 (2^32 - 1)
      if (ahi[i] < 1024) { alo[i] += 1; ahi[i] += 0xFFFFFFFFULL; }
      if (bhi[i] < 1024) { blo[i] += 1; bhi[i] += 0xFFFFFFFFULL; }
      alo[i] += 1ULL << 42; blo[i] += 1ULL << 42; ahi[i] -= 1024; bhi[i] -= 1024;
#endif
      bool flo = ((alo[i] ^ blo[i]) & 1) != 0, fhi = ((ahi[i] ^ bhi[i]) & 1) != 0;
      if (flo && fhi) { blo[i] += 1; bhi[i] += 0xFFFFFFFFULL; }
      else if (flo) { blo[i] += 1 + (1ULL << 32); bhi[i] += 0xFFFFFFFEULL; }
      else if (fhi) {
        if (bhi[i] >= 1) { blo[i] += 1ULL << 32; bhi[i] -= 1; }
        else { blo[i] += 2 + (1ULL << 32); bhi[i] += 2 * 0xFFFFFFFFULL - 1; }
      }
    }
    for (int i = 0; i < 3; i++) {
      // sigma = (a + b) / 2 per part; make sigma_i + sigma_{i+3} even in both parts by moving a_i (all increments even).  a_i, not
      // a_{i+3}: with equivalent constants only lane 0 differs from round to round, and so must everything derived from it
      bool olo = (((alo[i] + blo[i]) / 2 + (alo[i + 3] + blo[i + 3]) / 2) & 1) != 0;
      bool ohi = (((ahi[i] + bhi[i]) / 2 + (ahi[i + 3] + bhi[i + 3]) / 2) & 1) != 0;
      if (olo && ohi) { alo[i] += 2; ahi[i] += (1ULL << 33) - 2; }                        // + 2p
      else if (olo) { alo[i] += 2 + (1ULL << 33); ahi[i] += (1ULL << 33) - 4; }          // + 2p, 2^33 moved down
      else if (ohi) { alo[i] += 4 + (1ULL << 33); ahi[i] += (1ULL << 34) - 6; }          // both of the above
    }
    for (int i = 0; i < 6; i++) {
      t.v[r][L2_DL + i] = (double)(((long long)alo[i] - (long long)blo[i]) / 2);
      t.v[r][L2_DH + i] = (double)(((long long)ahi[i] - (long long)bhi[i]) / 2);
    }
    for (int i = 0; i < 3; i++) {
      long long sl0 = (long long)((alo[i] + blo[i]) / 2), sl3 = (long long)((alo[i + 3] + blo[i + 3]) / 2);
      long long sh0 = (long long)((ahi[i] + bhi[i]) / 2), sh3 = (long long)((ahi[i + 3] + bhi[i + 3]) / 2);
      t.v[r][L2_EL + i] = P2V_TWO52 / 16.0 + (double)((sl0 + sl3) / 2) / 16.0;
      t.v[r][L2_GL + i] = (double)((sl0 - sl3) / 2);
      t.v[r][L2_EH + i] = P2V_TWO52 / 16.0 + (double)((sh0 + sh3) / 2) / 16.0;
      t.v[r][L2_GH + i] = (double)((sh0 - sh3) / 2);
    }
  }
  return t;
}
static __constant__ PoseidonRcCrt64L2 c_rc64l2 = poseidon_make_rc_crt64_l2();
#if POSEIDON_EQUIV_RC
// rows 5..25 (the layers of rounds 4..24) differ only in the six seeds that lane 0 reaches: D'_0, e_0, g_0 of either part
__host__ __device__ constexpr bool poseidon_l2_seed_varies(int k) { return k == L2_DL || k == L2_DH || k == L2_EL || k == L2_GL || k == L2_EH || k == L2_GH; }
struct alignas(16) PoseidonRcL2Static {
  double v[24];
};
constexpr PoseidonRcL2Static poseidon_make_rc_l2_static() {
  constexpr PoseidonRcCrt64L2 t = poseidon_make_rc_crt64_l2();
  PoseidonRcL2Static r{};
  for (int k = 0; k < 24; k++) r.v[k] = t.v[5][k];
  return r;
}
constexpr bool poseidon_l2_static_ok() {
  constexpr PoseidonRcCrt64L2 t = poseidon_make_rc_crt64_l2();
  for (int r = 5; r <= 25; r++)
    for (int k = 0; k < 24; k++)
      if (!poseidon_l2_seed_varies(k) && t.v[r][k] != t.v[5][k]) return false;
  return true;
}
static_assert(poseidon_l2_static_ok(), "partial-round seeds that lane 0 does not reach must not depend on the round");
static __constant__ PoseidonRcL2Static c_l2s = poseidon_make_rc_l2_static();
#endif

// pair J: conversions, butterfly, the D' half accumulated at once (column-major), X+ kept for the second level
template <int J, int I>
__device__ __forceinline__ void poseidon_l2_dcol(double (&DL)[6], double (&DH)[6], double xmL, double xmH) {
  constexpr int Qc[6] = POSEIDON_MDS_QH;
  constexpr int d = (J - I + 12) % 12;
  constexpr double qc = (double)(d < 6 ? Qc[d] : -Qc[d - 6]);
  DL[I] = fma(xmL, qc, DL[I]);
  DH[I] = fma(xmH, qc, DH[I]);
  if constexpr (I + 1 < 6) poseidon_l2_dcol<J, I + 1>(DL, DH, xmL, xmH);
}
template <int J>
__device__ __forceinline__ void poseidon_l2_pair_d(double bjl, double bjh, double bkl, double bkh, double (&XL)[6], double (&XH)[6], double (&DL)[6],
                                                   double (&DH)[6]) {
  XL[J] = bjl + bkl;
  XH[J] = bjh + bkh;
  poseidon_l2_dcol<J, 0>(DL, DH, bjl - bkl, bjh - bkh);
}
template <int J>
__device__ __forceinline__ void poseidon_l2_pair(u64 xj, u64 xk, double (&XL)[6], double (&XH)[6], double (&DL)[6], double (&DH)[6]) {
  poseidon_l2_pair_d<J>(__uint2double_rn((u32)xj), __uint2double_rn((u32)(xj >> 32)), __uint2double_rn((u32)xk), __uint2double_rn((u32)(xk >> 32)), XL,
                        XH, DL, DH);
}
#if POSEIDON_FUSE_SBOX
// x^7 whose LAST product is handed to the layer unreduced (POSEIDON_FUSE_SBOX): x3 * x4 = w0 + 2^32 w1 + 2^64 w2 + 2^96 w3 with
// 2^64 = 2^32 - 1 and 2^96 = -1 (mod p) is  (w0 - w2 - w3) + 2^32 (w1 + w2): the two parts the layer works on anyway, as exact
// doubles in (-2^34, 2^32) and [0, 2^33).  4 conversions + 3 additions on the XU / FP64 pipes instead of the 10-instruction
// carry chain + 2 conversions; the layer's seeds carry the offset that keeps its low sums non-negative.
// (A reduction built from subtractions only — nothing but the wide multiplies on the FMA pipe, +3 instructions — measured 3.5%
// slower in K6a: what counts is the instruction count, not that pipe.)
__device__ __forceinline__ void poseidon_sbox_raw(u64 x, double &L, double &H) {
  u64 x2 = gl_mul(x, x);
  u64 x3 = gl_mul(x, x2);
  u64 x4 = gl_mul(x2, x2);
  unsigned __int128 m = (unsigned __int128)x3 * x4;
  u64 lo = (u64)m, hi = (u64)(m >> 64);
  double w0 = __uint2double_rn((u32)lo), w1 = __uint2double_rn((u32)(lo >> 32)), w2 = __uint2double_rn((u32)hi), w3 = __uint2double_rn((u32)(hi >> 32));
  L = w0 - (w2 + w3);
  H = w1 + w2;
}
#endif
// S' of one part from the six X+ (second CRT level)
__device__ __forceinline__ void poseidon_l2_s(const double (&X)[6], const double (&e)[3], const double (&g)[3], double (&S)[6]) {
  double A0 = X[0] + X[3], A1 = X[1] + X[4], A2 = X[2] + X[5];
  double B0 = X[0] - X[3], B1 = X[1] - X[4], B2 = X[2] - X[5];
  double sum = (A0 + A1) + A2;
  double E0 = (A2 + e[0]) + sum, E1 = (A0 + e[1]) + sum, E2 = (A1 + e[2]) + sum;
  // SD_i = sum_k m_{(k-i) mod 6} B_k, m = (-1,-2,8,1,2,-8)
  double SD0 = fma(B2, 8.0, fma(B1, -2.0, g[0] - B0));
  double SD1 = fma(B0, -8.0, fma(B2, -2.0, g[1] - B1));
  double SD2 = fma(B1, -8.0, fma(B0, 2.0, g[2] - B2));
  S[0] = fma(E0, 16.0, SD0); S[3] = fma(E0, 16.0, -SD0);
  S[1] = fma(E1, 16.0, SD1); S[4] = fma(E1, 16.0, -SD1);
  S[2] = fma(E2, 16.0, SD2); S[5] = fma(E2, 16.0, -SD2);
}
// ST: a partial round whose layer adds a constant on lane 0 only (rows 5..25 with equivalent constants): all seeds but six come
// from compile-time addresses
template <bool ST, int K>
__device__ __forceinline__ double poseidon_l2_seed(int next_round) {
#if POSEIDON_EQUIV_RC
  if constexpr (ST && !poseidon_l2_seed_varies(K)) return c_l2s.v[K];
#endif
  return c_rc64l2.v[next_round][K];
}
template <bool ST = false>
__device__ __forceinline__ void poseidon_l2_seed_d(double (&DL)[6], double (&DH)[6], int next_round) {
  DL[0] = poseidon_l2_seed<ST, L2_DL + 0>(next_round); DH[0] = poseidon_l2_seed<ST, L2_DH + 0>(next_round);
  DL[1] = poseidon_l2_seed<ST, L2_DL + 1>(next_round); DH[1] = poseidon_l2_seed<ST, L2_DH + 1>(next_round);
  DL[2] = poseidon_l2_seed<ST, L2_DL + 2>(next_round); DH[2] = poseidon_l2_seed<ST, L2_DH + 2>(next_round);
  DL[3] = poseidon_l2_seed<ST, L2_DL + 3>(next_round); DH[3] = poseidon_l2_seed<ST, L2_DH + 3>(next_round);
  DL[4] = poseidon_l2_seed<ST, L2_DL + 4>(next_round); DH[4] = poseidon_l2_seed<ST, L2_DH + 4>(next_round);
  DL[5] = poseidon_l2_seed<ST, L2_DL + 5>(next_round); DH[5] = poseidon_l2_seed<ST, L2_DH + 5>(next_round);
}
template <bool ST = false>
__device__ __forceinline__ void poseidon_l2_sums(const double (&XL)[6], const double (&XH)[6], int next_round, double (&SL)[6], double (&SH)[6]) {
  double el[3], gl[3], eh[3], gh[3];
  el[0] = poseidon_l2_seed<ST, L2_EL + 0>(next_round); el[1] = poseidon_l2_seed<ST, L2_EL + 1>(next_round); el[2] = poseidon_l2_seed<ST, L2_EL + 2>(next_round);
  gl[0] = poseidon_l2_seed<ST, L2_GL + 0>(next_round); gl[1] = poseidon_l2_seed<ST, L2_GL + 1>(next_round); gl[2] = poseidon_l2_seed<ST, L2_GL + 2>(next_round);
  eh[0] = poseidon_l2_seed<ST, L2_EH + 0>(next_round); eh[1] = poseidon_l2_seed<ST, L2_EH + 1>(next_round); eh[2] = poseidon_l2_seed<ST, L2_EH + 2>(next_round);
  gh[0] = poseidon_l2_seed<ST, L2_GH + 0>(next_round); gh[1] = poseidon_l2_seed<ST, L2_GH + 1>(next_round); gh[2] = poseidon_l2_seed<ST, L2_GH + 2>(next_round);
  poseidon_l2_s(XL, el, gl, SL);
  poseidon_l2_s(XH, eh, gh, SH);
}
template <bool ST = false>
__device__ __forceinline__ void poseidon_l2_finish(u64 (&s)[12], u64 x0, const double (&XL)[6], const double (&XH)[6], const double (&DL)[6],
                                                   const double (&DH)[6], int next_round) {
  double SL[6], SH[6];
  poseidon_l2_sums<ST>(XL, XH, next_round, SL, SH);
  poseidon_crt64_finish(s, x0, SL, DL, SH, DH);
}
// lane 0 given as its two parts (fused s-box)
__device__ __forceinline__ void poseidon_l2_finish_d(u64 (&s)[12], double l0, double h0, const double (&XL)[6], const double (&XH)[6],
                                                     const double (&DL)[6], const double (&DH)[6], int next_round) {
  double SL[6], SH[6];
  poseidon_l2_sums(XL, XH, next_round, SL, SH);
  poseidon_crt64_finish_d(s, l0, h0, SL, DL, SH, DH);
}
template <bool ST = false>
__device__ __forceinline__ void poseidon_mds_crt64_l2(u64 (&s)[12], int next_round) {
  double XL[6], XH[6], DL[6], DH[6];
  poseidon_l2_seed_d<ST>(DL, DH, next_round);
  poseidon_l2_pair<1>(s[1], s[7], XL, XH, DL, DH);
  poseidon_l2_pair<2>(s[2], s[8], XL, XH, DL, DH);
  poseidon_l2_pair<3>(s[3], s[9], XL, XH, DL, DH);
  poseidon_l2_pair<4>(s[4], s[10], XL, XH, DL, DH);
  poseidon_l2_pair<5>(s[5], s[11], XL, XH, DL, DH);
  poseidon_l2_pair<0>(s[0], s[6], XL, XH, DL, DH);
  poseidon_l2_finish<ST>(s, s[0], XL, XH, DL, DH, next_round);
}
__device__ __forceinline__ void poseidon_full_round_crt64_l2(u64 (&s)[12], int next_round) {
  double XL[6], XH[6], DL[6], DH[6];
  poseidon_l2_seed_d(DL, DH, next_round);
#if POSEIDON_FUSE_SBOX
  double x0l = 0, x0h = 0;
#define P2V_FR_PAIR(J)                                             \
  {                                                                \
    double al, ah, bl, bh;                                         \
    poseidon_sbox_raw(s[J], al, ah);                               \
    poseidon_sbox_raw(s[J + 6], bl, bh);                           \
    if (J == 0) { x0l = al; x0h = ah; }                            \
    poseidon_l2_pair_d<J>(al, ah, bl, bh, XL, XH, DL, DH);         \
  }
  P2V_FR_PAIR(1) P2V_FR_PAIR(2) P2V_FR_PAIR(3) P2V_FR_PAIR(4) P2V_FR_PAIR(5) P2V_FR_PAIR(0)
#undef P2V_FR_PAIR
  poseidon_l2_finish_d(s, x0l, x0h, XL, XH, DL, DH, next_round);
#else
  u64 x0 = 0;
#define P2V_FR_PAIR(J)                                             \
  {                                                                \
    u64 a = poseidon_sbox(s[J]), b = poseidon_sbox(s[J + 6]);      \
    if (J == 0) x0 = a;                                            \
    poseidon_l2_pair<J>(a, b, XL, XH, DL, DH);                     \
  }
  P2V_FR_PAIR(1) P2V_FR_PAIR(2) P2V_FR_PAIR(3) P2V_FR_PAIR(4) P2V_FR_PAIR(5) P2V_FR_PAIR(0)
#undef P2V_FR_PAIR
  poseidon_l2_finish(s, x0, XL, XH, DL, DH, next_round);
#endif
}
#endif

// ---- two partial rounds as ONE layer (POSEIDON_DOUBLE_ROUNDS) ------------------------------------------------------------
// In a partial round eleven of the twelve lanes only pass through the linear layer, so two consecutive partial rounds are one
// application of M^2 plus a rank-one correction for the s-box in between.  With v = (x_r, s_1..s_11), x_r = sbox(s_0), C = circ(c),
// M = C + 8 E00, t0 = (M v)_0 (ONE row of the first layer), x' = sbox(t0 + const) the s-box of the second round:
//     M (x', (M v)_1..11) = M^2 v + col0(M) (x' - t0),      M^2 v = C^2 v + 8 x_r col0(C) + 8 t0 e_0
//     =>  state after both rounds = C^2 v + col0(C) w + 8 x' e_0 (+ constants),     w = 8 x_r + x' - t0.
// C^2 = circ(c * c) goes through the same CRT butterfly as C (its coefficients sum to 2^16: word sums < 2^49, exact), col0(C) w
// is the contribution of a thirteenth input to the SINGLE layer's sums, and t0 costs 12 multiply-adds per part: 234 FP64
// operations for two rounds instead of 360 (with the second CRT level below), one set of conversions and folds instead of two.  Every coefficient that reaches an
// output is non-negative (c*c >= 4586 > 41 * 41 + ...), so only the unreduced low part of x' needs the 2^42 lift of the fused
// s-box.  Needs the equivalent round constants: the middle round adds a constant on lane 0 only.
#ifndef POSEIDON_DOUBLE_ROUNDS
#define POSEIDON_DOUBLE_ROUNDS 1
#endif
#if POSEIDON_DOUBLE_ROUNDS && POSEIDON_EQUIV_RC && POSEIDON_CRT_LEVEL2 && POSEIDON_FUSE_SBOX
struct PoseidonSqCoef {
  int p2h[6], q2h[6];  // halved CRT coefficients of c * c (cyclic): (c2_d + c2_{d+6}) / 2, (c2_d - c2_{d+6}) / 2
  int m0[12];          // row 0 of M
  int n3[3];           // (p2h_e - p2h_{e+3}) / 2
};
__host__ __device__ constexpr PoseidonSqCoef poseidon_make_sq_coef() {
  constexpr int c[12] = POSEIDON_MDS_ROW;
  PoseidonSqCoef t{};
  int c2[12] = {};
  for (int k = 0; k < 12; k++)
    for (int j = 0; j < 12; j++) c2[k] += c[j] * c[(k - j + 12) % 12];
  for (int d = 0; d < 6; d++) {
    t.p2h[d] = (c2[d] + c2[d + 6]) / 2;
    t.q2h[d] = (c2[d] - c2[d + 6]) / 2;
  }
  for (int j = 0; j < 12; j++) t.m0[j] = c[j] + (j == 0 ? 8 : 0);
  for (int e = 0; e < 3; e++) t.n3[e] = (t.p2h[e] - t.p2h[e + 3]) / 2;
  return t;
}
constexpr bool poseidon_sq_coef_ok() {
  constexpr int c[12] = POSEIDON_MDS_ROW;
  constexpr PoseidonSqCoef t = poseidon_make_sq_coef();
  int c2[12] = {};
  for (int k = 0; k < 12; k++)
    for (int j = 0; j < 12; j++) c2[k] += c[j] * c[(k - j + 12) % 12];
  for (int d = 0; d < 6; d++)
    if ((c2[d] + c2[d + 6]) % 2 || (c2[d] - c2[d + 6]) % 2) return false;
  // second level: p2h_e + p2h_{e+3} = 2048 (5,6,5) and the differences are even
  if (t.p2h[0] + t.p2h[3] != 2048 * 5 || t.p2h[1] + t.p2h[4] != 2048 * 6 || t.p2h[2] + t.p2h[5] != 2048 * 5) return false;
  for (int e = 0; e < 3; e++)
    if ((t.p2h[e] - t.p2h[e + 3]) % 2) return false;
  // every net coefficient of an input word in an output word is >= 0:  c2[(j-i)] - C[i][0] m0[j] (+ 8 C[i][0] for j = 0)
  for (int i = 0; i < 12; i++)
    for (int j = 0; j < 12; j++)
      if (c2[(j - i + 12) % 12] - c[(12 - i) % 12] * t.m0[j] + (j == 0 ? 8 * c[(12 - i) % 12] : 0) < 0) return false;
  return true;
}
static_assert(poseidon_sq_coef_ok(), "double partial rounds: coefficient conditions");

// Second CRT level on the cyclic half of C^2, as for the single layer: S_i = sum_j p2h_{(j-i) mod 6} X_j with A_k = X_k + X_{k+3},
// B_k = X_k - X_{k+3}:  p2h_e + p2h_{e+3} = 2048 (5,6,5),  p2h_e - p2h_{e+3} = (264,-480,-96), halved again:
//     S_i = 1024 E_i + SD_i,  S_{i+3} = 1024 E_i - SD_i,   E_i = 5 (A_0 + A_1 + A_2) + A_{(i+1) mod 3},   SD_i = sum_k n_{(k-i) mod 6} B_k,
//     n = (132,-240,-48,-132,240,48)                                        29 FP64 operations per part instead of 36.
// Seeds of double round k (rounds 4 + 2k and 5 + 2k), for the constants of round 6 + 2k: [0..5] D low, [6..11] D high, [12..14] e low,
// [15..17] g low, [18..20] e high, [21..23] g high with 1024 e_i + g_i = 2^52 + sigma_i, 1024 e_i - g_i = 2^52 + sigma_{i+3} (sigma = the S seed
// of the one-level form; e_i has ten fractional bits next to 2^42: exact), then the two parts of the seed of t0 (lane-0 constant of
// round 5 + 2k).
// POSEIDON_DBL_ONE_BODY: the last double round (k = 10) hands a FULL constant vector (round 26's) on to the closing rounds; built
// as a second loop body with every seed loaded (481 instructions, 7.7 KB) it pushes the code that the warps of an SM are spread
// over — full round, double round, last double round, the Merkle loop — to 35 KB, past the 32 KB instruction cache level.  With
// this option the seeds of k = 10 carry lane 0's constant only, like every other double round (ONE body, eleven iterations), and
// lanes 1..11 receive their constants behind the layer with eleven 5-instruction constant additions.
#ifndef POSEIDON_DBL_ONE_BODY
#define POSEIDON_DBL_ONE_BODY 1
#endif
struct alignas(16) PoseidonRcDbl {
  double v[11][26];
};
enum { DBL_DL = 0, DBL_DH = 6, DBL_EL = 12, DBL_GL = 15, DBL_EH = 18, DBL_GH = 21, DBL_TL = 24, DBL_TH = 25 };
__host__ __device__ constexpr PoseidonRcDbl poseidon_make_rc_dbl() {
  constexpr PoseidonEquivRc rce = poseidon_make_equiv_rc(true);
  constexpr u64 kadj = ((u64)P2V_F64_K << 33) - (u64)P2V_F64_K;
  constexpr u64 cadj = GL_P - kadj;
  PoseidonRcDbl t{};
  for (int k = 0; k < 11; k++) {
    const int r1 = 5 + 2 * k, r2 = 6 + 2 * k;
    u64 alo[6] = {}, ahi[6] = {}, blo[6] = {}, bhi[6] = {};
    for (int i = 0; i < 6; i++) {
#if POSEIDON_DBL_ONE_BODY
      // lane 0's constant only, also for the last pair (the rest of round 26's vector follows behind the layer)
      u64 a = poseidon_addmod_c(i == 0 ? rce.v[r2][0] : 0, cadj), b = poseidon_addmod_c(0, cadj);
#else
      u64 a = poseidon_addmod_c(rce.v[r2][i], cadj), b = poseidon_addmod_c(rce.v[r2][i + 6], cadj);
#endif
      alo[i] = a & 0xFFFFFFFFULL; ahi[i] = a >> 32; blo[i] = b & 0xFFFFFFFFULL; bhi[i] = b >> 32;
      // the unreduced low part of x' is in (-2^34, 2^32) and reaches every output with a coefficient <= 49: lift the low parts by
      // 2^42, take 2^10 off the high parts (same value); a high part below 2^10 is first moved up by p
      if (ahi[i] < 1024) { alo[i] += 1; ahi[i] += 0xFFFFFFFFULL; }
      if (bhi[i] < 1024) { blo[i] += 1; bhi[i] += 0xFFFFFFFFULL; }
      alo[i] += 1ULL << 42; blo[i] += 1ULL << 42; ahi[i] -= 1024; bhi[i] -= 1024;
      bool flo = ((alo[i] ^ blo[i]) & 1) != 0, fhi = ((ahi[i] ^ bhi[i]) & 1) != 0;
      if (flo && fhi) { blo[i] += 1; bhi[i] += 0xFFFFFFFFULL; }
      else if (flo) { blo[i] += 1 + (1ULL << 32); bhi[i] += 0xFFFFFFFEULL; }
      else if (fhi) { blo[i] += 1ULL << 32; bhi[i] -= 1; }
    }
    for (int i = 0; i < 3; i++) {  // sigma_i + sigma_{i+3} even in both parts, by moving a_i (lane 0 is the only one that varies)
      bool olo = (((alo[i] + blo[i]) / 2 + (alo[i + 3] + blo[i + 3]) / 2) & 1) != 0;
      bool ohi = (((ahi[i] + bhi[i]) / 2 + (ahi[i + 3] + bhi[i + 3]) / 2) & 1) != 0;
      if (olo && ohi) { alo[i] += 2; ahi[i] += (1ULL << 33) - 2; }
      else if (olo) { alo[i] += 2 + (1ULL << 33); ahi[i] += (1ULL << 33) - 4; }
      else if (ohi) { alo[i] += 4 + (1ULL << 33); ahi[i] += (1ULL << 34) - 6; }
    }
    for (int i = 0; i < 6; i++) {
      t.v[k][DBL_DL + i] = (double)(((long long)alo[i] - (long long)blo[i]) / 2);
      t.v[k][DBL_DH + i] = (double)(((long long)ahi[i] - (long long)bhi[i]) / 2);
    }
    for (int i = 0; i < 3; i++) {
      long long sl0 = (long long)((alo[i] + blo[i]) / 2), sl3 = (long long)((alo[i + 3] + blo[i + 3]) / 2);
      long long sh0 = (long long)((ahi[i] + bhi[i]) / 2), sh3 = (long long)((ahi[i + 3] + bhi[i + 3]) / 2);
      t.v[k][DBL_EL + i] = P2V_TWO52 / 1024.0 + (double)((sl0 + sl3) / 2) / 1024.0;
      t.v[k][DBL_GL + i] = (double)((sl0 - sl3) / 2);
      t.v[k][DBL_EH + i] = P2V_TWO52 / 1024.0 + (double)((sh0 + sh3) / 2) / 1024.0;
      t.v[k][DBL_GH + i] = (double)((sh0 - sh3) / 2);
    }
    u64 a = poseidon_addmod_c(rce.v[r1][0], cadj);
    t.v[k][DBL_TL] = P2V_TWO52 + (double)(a & 0xFFFFFFFFULL);
    t.v[k][DBL_TH] = P2V_TWO52 + (double)(a >> 32);
  }
  return t;
}
__host__ __device__ constexpr bool poseidon_dbl_seed_varies(int k) { return k == DBL_DL || k == DBL_DH || k == DBL_EL || k == DBL_GL || k == DBL_EH || k == DBL_GH || k >= DBL_TL; }
constexpr bool poseidon_dbl_static_ok() {
  constexpr PoseidonRcDbl t = poseidon_make_rc_dbl();
  constexpr PoseidonEquivRc rce = poseidon_make_equiv_rc(true);
  for (int k = 0; k < (POSEIDON_DBL_ONE_BODY ? 11 : 10); k++)
    for (int j = 0; j < 26; j++)
      if (!poseidon_dbl_seed_varies(j) && t.v[k][j] != t.v[0][j]) return false;
  for (int r = 5; r <= 25; r++)  // the middle rounds (and all but the last target) carry a constant on lane 0 only
    for (int i = 1; i < 12; i++)
      if (rce.v[r][i] != 0) return false;
  return true;
}
static_assert(poseidon_dbl_static_ok(), "double partial rounds: seeds that lane 0 does not reach must not depend on the round");
static __constant__ PoseidonRcDbl c_rcdbl = poseidon_make_rc_dbl();
struct alignas(16) PoseidonRcDblStatic {
  double v[24];
};
__host__ __device__ constexpr PoseidonRcDblStatic poseidon_make_rc_dbl_static() {
  constexpr PoseidonRcDbl t = poseidon_make_rc_dbl();
  PoseidonRcDblStatic r{};
  for (int j = 0; j < 24; j++) r.v[j] = t.v[0][j];
  return r;
}

// the round-independent seeds as compile-time values (ten of the D seeds and the four static g are zero)
constexpr PoseidonRcDblStatic k_rcdbls = poseidon_make_rc_dbl_static();
template <bool ST, int K>
__device__ __forceinline__ double poseidon_dbl_seed(int k) {
  if constexpr (ST && !poseidon_dbl_seed_varies(K)) {
    constexpr double v = k_rcdbls.v[K];
    return v;
  }
  return c_rcdbl.v[k][K];
}
// contribution of input pair J to the negacyclic half (D) of the C^2 sums, rows I..5
template <int J, int I>
__device__ __forceinline__ void poseidon_dbl_dcol(double (&DL)[6], double (&DH)[6], double xmL, double xmH) {
  constexpr PoseidonSqCoef cf = poseidon_make_sq_coef();
  constexpr int d = (J - I + 12) % 12;
  constexpr double qc = (double)(d < 6 ? cf.q2h[d] : -cf.q2h[d - 6]);
  DL[I] = fma(xmL, qc, DL[I]);
  DH[I] = fma(xmH, qc, DH[I]);
  if constexpr (I + 1 < 6) poseidon_dbl_dcol<J, I + 1>(DL, DH, xmL, xmH);
}
// pair (v_J, v_{J+6}): conversions, its share of t0, butterfly, D sums; X+ is kept for the second level
template <int J>
__device__ __forceinline__ void poseidon_dbl_pair(u64 xj, u64 xk, double &tL, double &tH, double (&XL)[6], double (&XH)[6], double (&DL)[6],
                                                  double (&DH)[6]) {
  constexpr PoseidonSqCoef cf = poseidon_make_sq_coef();
  double bjl = __uint2double_rn((u32)xj), bjh = __uint2double_rn((u32)(xj >> 32));
  double bkl = __uint2double_rn((u32)xk), bkh = __uint2double_rn((u32)(xk >> 32));
  tL = fma(bjl, (double)cf.m0[J], tL);
  tH = fma(bjh, (double)cf.m0[J], tH);
  tL = fma(bkl, (double)cf.m0[J + 6], tL);
  tH = fma(bkh, (double)cf.m0[J + 6], tH);
  XL[J] = bjl + bkl;
  XH[J] = bjh + bkh;
  poseidon_dbl_dcol<J, 0>(DL, DH, bjl - bkl, bjh - bkh);
}
// cyclic half of one part from the six X+ (second CRT level)
__device__ __forceinline__ void poseidon_dbl_s(const double (&X)[6], const double (&e)[3], const double (&g)[3], double (&S)[6]) {
  constexpr PoseidonSqCoef cf = poseidon_make_sq_coef();
  constexpr double n0 = (double)cf.n3[0], n1 = (double)cf.n3[1], n2 = (double)cf.n3[2];
  double A0 = X[0] + X[3], A1 = X[1] + X[4], A2 = X[2] + X[5];
  double B0 = X[0] - X[3], B1 = X[1] - X[4], B2 = X[2] - X[5];
  double sum = (A0 + A1) + A2;
  double E0 = fma(sum, 5.0, A1 + e[0]), E1 = fma(sum, 5.0, A2 + e[1]), E2 = fma(sum, 5.0, A0 + e[2]);
  // SD_i = sum_k n_{(k-i) mod 6} B_k,  n_{e+3} = -n_e
  double SD0 = fma(B2, n2, fma(B1, n1, fma(B0, n0, g[0])));
  double SD1 = fma(B0, -n2, fma(B2, n1, fma(B1, n0, g[1])));
  double SD2 = fma(B1, -n2, fma(B0, -n1, fma(B2, n0, g[2])));
  S[0] = fma(E0, 1024.0, SD0); S[3] = fma(E0, 1024.0, -SD0);
  S[1] = fma(E1, 1024.0, SD1); S[4] = fma(E1, 1024.0, -SD1);
  S[2] = fma(E2, 1024.0, SD2); S[5] = fma(E2, 1024.0, -SD2);
}
// rounds 4 + 2k and 5 + 2k.  ST: all seeds but the six that the two lane-0 constants reach are compile-time addresses
template <bool ST>
__device__ __forceinline__ void poseidon_double_partial_round(u64 (&s)[12], int k) {
  double XL[6], XH[6], DL[6], DH[6];
  DL[0] = poseidon_dbl_seed<ST, DBL_DL + 0>(k); DH[0] = poseidon_dbl_seed<ST, DBL_DH + 0>(k);
  DL[1] = poseidon_dbl_seed<ST, DBL_DL + 1>(k); DH[1] = poseidon_dbl_seed<ST, DBL_DH + 1>(k);
  DL[2] = poseidon_dbl_seed<ST, DBL_DL + 2>(k); DH[2] = poseidon_dbl_seed<ST, DBL_DH + 2>(k);
  DL[3] = poseidon_dbl_seed<ST, DBL_DL + 3>(k); DH[3] = poseidon_dbl_seed<ST, DBL_DH + 3>(k);
  DL[4] = poseidon_dbl_seed<ST, DBL_DL + 4>(k); DH[4] = poseidon_dbl_seed<ST, DBL_DH + 4>(k);
  DL[5] = poseidon_dbl_seed<ST, DBL_DL + 5>(k); DH[5] = poseidon_dbl_seed<ST, DBL_DH + 5>(k);
  const double stL = c_rcdbl.v[k][DBL_TL], stH = c_rcdbl.v[k][DBL_TH];
  double tL = stL, tH = stH;
  poseidon_dbl_pair<1>(s[1], s[7], tL, tH, XL, XH, DL, DH);
  poseidon_dbl_pair<2>(s[2], s[8], tL, tH, XL, XH, DL, DH);
  poseidon_dbl_pair<3>(s[3], s[9], tL, tH, XL, XH, DL, DH);
  poseidon_dbl_pair<4>(s[4], s[10], tL, tH, XL, XH, DL, DH);
  poseidon_dbl_pair<5>(s[5], s[11], tL, tH, XL, XH, DL, DH);
  const u64 xr = poseidon_sbox(s[0]);  // first round's s-box
  poseidon_dbl_pair<0>(xr, s[6], tL, tH, XL, XH, DL, DH);
  // second round's s-box on lane 0 of the first layer; its result is only needed as the two parts of an unreduced product
  double x1L, x1H;
  poseidon_sbox_raw(poseidon_crt64_fold(tL, tH), x1L, x1H);
  // cyclic half of C^2 v (independent of the second s-box)
  double SL[6], SH[6];
  {
    double el[3], gl[3], eh[3], gh[3];
    el[0] = poseidon_dbl_seed<ST, DBL_EL + 0>(k); el[1] = poseidon_dbl_seed<ST, DBL_EL + 1>(k); el[2] = poseidon_dbl_seed<ST, DBL_EL + 2>(k);
    gl[0] = poseidon_dbl_seed<ST, DBL_GL + 0>(k); gl[1] = poseidon_dbl_seed<ST, DBL_GL + 1>(k); gl[2] = poseidon_dbl_seed<ST, DBL_GL + 2>(k);
    eh[0] = poseidon_dbl_seed<ST, DBL_EH + 0>(k); eh[1] = poseidon_dbl_seed<ST, DBL_EH + 1>(k); eh[2] = poseidon_dbl_seed<ST, DBL_EH + 2>(k);
    gh[0] = poseidon_dbl_seed<ST, DBL_GH + 0>(k); gh[1] = poseidon_dbl_seed<ST, DBL_GH + 1>(k); gh[2] = poseidon_dbl_seed<ST, DBL_GH + 2>(k);
    poseidon_dbl_s(XL, el, gl, SL);
    poseidon_dbl_s(XH, eh, gh, SH);
  }
  // w = 8 x_r + x' - t0 per part (t0 = t - seed, exact), the thirteenth input of the single layer's sums
  const double wL = fma(__uint2double_rn((u32)xr), 8.0, stL - tL) + x1L;
  const double wH = fma(__uint2double_rn((u32)(xr >> 32)), 8.0, stH - tH) + x1H;
  poseidon_crt64_col<0, 0>(SL, DL, SH, DH, wL, wL, wH, wH);
  poseidon_crt64_finish_d(s, x1L, x1H, SL, DL, SH, DH);  // + 8 x' on lane 0, recombination, folds
}
#define POSEIDON_HAVE_DOUBLE_ROUNDS 1
#else
#define POSEIDON_HAVE_DOUBLE_ROUNDS 0
#endif

// full round with the s-boxes and the layer in ONE basic block: the FP64/ALU work of a pair is independent of the
// s-boxes still to come, so ptxas can interleave it with their wide multiplies (POSEIDON_SPLIT_ROUNDS)
__device__ __forceinline__ void poseidon_full_round_crt64(u64 (&s)[12], int next_round) {
  double SL[6], DL[6], SH[6], DH[6];
  poseidon_crt64_seed(SL, DL, SH, DH, next_round);
  u64 x0 = 0;
#define P2V_FR_PAIR(J)                                             \
  {                                                                \
    u64 a = poseidon_sbox(s[J]), b = poseidon_sbox(s[J + 6]);      \
    if (J == 0) x0 = a;                                            \
    poseidon_crt64_pair<J>(a, b, SL, DL, SH, DH);                  \
  }
  P2V_FR_PAIR(1) P2V_FR_PAIR(2) P2V_FR_PAIR(3) P2V_FR_PAIR(4) P2V_FR_PAIR(5) P2V_FR_PAIR(0)
#undef P2V_FR_PAIR
  poseidon_crt64_finish(s, x0, SL, DL, SH, DH);
}

#ifndef POSEIDON_MDS_SEPARATE_OUT
#define POSEIDON_MDS_SEPARATE_OUT 0
#endif
#ifndef POSEIDON_MDS_F64
#define POSEIDON_MDS_F64 4
#endif

// The permutation.  Input: lazy u64 (any values); output: lazy u64 (apply gl_canon before use as data).
#ifndef POSEIDON_SPLIT_ROUNDS
#define POSEIDON_SPLIT_ROUNDS 1
#endif
__device__ __forceinline__ void poseidon_mds_layer(u64 (&s)[12], int next_round) {
#if POSEIDON_MDS_F64 == 4 && POSEIDON_CRT_LEVEL2
  poseidon_mds_crt64_l2(s, next_round);
#elif POSEIDON_MDS_F64 == 4
  poseidon_mds_crt64(s, next_round);
#elif POSEIDON_MDS_F64 == 3
  poseidon_mds_crt(s, next_round);
#elif POSEIDON_MDS_F64 == 2
  poseidon_mds_mixed(s, next_round);
#elif POSEIDON_MDS_F64
  poseidon_mds_f64(s, next_round);
#else
  poseidon_mds(s, next_round);
#endif
}
// a + c for a CANONICAL constant c (< p), a lazy: after a wrap the low word is a + c - 2^64 <= p - 2, so the single correction
// + EPS cannot wrap again — 5 instructions against the 12 of gl_add with its two possible wraps
constexpr bool poseidon_rc_canonical() {
  constexpr u64 rc[360] = P2V_ALL_ROUND_CONSTANTS;
  for (int i = 0; i < 360; i++)
    if (rc[i] >= GL_P) return false;
  return true;
}
static_assert(poseidon_rc_canonical(), "round constants must be canonical");
__device__ __forceinline__ u64 poseidon_add_rc(u64 a, u64 c) {
  u64 r = a + c;
  return r < a ? r + GL_EPS : r;
}
__device__ __forceinline__ void poseidon_permute(u64 (&s)[12]) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = poseidon_add_rc(s[i], c_pt.rc[0][i]);
#if POSEIDON_SPLIT_ROUNDS
  // Two loop bodies instead of one with a branch: a full round (12 s-boxes + layer in one basic block, so that the layer's
  // FP64/ALU work is scheduled into the shadow of the wide multiplies) and a partial round.
#pragma unroll 1
  for (int ph = 0; ph < 3; ph++) {
    if (ph != 1) {
      int r0 = ph ? 26 : 0;
#pragma unroll 1
      for (int r = r0; r < r0 + 4; r++) {
#if POSEIDON_MDS_F64 == 4 && POSEIDON_CRT_LEVEL2
        poseidon_full_round_crt64_l2(s, r + 1);
#elif POSEIDON_MDS_F64 == 4
        poseidon_full_round_crt64(s, r + 1);
#else
#pragma unroll
        for (int k = 0; k < 12; k++) s[k] = poseidon_sbox(s[k]);
        poseidon_mds_layer(s, r + 1);
#endif
      }
    } else {
#if POSEIDON_MDS_F64 == 4 && POSEIDON_HAVE_DOUBLE_ROUNDS
      // the 22 partial rounds two at a time; the last pair hands the full constant vector on to the closing rounds
#if POSEIDON_DBL_ONE_BODY
#pragma unroll 1
      for (int k = 0; k < 11; k++) poseidon_double_partial_round<true>(s, k);
      {  // lanes 1..11 of round 26's constants, behind the layer (compile-time values)
        constexpr PoseidonEquivRc rce = poseidon_make_equiv_rc(true);
#pragma unroll
        for (int i = 1; i < 12; i++) s[i] = poseidon_add_rc(s[i], rce.v[26][i]);
      }
#else
#pragma unroll 1
      for (int k = 0; k < 10; k++) poseidon_double_partial_round<true>(s, k);
      poseidon_double_partial_round<false>(s, 10);
#endif
#elif POSEIDON_MDS_F64 == 4 && POSEIDON_CRT_LEVEL2 && POSEIDON_EQUIV_RC
      // rounds 4..24 add a constant on lane 0 only (equivalent constants); round 25 hands the full vector on to the closing rounds
#pragma unroll 1
      for (int r = 4; r < 25; r++) {
        s[0] = poseidon_sbox(s[0]);
        poseidon_mds_crt64_l2<true>(s, r + 1);
      }
      s[0] = poseidon_sbox(s[0]);
      poseidon_mds_crt64_l2<false>(s, 26);
#else
#pragma unroll 1
      for (int r = 4; r < 26; r++) {
        s[0] = poseidon_sbox(s[0]);
        poseidon_mds_layer(s, r + 1);
      }
#endif
    }
  }
#else
#pragma unroll 1
  for (int r = 0; r < 30; r++) {
    if (r < 4 || r >= 26) {
      // full round: 12 s-boxes, as a register-rotating loop so that the body stays small for the
      // instruction cache while every register index remains static
#pragma unroll 1
      for (int g = 0; g < 12 / POSEIDON_SBOX_GROUP; g++) {
        u64 t[POSEIDON_SBOX_GROUP];
#pragma unroll
        for (int k = 0; k < POSEIDON_SBOX_GROUP; k++) t[k] = poseidon_sbox(s[k]);
#pragma unroll
        for (int k = 0; k + POSEIDON_SBOX_GROUP < 12; k++) s[k] = s[k + POSEIDON_SBOX_GROUP];
#pragma unroll
        for (int k = 0; k < POSEIDON_SBOX_GROUP; k++) s[12 - POSEIDON_SBOX_GROUP + k] = t[k];
      }
    } else {
      s[0] = poseidon_sbox(s[0]);
    }
    poseidon_mds_layer(s, r + 1);
  }
#endif
}
