// Goldilocks (p = 2^64 - 2^32 + 1) and its quadratic extension F[X]/(X^2-7) on sm_100a.
//
// Device restatement of what the reference does with `Integer` + `mod`
// (src/Algebra/Goldilocks.hs:140-175, src/Algebra/GoldilocksExt.hs:54-99).
//
// Representation: a field element is a u64 in [0, 2^64) ("lazy": value mod p is what counts).
// Every routine here accepts lazy inputs and returns a lazy output unless its name says
// otherwise; `gl_canon` gives the canonical representative in [0,p) and MUST be applied before
// any comparison, bit test, index derivation or store to caller-visible memory
// (SURVEY.md App. B.1).
//
// The reduction uses the special form 2^64 = 2^32 - 1 (mod p), 2^96 = -1 (mod p):
//   lo + 2^64*(hh*2^32 + hl)  =  lo - hh + hl*(2^32-1)   (mod p)
// i.e. only 32x32->64 IMAD.WIDE, IADD3 and predicated fix-ups; no division, no Montgomery form.
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_MUL_GEN_C 0xc65c18b67785d900ULL /* mulGen, Algebra/Goldilocks.hs:135 */

__device__ __forceinline__ u64 gl_canon(u64 x) { return x >= GL_P ? x - GL_P : x; }

// a + b, lazy inputs.  Two possible wraps (see DESIGN.md "lazy arithmetic bounds").
__device__ __forceinline__ u64 gl_add(u64 a, u64 b) {
  u64 r = a + b;
  if (r < a) {
    r += GL_EPS;
    if (r < GL_EPS) r += GL_EPS;
  }
  return r;
}
// a - b, lazy inputs.
__device__ __forceinline__ u64 gl_sub(u64 a, u64 b) {
  u64 r = a - b;
  if (a < b) {
    u64 t = r - GL_EPS;
    if (r < GL_EPS) t -= GL_EPS;
    r = t;
  }
  return r;
}
__device__ __forceinline__ u64 gl_neg(u64 a) { return gl_sub(0, a); }

// (hi:lo) mod p, 128-bit input, lazy 64-bit output: carry-chain form (no compares/selects)
//   t = (w1:w0) - w3, minus EPS on borrow;  r = t + w2*(2^32-1), plus EPS on carry.
// The multiply-add pair mad.lo.cc / madc.hi.cc becomes ONE IMAD.HI.U32 with carry-out (+ an IMAD.IADD for the low
// word, which is just t0 - w2).  The carry c is applied as r0 = t0 - c, r1 = t1 - borrow + c (3 instructions; the
// mask form add.cc m / addc 0 compiled to 5).  10 SASS instructions per reduction instead of 15.  ptxas equalises
// the instruction COUNTS of the FMA and ALU pipes (IMAD.X / IMAD.MOV / IMAD.IADD stand in for IADD3 / MOV) as if
// IMAD.WIDE and IMAD.HI cost one slot; they cost two, so what the s-box phase minimises is (instructions + 2 x wide
// multiplies) — measured on a 4-s-box loop: 356+128 (shift form) -> 308+160 -> 276+160.
// Neither correction can wrap twice: after a borrow t >= 2^64 - 2^32 + 1, and a wrapped r is < u <= 2^64 - 2^33 + 1.
// NOTE on flags: `subc m,0,0` right after a SUB chain yields the borrow mask, but after an ADD chain ptxas feeds the
// raw hardware carry into it (inverted meaning), so carries are materialised with addc + neg instead.
__device__ __forceinline__ u64 gl_reduce128(u64 lo, u64 hi) {
  u32 w0 = (u32)lo, w1 = (u32)(lo >> 32), w2 = (u32)hi, w3 = (u32)(hi >> 32), r0, r1;
  asm("{\n\t.reg .u32 m,t0,t1;\n\t"
      "sub.cc.u32 t0,%2,%5;\n\tsubc.cc.u32 t1,%3,0;\n\tsubc.u32 m,0,0;\n\t"
      "sub.cc.u32 t0,t0,m;\n\tsubc.u32 t1,t1,0;\n\t"
      "mad.lo.cc.u32 t0,%4,0xffffffff,t0;\n\tmadc.hi.cc.u32 t1,%4,0xffffffff,t1;\n\taddc.u32 m,0,0;\n\t"
      "sub.cc.u32 %0,t0,m;\n\tsubc.u32 t1,t1,0;\n\tadd.u32 %1,t1,m;\n\t}"   // + c*EPS = + (c << 32) - c
      : "=r"(r0), "=r"(r1)
      : "r"(w0), "r"(w1), "r"(w2), "r"(w3));
  return ((u64)r1 << 32) | r0;
}
// one 128-bit product: ptxas shares the partial products between the low and the high half (4 IMAD.WIDE + 3)
__device__ __forceinline__ u64 gl_mul(u64 a, u64 b) {
  unsigned __int128 m = (unsigned __int128)a * b;
  return gl_reduce128((u64)m, (u64)(m >> 64));
}
__device__ __forceinline__ u64 gl_sqr(u64 a) { return gl_mul(a, a); }
// a * small constant c (c < 2^32)
__device__ __forceinline__ u64 gl_mul_small(u64 a, u32 c) {
  u64 lo = (u64)(u32)a * c;
  u64 hi = (u64)(u32)(a >> 32) * c;  // value = lo + 2^32*hi  (hi < 2^64)
  // 2^32*hi = (hi_lo << 32) + hi_hi*2^64 = (hi_lo<<32) + hi_hi*(2^32-1)
  u32 hi_lo = (u32)hi, hi_hi = (u32)(hi >> 32);
  u64 a0 = lo + (u64)hi_hi * 0xFFFFFFFFu;  // lo <= (2^32-1)^2, second term likewise: may wrap once
  if (a0 < lo) a0 += GL_EPS;
  u64 r = a0 + ((u64)hi_lo << 32);
  if (r < a0) {
    r += GL_EPS;
    if (r < GL_EPS) r += GL_EPS;
  }
  return r;
}

__device__ __forceinline__ u64 gl_pow(u64 x, u64 e) {
  u64 acc = 1, s = x;
  while (e) {
    if (e & 1) acc = gl_mul(acc, s);
    s = gl_sqr(s);
    e >>= 1;
  }
  return acc;
}
// x^(p-2): inv 0 = 0 exactly like the reference (Algebra/Goldilocks.hs:155-156).
// Addition chain: 63 squarings + 9 multiplications.
__device__ __forceinline__ u64 gl_exp_acc(u64 base, u64 tail, int n) {
  for (int i = 0; i < n; i++) base = gl_sqr(base);
  return gl_mul(base, tail);
}
static __device__ __noinline__ u64 gl_inv(u64 x) {
  // x^(2^k - 1) ladders
  u64 t2 = gl_exp_acc(x, x, 1);      // x^(2^2-1)
  u64 t3 = gl_exp_acc(t2, x, 1);     // 2^3-1
  u64 t6 = gl_exp_acc(t3, t3, 3);    // 2^6-1
  u64 t12 = gl_exp_acc(t6, t6, 6);   // 2^12-1
  u64 t24 = gl_exp_acc(t12, t12, 12);  // 2^24-1
  u64 t30 = gl_exp_acc(t24, t6, 6);  // 2^30-1
  u64 t31 = gl_exp_acc(t30, x, 1);   // 2^31-1
  u64 t32 = gl_exp_acc(t31, x, 1);   // 2^32-1
  // p - 2 = 0xFFFFFFFE_FFFFFFFF = (2^31-1)*2^33 + (2^32-1)
  u64 t = t31;
  for (int i = 0; i < 33; i++) t = gl_sqr(t);
  return gl_mul(t, t32);
}

// ---- quadratic extension ---------------------------------------------------------------
struct gl2 {
  u64 a, b;  // a + b*X
};
__device__ __forceinline__ gl2 gl2_make(u64 a, u64 b) { gl2 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ gl2 gl2_add(gl2 x, gl2 y) { return gl2_make(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
__device__ __forceinline__ gl2 gl2_sub(gl2 x, gl2 y) { return gl2_make(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
__device__ __forceinline__ gl2 gl2_add_base(gl2 x, u64 y) { return gl2_make(gl_add(x.a, y), x.b); }
__device__ __forceinline__ gl2 gl2_sub_base(gl2 x, u64 y) { return gl2_make(gl_sub(x.a, y), x.b); }
// (r1*r2 + 7*i1*i2, r1*i2 + r2*i1)  GoldilocksExt.hs:59
__device__ __forceinline__ gl2 gl2_mul(gl2 x, gl2 y) {
  u64 rr = gl_mul(x.a, y.a);
  u64 ii = gl_mul(x.b, y.b);
  u64 ri = gl_mul(x.a, y.b);
  u64 ir = gl_mul(x.b, y.a);
  return gl2_make(gl_add(rr, gl_mul_small(ii, 7)), gl_add(ri, ir));
}
__device__ __forceinline__ gl2 gl2_sqr(gl2 x) { return gl2_mul(x, x); }
__device__ __forceinline__ gl2 gl2_scale(u64 s, gl2 x) { return gl2_make(gl_mul(s, x.a), gl_mul(s, x.b)); }
__device__ __forceinline__ gl2 gl2_mul_small(gl2 x, u32 c) { return gl2_make(gl_mul_small(x.a, c), gl_mul_small(x.b, c)); }
// multiply by X: (a + bX)*X = 7b + aX
__device__ __forceinline__ gl2 gl2_mul_x(gl2 x) { return gl2_make(gl_mul_small(x.b, 7), x.a); }
__device__ __forceinline__ gl2 gl2_neg(gl2 x) { return gl2_make(gl_neg(x.a), gl_neg(x.b)); }
__device__ __forceinline__ gl2 gl2_canon(gl2 x) { return gl2_make(gl_canon(x.a), gl_canon(x.b)); }
__device__ __forceinline__ bool gl2_eq(gl2 x, gl2 y) { return gl_canon(x.a) == gl_canon(y.a) && gl_canon(x.b) == gl_canon(y.b); }
// invExt, GoldilocksExt.hs:75-80
__device__ __forceinline__ gl2 gl2_inv(gl2 x) {
  u64 denom = gl_inv(gl_sub(gl_sqr(x.a), gl_mul_small(gl_sqr(x.b), 7)));
  return gl2_make(gl_mul(x.a, denom), gl_mul(gl_neg(x.b), denom));
}
__device__ __forceinline__ gl2 gl2_pow(gl2 x, u64 e) {
  gl2 acc = gl2_make(1, 0), s = x;
  while (e) {
    if (e & 1) acc = gl2_mul(acc, s);
    s = gl2_sqr(s);
    e >>= 1;
  }
  return acc;
}
