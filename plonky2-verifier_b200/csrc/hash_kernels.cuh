// K1-K3: batched Poseidon permutation, leaf sponge, 2-to-1 compression, Merkle path checks and
// the synthetic-tree builder.  One thread per state / sponge / path; every plane access is
// [word][item] so a warp reads 256 contiguous bytes per word.
#pragma once
#include "poseidon.cuh"

// K1: `permutation`, Hash/Poseidon.hs:42.  in/out SoA [12][n].
__global__ void __launch_bounds__(256) k_poseidon_permute(const u64 *__restrict__ in, u64 *__restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = in[(size_t)i * n + t];
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 12; i++) out[(size_t)i * n + t] = gl_canon(s[i]);
  }
}

// Sponge over a strided column of words: word j of item t is at base[j*stride + t].
// `sponge`, Hash/Sponge.hs:26-31: overwrite mode, rate 8, no padding, w = 0 -> zero digest.
// Returns the state; digest = s[0..3] (lazy).
__device__ __forceinline__ void sponge_absorb_strided(u64 (&s)[12], const u64 *__restrict__ base, size_t stride, u32 w) {
#pragma unroll
  for (int i = 0; i < 12; i++) s[i] = 0;
  u32 nblk = (w + 7) / 8;
#pragma unroll 1
  for (u32 b = 0; b < nblk; b++) {
    u32 k = w - b * 8;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if ((u32)i < k) s[i] = base[(size_t)(b * 8 + i) * stride];
    poseidon_permute(s);
  }
}

// K2: leaf hashing.  leaves SoA [w][n] -> digests SoA [4][n]
__global__ void __launch_bounds__(256) k_hash_leaves(const u64 *__restrict__ leaves, u32 w, size_t n, u64 *__restrict__ digests) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    u64 s[12];
    sponge_absorb_strided(s, leaves + t, n, w);
#pragma unroll
    for (int i = 0; i < 4; i++) digests[(size_t)i * n + t] = gl_canon(s[i]);
  }
}

// `compress`, Hash/Merkle.hs:21-23.  out may alias neither input.
// General strided form used by the tree builder: left word k of item t at left[k*ls + t*lm].
__global__ void __launch_bounds__(256) k_compress(const u64 *__restrict__ left, const u64 *__restrict__ right, size_t in_stride,
                                                  size_t in_mul, u64 *__restrict__ out, size_t out_stride, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      s[i] = left[(size_t)i * in_stride + t * in_mul];
      s[4 + i] = right[(size_t)i * in_stride + t * in_mul];
      s[8 + i] = 0;
    }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[(size_t)i * out_stride + t] = gl_canon(s[i]);
  }
}

// K3: `checkMerkleProof cap idx leaf proof`, Hash/Merkle.hs:27-42, one thread per opening.
// A single permutation call site: iteration 0..nblk-1 absorb the leaf, then one compression per
// sibling (even idx => node is the left input, Merkle.hs:34-36).
__global__ void __launch_bounds__(256) k_merkle_verify(const u64 *__restrict__ leaves, u32 w, const u32 *__restrict__ idx,
                                                       const u64 *__restrict__ siblings, u32 path_len,
                                                       const u64 *__restrict__ cap, u32 cap_height, size_t n,
                                                       u32 *__restrict__ ok_bits, u64 *__restrict__ roots_out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n_round = (n + 31) / 32 * 32;  // whole warps stay alive for the ballot
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_round; t += stride) {
    bool live = t < n;
    size_t tt = live ? t : n - 1;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    u32 nblk = (w + 7) / 8;
    u32 index = idx[tt];
    u32 iters = nblk + path_len;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
      if (it < nblk) {
        u32 k = w - it * 8;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if ((u32)i < k) s[i] = leaves[(size_t)(it * 8 + i) * n + tt];
      } else {
        u32 l = it - nblk;
        u64 sib[4];
#pragma unroll
        for (int i = 0; i < 4; i++) sib[i] = siblings[(size_t)(l * 4 + i) * n + tt];
        bool even = (index & 1u) == 0;
        index >>= 1;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          u64 node = s[i];
          s[i] = even ? node : sib[i];
          s[4 + i] = even ? sib[i] : node;
          s[8 + i] = 0;
        }
      }
      poseidon_permute(s);
    }
    // w == 0 and path_len == 0: digest is the zero digest (Sponge.hs:28)
    bool ok = index < (1u << cap_height);
    u32 ci = ok ? index : 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      u64 v = gl_canon(s[i]);
      if (roots_out && live) roots_out[(size_t)i * n + t] = v;
      ok = ok && (v == gl_canon(cap[(size_t)ci * 4 + i]));
    }
    u32 ballot = __ballot_sync(0xffffffffu, ok && live);
    if ((threadIdx.x & 31) == 0) ok_bits[t / 32] = ballot;
  }
}

// Gather openings from a built tree (synthetic config 2).
__global__ void k_merkle_open(const u64 *__restrict__ leaves, u32 w, u32 log_n, u32 cap_height,
                              const u64 *__restrict__ digests, const u32 *__restrict__ idx, size_t n,
                              u64 *__restrict__ leaves_out, u64 *__restrict__ siblings_out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n_leaves = (size_t)1 << log_n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    u32 index = idx[t];
    for (u32 j = 0; j < w; j++) leaves_out[(size_t)j * n + t] = leaves[(size_t)j * n_leaves + index];
    size_t level_off = 0;  // in digests (4 words each, SoA per level)
    u32 path_len = log_n - cap_height;
    for (u32 l = 0; l < path_len; l++) {
      size_t level_n = n_leaves >> l;
      u32 sib = (index >> l) ^ 1u;
      for (int k = 0; k < 4; k++) siblings_out[(size_t)(l * 4 + k) * n + t] = digests[level_off + (size_t)k * level_n + sib];
      level_off += 4 * level_n;
    }
  }
}

// cap level (SoA [4][2^cap_height]) -> row-major [2^cap_height][4]
__global__ void k_cap_transpose(const u64 *__restrict__ level, u32 ncap, u64 *__restrict__ cap_out) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ncap * 4) cap_out[(t % ncap) * 4 + t / ncap] = level[t];
}

// ---- integer-pipe peak microbenchmark (measurement helper) --------------------------------
// 8 independent dependent-chains per thread; every step of a chain is a GROUP of instructions:
//   mode 0: LOP3 + IMAD.WIDE.U32   (both halves of the product feed the next step, so ptxas cannot
//                                   narrow it to a 32-bit IMAD or hoist it)
//   mode 1: LOP3 + LOP3 + IMAD.WIDE.U32
//   mode 2: LOP3 + IMAD (32-bit mul.lo)
//   mode 3: LOP3 + IADD3           (ALU pipe only)
//   mode 4: IMAD.WIDE.U32 + IMAD.WIDE.U32 + LOP3
//   mode 5: 2 x IMAD (32-bit)   6: 2 x LOP3   7: 1 x IMAD.WIDE.U32   8: 2 x SHF   9: IADD3 + IADD3.X (carry chain)
//   mode 10: DFMA   11: DFMA + LOP3 + IMAD (three pipes)   12: DADD
// p2v_int_pipe_peak reports GROUPS per second (x32 threads).  If IMAD.WIDE issues every 2 cycles per
// SM sub-partition, modes 0, 2 and 3 give the same rate (64 groups/clk/SM); mode 4 then runs at half
// that rate, and if IMAD.WIDE were half rate mode 0 would already be at half.
template <int MODE>
__global__ void __launch_bounds__(256) k_int_pipe(u64 *out, u32 iters, u32 seed) {
  u32 x = threadIdx.x * 2654435761u + seed;
  u64 a[8];
  u32 b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = (u64)x * (i + 3) + i;
    b[i] = x * (i + 7);
  }
  if (MODE >= 10 && MODE != 13 && MODE != 15) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 0x3FF0000000000000ULL | (a[i] & 0xFFFFFFFFFULL);  // doubles in [1,2)
  }
#pragma unroll 1
  for (u32 it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (MODE == 0) {
          asm volatile("{.reg .u32 lo,hi,t; mov.b64 {lo,hi},%0; xor.b32 t,lo,hi; mul.wide.u32 %0,t,%1;}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 1) {
          asm volatile("{.reg .u32 lo,hi,t; mov.b64 {lo,hi},%0; xor.b32 t,lo,hi; and.b32 t,t,%1; mul.wide.u32 %0,t,%1;}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 2) {
          asm volatile("{.reg .u32 lo,hi,t; mov.b64 {lo,hi},%0; xor.b32 t,lo,%1; mul.lo.u32 lo,t,%1; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 3) {
          asm volatile("{.reg .u32 lo,hi,t; mov.b64 {lo,hi},%0; xor.b32 t,lo,%1; add.u32 lo,lo,t; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 4) {
          asm volatile("{.reg .u32 lo,hi,t; .reg .u64 w; mov.b64 {lo,hi},%0; xor.b32 t,lo,hi; mul.wide.u32 w,t,%1; mov.b64 {lo,hi},w; mul.wide.u32 %0,lo,hi;}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 5) {  // 32-bit IMAD only (both halves are independent chains)
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; mad.lo.u32 lo,lo,lo,%1; mad.lo.u32 hi,hi,hi,%1; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 6) {  // LOP3 only
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; lop3.b32 lo,lo,hi,%1,0x96; lop3.b32 hi,hi,lo,%1,0xe8; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 7) {  // IMAD.WIDE only
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; mul.wide.u32 %0,lo,hi;}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 8) {  // funnel shifts only
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; shf.l.wrap.b32 lo,lo,hi,%1; shf.r.wrap.b32 hi,hi,lo,%1; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 9) {  // 64-bit add with carry (IADD3 + IADD3.X)
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; add.cc.u32 lo,lo,hi; addc.u32 hi,hi,%1; mov.b64 %0,{lo,hi};}" : "+l"(a[i]) : "r"(x));
        } else if (MODE == 10) {  // DFMA only (FP64 pipe)
          asm volatile("{.reg .f64 d; mov.b64 d,%0; fma.rn.f64 d,d,0d3FF0000000000001,0d3FF8000000000000; mov.b64 %0,d;}" : "+l"(a[i]));
        } else if (MODE == 11) {  // DFMA + IMAD32 + LOP3: three pipes
          asm volatile("{.reg .f64 d; mov.b64 d,%0; fma.rn.f64 d,d,0d3FF0000000000001,0d3FF8000000000000; mov.b64 %0,d;}" : "+l"(a[i]));
          asm volatile("{.reg .u32 t; xor.b32 t,%0,%1; mul.lo.u32 %0,t,%1;}" : "+r"(b[i]) : "r"(x));
        } else if (MODE == 12) {  // DADD only
          asm volatile("{.reg .f64 d; mov.b64 d,%0; add.rn.f64 d,d,0d3FF8000000000000; mov.b64 %0,d;}" : "+l"(a[i]));
        } else if (MODE == 13) {  // I2F.F64.U32 + LOP3 (which pipe converts, and how fast?)
          asm volatile("{.reg .f64 d; .reg .u32 lo,hi; cvt.rn.f64.u32 d,%0; mov.b64 {lo,hi},d; xor.b32 %0,lo,hi;}" : "+r"(b[i]));
        } else if (MODE == 14) {  // I2F.F64.U32 + LOP3 next to an independent DFMA chain
          asm volatile("{.reg .f64 d; .reg .u32 lo,hi; cvt.rn.f64.u32 d,%0; mov.b64 {lo,hi},d; xor.b32 %0,lo,hi;}" : "+r"(b[i]));
          asm volatile("{.reg .f64 d; mov.b64 d,%0; fma.rn.f64 d,d,0d3FF0000000000001,0d3FF8000000000000; mov.b64 %0,d;}" : "+l"(a[i]));
        } else {                  // 15: I2F.F64.U32 + LOP3 next to an independent IMAD.WIDE chain
          asm volatile("{.reg .f64 d; .reg .u32 lo,hi; cvt.rn.f64.u32 d,%0; mov.b64 {lo,hi},d; xor.b32 %0,lo,hi;}" : "+r"(b[i]));
          asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%0; mul.wide.u32 %0,lo,hi;}" : "+l"(a[i]) : "r"(x));
        }
      }
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc += a[i] + b[i];
  if (acc == 0x1234567812345678ULL) out[0] = acc;  // keep the chains alive
}

// Register-file / operand-bandwidth probe (p2v_int_pipe_peak modes 16..23): the same pipes as k_int_pipe but with THREE
// distinct register sources per instruction, spread over 24 live registers, the way real code reads operands.
//   16: IADD3 a = a + b + c     17: LOP3 a = f(a, b, c)     18: DFMA d = d * e + f (three 64-bit registers)
//   19: DFMA d = e * imm + d    20: 18 + 16 interleaved      21: 19 + 16 interleaved
//   22: IMAD.WIDE acc = x * y + acc (two 32-bit + one 64-bit register)   23: 22 + 19 + 16 interleaved
template <int MODE>
__global__ void __launch_bounds__(256) k_rf_probe(u64 *out, u32 iters, u32 seed) {
  u32 x = threadIdx.x * 2654435761u + seed;
  u32 a[8], b[8], c[8];
  double d[8], e[8], f[8];
  u64 w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = x * (i + 3); b[i] = x * (i + 11) + 1; c[i] = x * (i + 19) + 2;
    d[i] = 1.0 + 1e-9 * (double)(a[i] & 0xffff); e[i] = 1.0 + 1e-12 * (double)(b[i] & 0xffff); f[i] = 1e-9 * (double)(c[i] & 0xffff);
    w[i] = (u64)a[i] * b[i];
  }
#pragma unroll 1
  for (u32 it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int j = (i + 3) & 7, k = (i + 5) & 7;
        if (MODE == 16 || MODE == 20 || MODE == 21 || MODE == 23) asm volatile("{.reg .u32 t; add.u32 t,%1,%2; add.u32 %0,%0,t;}" : "+r"(a[i]) : "r"(b[j]), "r"(c[k]));
        if (MODE == 17) asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(a[i]) : "r"(b[j]), "r"(c[k]));
        if (MODE == 18 || MODE == 20) asm volatile("fma.rn.f64 %0,%0,%1,%2;" : "+d"(d[i]) : "d"(e[j]), "d"(f[k]));
        if (MODE == 19 || MODE == 21 || MODE == 23) asm volatile("fma.rn.f64 %0,%1,0d4000000000000000,%0;" : "+d"(d[i]) : "d"(e[j]));
        if (MODE == 22 || MODE == 23) asm volatile("{.reg .u32 lo,hi; mov.b64 {lo,hi},%1; mad.wide.u32 %0,lo,%2,%0;}" : "+l"(w[i]) : "l"(w[j]), "r"(c[k]));
      }
    }
  }
  u64 acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc += a[i] + (u64)__double_as_longlong(d[i]) + w[i];
  if (acc == 0x1234567812345678ULL) out[0] = acc;
}

// Test hook (p2v_debug_field_op): the device field routines of gl.cuh one by one on caller-chosen operands, so that the
// lazy-representation edge values (0, 1, p-1, p, p+1, 2^64-1, 2^32+-1) reach every routine directly and `inv 0 = 0`
// (Algebra/Goldilocks.hs:155, GoldilocksExt.hs:75-80) is tested on the GPU.  Outputs are canonical.
__global__ void k_field_op(int op, const u64 *__restrict__ a, const u64 *__restrict__ b, u64 *__restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    if (op < 16) {
      u64 x = a[t], y = b ? b[t] : 0, r = 0;
      switch (op) {
        case 0: r = gl_add(x, y); break;
        case 1: r = gl_sub(x, y); break;
        case 2: r = gl_mul(x, y); break;
        case 3: r = gl_inv(x); break;
        case 4: r = gl_neg(x); break;
        case 5: r = gl_mul_small(x, (u32)y); break;
        case 6: r = gl_reduce128(x, y); break;
        case 7: r = gl_pow(x, y); break;
        case 8: r = x; break;
        default: r = poseidon_sbox(x); break;
      }
      out[t] = gl_canon(r);
    } else {
      gl2 x = gl2_make(a[t], a[n + t]), y = b ? gl2_make(b[t], b[n + t]) : gl2_make(0, 0), r = gl2_make(0, 0);
      switch (op) {
        case 16: r = gl2_mul(x, y); break;
        case 17: r = gl2_inv(x); break;
        case 18: r = gl2_add(x, y); break;
        case 19: r = gl2_sub(x, y); break;
        case 20: r = gl2_sqr(x); break;
        case 21: r = gl2_scale(y.a, x); break;
        case 22: r = gl2_mul_x(x); break;
        case 23: r = gl2_pow(x, y.a); break;
        default: r = gl2_neg(x); break;
      }
      r = gl2_canon(r);
      out[t] = r.a;
      out[n + t] = r.b;
    }
  }
}

