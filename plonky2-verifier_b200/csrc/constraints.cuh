// K5: all Plonk constraints at zeta, combined with powers of each alpha, and the quotient identity.
//   evalAllPlonkConstraints / combineWithPowersOfAlpha   src/Plonk/Vanishing.hs:54-111
//   checkCombinedPlonkEquations'                          src/Plonk/Verifier.hs:35-52
//   gate constraint programs                              src/Gate/Constraints.hs:40-128, src/Gate/Custom/*.hs
//   selector filters                                      src/Gate/Selector.hs:62-95
//   lookup equations                                      src/Plonk/Lookups.hs:45-132
//
// One thread per proof.  The reference compiles each gate to a symbolic straight-line program and
// interprets it over FExt (Gate/Computation.hs:157-164); field arithmetic is exact, so here each
// gate is a device function evaluating the same formulas directly, emitting its constraints IN THE
// SAME ORDER.  Because the final value is
//     combined_j = sum_k alpha_j^k term_k,   terms = zs1 ++ pp ++ lookups ++ (sum_g filter_g * c_g)
// every emitted constraint is folded straight into r running sums and never stored.
#pragma once
#include "verify_kernels.cuh"

// "ext of ext": Ext (Expr v) evaluated over FExt (Gate/Vars.hs:56-57, GoldilocksExt.hs:59)
struct gl4 {
  gl2 r, i;
};
__device__ __forceinline__ gl4 gl4_make(gl2 r, gl2 i) { gl4 x; x.r = r; x.i = i; return x; }
__device__ __forceinline__ gl4 gl4_add(gl4 x, gl4 y) { return gl4_make(gl2_add(x.r, y.r), gl2_add(x.i, y.i)); }
__device__ __forceinline__ gl4 gl4_sub(gl4 x, gl4 y) { return gl4_make(gl2_sub(x.r, y.r), gl2_sub(x.i, y.i)); }
__device__ __forceinline__ gl4 gl4_mul(gl4 x, gl4 y) {
  gl2 rr = gl2_mul(x.r, y.r), ii = gl2_mul(x.i, y.i);
  return gl4_make(gl2_add(rr, gl2_mul_small(ii, 7)), gl2_add(gl2_mul(x.r, y.i), gl2_mul(y.r, x.i)));
}
__device__ __forceinline__ gl4 gl4_scale(gl2 s, gl4 x) { return gl4_make(gl2_mul(s, x.r), gl2_mul(s, x.i)); }
__device__ __forceinline__ gl4 gl4_scale_base(u64 s, gl4 x) { return gl4_make(gl2_scale(s, x.r), gl2_scale(s, x.i)); }

struct ConstraintCtx {
  const u64 *__restrict__ pp;
  size_t n, p;
  int off_wires, off_consts;  // plane offsets of opening_wires and of the gate constants
  u64 pih[4];
  int r;
  u64 alpha[P2V_MAX_CHALLENGES];
  u64 gpow[P2V_MAX_CHALLENGES];  // alpha_j^(index of the next constraint)
  gl2 gacc[P2V_MAX_CHALLENGES];  // running sum_i alpha_j^i c_i of the current scope

  __device__ __forceinline__ gl2 ext_at(int off) const { return gl2_make(pp[(size_t)off * n + p], pp[(size_t)(off + 1) * n + p]); }
  __device__ __forceinline__ gl2 wire(int i) const { return ext_at(off_wires + 2 * i); }
  __device__ __forceinline__ gl2 cnst(int i) const { return ext_at(off_consts + 2 * i); }
  __device__ __forceinline__ gl4 wireExt(int i) const { return gl4_make(wire(i), wire(i + 1)); }
  __device__ __forceinline__ void commit(gl2 x) {
#pragma unroll
    for (int j = 0; j < P2V_MAX_CHALLENGES; j++)
      if (j < r) {
        gacc[j] = gl2_add(gacc[j], gl2_scale(gpow[j], x));
        gpow[j] = gl_mul(gpow[j], alpha[j]);
      }
  }
  __device__ __forceinline__ void commitExt(gl4 x) { commit(x.r); commit(x.i); }
};

__device__ __forceinline__ gl2 gate_sbox(gl2 x) {  // Gate/Custom/Poseidon.hs:23-30
  gl2 x2 = gl2_sqr(x), x3 = gl2_mul(x, x2), x4 = gl2_sqr(x2);
  return gl2_mul(x3, x4);
}

// state <- MDS * state  (mds, Gate/Custom/Poseidon.hs:111-115)
__device__ __forceinline__ void gate_mds(gl2 (&st)[12]) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
  gl2 o[12];
#pragma unroll
  for (int i = 0; i < 12; i++) {
    gl2 acc = gl2_make(0, 0);
#pragma unroll
    for (int j = 0; j < 12; j++) {
      u32 cf = C[(j - i + 12) % 12] + ((i == 0 && j == 0) ? 8u : 0u);
      acc = gl2_add(acc, gl2_mul_small(st[j], cf));
    }
    o[i] = acc;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = o[i];
}

// poseidonGateConstraints, Gate/Custom/Poseidon.hs:63-150
__device__ __noinline__ void gate_poseidon(ConstraintCtx &g) {
  gl2 swap = g.wire(24);
  g.commit(gl2_mul(swap, gl2_sub_base(swap, 1)));
#pragma unroll 1
  for (int i = 0; i < 4; i++) g.commit(gl2_sub(gl2_mul(swap, gl2_sub(g.wire(i + 4), g.wire(i))), g.wire(25 + i)));
  gl2 st[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    gl2 d = g.wire(25 + i);
    st[i] = gl2_add(g.wire(i), d);
    st[i + 4] = gl2_sub(g.wire(i + 4), d);
    st[i + 8] = g.wire(i + 8);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = gl2_add_base(st[i], c_pt.rc[r][i]);
    if (r != 0) {
#pragma unroll
      for (int i = 0; i < 12; i++) {
        gl2 sin = g.wire(29 + 12 * (r - 1) + i);
        g.commit(gl2_sub(st[i], sin));
        st[i] = sin;
      }
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = gate_sbox(st[i]);
    gate_mds(st);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) st[i] = gl2_add_base(st[i], c_pt.first_rc[i]);
  {  // mdsInitPartial :121-125 ; register-rotating loop keeps indices static
    gl2 acc[11];
#pragma unroll
    for (int i = 0; i < 11; i++) acc[i] = gl2_make(0, 0);
#pragma unroll 1
    for (int j = 0; j < 11; j++) {
      gl2 x = st[1];
#pragma unroll
      for (int i = 1; i < 11; i++) st[i] = st[i + 1];
      st[11] = x;
#pragma unroll
      for (int i = 0; i < 11; i++) acc[i] = gl2_add(acc[i], gl2_scale(c_pt.init_mat[j][i], x));
    }
#pragma unroll
    for (int i = 0; i < 11; i++) st[i + 1] = acc[i];
  }
#pragma unroll 1
  for (int r = 0; r < 22; r++) {
    gl2 sin = g.wire(29 + 36 + r);
    g.commit(gl2_sub(st[0], sin));
    gl2 y = gate_sbox(sin);
    y = gl2_add_base(y, c_pt.partial_rc[r]);  // entry 21 is 0  (:97)
    gl2 d = gl2_mul_small(y, 25u);
#pragma unroll
    for (int i = 0; i < 11; i++) {
      d = gl2_add(d, gl2_scale(c_pt.w_hats[r][i], st[i + 1]));
      st[i + 1] = gl2_add(st[i + 1], gl2_scale(c_pt.vs[r][i], y));
    }
    st[0] = d;
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 12; i++) {
      gl2 sin = g.wire(29 + 36 + 22 + 12 * r + i);
      g.commit(gl2_sub(gl2_add_base(st[i], c_pt.rc[26 + r][i]), sin));
      st[i] = gate_sbox(sin);
    }
    gate_mds(st);
  }
#pragma unroll
  for (int i = 0; i < 12; i++) g.commit(gl2_sub(st[i], g.wire(12 + i)));
}

// poseidonMdsGateConstraints, Gate/Custom/Poseidon.hs:49-59
__device__ __noinline__ void gate_poseidon_mds(ConstraintCtx &g) {
  constexpr u32 C[12] = POSEIDON_MDS_ROW;
#pragma unroll 1
  for (int i = 0; i < 12; i++) {
    gl4 acc = gl4_make(gl2_make(0, 0), gl2_make(0, 0));
#pragma unroll 1
    for (int j = 0; j < 12; j++) {
      u32 cf = C[(j - i + 12) % 12] + ((i == 0 && j == 0) ? 8u : 0u);
      gl4 in = g.wireExt(2 * j);
      acc = gl4_add(acc, gl4_make(gl2_mul_small(in.r, cf), gl2_mul_small(in.i, cf)));
    }
    g.commitExt(gl4_sub(g.wireExt(2 * (i + 12)), acc));
  }
}

// cosetInterpolationGateConstraints, Gate/Custom/CosetInterp.hs:51-121
__device__ __noinline__ void gate_coset_interp(ConstraintCtx &g, int bits, int degree, const u64 *__restrict__ weights, int nweights) {
  int n_points = 1 << bits;
  int n_int = (n_points - 2) / (degree - 1);
  gl2 shift = g.wire(0);
  int base = 1 + 2 * (n_points + 2);
  gl4 shifted_loc = g.wireExt(base + 4 * n_int);
  g.commitExt(gl4_sub(g.wireExt(1 + 2 * n_points), gl4_scale(shift, shifted_loc)));
  // subgroup generator of order 2^bits: rootsOfUnity!bits = twoAdicGen^(2^(32-bits))
  u64 gen = 0x64fdd1a46201e246ULL;
  for (int i = 0; i < 32 - bits; i++) gen = gl_sqr(gen);
  // chunks: first `degree` points, then groups of degree-1 (:121); zip3 truncates to the shortest list
  int npts = n_points < nweights ? n_points : nweights;  // weights list may be shorter/longer than the domain
  auto nchunks_of = [&](int total) { return 1 + (total > degree ? (total - degree + degree - 2) / (degree - 1) : 0); };
  int nchunks = nchunks_of(n_points) < nchunks_of(nweights) ? nchunks_of(n_points) : nchunks_of(nweights);
  int nstuff = nchunks < n_int + 1 ? nchunks : n_int + 1;
  u64 xi = 1;  // domain[idx]
  int idx = 0;
  gl4 eval, prod;
#pragma unroll 1
  for (int ck = 0; ck < nstuff; ck++) {
    if (ck == 0) {
      eval = gl4_make(gl2_make(0, 0), gl2_make(0, 0));
      prod = gl4_make(gl2_make(1, 0), gl2_make(0, 0));
    } else {
      eval = g.wireExt(base + 2 * (ck - 1));
      prod = g.wireExt(base + 2 * (n_int + ck - 1));
    }
    int len = ck == 0 ? degree : degree - 1;
#pragma unroll 1
    for (int k = 0; k < len && idx < npts; k++, idx++) {
      gl4 val = gl4_scale_base(__ldg(weights + idx), g.wireExt(1 + 2 * idx));
      gl4 term = shifted_loc;
      term.r = gl2_sub_base(term.r, xi);
      gl4 ne = gl4_add(gl4_mul(term, eval), gl4_mul(val, prod));
      prod = gl4_mul(term, prod);
      eval = ne;
      xi = gl_mul(xi, gen);
    }
    if (ck + 1 < nstuff) {
      g.commitExt(gl4_sub(g.wireExt(base + 2 * ck), eval));
      g.commitExt(gl4_sub(g.wireExt(base + 2 * (n_int + ck)), prod));
    }
  }
  g.commitExt(gl4_sub(g.wireExt(1 + 2 * n_points + 2), eval));
}

// randomAccessGateConstraints, Gate/Custom/RandomAccess.hs:47-88
__device__ __noinline__ void gate_random_access(ConstraintCtx &g, int nbits, int copies, int extra) {
  int veclen = 1 << nbits, width = 2 + veclen;
  int bits_start = width * copies + extra;
#pragma unroll 1
  for (int k = 0; k < copies; k++) {
#pragma unroll 1
    for (int j = 0; j < nbits; j++) {
      gl2 b = g.wire(bits_start + k * nbits + j);
      g.commit(gl2_mul(b, gl2_sub_base(b, 1)));
    }
    gl2 rec = gl2_make(0, 0);
#pragma unroll 1
    for (int j = nbits - 1; j >= 0; j--) rec = gl2_add(gl2_mul_small(rec, 2), g.wire(bits_start + k * nbits + j));
    g.commit(gl2_sub(rec, g.wire(k * width)));
    // mux tree, evaluated recursively without a value array: node(level, i) over bit `level-1`
    // value(L, i): L == 0 -> inputs[i]; else x + b_{L-1} * (y - x) with x = value(L-1, 2i), y = value(L-1, 2i+1)
    // iterative post-order with an explicit stack of depth nbits+1
    gl2 stack[9];
    int slevel[9];
    int sp = 0;
    for (int leaf = 0; leaf < veclen; leaf++) {
      gl2 v = g.wire(k * width + 2 + leaf);
      int lvl = 0;
      while (sp > 0 && slevel[sp - 1] == lvl) {
        gl2 x = stack[--sp];
        gl2 b = g.wire(bits_start + k * nbits + lvl);
        v = gl2_add(x, gl2_mul(b, gl2_sub(v, x)));
        lvl++;
      }
      stack[sp] = v;
      slevel[sp] = lvl;
      sp++;
    }
    g.commit(gl2_sub(stack[0], g.wire(k * width + 1)));
  }
#pragma unroll 1
  for (int j = 0; j < extra; j++) g.commit(gl2_sub(g.cnst(j), g.wire(copies * width + j)));
}

// reducingGateConstraints / reducingExtensionGateConstraints, Gate/Custom/Reducing.hs:28-60
__device__ __noinline__ void gate_reducing(ConstraintCtx &g, int nc, bool ext) {
  gl4 output = g.wireExt(0), alpha = g.wireExt(2), prev = g.wireExt(4);
  int acc_start = 6 + (ext ? 2 * nc : nc);
#pragma unroll 1
  for (int i = 0; i < nc; i++) {
    gl4 accum = i < nc - 1 ? g.wireExt(acc_start + 2 * i) : output;
    gl4 coeff = ext ? g.wireExt(6 + 2 * i) : gl4_make(g.wire(6 + i), gl2_make(0, 0));
    g.commitExt(gl4_sub(gl4_add(gl4_mul(prev, alpha), coeff), accum));
    prev = accum;
  }
}

// The simple gates of Gate/Constraints.hs:40-128
__device__ __noinline__ void gate_simple(ConstraintCtx &g, const p2v_gate &gt) {
  switch (gt.kind) {
    case P2V_GATE_ARITHMETIC: {  // :45-46
      gl2 c0 = g.cnst(0), c1 = g.cnst(1);
#pragma unroll 1
      for (int i = 0; i < gt.p0; i++) {
        int j = 4 * i;
        g.commit(gl2_sub(gl2_sub(g.wire(j + 3), gl2_mul(gl2_mul(c0, g.wire(j)), g.wire(j + 1))), gl2_mul(c1, g.wire(j + 2))));
      }
      break;
    }
    case P2V_GATE_ARITHMETIC_EXT: {  // :49-54
      gl2 c0 = g.cnst(0), c1 = g.cnst(1);
#pragma unroll 1
      for (int i = 0; i < gt.p0; i++) {
        int j = 8 * i;
        gl4 t = gl4_mul(gl4_scale(c0, g.wireExt(j)), g.wireExt(j + 2));
        g.commitExt(gl4_sub(gl4_sub(g.wireExt(j + 6), t), gl4_scale(c1, g.wireExt(j + 4))));
      }
      break;
    }
    case P2V_GATE_MUL_EXT: {  // :80-83
      gl2 c0 = g.cnst(0);
#pragma unroll 1
      for (int i = 0; i < gt.p0; i++) {
        int j = 6 * i;
        g.commitExt(gl4_sub(g.wireExt(j + 4), gl4_mul(gl4_scale(c0, g.wireExt(j)), g.wireExt(j + 2))));
      }
      break;
    }
    case P2V_GATE_BASE_SUM: {  // :57-62
      int L = gt.p0, B = gt.p1;
      u64 Bf = gl_canon((u64)(u32)B);
      gl2 h = g.wire(L);  // limb L-1
#pragma unroll 1
      for (int k = L - 2; k >= 0; k--) h = gl2_add(g.wire(k + 1), gl2_scale(Bf, h));
      g.commit(gl2_sub(h, g.wire(0)));
#pragma unroll 1
      for (int i = 0; i < L; i++) {
        gl2 limb = g.wire(i + 1);
        gl2 prod = gl2_make(1, 0);
#pragma unroll 1
        for (int k = 0; k < B; k++) prod = gl2_mul(prod, gl2_sub_base(limb, (u64)k));
        g.commit(prod);
      }
      break;
    }
    case P2V_GATE_CONSTANT:  // :68-69
#pragma unroll 1
      for (int i = 0; i < gt.p0; i++) g.commit(gl2_sub(g.cnst(i), g.wire(i)));
      break;
    case P2V_GATE_PUBLIC_INPUT:  // :88-89 ; PIV lifted with fromBase (Computation.hs:211)
#pragma unroll 1
      for (int i = 0; i < 4; i++) g.commit(gl2_sub_base(g.wire(i), g.pih[i]));
      break;
    case P2V_GATE_EXPONENTIATION: {  // :114-128
      int nb = gt.p0;
      gl2 base = g.wire(0);
      gl2 prev = gl2_make(1, 0);
#pragma unroll 1
      for (int i = 0; i < nb; i++) {
        gl2 cur = g.wire(nb - i);  // exp_bit (n-1-i) = wire (n-i)
        gl2 tmp = g.wire(nb + 2 + i);
        gl2 one_minus = gl2_sub(gl2_make(1, 0), cur);
        g.commit(gl2_sub(gl2_mul(prev, gl2_add(gl2_mul(cur, base), one_minus)), tmp));
        prev = gl2_sqr(tmp);
      }
      g.commit(gl2_sub(g.wire(nb + 1), g.wire(nb + 2 + nb - 1)));
      break;
    }
    default: break;  // Noop, Lookup, LookupTable: no constraints (:76-77,85)
  }
}

__device__ __forceinline__ void run_gate(ConstraintCtx &g, const DevCircuit &c, int k) {
  const p2v_gate &gt = c.gates[k];
  switch (gt.kind) {
    case P2V_GATE_POSEIDON: gate_poseidon(g); break;
    case P2V_GATE_POSEIDON_MDS: gate_poseidon_mds(g); break;
    case P2V_GATE_COSET_INTERP: gate_coset_interp(g, gt.p0, gt.p1, c.weights + gt.weights_off, gt.weights_len); break;
    case P2V_GATE_RANDOM_ACCESS: gate_random_access(g, gt.p0, gt.p1, gt.p2); break;
    case P2V_GATE_REDUCING: gate_reducing(g, gt.p0, false); break;
    case P2V_GATE_REDUCING_EXT: gate_reducing(g, gt.p0, true); break;
    default: gate_simple(g, gt); break;
  }
}

// A term of the global constraint list (zs1, pp checks, lookup equations): same accumulator as commit()
// but against the global running powers.
struct TermAcc {
  int r;
  u64 alpha[P2V_MAX_CHALLENGES];
  u64 apow[P2V_MAX_CHALLENGES];
  gl2 total[P2V_MAX_CHALLENGES];
  __device__ __forceinline__ void term(gl2 x) {
#pragma unroll
    for (int j = 0; j < P2V_MAX_CHALLENGES; j++)
      if (j < r) {
        total[j] = gl2_add(total[j], gl2_scale(apow[j], x));
        apow[j] = gl_mul(apow[j], alpha[j]);
      }
  }
};

// evalLookupEquations, Plonk/Lookups.hs:45-132 (only when the circuit has lookup tables)
__device__ __noinline__ void lookup_equations(const DevCircuit &c, const u64 *__restrict__ pp, const u64 *__restrict__ ch, size_t n, size_t p,
                                              TermAcc &T) {
  const p2v_layout &L = c.L;
  auto ext_at = [&](int off) { return gl2_make(pp[(size_t)off * n + p], pp[(size_t)(off + 1) * n + p]); };
  auto wire = [&](int i) { return ext_at(L.off_open_wires + 2 * i); };
  int off_lsel = L.off_open_constants + 2 * c.num_groups;
  auto selector = [&](int k) { return ext_at(off_lsel + 2 * k); };
  int nlp = c.num_lookup_polys;
  int num_lu_slots = c.num_routed / 2, num_lut_slots = c.num_routed / 3;
  int num_sldc = nlp - 1;
  int lu_degree = c.qdf - 1;
  int lut_degree = (num_lut_slots + num_sldc - 1) / num_sldc;
  // the wires are chunked over ALL opening wires (partition 2 / partition 3 of opening_wires), only
  // complete chunks match the list patterns [inp,out] / [inp,out,mult]
  int n_lu = num_lu_slots < c.num_wires / 2 ? num_lu_slots : c.num_wires / 2;
  int n_lut = num_lut_slots < c.num_wires / 3 ? num_lut_slots : c.num_wires / 3;
#pragma unroll 1
  for (int rd = 0; rd < c.r; rd++) {
    u64 dA = ch[(size_t)(c.ch_deltas + 4 * rd + 0) * n + p], dB = ch[(size_t)(c.ch_deltas + 4 * rd + 1) * n + p];
    u64 dAlpha = ch[(size_t)(c.ch_deltas + 4 * rd + 2) * n + p], dDelta = ch[(size_t)(c.ch_deltas + 4 * rd + 3) * n + p];
    int col0 = rd * nlp;  // this round's columns: re, then sldc[0..num_sldc)
    gl2 re = ext_at(L.off_open_lookup_zs + 2 * col0), re_next = ext_at(L.off_open_lookup_zs_next + 2 * col0);
    auto sldc = [&](int i) { return ext_at(L.off_open_lookup_zs + 2 * (col0 + 1 + i)); };
    auto sldc_next = [&](int i) { return ext_at(L.off_open_lookup_zs_next + 2 * (col0 + 1 + i)); };
    T.term(gl2_mul(selector(3), sldc(num_sldc - 1)));  // eq_last_sldc
    T.term(gl2_mul(selector(2), sldc(0)));              // eq_ini_sum
    T.term(gl2_mul(selector(2), re));                   // eq_ini_re
#pragma unroll 1
    for (int k = 0; k < c.num_luts; k++) {              // eq_finals_re
      int len = c.lut_off[k + 1] - c.lut_off[k];
      int rows = (len + num_lut_slots - 1) / num_lut_slots;
      int padded = rows * num_lut_slots;
      u64 cur = 0;
#pragma unroll 1
      for (int t = 0; t < padded; t++) {
        int e = c.lut_off[k] + (t < len ? t : 0);  // padding repeats the first entry (:107)
        u64 inp = __ldg(c.lut_pairs + 2 * e), out = __ldg(c.lut_pairs + 2 * e + 1);
        cur = gl_add(gl_mul(dDelta, cur), gl_add(gl_canon(inp), gl_mul(dB, out)));
      }
      T.term(gl2_mul(selector(4 + k), gl2_sub_base(re, cur)));
    }
    {  // eq_re_trans
      gl2 cur_sum = re_next;
#pragma unroll 1
      for (int i = 0; i < n_lut; i++) {
        gl2 combo = gl2_add(wire(3 * i), gl2_scale(dB, wire(3 * i + 1)));
        cur_sum = gl2_add(gl2_scale(dDelta, cur_sum), combo);
      }
      T.term(gl2_mul(selector(0), gl2_sub(re, cur_sum)));
    }
    // eqs_sldc: zip (pairs (last sldc_next : sldc)) (zip3 lu chunks, lut chunks, mult chunks)
    int n_lu_chunks = (n_lu + lu_degree - 1) / lu_degree;
    int n_lut_chunks = (n_lut + lut_degree - 1) / lut_degree;
    int n_mult_chunks = (num_lut_slots + lut_degree - 1) / lut_degree;
    int nz = num_sldc;
    if (n_lu_chunks < nz) nz = n_lu_chunks;
    if (n_lut_chunks < nz) nz = n_lut_chunks;
    if (n_mult_chunks < nz) nz = n_mult_chunks;
#pragma unroll 1
    for (int t = 0; t < nz; t++) {
      gl2 prev = t == 0 ? sldc_next(num_sldc - 1) : sldc(t - 1);
      gl2 cur = sldc(t);
      int lu0 = t * lu_degree, lu1 = lu0 + lu_degree < n_lu ? lu0 + lu_degree : n_lu;
      int lt0 = t * lut_degree, lt1 = lt0 + lut_degree < n_lut ? lt0 + lut_degree : n_lut;
      int m1 = lt0 + lut_degree < num_lut_slots ? lt0 + lut_degree : num_lut_slots;
      auto lu_factor = [&](int i) { return gl2_sub(gl2_make(dAlpha, 0), gl2_add(wire(2 * i), gl2_scale(dA, wire(2 * i + 1)))); };
      auto lut_factor = [&](int i) { return gl2_sub(gl2_make(dAlpha, 0), gl2_add(wire(3 * i), gl2_scale(dA, wire(3 * i + 1)))); };
      gl2 lu_prod = gl2_make(1, 0), lu_sum = gl2_make(0, 0);
#pragma unroll 1
      for (int i = lu0; i < lu1; i++) lu_prod = gl2_mul(lu_prod, lu_factor(i));
#pragma unroll 1
      for (int i = lu0; i < lu1; i++) {
        gl2 pr = gl2_make(1, 0);
#pragma unroll 1
        for (int k = lu0; k < lu1; k++)
          if (k != i) pr = gl2_mul(pr, lu_factor(k));
        lu_sum = gl2_add(lu_sum, pr);
      }
      gl2 lut_prod = gl2_make(1, 0), lut_sum = gl2_make(0, 0);
#pragma unroll 1
      for (int i = lt0; i < lt1; i++) lut_prod = gl2_mul(lut_prod, lut_factor(i));
      int nm = (m1 - lt0) < (lt1 - lt0) ? (m1 - lt0) : (lt1 - lt0);  // zip mults (remove1 lut_combos)
#pragma unroll 1
      for (int i = lt0; i < lt0 + nm; i++) {
        gl2 pr = wire(3 * i + 2);
#pragma unroll 1
        for (int k = lt0; k < lt1; k++)
          if (k != i) pr = gl2_mul(pr, lut_factor(k));
        lut_sum = gl2_add(lut_sum, pr);
      }
      gl2 diff = gl2_sub(cur, prev);
      T.term(gl2_mul(selector(0), gl2_sub(gl2_mul(lut_prod, diff), lut_sum)));  // eq_sum_trans
      T.term(gl2_mul(selector(1), gl2_add(gl2_mul(lu_prod, diff), lu_sum)));    // eq_ldc_trans
    }
  }
}

__global__ void __launch_bounds__(128) k_constraints(const __grid_constant__ DevCircuit c, Workspace ws, size_t n) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const u64 *__restrict__ pp = ws.pp;
  const p2v_layout &L = c.L;
  auto ext_at = [&](int off) { return gl2_make(pp[(size_t)off * n + p], pp[(size_t)(off + 1) * n + p]); };
  TermAcc T;
  T.r = c.r;
#pragma unroll
  for (int j = 0; j < P2V_MAX_CHALLENGES; j++) {
    T.alpha[j] = j < c.r ? ws.ch[(size_t)(c.ch_alphas + j) * n + p] : 0;
    T.apow[j] = 1;
    T.total[j] = gl2_make(0, 0);
  }
  gl2 zeta = gl2_make(ws.ch[(size_t)c.ch_zeta * n + p], ws.ch[(size_t)(c.ch_zeta + 1) * n + p]);
  // zeta^N by degree_bits squarings (powExt_ zeta nn)
  gl2 zeta_n = zeta;
#pragma unroll 1
  for (int i = 0; i < c.degree_bits; i++) zeta_n = gl2_sqr(zeta_n);
  // zs1 = L0(zeta) * (z - 1), evalLagrange0 (Algebra/Poly.hs:14-17)
  {
    gl2 l0;
    if (gl2_eq(zeta, gl2_make(1, 0))) l0 = gl2_make(1, 0);
    else {
      u64 nn = (u64)1 << c.degree_bits;
      l0 = gl2_mul(gl2_sub_base(zeta_n, 1), gl2_inv(gl2_scale(nn, gl2_sub_base(zeta, 1))));
    }
#pragma unroll 1
    for (int j = 0; j < L.n_open_zs; j++) T.term(gl2_mul(l0, gl2_sub_base(ext_at(L.off_open_zs + 2 * j), 1)));
  }
  // partial-product checks (Vanishing.hs:98-111)
  if (c.num_pp > 0) {
    int n_chunks_pp = L.n_open_pp / c.num_pp + (L.n_open_pp % c.num_pp ? 1 : 0);  // partition num_pp partial_products
    int nrounds = c.r < n_chunks_pp ? c.r : n_chunks_pp;
    int nk = c.num_routed < c.num_wires ? c.num_routed : c.num_wires;             // zipWith truncation (:107-108)
    int nwchunks = (nk + c.qdf - 1) / c.qdf;
#pragma unroll 1
    for (int rd = 0; rd < nrounds; rd++) {
      u64 beta = ws.ch[(size_t)(c.ch_betas + rd) * n + p], gamma = ws.ch[(size_t)(c.ch_gammas + rd) * n + p];
      int chunk_len = c.num_pp < L.n_open_pp - rd * c.num_pp ? c.num_pp : L.n_open_pp - rd * c.num_pp;
      int np = chunk_len + 1 < nwchunks ? chunk_len + 1 : nwchunks;  // zipWith3 (pairs current) numers denoms
      gl2 prev = ext_at(L.off_open_zs + 2 * rd);
#pragma unroll 1
      for (int t = 0; t < np; t++) {
        gl2 next = t < chunk_len ? ext_at(L.off_open_pp + 2 * (rd * c.num_pp + t)) : ext_at(L.off_open_zs_next + 2 * rd);
        gl2 pn = gl2_make(1, 0), pd = gl2_make(1, 0);
        int hi = (t + 1) * c.qdf < nk ? (t + 1) * c.qdf : nk;
#pragma unroll 1
        for (int i = t * c.qdf; i < hi; i++) {
          gl2 w = ext_at(L.off_open_wires + 2 * i);
          gl2 sg = ext_at(L.off_open_sigmas + 2 * i);
          u64 bk = gl_mul(beta, __ldg(c.k_is + i));
          pn = gl2_mul(pn, gl2_add_base(gl2_add(w, gl2_scale(bk, zeta)), gamma));
          pd = gl2_mul(pd, gl2_add_base(gl2_add(w, gl2_scale(beta, sg)), gamma));
        }
        T.term(gl2_sub(gl2_mul(prev, pn), gl2_mul(next, pd)));
        prev = next;
      }
    }
  }
  // lookup equations (skipped when luts is empty, Vanishing.hs:74-76)
  if (c.num_luts > 0) lookup_equations(c, pp, ws.ch, n, p, T);
  // gates: sum_g filter_g * sum_i alpha^(K+i) c_{g,i}   (Vanishing.hs:87-94,124-125)
  {
    ConstraintCtx g;
    g.pp = pp; g.n = n; g.p = p;
    g.off_wires = L.off_open_wires;
    g.off_consts = L.off_open_constants + 2 * (c.num_groups + c.num_lookup_sel);  // splitConstantColumns (Selector.hs:72-74)
    g.r = c.r;
#pragma unroll
    for (int i = 0; i < 4; i++) g.pih[i] = ws.pih[(size_t)i * n + p];
#pragma unroll
    for (int j = 0; j < P2V_MAX_CHALLENGES; j++) g.alpha[j] = T.alpha[j];
#pragma unroll 1
    for (int k = 0; k < c.num_gates; k++) {
#pragma unroll
      for (int j = 0; j < P2V_MAX_CHALLENGES; j++) {
        g.gpow[j] = T.apow[j];
        g.gacc[j] = gl2_make(0, 0);
      }
      run_gate(g, c, k);
      // evalGateSelectorPoly, Selector.hs:83-89
      int grp = c.gates[k].group;
      gl2 x = ext_at(L.off_open_constants + 2 * grp);
      gl2 f = c.num_groups > 1 ? gl2_sub(gl2_make(0xFFFFFFFFULL, 0), x) : gl2_make(1, 0);
#pragma unroll 1
      for (int j = c.group_start[grp]; j < c.group_end[grp]; j++)
        if (j != k) f = gl2_mul(f, gl2_sub(gl2_make((u64)j, 0), x));
#pragma unroll
      for (int j = 0; j < P2V_MAX_CHALLENGES; j++)
        if (j < c.r) T.total[j] = gl2_add(T.total[j], gl2_mul(f, g.gacc[j]));
    }
  }
  // quotient identity, Plonk/Verifier.hs:44-52: (sum_k zeta^{nk} q_k) * (zeta^n - 1) == combined_j
  u32 okmask = 0;
  gl2 zh = gl2_sub_base(zeta_n, 1);
#pragma unroll
  for (int j = 0; j < P2V_MAX_CHALLENGES; j++)
    if (j < c.r) {
      gl2 q = gl2_make(0, 0);
      int lo = j * c.qdf, hi = (j + 1) * c.qdf < L.n_open_quotient ? (j + 1) * c.qdf : L.n_open_quotient;
#pragma unroll 1
      for (int k = hi - 1; k >= lo; k--) q = gl2_add(ext_at(L.off_open_quotient + 2 * k), gl2_mul(zeta_n, q));
      gl2 tot = gl2_canon(T.total[j]);
      if (gl2_eq(gl2_mul(q, zh), tot)) okmask |= 1u << j;
      if (ws.comb) {
        ws.comb[(size_t)(2 * j) * n + p] = tot.a;
        ws.comb[(size_t)(2 * j + 1) * n + p] = tot.b;
      }
    }
  ws.eqmask[p] = (uint8_t)okmask;
}
