// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the
// product path (libp2v.so / plonky2-verifier_b200).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the reference's field arithmetic:
//   src/Algebra/Goldilocks.hs     (F = Z/p, p = 2^64 - 2^32 + 1, always canonical)
//   src/Algebra/GoldilocksExt.hs  (Ext a = a[X]/(X^2 - 7))
//   src/Algebra/FFT.hs:20-45      (bit reversal, powersOf)
//   src/Algebra/Poly.hs           (L_0, Z_H)
// Parity status: the reference has no tests for these modules ("parity unpinned" by the
// reference itself); pinned here transitively by the Poseidon KAT (Hash/Poseidon.hs:27-32)
// and by the independent Python twin oracle/pyref.py.
#pragma once
#include <cstdint>
#include <vector>
#include <stdexcept>

namespace orc {

typedef unsigned __int128 u128;
typedef uint64_t u64;

// goldilocksPrime, Algebra/Goldilocks.hs:125-126
static const u64 P = 0xFFFFFFFF00000001ULL;

// `newtype Goldilocks = Goldilocks Integer`, invariant 0 <= x < p (mkGoldilocks :132-133)
struct F {
  u64 v;
  F() : v(0) {}
  F(u64 x) : v(x >= P ? x - P : x) {}  // mkGoldilocks for inputs < 2^64 (2^64 < 2p)
  static F fromInt(long long k) {      // intToF / fromInteger for small signed ints
    if (k >= 0) return F((u64)k);
    return F(P - (u64)(-k));
  }
  bool operator==(const F &o) const { return v == o.v; }
  bool operator!=(const F &o) const { return v != o.v; }
};

// add/sub/neg/mul: Algebra/Goldilocks.hs:140-153 (Integer op followed by `mod p`)
inline F operator+(F a, F b) {
  u128 s = (u128)a.v + b.v;
  if (s >= P) s -= P;
  F r; r.v = (u64)s; return r;
}
inline F operator-(F a, F b) {
  F r; r.v = a.v >= b.v ? a.v - b.v : a.v + (P - b.v); return r;
}
inline F neg(F a) { F r; r.v = a.v ? P - a.v : 0; return r; }
inline F operator*(F a, F b) {
  u128 m = (u128)a.v * b.v;
  F r; r.v = (u64)(m % P); return r;
}
inline F sqr(F a) { return a * a; }

// pow, Algebra/Goldilocks.hs:166-175 (non-negative exponent part); inv = x^(p-2) :155 (so inv 0 = 0)
inline F powu(F x, u64 e) {
  F acc(1), s = x;
  while (e) {
    if (e & 1) acc = acc * s;
    s = sqr(s);
    e >>= 1;
  }
  return acc;
}
inline F inv(F x) { return powu(x, P - 2); }
// negative exponents invert first, Algebra/Goldilocks.hs:169
inline F powi(F x, long long e) {
  if (e == 0) return F(1);
  if (e < 0) return powu(inv(x), (u64)(-e));
  return powu(x, (u64)e);
}
inline F operator/(F a, F b) { return a * inv(b); }

// mulGen / multGen :48-49,135 ; twoAdicGen :54-55
static const u64 MUL_GEN = 0xc65c18b67785d900ULL;
static const u64 TWO_ADIC_GEN = 0x64fdd1a46201e246ULL;

// rootsOfUnity / subgroupGenerator, Algebra/Goldilocks.hs:68-74:
//   list = reverse (go twoAdicGen), go x = x : go (x*x), go 1 = [1]  => rootsOfUnity!k has order 2^k
inline F subgroupGenerator(int k) {
  if (k < 0 || k > 32) throw std::runtime_error("subgroupGenerator: out of range");
  F x(TWO_ADIC_GEN);
  for (int i = 0; i < 32 - k; i++) x = sqr(x);
  return x;
}
// enumerateSubgroup :76-79
inline std::vector<F> enumerateSubgroup(int logSize) {
  F g = subgroupGenerator(logSize);
  std::vector<F> out;
  F x(1);
  for (int i = 0; i < (1 << logSize); i++) { out.push_back(x); x = x * g; }
  return out;
}

// ---------------------------------------------------------------------------------------
// `data Ext a = MkExt !a !a`, Algebra/GoldilocksExt.hs:28-31, generic over the coefficient ring
template <class A> struct ExtT {
  A r, i;
  ExtT() : r(), i() {}
  ExtT(A a, A b) : r(a), i(b) {}
  bool operator==(const ExtT &o) const { return r == o.r && i == o.i; }
  bool operator!=(const ExtT &o) const { return !(*this == o); }
};
typedef ExtT<F> FExt;

inline F seven() { return F(7); }

// fromBase :33-34
inline FExt fromBase(F x) { return FExt(x, F(0)); }
// Num instance :54-61
template <class A> inline ExtT<A> operator+(const ExtT<A> &a, const ExtT<A> &b) { return ExtT<A>(a.r + b.r, a.i + b.i); }
template <class A> inline ExtT<A> operator-(const ExtT<A> &a, const ExtT<A> &b) { return ExtT<A>(a.r - b.r, a.i - b.i); }
inline FExt operator*(const FExt &a, const FExt &b) {
  // (r1*r2 + 7*i1*i2 , r1*i2 + r2*i1)  GoldilocksExt.hs:59
  return FExt(a.r * b.r + seven() * a.i * b.i, a.r * b.i + b.r * a.i);
}
inline FExt negE(const FExt &a) { return FExt(neg(a.r), neg(a.i)); }
// scaleExt :70-71
inline FExt scaleExt(F s, const FExt &a) { return FExt(s * a.r, s * a.i); }
inline FExt sqrExt(const FExt &a) { return a * a; }
// invExt :75-80   denom = recip (a*a - 7*b*b)
inline FExt invExt(const FExt &x) {
  F denom = inv(x.r * x.r - seven() * x.i * x.i);
  return FExt(x.r * denom, neg(x.i) * denom);
}
inline FExt operator/(const FExt &a, const FExt &b) { return a * invExt(b); }
// powExt :89-99
inline FExt powExtU(FExt x, u64 e) {
  FExt acc(F(1), F(0)), s = x;
  while (e) {
    if (e & 1) acc = acc * s;
    s = sqrExt(s);
    e >>= 1;
  }
  return acc;
}
inline FExt powExtI(FExt x, long long e) {
  if (e == 0) return FExt(F(1), F(0));
  if (e < 0) return powExtU(invExt(x), (u64)(-e));
  return powExtU(x, (u64)e);
}
// flattenExt :102-106
inline std::vector<F> flattenExt(const std::vector<FExt> &xs) {
  std::vector<F> out;
  for (auto &x : xs) { out.push_back(x.r); out.push_back(x.i); }
  return out;
}

// "ext of ext": the gates multiply `Ext (Expr v)` values whose components evaluate to FExt
// (Gate/Vars.hs:56-57, Algebra/GoldilocksExt.hs:59 instantiated at a = Expr, evaluated by
// Algebra/Expr.hs:130-138).  Numerically that is ExtT<FExt> with the same product rule.
typedef ExtT<FExt> EExt;
inline EExt operator*(const EExt &a, const EExt &b) {
  FExt sev = fromBase(F(7));
  return EExt(a.r * b.r + sev * a.i * b.i, a.r * b.i + b.r * a.i);
}
inline EExt scaleEE(const FExt &s, const EExt &a) { return EExt(s * a.r, s * a.i); }
inline EExt fromBaseEE(const FExt &x) { return EExt(x, FExt()); }

// ---------------------------------------------------------------------------------------
// reverseBits / reverseBitsInt, Algebra/FFT.hs:20-28
inline u64 reverseBits(int n, u64 w) {
  u64 r = 0;
  for (int k = 0; k < n; k++) r |= ((w >> k) & 1) << (n - k - 1);
  return r;
}
inline int reverseBitsInt(int n, int w) { return (int)reverseBits(n, (u64)w); }
// reverseIndexBitsList, Algebra/FFT.hs:30-41: arr2[rev i] = arr1[i]
template <class T> inline std::vector<T> reverseIndexBitsList(const std::vector<T> &xs) {
  size_t n = xs.size();
  int k = 0;
  while ((size_t(1) << k) < n) k++;
  if ((size_t(1) << k) != n) throw std::runtime_error("safeLog2: input is not a power of two");
  std::vector<T> out(n);
  for (size_t i = 0; i < n; i++) out[reverseBits(k, i)] = xs[i];
  return out;
}

// reduceWithPowers, Algebra/Goldilocks.hs:179-183:  go (x:xs) = x + alpha * go xs
inline FExt reduceWithPowers(const FExt &alpha, const std::vector<FExt> &xs) {
  FExt acc;
  for (size_t k = xs.size(); k-- > 0;) acc = xs[k] + alpha * acc;
  return acc;
}

// evalLagrange0 / evalZeroPoly, Algebra/Poly.hs:13-22
inline FExt evalLagrange0(u64 nn, const FExt &zeta) {
  FExt one(F(1), F(0));
  if (zeta == one) return one;
  return (powExtU(zeta, nn) - one) / (fromBase(F(nn)) * (zeta - one));
}
inline FExt evalZeroPoly(u64 nn, const FExt &zeta) { return powExtU(zeta, nn) - FExt(F(1), F(0)); }

}  // namespace orc
