// ec.cuh -- short-Weierstrass (a = 0) group law for the Pasta and BN254/Grumpkin cycles, sm_100a.
//
// Bucket accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity ZZ = 0):
// a mixed add with an affine base costs 8M + 2S and never needs an inversion.  Affine inputs are the
// 64-byte {x, y} Montgomery pairs of halo2curves / pasta_curves (identity encoded as (0, 0));
// results leave the library as Jacobian {X, Y, Z} (96 bytes), the in-memory layout of
// `pallas::Point` / `bn256::G1`.
//
// Replaces: the Jacobian bucket arithmetic inside nova-snark 0.23.0 `cpu_best_multiexp` and
// pasta-msm's Pippenger, reached from CommitmentEngine::commit.
#pragma once
#include "fp.cuh"

namespace vimz {

template <class C>
struct Affine {
  using F = Fp<typename C::Fb>;
  F x, y;
  VIMZ_DI bool is_identity() const { return x.is_zero() && y.is_zero(); }
  VIMZ_DI static Affine load(const void* p) {
    Affine a;
    a.x = F::load(p);
    a.y = F::load(reinterpret_cast<const char*>(p) + 32);
    return a;
  }
  VIMZ_DI static Affine load_nc(const void* p) {
    Affine a;
    a.x = F::load_nc(p);
    a.y = F::load_nc(reinterpret_cast<const char*>(p) + 32);
    return a;
  }
  VIMZ_DI void store(void* p) const {
    x.store(p);
    y.store(reinterpret_cast<char*>(p) + 32);
  }
};

template <class C>
struct Xyzz {
  using F = Fp<typename C::Fb>;
  F x, y, zz, zzz;
  VIMZ_DI static Xyzz identity() {
    Xyzz r;
    r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero();
    return r;
  }
  VIMZ_DI bool is_identity() const { return zz.is_zero(); }
  VIMZ_DI static Xyzz from_affine(const Affine<C>& a) {
    Xyzz r;
    if (a.is_identity()) return identity();
    r.x = a.x; r.y = a.y; r.zz = F::one(); r.zzz = F::one();
    return r;
  }
  VIMZ_DI static Xyzz load(const void* p) {
    const char* q = reinterpret_cast<const char*>(p);
    Xyzz r;
    r.x = F::load(q); r.y = F::load(q + 32); r.zz = F::load(q + 64); r.zzz = F::load(q + 96);
    return r;
  }
  VIMZ_DI void store(void* p) const {
    char* q = reinterpret_cast<char*>(p);
    x.store(q); y.store(q + 32); zz.store(q + 64); zzz.store(q + 96);
  }
  VIMZ_DI Xyzz neg() const {
    Xyzz r = *this;
    r.y = fp_neg(y);
    return r;
  }
};

// Multiplication policy.  The throughput kernel (bucket accumulation) inlines every field multiplication
// so ptxas can interleave independent products; the latency-bound kernels (bucket reduction trees,
// scalar multiplications, table expansion) instead CALL one shared copy: their warps run alone, and with
// ~3 KB per inlined product an EC addition is ~45 KB of straight-line code that never fits the
// instruction caches (measured: 10 us per addition inlined vs ~3 us called).
template <class F>
struct FpPair { Fp<F> a, b; };
template <class F>
__device__ __noinline__ Fp<F> fp_mul_call(Fp<F> a, Fp<F> b) { return fp_mul(a, b); }
// two independent products in one call: a lone warp is bound by the dependent-issue latency of one
// Montgomery chain, two interleaved chains nearly halve the time per product
template <class F>
__device__ __noinline__ FpPair<F> fp_mul2_call(Fp<F> a1, Fp<F> b1, Fp<F> a2, Fp<F> b2) {
  FpPair<F> r;
  fp_mul2(r.a, a1, b1, r.b, a2, b2);
  return r;
}
struct MulInline {
  template <class F> static VIMZ_DI Fp<F> mul(const Fp<F>& a, const Fp<F>& b) { return fp_mul(a, b); }
  template <class F> static VIMZ_DI void mul2(Fp<F>& r1, const Fp<F>& a1, const Fp<F>& b1, Fp<F>& r2, const Fp<F>& a2, const Fp<F>& b2) {
#ifdef VIMZ_INLINE_MUL2_INTERLEAVED   // A/B: rows of the two products interleaved in program order (more registers live)
    fp_mul2(r1, a1, b1, r2, a2, b2);
#else
    r1 = fp_mul(a1, b1);
    r2 = fp_mul(a2, b2);
#endif
  }
  template <class F> static VIMZ_DI void sqr2(Fp<F>& r1, const Fp<F>& a1, Fp<F>& r2, const Fp<F>& a2) {
    r1 = fp_sqr(a1);
    r2 = fp_sqr(a2);
  }
};
struct MulCall {
  template <class F> static VIMZ_DI Fp<F> mul(const Fp<F>& a, const Fp<F>& b) { return fp_mul_call<F>(a, b); }
  template <class F> static VIMZ_DI void mul2(Fp<F>& r1, const Fp<F>& a1, const Fp<F>& b1, Fp<F>& r2, const Fp<F>& a2, const Fp<F>& b2) {
    FpPair<F> r = fp_mul2_call<F>(a1, b1, a2, b2);
    r1 = r.a;
    r2 = r.b;
  }
  template <class F> static VIMZ_DI void sqr2(Fp<F>& r1, const Fp<F>& a1, Fp<F>& r2, const Fp<F>& a2) { mul2(r1, a1, a1, r2, a2, a2); }
};

// 2 * (affine point) -> XYZZ   ("mdbl-2008-s-1", a = 0)
template <class C, class M = MulInline>
VIMZ_DI Xyzz<C> xyzz_dbl_affine(const Affine<C>& p) {
  using F = Fp<typename C::Fb>;
  Xyzz<C> r;
  if (p.y.is_zero()) return Xyzz<C>::identity();  // identity (0,0); no finite point has y = 0 on these curves
  F u = fp_dbl(p.y);
  F v, xx, w, s, mm, wy;
  M::mul2(v, u, u, xx, p.x, p.x);
  M::mul2(w, u, v, s, p.x, v);
  F m = fp_add(fp_dbl(xx), xx);
  M::mul2(mm, m, m, wy, w, p.y);
  r.x = fp_sub(mm, fp_dbl(s));
  r.y = fp_sub(M::mul(m, fp_sub(s, r.x)), wy);
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * XYZZ ("dbl-2008-s-1", a = 0)
template <class C, class M = MulInline>
VIMZ_DI Xyzz<C> xyzz_dbl(const Xyzz<C>& p) {
  using F = Fp<typename C::Fb>;
  if (p.is_identity() || p.y.is_zero()) return Xyzz<C>::identity();
  Xyzz<C> r;
  F u = fp_dbl(p.y);
  F v, xx, w, s, mm, wy, t;
  M::mul2(v, u, u, xx, p.x, p.x);
  M::mul2(w, u, v, s, p.x, v);
  F m = fp_add(fp_dbl(xx), xx);
  M::mul2(mm, m, m, wy, w, p.y);
  r.x = fp_sub(mm, fp_dbl(s));
  M::mul2(t, m, fp_sub(s, r.x), r.zz, v, p.zz);
  r.y = fp_sub(t, wy);
  r.zzz = M::mul(w, p.zzz);
  return r;
}

// acc += (neg ? -q : q), q affine  ("madd-2008-s", 8M + 2S).  All exceptional cases are handled
// (acc = identity, q = identity, q = +-acc) so results are exact for any input, including repeated bases.
template <class C, class M = MulInline>
VIMZ_DI void xyzz_madd(Xyzz<C>& acc, const Affine<C>& q, bool neg) {
  using F = Fp<typename C::Fb>;
  if (q.is_identity()) return;
  F qy = fp_cneg(q.y, neg);
  if (acc.is_identity()) {
    acc.x = q.x; acc.y = qy; acc.zz = F::one(); acc.zzz = F::one();
    return;
  }
  F u2, s2;
  M::mul2(u2, q.x, acc.zz, s2, qy, acc.zzz);
  F p = fp_sub(u2, acc.x);
  F r = fp_sub(s2, acc.y);
  if (p.is_zero()) {
    if (r.is_zero()) {
      Affine<C> t;
      t.x = q.x; t.y = qy;
      acc = xyzz_dbl_affine<C, MulCall>(t);  // rare (q == acc): keep it out of line, out of the hot loop's I-cache footprint
    } else {
      acc = Xyzz<C>::identity();
    }
    return;
  }
  F pp, rr, ppp, qq, t1, t2;
  M::sqr2(pp, p, rr, r);
  M::mul2(ppp, p, pp, qq, acc.x, pp);
  F x3 = fp_sub(fp_sub(rr, ppp), fp_dbl(qq));
  M::mul2(t1, r, fp_sub(qq, x3), t2, acc.y, ppp);
  acc.x = x3;
  acc.y = fp_sub(t1, t2);
  M::mul2(acc.zz, acc.zz, pp, acc.zzz, acc.zzz, ppp);
}

// acc += q, both XYZZ ("add-2008-s", 12M + 2S)
template <class C, class M = MulInline>
VIMZ_DI void xyzz_add(Xyzz<C>& acc, const Xyzz<C>& q) {
  using F = Fp<typename C::Fb>;
  if (q.is_identity()) return;
  if (acc.is_identity()) { acc = q; return; }
  F u1, u2, s1, s2;
  M::mul2(u1, acc.x, q.zz, u2, q.x, acc.zz);
  M::mul2(s1, acc.y, q.zzz, s2, q.y, acc.zzz);
  F p = fp_sub(u2, u1);
  F r = fp_sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_dbl<C, MulCall>(acc);  // rare: out of line
    else acc = Xyzz<C>::identity();
    return;
  }
  F pp, rr, ppp, qq, zzq, zzzq, t1, t2;
  M::mul2(pp, p, p, rr, r, r);
  M::mul2(zzq, acc.zz, q.zz, zzzq, acc.zzz, q.zzz);
  M::mul2(ppp, p, pp, qq, u1, pp);
  F x3 = fp_sub(fp_sub(rr, ppp), fp_dbl(qq));
  M::mul2(t1, r, fp_sub(qq, x3), t2, s1, ppp);
  acc.x = x3;
  acc.y = fp_sub(t1, t2);
  M::mul2(acc.zz, zzq, pp, acc.zzz, zzzq, ppp);
}

// XYZZ -> Jacobian without inversion: Z = ZZ*ZZZ  =>  Z^2 = ZZ^5, Z^3 = ZZZ^5 (using ZZ^3 = ZZZ^2),
// so X_j = X*ZZ^4, Y_j = Y*ZZZ^4.  Identity -> (0, 1, 0) like halo2curves.
template <class C, class M = MulInline>
VIMZ_DI void xyzz_to_jacobian(const Xyzz<C>& p, Fp<typename C::Fb>& X, Fp<typename C::Fb>& Y, Fp<typename C::Fb>& Z) {
  using F = Fp<typename C::Fb>;
  if (p.is_identity()) {
    X = F::zero(); Y = F::one(); Z = F::zero();
    return;
  }
  F zz2 = M::mul(p.zz, p.zz), zzz2 = M::mul(p.zzz, p.zzz);
  X = M::mul(p.x, M::mul(zz2, zz2));
  Y = M::mul(p.y, M::mul(zzz2, zzz2));
  Z = M::mul(p.zz, p.zzz);
}

// XYZZ -> canonical affine (one inversion; identity -> (0,0)).
template <class C>
__device__ __noinline__ Affine<C> xyzz_to_affine(const Xyzz<C>& p) {
  using F = Fp<typename C::Fb>;
  using M = MulCall;
  Affine<C> a;
  if (p.is_identity()) { a.x = F::zero(); a.y = F::zero(); return a; }
  F zi = fp_inv(p.zzz);              // 1/ZZZ
  F t = M::mul(p.zz, zi);
  F zzi = M::mul(t, t);              // (ZZ/ZZZ)^2 = ZZ^2/ZZ^3 = 1/ZZ
  a.x = M::mul(p.x, zzi);
  a.y = M::mul(p.y, zi);
  return a;
}

// Jacobian {X,Y,Z} (Z may be anything, identity Z = 0) -> XYZZ (ZZ = Z^2, ZZZ = Z^3).
template <class C, class M = MulInline>
VIMZ_DI Xyzz<C> xyzz_from_jacobian(const Fp<typename C::Fb>& X, const Fp<typename C::Fb>& Y, const Fp<typename C::Fb>& Z) {
  Xyzz<C> r;
  if (Z.is_zero()) return Xyzz<C>::identity();
  r.x = X; r.y = Y;
  r.zz = M::mul(Z, Z);
  r.zzz = M::mul(r.zz, Z);
  return r;
}

// Group operations for the latency-bound kernels: formulas inlined, every field product a call to the one
// shared fp_mul_call (operands by value, in registers), ~10 KB per addition instead of ~45 KB.
template <class C>
VIMZ_DI void xyzz_add_call(Xyzz<C>& acc, const Xyzz<C>& q) { xyzz_add<C, MulCall>(acc, q); }
template <class C>
VIMZ_DI void xyzz_dbl_call(Xyzz<C>& p) { p = xyzz_dbl<C, MulCall>(p); }
template <class C>
VIMZ_DI void xyzz_madd_call(Xyzz<C>& acc, const Affine<C>& q, bool neg) { xyzz_madd<C, MulCall>(acc, q, neg); }

}  // namespace vimz
