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

// 2 * (affine point) -> XYZZ   ("mdbl-2008-s-1", a = 0)
template <class C>
VIMZ_DI Xyzz<C> xyzz_dbl_affine(const Affine<C>& p) {
  using F = Fp<typename C::Fb>;
  Xyzz<C> r;
  if (p.y.is_zero()) return Xyzz<C>::identity();  // identity (0,0); no finite point has y = 0 on these curves
  F u = fp_dbl(p.y);
  F v = fp_sqr(u);
  F w = fp_mul(u, v);
  F s = fp_mul(p.x, v);
  F xx = fp_sqr(p.x);
  F m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fp_sqr(m), fp_dbl(s));
  r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

// 2 * XYZZ ("dbl-2008-s-1", a = 0)
template <class C>
VIMZ_DI Xyzz<C> xyzz_dbl(const Xyzz<C>& p) {
  using F = Fp<typename C::Fb>;
  if (p.is_identity() || p.y.is_zero()) return Xyzz<C>::identity();
  Xyzz<C> r;
  F u = fp_dbl(p.y);
  F v = fp_sqr(u);
  F w = fp_mul(u, v);
  F s = fp_mul(p.x, v);
  F xx = fp_sqr(p.x);
  F m = fp_add(fp_dbl(xx), xx);
  r.x = fp_sub(fp_sqr(m), fp_dbl(s));
  r.y = fp_sub(fp_mul(m, fp_sub(s, r.x)), fp_mul(w, p.y));
  r.zz = fp_mul(v, p.zz);
  r.zzz = fp_mul(w, p.zzz);
  return r;
}

// acc += (neg ? -q : q), q affine  ("madd-2008-s", 8M + 2S).  All exceptional cases are handled
// (acc = identity, q = identity, q = +-acc) so results are exact for any input, including repeated bases.
template <class C>
VIMZ_DI void xyzz_madd(Xyzz<C>& acc, const Affine<C>& q, bool neg) {
  using F = Fp<typename C::Fb>;
  if (q.is_identity()) return;
  F qy = fp_cneg(q.y, neg);
  if (acc.is_identity()) {
    acc.x = q.x; acc.y = qy; acc.zz = F::one(); acc.zzz = F::one();
    return;
  }
  F u2 = fp_mul(q.x, acc.zz);
  F s2 = fp_mul(qy, acc.zzz);
  F p = fp_sub(u2, acc.x);
  F r = fp_sub(s2, acc.y);
  if (p.is_zero()) {
    if (r.is_zero()) {
      Affine<C> t;
      t.x = q.x; t.y = qy;
      acc = xyzz_dbl_affine<C>(t);
    } else {
      acc = Xyzz<C>::identity();
    }
    return;
  }
  F pp = fp_sqr(p);
  F ppp = fp_mul(p, pp);
  F qq = fp_mul(acc.x, pp);
  F x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
  F y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul(acc.zz, pp);
  acc.zzz = fp_mul(acc.zzz, ppp);
}

// acc += q, both XYZZ ("add-2008-s", 12M + 2S)
template <class C>
VIMZ_DI void xyzz_add(Xyzz<C>& acc, const Xyzz<C>& q) {
  using F = Fp<typename C::Fb>;
  if (q.is_identity()) return;
  if (acc.is_identity()) { acc = q; return; }
  F u1 = fp_mul(acc.x, q.zz);
  F u2 = fp_mul(q.x, acc.zz);
  F s1 = fp_mul(acc.y, q.zzz);
  F s2 = fp_mul(q.y, acc.zzz);
  F p = fp_sub(u2, u1);
  F r = fp_sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) acc = xyzz_dbl<C>(acc);
    else acc = Xyzz<C>::identity();
    return;
  }
  F pp = fp_sqr(p);
  F ppp = fp_mul(p, pp);
  F qq = fp_mul(u1, pp);
  F x3 = fp_sub(fp_sub(fp_sqr(r), ppp), fp_dbl(qq));
  F y3 = fp_sub(fp_mul(r, fp_sub(qq, x3)), fp_mul(s1, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fp_mul(fp_mul(acc.zz, q.zz), pp);
  acc.zzz = fp_mul(fp_mul(acc.zzz, q.zzz), ppp);
}

// XYZZ -> Jacobian without inversion: Z = ZZ*ZZZ  =>  Z^2 = ZZ^5, Z^3 = ZZZ^5 (using ZZ^3 = ZZZ^2),
// so X_j = X*ZZ^4, Y_j = Y*ZZZ^4.  Identity -> (0, 1, 0) like halo2curves.
template <class C>
VIMZ_DI void xyzz_to_jacobian(const Xyzz<C>& p, Fp<typename C::Fb>& X, Fp<typename C::Fb>& Y, Fp<typename C::Fb>& Z) {
  using F = Fp<typename C::Fb>;
  if (p.is_identity()) {
    X = F::zero(); Y = F::one(); Z = F::zero();
    return;
  }
  F zz2 = fp_sqr(p.zz), zzz2 = fp_sqr(p.zzz);
  X = fp_mul(p.x, fp_sqr(zz2));
  Y = fp_mul(p.y, fp_sqr(zzz2));
  Z = fp_mul(p.zz, p.zzz);
}

// XYZZ -> canonical affine (one inversion; identity -> (0,0)).
template <class C>
__device__ __noinline__ Affine<C> xyzz_to_affine(const Xyzz<C>& p) {
  using F = Fp<typename C::Fb>;
  Affine<C> a;
  if (p.is_identity()) { a.x = F::zero(); a.y = F::zero(); return a; }
  F zi = fp_inv(p.zzz);              // 1/ZZZ
  F zzi = fp_sqr(fp_mul(p.zz, zi));  // (ZZ/ZZZ)^2 = ZZ^2/ZZ^3 = 1/ZZ
  a.x = fp_mul(p.x, zzi);
  a.y = fp_mul(p.y, zi);
  return a;
}

// Jacobian {X,Y,Z} (Z may be anything, identity Z = 0) -> XYZZ (ZZ = Z^2, ZZZ = Z^3).
template <class C>
VIMZ_DI Xyzz<C> xyzz_from_jacobian(const Fp<typename C::Fb>& X, const Fp<typename C::Fb>& Y, const Fp<typename C::Fb>& Z) {
  Xyzz<C> r;
  if (Z.is_zero()) return Xyzz<C>::identity();
  r.x = X; r.y = Y;
  r.zz = fp_sqr(Z);
  r.zzz = fp_mul(r.zz, Z);
  return r;
}

}  // namespace vimz
