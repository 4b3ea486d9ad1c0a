// ecq.cuh -- quad-cooperative XYZZ group operations for the LATENCY-bound kernels.
//
// A field multiplication executed by a lone warp takes ~0.4 us (a single dependent carry chain, one
// IMAD.WIDE every ~4 cycles), so an EC addition done by one thread costs ~5.6 us no matter how idle the
// SM is.  The bucket reduction, the oversize-bucket trees and the commitment folds are chains of such
// additions.  Here FOUR adjacent lanes own one point -- lane (l & 3) holds coordinate X, Y, ZZ or ZZZ --
// and the 14 products of an addition are spread over the quad: 4 rounds of one product per lane with
// operands moved by warp shuffles (3 rounds for a doubling).  This is the "warp-cooperative" arrangement
// BASELINE.json asks for, applied where it pays: at EC-operation granularity on the serial tails.
//
// All functions must be called by all 32 lanes of a warp (full-mask shuffles); the eight quads of a warp
// work on eight independent points.  Exceptional cases (identity operands, P + P, P - P) are resolved with
// quad-wide flags and selects, so results are exact for any input.
#pragma once
#include "ec.cuh"

namespace vimz {

constexpr unsigned FULL = 0xffffffffu;

template <class F>
VIMZ_DI Fp<F> q_fetch(const Fp<F>& v, int src_in_quad) {
  int src = (threadIdx.x & 28) | src_in_quad;  // same quad, lane `src_in_quad`
  Fp<F> r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = __shfl_sync(FULL, v.v[k], src);
  return r;
}
VIMZ_DI bool q_flag(bool f, int src_in_quad) {
  return __shfl_sync(FULL, (int)f, (threadIdx.x & 28) | src_in_quad) != 0;
}
template <class F>
VIMZ_DI Fp<F> fp_select(bool c, const Fp<F>& a, const Fp<F>& b) {
  Fp<F> r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.v[k] = c ? a.v[k] : b.v[k];
  return r;
}

// One coordinate of a quad-distributed XYZZ point: lane role = threadIdx.x & 3 (0 X, 1 Y, 2 ZZ, 3 ZZZ).
template <class C>
struct QPoint {
  using F = Fp<typename C::Fb>;
  F c;
  VIMZ_DI static QPoint identity() {
    QPoint p;
    p.c = F::zero();
    return p;
  }
  // 128-byte XYZZ record: each lane reads its own 32-byte coordinate (one coalesced 128-byte access per quad)
  VIMZ_DI static QPoint load(const void* rec) {
    QPoint p;
    p.c = F::load(reinterpret_cast<const char*>(rec) + 32 * (threadIdx.x & 3));
    return p;
  }
  VIMZ_DI static QPoint load_cg(const void* rec) {
    QPoint p;
    p.c = F::load_cg(reinterpret_cast<const char*>(rec) + 32 * (threadIdx.x & 3));
    return p;
  }
  VIMZ_DI void store(void* rec) const { c.store(reinterpret_cast<char*>(rec) + 32 * (threadIdx.x & 3)); }
  VIMZ_DI bool is_identity() const { return q_flag(c.is_zero(), 2); }  // ZZ == 0
  // the same point held by the quad `delta` quads further up the warp (garbage beyond the warp: caller masks)
  VIMZ_DI QPoint from_quad_above(int delta) const {
    QPoint p;
#pragma unroll
    for (int k = 0; k < 8; k++) p.c.v[k] = __shfl_down_sync(FULL, c.v[k], 4 * delta);
    return p;
  }
};

// The kernels that use these operations run a few warps through long dependent chains: their time is set by
// latency, and with every operation inlined (~15 KB of SASS per addition, 100-180 KB per kernel) the top stall was
// instruction fetch (`no_instruction`).  q_add / q_dbl are therefore OUT OF LINE: one copy per kernel that stays
// in the instruction caches (L0 ~6 KB, L1.5 32 KB) across the iterations of a chain.
template <class C>
__device__ __noinline__ QPoint<C> q_dbl(QPoint<C> p);

// 2 * P   (dbl-2008-s-1, a = 0): 3 product rounds
template <class C>
VIMZ_DI QPoint<C> q_dbl_inl(const QPoint<C>& p) {
  using F = Fp<typename C::Fb>;
  const int role = threadIdx.x & 3;
  const bool ident = p.is_identity();  // prime-order curves: no finite point with y = 0
  // round 1: L0 XX = X^2 ; L1 V = U^2 (U = 2Y)
  F u = fp_dbl(p.c);                   // meaningful on L1
  F a1 = role == 1 ? u : p.c;
  F r1 = fp_mul_call<typename C::Fb>(a1, a1);     // L0: XX, L1: V (L2/L3: unused squares)
  // round 2: L0 S = X*V ; L1 W = U*V ; L2 MM = M^2 (M = 3XX) ; L3 ZZ3 = V*ZZ
  F v = q_fetch(r1, 1);
  F xx = q_fetch(r1, 0);
  F m = fp_add(fp_dbl(xx), xx);
  F zz = q_fetch(p.c, 2);
  F a2 = role == 0 ? p.c : (role == 1 ? u : (role == 2 ? m : zz));
  F b2 = role == 2 ? m : v;
  F r2 = fp_mul_call<typename C::Fb>(a2, b2);     // L0: S, L1: W, L2: MM, L3: ZZ3
  // round 3: L0 T = M*(S - X3), X3 = MM - 2S ; L1 WY = W*Y ; L3 ZZZ3 = W*ZZZ
  F mm = q_fetch(r2, 2);
  F w = q_fetch(r2, 1);
  F x3 = fp_sub(mm, fp_dbl(r2));       // meaningful on L0 (r2 = S)
  F a3 = role == 0 ? m : w;
  F b3 = role == 0 ? fp_sub(r2, x3) : p.c;        // L1: Y, L3: ZZZ
  F r3 = fp_mul_call<typename C::Fb>(a3, b3);     // L0: T, L1: WY, L3: ZZZ3
  F t = q_fetch(r3, 0);
  F zz3 = q_fetch(r2, 3);
  QPoint<C> out;
  out.c = role == 0 ? x3 : (role == 1 ? fp_sub(t, r3) : (role == 2 ? zz3 : r3));
  if (ident) out.c = F::zero();
  return out;
}

template <class C>
__device__ __noinline__ QPoint<C> q_dbl(QPoint<C> p) { return q_dbl_inl<C>(p); }

// P1 + P2   (add-2008-s): 4 product rounds; all exceptional cases exact
template <class C>
VIMZ_DI QPoint<C> q_add_inl(const QPoint<C>& p1, const QPoint<C>& p2) {
  using F = Fp<typename C::Fb>;
  const int role = threadIdx.x & 3;
  const bool id1 = p1.is_identity(), id2 = p2.is_identity();
  // round 1: L0 U1 = X1*ZZ2 ; L1 S1 = Y1*ZZZ2 ; L2 U2 = X2*ZZ1 ; L3 S2 = Y2*ZZZ1
  F o2 = q_fetch(p2.c, role ^ 2);      // L0<-ZZ2, L1<-ZZZ2, L2<-X2, L3<-Y2
  F r1 = fp_mul_call<typename C::Fb>(p1.c, o2);
  F hi = q_fetch(r1, role | 2);        // L0,L2 <- U2 ; L1,L3 <- S2
  F lo = q_fetch(r1, role & 1);        // L0,L2 <- U1 ; L1,L3 <- S1
  F d = fp_sub(hi, lo);                // even lanes: P = U2 - U1 ; odd lanes: R = S2 - S1
  const bool pz = q_flag(d.is_zero(), 0), rz = q_flag(d.is_zero(), 1);
  // round 2: L0 PP = P^2 ; L1 RR = R^2 ; L2 ZZq = ZZ1*ZZ2 ; L3 ZZZq = ZZZ1*ZZZ2
  F a2 = role < 2 ? d : p1.c;
  F b2 = role < 2 ? d : p2.c;
  F r2 = fp_mul_call<typename C::Fb>(a2, b2);
  // round 3: L0 PPP = P*PP ; L1 Q = U1*PP ; L2 ZZ3 = ZZq*PP ; (L3 idle)
  F pp = q_fetch(r2, 0);
  F u1 = q_fetch(r1, 0);
  F a3 = role == 0 ? d : (role == 1 ? u1 : r2);
  F r3 = fp_mul_call<typename C::Fb>(a3, pp);
  // round 4: L0 T2 = S1*PPP ; L1 T1 = R*(Q - X3), X3 = RR - PPP - 2Q ; L3 ZZZ3 = ZZZq*PPP
  F ppp = q_fetch(r3, 0);
  F s1 = q_fetch(r1, 1);
  F x3 = fp_sub(fp_sub(r2, ppp), fp_dbl(r3));  // meaningful on L1 (r2 = RR, r3 = Q)
  F a4 = role == 0 ? s1 : (role == 1 ? d : r2);
  F b4 = role == 1 ? fp_sub(r3, x3) : ppp;
  F r4 = fp_mul_call<typename C::Fb>(a4, b4);
  F t2 = q_fetch(r4, 0);
  F x3b = q_fetch(x3, 1);
  QPoint<C> out;
  out.c = role == 0 ? x3b : (role == 1 ? fp_sub(r4, t2) : (role == 2 ? r3 : r4));
  // exceptional cases (quad-uniform flags; the doubling is executed only if some quad of the warp needs it)
  const bool need_dbl = !id1 && !id2 && pz && rz;
  if (__any_sync(FULL, need_dbl)) {
    QPoint<C> dd = q_dbl<C>(p1);
    if (need_dbl) out = dd;
  }
  if (!id1 && !id2 && pz && !rz) out.c = F::zero();
  if (id2) out = p1;
  else if (id1) out = p2;
  return out;
}
template <class C>
__device__ __noinline__ QPoint<C> q_add(QPoint<C> p1, QPoint<C> p2) { return q_add_inl<C>(p1, p2); }

// TWO independent additions (p1 + p2, s1 + s2) with their product rounds interleaved (fp_mul2_call: two Montgomery chains in flight
// in one warp -- a lone warp pays ~1.45x the time of one addition for both).  For the running-sum chains of the bucket reduction,
// where "running += bucket" and "weighted += running" of consecutive steps are independent of each other.
template <class C>
struct QPair {
  QPoint<C> a, b;
};
template <class C>
__device__ __noinline__ QPair<C> q_add2(QPoint<C> p1, QPoint<C> p2, QPoint<C> s1, QPoint<C> s2) {
  using F = Fp<typename C::Fb>;
  using Fb = typename C::Fb;
  const int role = threadIdx.x & 3;
  const bool id1 = p1.is_identity(), id2 = p2.is_identity(), jd1 = s1.is_identity(), jd2 = s2.is_identity();
  // round 1 (see q_add_inl for the lane roles)
  FpPair<Fb> r1 = fp_mul2_call<Fb>(p1.c, q_fetch(p2.c, role ^ 2), s1.c, q_fetch(s2.c, role ^ 2));
  F dA = fp_sub(q_fetch(r1.a, role | 2), q_fetch(r1.a, role & 1));
  F dB = fp_sub(q_fetch(r1.b, role | 2), q_fetch(r1.b, role & 1));
  const bool pzA = q_flag(dA.is_zero(), 0), rzA = q_flag(dA.is_zero(), 1);
  const bool pzB = q_flag(dB.is_zero(), 0), rzB = q_flag(dB.is_zero(), 1);
  // round 2
  FpPair<Fb> r2 = fp_mul2_call<Fb>(role < 2 ? dA : p1.c, role < 2 ? dA : p2.c, role < 2 ? dB : s1.c, role < 2 ? dB : s2.c);
  // round 3
  F ppA = q_fetch(r2.a, 0), u1A = q_fetch(r1.a, 0);
  F ppB = q_fetch(r2.b, 0), u1B = q_fetch(r1.b, 0);
  FpPair<Fb> r3 = fp_mul2_call<Fb>(role == 0 ? dA : (role == 1 ? u1A : r2.a), ppA, role == 0 ? dB : (role == 1 ? u1B : r2.b), ppB);
  // round 4
  F pppA = q_fetch(r3.a, 0), s1A = q_fetch(r1.a, 1);
  F pppB = q_fetch(r3.b, 0), s1B = q_fetch(r1.b, 1);
  F x3A = fp_sub(fp_sub(r2.a, pppA), fp_dbl(r3.a));
  F x3B = fp_sub(fp_sub(r2.b, pppB), fp_dbl(r3.b));
  FpPair<Fb> r4 = fp_mul2_call<Fb>(role == 0 ? s1A : (role == 1 ? dA : r2.a), role == 1 ? fp_sub(r3.a, x3A) : pppA,
                                   role == 0 ? s1B : (role == 1 ? dB : r2.b), role == 1 ? fp_sub(r3.b, x3B) : pppB);
  QPair<C> out;
  {
    F t2 = q_fetch(r4.a, 0), x3b = q_fetch(x3A, 1);
    out.a.c = role == 0 ? x3b : (role == 1 ? fp_sub(r4.a, t2) : (role == 2 ? r3.a : r4.a));
  }
  {
    F t2 = q_fetch(r4.b, 0), x3b = q_fetch(x3B, 1);
    out.b.c = role == 0 ? x3b : (role == 1 ? fp_sub(r4.b, t2) : (role == 2 ? r3.b : r4.b));
  }
  // exceptional cases, as in q_add_inl
  const bool dblA = !id1 && !id2 && pzA && rzA, dblB = !jd1 && !jd2 && pzB && rzB;
  if (__any_sync(FULL, dblA)) {
    QPoint<C> dd = q_dbl<C>(p1);
    if (dblA) out.a = dd;
  }
  if (__any_sync(FULL, dblB)) {
    QPoint<C> dd = q_dbl<C>(s1);
    if (dblB) out.b = dd;
  }
  if (!id1 && !id2 && pzA && !rzA) out.a.c = F::zero();
  if (!jd1 && !jd2 && pzB && !rzB) out.b.c = F::zero();
  if (id2) out.a = p1;
  else if (id1) out.a = p2;
  if (jd2) out.b = s1;
  else if (jd1) out.b = s2;
  return out;
}

// acc (quad) -> Jacobian {X*ZZ^4, Y*ZZZ^4, ZZ*ZZZ}; identity -> (0, R, 0).  Writes 96 bytes from the quad.
template <class C>
VIMZ_DI void q_store_jacobian(const QPoint<C>& p, void* out, bool do_store = true, void* out2 = nullptr) {
  using F = Fp<typename C::Fb>;
  const int role = threadIdx.x & 3;
  const bool ident = p.is_identity();
  F zz = q_fetch(p.c, 2), zzz = q_fetch(p.c, 3);
  F base = (role & 1) ? zzz : zz;                         // L0: ZZ, L1: ZZZ, L2: ZZ, L3: ZZZ
  F sq = fp_mul_call<typename C::Fb>(base, base);         // ^2
  F q4 = fp_mul_call<typename C::Fb>(sq, sq);             // ^4
  F other = role == 2 ? zzz : p.c;                        // L0: X, L1: Y, L2: ZZZ
  F res = fp_mul_call<typename C::Fb>(role == 2 ? zz : q4, other);  // L0: X*ZZ^4, L1: Y*ZZZ^4, L2: ZZ*ZZZ
  if (ident) res = role == 1 ? F::one() : F::zero();
  if (role < 3 && do_store) {
    res.store(reinterpret_cast<char*>(out) + 32 * role);
    if (out2) res.store(reinterpret_cast<char*>(out2) + 32 * role);  // second copy: page-locked host memory the caller polls for
  }
}

// Jacobian {X, Y, Z} (96 bytes) -> quad XYZZ (ZZ = Z^2, ZZZ = Z^3); Z = 0 -> identity
template <class C>
VIMZ_DI QPoint<C> q_load_jacobian(const void* in) {
  using F = Fp<typename C::Fb>;
  const int role = threadIdx.x & 3;
  F z = F::load(reinterpret_cast<const char*>(in) + 64);
  F mine = F::load(reinterpret_cast<const char*>(in) + 32 * (role & 1));  // L0,L2: X ; L1,L3: Y
  F zz = fp_mul_call<typename C::Fb>(z, z);
  F zzz = fp_mul_call<typename C::Fb>(zz, z);
  QPoint<C> p;
  p.c = role == 0 ? mine : (role == 1 ? mine : (role == 2 ? zz : zzz));
  if (z.is_zero()) p.c = F::zero();
  return p;
}

// Tree-sum of the eight points held by the eight quads of a warp; result valid in quad 0 (lanes 0..3).
// first_delta = 2: only quads 0..3 hold points (two levels instead of three).
template <class C>
VIMZ_DI QPoint<C> q_warp_reduce(QPoint<C> acc, int first_delta = 4) {
  using F = Fp<typename C::Fb>;
#pragma unroll 1
  for (int delta = first_delta; delta > 0; delta >>= 1) {
    QPoint<C> other = acc.from_quad_above(delta);
    if ((int)(threadIdx.x & 31) + 4 * delta >= 32) other.c = F::zero();  // no partner: add the identity
    acc = q_add<C>(acc, other);
  }
  return acc;
}

// broadcast quad 0's point to every quad of the warp
template <class C>
VIMZ_DI QPoint<C> q_fetch_quad0(const QPoint<C>& p) {
  QPoint<C> r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.c.v[k] = __shfl_sync(FULL, p.c.v[k], threadIdx.x & 3);
  return r;
}

// the full XYZZ point held by lane j of each quad (thread-level layout), redistributed over the quad
template <class C>
VIMZ_DI QPoint<C> q_from_lane(const Xyzz<C>& mine, int j) {
  const int role = threadIdx.x & 3;
  const int src = (threadIdx.x & 28) | j;
  QPoint<C> r;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t x = __shfl_sync(FULL, mine.x.v[k], src);
    uint32_t y = __shfl_sync(FULL, mine.y.v[k], src);
    uint32_t zz = __shfl_sync(FULL, mine.zz.v[k], src);
    uint32_t zzz = __shfl_sync(FULL, mine.zzz.v[k], src);
    r.c.v[k] = role == 0 ? x : (role == 1 ? y : (role == 2 ? zz : zzz));
  }
  return r;
}

}  // namespace vimz
