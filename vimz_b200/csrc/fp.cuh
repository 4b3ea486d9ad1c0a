// fp.cuh -- 254/255-bit prime-field arithmetic for sm_100a, one element per thread.
//
// Representation: 8 x 32-bit little-endian limbs, Montgomery form with R = 2^256, always fully
// reduced to [0, p).  This is bit-for-bit the in-memory layout of halo2curves 0.1.0 / pasta_curves
// 0.5.1 field types ([u64;4] little-endian Montgomery limbs), so host buffers
// cross the C ABI without conversion.
//
// Multiplication: operand-scanning Montgomery with two interleaved 64-bit-lane accumulators
// ("even" lanes at limb positions (0,1),(2,3).. and "odd" lanes at (1,2),(3,4)..).  Every
// mad.lo.cc/madc.hi.cc pair below is fused by ptxas into ONE IMAD.WIDE.U32.X with a predicate
// carry (checked with cuobjdump -sass), so a product costs 64 + (nonzero limbs of p)*8 wide
// IMADs instead of 272 32-bit ones.  Zero limbs of the modulus (Pasta: p[4..6] = 0) degrade to
// plain add-with-carry on the ALU pipe.
//
// Replaces: the 4x64-bit Montgomery arithmetic of halo2curves / pasta_curves that nova-snark 0.23.0
// runs under RecursiveSNARK::prove_step (call site /root/reference/vimz/src/nova_snark_backend/folding.rs:35).
#pragma once
#include <cstdint>
#include "field_constants.cuh"

namespace vimz {

#define VIMZ_DI __device__ __forceinline__

template <class F>
struct Fp {
  uint32_t v[8];

  VIMZ_DI static Fp zero() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
    return r;
  }
  VIMZ_DI static Fp one() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = F::one(i);
    return r;
  }
  VIMZ_DI static Fp r2() {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = F::r2(i);
    return r;
  }
  VIMZ_DI bool is_zero() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i];
    return o == 0;
  }
  VIMZ_DI bool operator==(const Fp& b) const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i];
    return o == 0;
  }
  VIMZ_DI bool operator!=(const Fp& b) const { return !(*this == b); }

  // 2 x 128-bit vector access (32-byte aligned AoS elements).
  VIMZ_DI static Fp load(const void* ptr) {
    const uint4* q = reinterpret_cast<const uint4*>(ptr);
    uint4 a = q[0], b = q[1];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  VIMZ_DI static Fp load_nc(const void* ptr) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(ptr);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  VIMZ_DI static Fp load_cg(const void* ptr) {  // through L2: data published by another block of the same launch
    const uint4* q = reinterpret_cast<const uint4*>(ptr);
    uint4 a = __ldcg(q), b = __ldcg(q + 1);
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  VIMZ_DI void store(void* ptr) const {
    uint4* q = reinterpret_cast<uint4*>(ptr);
    q[0] = make_uint4(v[0], v[1], v[2], v[3]);
    q[1] = make_uint4(v[4], v[5], v[6], v[7]);
  }
};

// r = a - p if a >= p (a < 2p assumed; `carry` = bit 256 of a, only possible for 256-bit sums).
template <class F>
VIMZ_DI void fp_final_sub(Fp<F>& a, uint32_t carry = 0) {
  uint32_t t[8], borrow = carry;
#pragma unroll
  for (int i = 0; i < 8; i++) t[i] = a.v[i];
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t"
      "subc.cc.u32 %4, %4, %13;\n\t"
      "subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t"
      "subc.cc.u32 %7, %7, %16;\n\t"
      "subc.u32 %8, %8, 0;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(borrow)
      : "r"(F::p(0)), "r"(F::p(1)), "r"(F::p(2)), "r"(F::p(3)), "r"(F::p(4)), "r"(F::p(5)), "r"(F::p(6)), "r"(F::p(7)));
  // borrow == 0  <=>  (carry:a) >= p  -> keep t
  bool ge = (borrow == 0);
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = ge ? t[i] : a.v[i];
}

template <class F>
VIMZ_DI Fp<F> fp_add(const Fp<F>& a, const Fp<F>& b) {
  Fp<F> r = a;
  uint32_t carry = 0;
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "addc.cc.u32 %1, %1, %10;\n\t"
      "addc.cc.u32 %2, %2, %11;\n\t"
      "addc.cc.u32 %3, %3, %12;\n\t"
      "addc.cc.u32 %4, %4, %13;\n\t"
      "addc.cc.u32 %5, %5, %14;\n\t"
      "addc.cc.u32 %6, %6, %15;\n\t"
      "addc.cc.u32 %7, %7, %16;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "+r"(carry)
      : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  if (F::BITS <= 255) carry = 0;  // a + b < 2p < 2^256
  fp_final_sub(r, carry);
  return r;
}

template <class F>
VIMZ_DI Fp<F> fp_sub(const Fp<F>& a, const Fp<F>& b) {
  Fp<F> r = a;
  uint32_t borrow = 0;
  asm("sub.cc.u32 %0, %0, %9;\n\t"
      "subc.cc.u32 %1, %1, %10;\n\t"
      "subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t"
      "subc.cc.u32 %4, %4, %13;\n\t"
      "subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t"
      "subc.cc.u32 %7, %7, %16;\n\t"
      "subc.u32 %8, %8, 0;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7]), "+r"(borrow)
      : "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
  // borrow = 0xffffffff if a < b: add p back (masked).
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(F::p(0) & borrow), "r"(F::p(1) & borrow), "r"(F::p(2) & borrow), "r"(F::p(3) & borrow),
        "r"(F::p(4) & borrow), "r"(F::p(5) & borrow), "r"(F::p(6) & borrow), "r"(F::p(7) & borrow));
  return r;
}

template <class F>
VIMZ_DI Fp<F> fp_neg(const Fp<F>& a) {
  Fp<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = F::p(i);
  asm("sub.cc.u32 %0, %0, %8;\n\t"
      "subc.cc.u32 %1, %1, %9;\n\t"
      "subc.cc.u32 %2, %2, %10;\n\t"
      "subc.cc.u32 %3, %3, %11;\n\t"
      "subc.cc.u32 %4, %4, %12;\n\t"
      "subc.cc.u32 %5, %5, %13;\n\t"
      "subc.cc.u32 %6, %6, %14;\n\t"
      "subc.u32 %7, %7, %15;"
      : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
      : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  bool z = a.is_zero();
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : r.v[i];
  return r;
}

// r = neg ? -a : a
template <class F>
VIMZ_DI Fp<F> fp_cneg(const Fp<F>& a, bool neg) {
  Fp<F> n = fp_neg(a);
  Fp<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = neg ? n.v[i] : a.v[i];
  return r;
}

template <class F>
VIMZ_DI Fp<F> fp_dbl(const Fp<F>& a) { return fp_add(a, a); }

// a > (p-1)/2 for a canonical (non-Montgomery) value a: the "negative half" of the field
template <class F>
VIMZ_DI bool fp_gt_half(const Fp<F>& a) {
  uint32_t borrow = 0;
  // (p-1)/2 - a  borrows  <=>  a > (p-1)/2
  uint32_t t;
  asm("sub.cc.u32 %0, %2, %10;\n\t"
      "subc.cc.u32 %0, %3, %11;\n\t"
      "subc.cc.u32 %0, %4, %12;\n\t"
      "subc.cc.u32 %0, %5, %13;\n\t"
      "subc.cc.u32 %0, %6, %14;\n\t"
      "subc.cc.u32 %0, %7, %15;\n\t"
      "subc.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %0, %9, %17;\n\t"
      "subc.u32 %1, 0, 0;"
      : "=&r"(t), "=&r"(borrow)
      : "r"((F::p(0) >> 1) | (F::p(1) << 31)), "r"((F::p(1) >> 1) | (F::p(2) << 31)), "r"((F::p(2) >> 1) | (F::p(3) << 31)),
        "r"((F::p(3) >> 1) | (F::p(4) << 31)), "r"((F::p(4) >> 1) | (F::p(5) << 31)), "r"((F::p(5) >> 1) | (F::p(6) << 31)),
        "r"((F::p(6) >> 1) | (F::p(7) << 31)), "r"(F::p(7) >> 1),
        "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]));
  return borrow != 0;
}

// A/B knob (default OFF, measured slower): reduction constants and shift counts of the Pasta fields read from CONSTANT MEMORY.
// With the limbs of p as immediates ptxas splits every m * p[k] into a half-rate IMAD.HI plus an IMAD and turns the two shifts
// of m * 2^254 back into a wide multiply by 2^30; opaque values remove 12 of the 227 multiplier-pipe issue units of a product
// (IMAD.WIDE / IMAD.HI run at half rate, profiles/int_peak.json) but add 18 instructions: k_msm_accumulate 2.604 -> 2.654 ms at
// 2^20 -- the kernel is bound by issue slots and dependent-issue stalls, not by the multiplier pipe alone.  {p[1], p[2], p[3], 30, 2}
#ifndef VIMZ_OPAQUE_P
#define VIMZ_OPAQUE_P 0
#endif
template <class F>
__constant__ uint32_t red_const[5] = {F::p(1), F::p(2), F::p(3), 30u, 2u};

// ---- Montgomery multiplication ----------------------------------------------------------------
// One row: acc += a * bi ; m = acc[0] * INV ; acc += m * p ; acc >>= 32, with the accumulator split
// into an even-aligned (ev) and an odd-aligned (od) set of 64-bit lanes.  The 32-bit right shift
// swaps their roles, so the caller alternates (ev, od) between rows; on entry (non-first rows) `od`
// still holds the previous row's even lanes, which are consumed shifted down by one lane.
template <class F, bool FIRST>
VIMZ_DI void mont_row(uint32_t (&ev)[8], uint32_t (&od)[8], const uint32_t (&a)[8], uint32_t bi) {
  if (FIRST) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint64_t po = (uint64_t)a[2 * k + 1] * bi, pe = (uint64_t)a[2 * k] * bi;  // IMAD.WIDE.U32
      od[2 * k] = (uint32_t)po; od[2 * k + 1] = (uint32_t)(po >> 32);
      ev[2 * k] = (uint32_t)pe; ev[2 * k + 1] = (uint32_t)(pe >> 32);
    }
  } else {
    // position 0 of the new frame: old even lane 0 high half lands on ev[0]; its carry enters the
    // odd chain at position 1.  od[j] <- a[j+1..]*bi + od[j+2] implements the lane shift for free.
    asm("add.cc.u32 %0, %0, %2;\n\t"
        "madc.lo.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.hi.cc.u32 %2, %9, %13, %4;\n\t"
        "madc.lo.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.hi.cc.u32 %4, %10, %13, %6;\n\t"
        "madc.lo.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.hi.cc.u32 %6, %11, %13, %8;\n\t"
        "madc.lo.cc.u32 %7, %12, %13, 0;\n\t"
        "madc.hi.u32 %8, %12, %13, 0;"
        : "+r"(ev[0]), "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(a[1]), "r"(a[3]), "r"(a[5]), "r"(a[7]), "r"(bi));
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(od[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(bi));
  }
  uint32_t m = ev[0] * F::INV;
  constexpr bool SPARSE = (F::p(4) == 0 && F::p(5) == 0 && F::p(6) == 0);  // Pasta: p = 2^254 + t, t < 2^128
  constexpr bool PASTA = SPARSE && F::p(0) == 1 && F::p(7) == 0x40000000u && F::INV == 0xffffffffu;
  if (PASTA) {
    // p = 2^254 + t with t odd < 2^128 and -p^-1 = -1 mod 2^32:  m = -acc0, m*p[0] = m just cancels limb 0
    // (carry = acc0 != 0) and m*p[7] = m << 30 is two shifts -- only p[1], p[2], p[3] need the multiplier.
    // (Left to ptxas, the constants 1 and 2^30 break the lo/hi pairs into half-rate IMAD.HI plus carries.)
    m = 0u - ev[0];
#if VIMZ_OPAQUE_P
    const uint32_t P1 = red_const<F>[0], P2 = red_const<F>[1], P3 = red_const<F>[2];
    uint32_t mlo = m << red_const<F>[3], mhi = m >> red_const<F>[4];
#else
    const uint32_t P1 = F::p(1), P2 = F::p(2), P3 = F::p(3);
    uint32_t mlo = m << 30, mhi = m >> 2;
#endif
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, %11;\n\t"
        "addc.u32 %7, %7, %12;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(m), "r"(P1), "r"(P3), "r"(mlo), "r"(mhi));
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "madc.lo.cc.u32 %2, %9, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %10, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(od[7])
        : "r"(m), "r"(P2));
  } else if (SPARSE) {
    // odd lanes += m * p[1,3,-,7]; even lanes += m * p[0,2,-,-]; zero limbs only ripple the carry.
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "madc.lo.cc.u32 %6, %8, %11, %6;\n\t"
        "madc.hi.u32 %7, %8, %11, %7;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(m), "r"(F::p(1)), "r"(F::p(3)), "r"(F::p(7)));
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(od[7])
        : "r"(m), "r"(F::p(0)), "r"(F::p(2)));
  } else {
    // odd lanes += m * p[1,3,5,7]   (no carry out: total < 2^288)
    asm("mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, %12, %6;\n\t"
        "madc.hi.u32 %7, %8, %12, %7;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7])
        : "r"(m), "r"(F::p(1)), "r"(F::p(3)), "r"(F::p(5)), "r"(F::p(7)));
    // even lanes += m * p[0,2,4,6] ; carry out of position 7 goes to position 8 = od[7]
    asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %9, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]), "+r"(ev[6]), "+r"(ev[7]), "+r"(od[7])
        : "r"(m), "r"(F::p(0)), "r"(F::p(2)), "r"(F::p(4)), "r"(F::p(6)));
  }
}

template <class F>
VIMZ_DI Fp<F> fp_mul(const Fp<F>& a, const Fp<F>& b) {
  uint32_t e[8], o[8];
  mont_row<F, true>(e, o, a.v, b.v[0]);
  mont_row<F, false>(o, e, a.v, b.v[1]);
  mont_row<F, false>(e, o, a.v, b.v[2]);
  mont_row<F, false>(o, e, a.v, b.v[3]);
  mont_row<F, false>(e, o, a.v, b.v[4]);
  mont_row<F, false>(o, e, a.v, b.v[5]);
  mont_row<F, false>(e, o, a.v, b.v[6]);
  mont_row<F, false>(o, e, a.v, b.v[7]);
  // final frame: even lanes = e, pending (already shifted) odd lanes = o[1..7]
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(e[0]), "+r"(e[1]), "+r"(e[2]), "+r"(e[3]), "+r"(e[4]), "+r"(e[5]), "+r"(e[6]), "+r"(e[7])
      : "r"(o[1]), "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]));
  Fp<F> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = e[i];
  fp_final_sub(r);
  return r;
}

// Two independent products with their Montgomery rows interleaved in program order, so a lone warp has
// two carry chains in flight instead of one (used by the latency-bound kernels through fp_mul2_call).
template <class F>
VIMZ_DI void fp_mul2(Fp<F>& r1, const Fp<F>& a1, const Fp<F>& b1, Fp<F>& r2, const Fp<F>& a2, const Fp<F>& b2) {
  uint32_t e1[8], o1[8], e2[8], o2[8];
  mont_row<F, true>(e1, o1, a1.v, b1.v[0]);
  mont_row<F, true>(e2, o2, a2.v, b2.v[0]);
  mont_row<F, false>(o1, e1, a1.v, b1.v[1]);
  mont_row<F, false>(o2, e2, a2.v, b2.v[1]);
  mont_row<F, false>(e1, o1, a1.v, b1.v[2]);
  mont_row<F, false>(e2, o2, a2.v, b2.v[2]);
  mont_row<F, false>(o1, e1, a1.v, b1.v[3]);
  mont_row<F, false>(o2, e2, a2.v, b2.v[3]);
  mont_row<F, false>(e1, o1, a1.v, b1.v[4]);
  mont_row<F, false>(e2, o2, a2.v, b2.v[4]);
  mont_row<F, false>(o1, e1, a1.v, b1.v[5]);
  mont_row<F, false>(o2, e2, a2.v, b2.v[5]);
  mont_row<F, false>(e1, o1, a1.v, b1.v[6]);
  mont_row<F, false>(e2, o2, a2.v, b2.v[6]);
  mont_row<F, false>(o1, e1, a1.v, b1.v[7]);
  mont_row<F, false>(o2, e2, a2.v, b2.v[7]);
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(e1[0]), "+r"(e1[1]), "+r"(e1[2]), "+r"(e1[3]), "+r"(e1[4]), "+r"(e1[5]), "+r"(e1[6]), "+r"(e1[7])
      : "r"(o1[1]), "r"(o1[2]), "r"(o1[3]), "r"(o1[4]), "r"(o1[5]), "r"(o1[6]), "r"(o1[7]));
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, 0;"
      : "+r"(e2[0]), "+r"(e2[1]), "+r"(e2[2]), "+r"(e2[3]), "+r"(e2[4]), "+r"(e2[5]), "+r"(e2[6]), "+r"(e2[7])
      : "r"(o2[1]), "r"(o2[2]), "r"(o2[3]), "r"(o2[4]), "r"(o2[5]), "r"(o2[6]), "r"(o2[7]));
#pragma unroll
  for (int i = 0; i < 8; i++) { r1.v[i] = e1[i]; r2.v[i] = e2[i]; }
  fp_final_sub(r1);
  fp_final_sub(r2);
}

}  // namespace vimz
#include "fp_sqr.cuh"
namespace vimz {
// Squaring.  Pasta base / scalar fields with -DVIMZ_FP_SQR=1: the dedicated routine of fp_sqr.cuh (generated by tools/gen_fp_sqr.py:
// off-diagonal triangle once + doubling + diagonal, then reduction-only rows: 60 wide products instead of 88), bit-identical to
// fp_mul(a, a) (tests/test_gpu_field.py); otherwise the plain product.
#ifndef VIMZ_FP_SQR
#define VIMZ_FP_SQR 1
#endif
template <class F>
VIMZ_DI Fp<F> fp_sqr(const Fp<F>& a) {
#if VIMZ_FP_SQR
  constexpr bool PASTA = F::p(4) == 0 && F::p(5) == 0 && F::p(6) == 0 && F::p(0) == 1 && F::p(7) == 0x40000000u && F::INV == 0xffffffffu;
  if constexpr (PASTA) {
    Fp<F> r;
    fp_sqr_pasta_limbs<F>(r.v, a.v);
    fp_final_sub(r);  // the routine returns a value below 2p
    return r;
  }
#endif
  return fp_mul(a, a);
}

template <class F>
VIMZ_DI Fp<F> fp_from_mont(const Fp<F>& a) {
  Fp<F> one_raw = Fp<F>::zero();
  one_raw.v[0] = 1;
  return fp_mul(a, one_raw);
}
template <class F>
VIMZ_DI Fp<F> fp_to_mont(const Fp<F>& a) { return fp_mul(a, Fp<F>::r2()); }

template <class F>
__device__ __noinline__ Fp<F> fp_mul_noinline(Fp<F> a, Fp<F> b) { return fp_mul(a, b); }

// a^(p-2): only used by affine normalisation (one inversion per thread), never on the hot loop.
template <class F>
__device__ __noinline__ Fp<F> fp_inv(const Fp<F>& a) {
  Fp<F> r = Fp<F>::one();
  for (int i = 7; i >= 0; i--) {
    uint32_t w = F::pm2(0);
    switch (i) {
      case 7: w = F::pm2(7); break; case 6: w = F::pm2(6); break; case 5: w = F::pm2(5); break;
      case 4: w = F::pm2(4); break; case 3: w = F::pm2(3); break; case 2: w = F::pm2(2); break;
      case 1: w = F::pm2(1); break; default: break;
    }
    for (int bit = 31; bit >= 0; bit--) {
      r = fp_mul_noinline<F>(r, r);
      if ((w >> bit) & 1) r = fp_mul_noinline<F>(r, a);
    }
  }
  return r;
}

}  // namespace vimz
