// fp_lat.cuh -- LATENCY-oriented Montgomery multiplication for the kernels whose warps run alone.
//
// fp_mul (fp.cuh) is written for throughput: every row of the operand-scanning product is two PTX carry chains
// (mad.lo.cc / madc.hi.cc), which ptxas turns into the fewest multiplier-pipe instructions (IMAD.WIDE.U32.X).  PTX has
// ONE carry flag, though, so those chains are issued strictly one after the other and a lone warp sees a product as
// a ~200-instruction dependent sequence (~780 cycles, tools/mul_latency.cu).  The bucket-reduction trees, the combine of
// cut buckets, the trees of k_msm_direct and the commitment folds are chains of EC additions run by a few warps per SM:
// their time IS that latency.
//
// Here the same CIOS product is written with 64-bit integer arithmetic and no inline carry chains: every partial product
// a_j * b_i + t_j is an independent IMAD.WIDE (the addend rides along, the sum cannot overflow 64 bits), and only the
// carry ripple -- one 64-bit add per limb, on the ALU pipe -- is sequential.  ptxas allocates its own predicates for
// those adds, so the ripple of the a*b_i half, the ripple of the m*p half and the partial products of the NEXT row
// overlap.  More instructions (it would lose in the throughput kernel), far fewer dependent cycles.
// Bit-identical results: both compute a*b*R^-1 mod p fully reduced.
#pragma once
#include "fp.cuh"

namespace vimz {

template <class F>
VIMZ_DI Fp<F> fp_mul_lat(const Fp<F>& a, const Fp<F>& b) {
  uint32_t t[8];
  uint32_t t8 = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // t += a * b_i : partial products first (independent), then the carry ripple
    uint64_t P[8];
#pragma unroll
    for (int j = 0; j < 8; j++) P[j] = (uint64_t)a.v[j] * b.v[i] + t[j];
    uint32_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t s = P[j] + c;  // <= (2^32-1)^2 + 2(2^32-1) = 2^64 - 1
      t[j] = (uint32_t)s;
      c = (uint32_t)(s >> 32);
    }
    uint64_t top = (uint64_t)t8 + c;  // t < 2p + a*b_i < 2^(256+33): one extra limb and a bit
    // t += m * p ; t >>= 32
    const uint32_t m = t[0] * F::INV;
    uint64_t Q[8];
#pragma unroll
    for (int j = 0; j < 8; j++) Q[j] = (uint64_t)m * F::p(j) + t[j];
    c = (uint32_t)(Q[0] >> 32);  // low word cancels by construction
#pragma unroll
    for (int j = 1; j < 8; j++) {
      uint64_t s = Q[j] + c;
      t[j - 1] = (uint32_t)s;
      c = (uint32_t)(s >> 32);
    }
    top += c;
    t[7] = (uint32_t)top;
    t8 = (uint32_t)(top >> 32);
  }
  // a, b < p < 2^255  =>  the CIOS result is < 2p < 2^256: t8 = 0
  Fp<F> r;
#pragma unroll
  for (int j = 0; j < 8; j++) r.v[j] = t[j];
  fp_final_sub(r, t8);
  return r;
}

}  // namespace vimz
