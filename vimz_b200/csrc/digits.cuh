// digits.cuh -- scalar recoding shared by the MSM kernels and by the cross-term kernels (which histogram the
// digits of T while it is still in registers, saving the MSM's own counting pass).
#pragma once
#include "fp.cuh"

namespace vimz {

// ---- signed-digit recoding -------------------------------------------------------------------
// raw scalar (canonical, NOT Montgomery) -> digits d_j in [-2^(c-1), 2^(c-1)], j < nwin.
// f(j, magnitude, negative) is called for every non-zero digit.
// The limbs are consumed in order through a 64-bit bit buffer, so a window costs a handful of instructions
// (no dynamic limb indexing); the trip counts depend only on (c, nwin), i.e. they are uniform across a warp.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t (&s)[8], int c, int nwin, Fn f) {
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1);
  uint64_t buf = 0;
  int have = 0, j = 0;
  uint32_t carry = 0;
  auto emit = [&]() {
    uint32_t d = ((uint32_t)buf & mask) + carry;
    buf >>= c;
    carry = 0;
    bool neg = false;
    if (d > half && j != nwin - 1) {
      d = (1u << c) - d;
      neg = true;
      carry = 1;
    }
    if (d != 0) f(j, d, neg);
    j++;
  };
#pragma unroll
  for (int k = 0; k < 8; k++) {
    buf |= (uint64_t)s[k] << have;  // have < c <= 32 here, so the 32 new bits fit
    have += 32;
    while (have >= c && j < nwin) {
      emit();
      have -= c;
    }
  }
  while (j < nwin) emit();  // windows reaching past bit 255: the missing bits are zero
}


// Warp-aggregated bucket counters.  A witness vector is ~90 % 0/1, so most lanes of a warp hit the SAME bucket
// (digit 1 of window 0) and T's top window lands in a few dozen buckets: same-address atomics serialise in L2 and
// made the histogram atomics-bound (k_msm_scatter aggregates its returning atomics the same way).  The lanes that are converged here and target the same counter
// are matched (MATCH.ANY); one of them adds the group's size.
constexpr uint32_t AGG_MAX_DIGIT = 16;  // digits up to this magnitude take the warp-aggregated path
__device__ __forceinline__ void warp_count_add(uint32_t* counter) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, (unsigned long long)counter);
  if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(counter, (uint32_t)__popc(peers));
}
// histogram the bucket digits of one Montgomery-form scalar (the counting half of the counting sort)
// The recoded digits of scalar i, written ONCE: digits[j * stride + i] = magnitude | sign << 31 (0 = no insertion),
// and histogrammed into counts[magnitude - 1].  The counting sort then runs one thread per (scalar, window) entry
// (k_msm_scatter) instead of one thread per scalar walking its windows behind a chain of atomic round trips.
// s*P = (q-s)*(-P): scalars above (q-1)/2 are recoded by their (smaller) negative, with the point's sign flipped.
// AGG: warp-aggregate the counter updates (pays for witness vectors; costs ~20 % on uniform digits).
template <class F, bool AGG = false>
__device__ __forceinline__ void recode_scalar(const Fp<F>& mont, int c, int nwin, uint32_t* __restrict__ counts,
                                              uint32_t* __restrict__ digits, size_t stride, size_t i) {
  Fp<F> s = fp_from_mont(mont);
  const bool flip = fp_gt_half(s);
  if (flip) s = fp_neg(s);
  int next = 0;  // windows below `next` are written
  for_each_digit(s.v, c, nwin, [&](int j, uint32_t mag, bool neg) {
    for (; next < j; next++) digits[(size_t)next * stride + i] = 0;
    digits[(size_t)j * stride + i] = mag | ((neg != flip) ? 0x80000000u : 0u);
    next = j + 1;
    if (counts) {  // nullptr: digits only (direct-table keys have no buckets to size)
      // Only SMALL digits are shared by many lanes (the 0/1 wires of a witness: digit 1 of window 0; the one- or two-bit top
      // window of the cross term): those are matched and added once per group.  The other windows hold uniform c-bit digits,
      // for which MATCH.ANY found nothing to merge and cost more than the atomics it saved (k_cross_finish: 45 % of its stall
      // samples were the short scoreboard of the match unit, 52 us for 131 k rows) -- they go out as plain reductions.
      if (AGG && mag <= AGG_MAX_DIGIT) warp_count_add(&counts[mag - 1]);
      else atomicAdd(&counts[mag - 1], 1u);
    }
  });
  for (; next < nwin; next++) digits[(size_t)next * stride + i] = 0;
}

}  // namespace vimz
