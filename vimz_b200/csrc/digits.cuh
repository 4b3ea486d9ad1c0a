// digits.cuh -- scalar recoding shared by the MSM kernels and by the cross-term kernels (which histogram the
// digits of T while it is still in registers, saving the MSM's own counting pass).
#pragma once
#include "fp.cuh"

namespace vimz {

// ---- signed-digit recoding -------------------------------------------------------------------
// raw scalar (canonical, NOT Montgomery) -> digits d_j in [-2^(c-1), 2^(c-1)], j < nwin.
// f(j, magnitude, negative) is called for every non-zero digit.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t (&s)[8], int c, int nwin, Fn f) {
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1);
  uint32_t carry = 0;
  for (int j = 0; j < nwin; j++) {
    int pos = j * c;
    int limb = pos >> 5, off = pos & 31;
    uint64_t lo = 0;
    // dynamic limb index resolved with selects (registers cannot be indexed)
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (k == limb) lo |= (uint64_t)s[k];
      if (k == limb + 1) lo |= (uint64_t)s[k] << 32;
    }
    uint32_t d = (uint32_t)((lo >> off) & mask) + carry;
    carry = 0;
    bool neg = false;
    if (d > half && j != nwin - 1) {
      d = (1u << c) - d;
      neg = true;
      carry = 1;
    }
    if (d != 0) f(j, d, neg);
  }
}


// histogram the bucket digits of one Montgomery-form scalar (the counting half of the counting sort)
template <class F>
__device__ __forceinline__ void count_scalar_digits(const Fp<F>& mont, int c, int nwin, uint32_t* __restrict__ counts) {
  Fp<F> s = fp_from_mont(mont);
  if (fp_gt_half(s)) s = fp_neg(s);  // s*P = (q-s)*(-P): digits of the smaller magnitude
  for_each_digit(s.v, c, nwin, [&](int, uint32_t mag, bool) { atomicAdd(&counts[mag - 1], 1u); });
}

}  // namespace vimz
