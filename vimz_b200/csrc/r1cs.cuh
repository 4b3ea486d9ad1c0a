// r1cs.cuh -- relaxed-R1CS vector kernels: CSR mat-vec triple, fused cross-term, witness fold.
//
// Replaces ([EXT nova-snark 0.23.0] src/r1cs.rs):
//   R1CSShape::multiply_vec   -> k_spmv3
//   R1CSShape::commit_T (T)   -> k_cross_term (A.z1, B.z1, C.z1, A.z2, B.z2, C.z2 and T in ONE pass
//                                over the matrices: the six products are never written to HBM)
//   RelaxedR1CSWitness::fold  -> k_axpy
// All of these are HBM-bound: per non-zero 4 B column + 32 B value are streamed once, z is gathered
// (it is 4-22 MB and lives in the 126 MB L2), outputs are written once with 128-bit stores.
#pragma once
#include "common.cuh"
#include "digits.cuh"
#include "fp.cuh"

namespace vimz {

// z = (W || tail) where tail = (u, X_0, ..) : column `col` of the R1CS variable space.
template <class F>
VIMZ_DI Fp<F> load_z(const void* __restrict__ W, const void* __restrict__ tail, uint32_t n, uint32_t col) {
  const char* p = col < n ? reinterpret_cast<const char*>(W) + (size_t)col * 32
                          : reinterpret_cast<const char*>(tail) + (size_t)(col - n) * 32;
  return Fp<F>::load(p);
}

struct CsrView {
  const uint32_t* rowptr;
  const uint32_t* col;
  const void* val;
};

// optional fusion with the MSM that follows: recode (digits != nullptr) and histogram (counts != nullptr: bucket path;
// a direct-table key needs no histogram) T's digits here
struct DigitCount {
  uint32_t* counts;
  int c, nwin;
  uint32_t* digits;  // [nwin][stride] recoded digits of T (read by the MSM's scatter)
  size_t stride;
};

// one thread per (matrix, row); blockIdx.y selects A/B/C
template <class F>
__global__ void __launch_bounds__(256) k_spmv3(CsrView A, CsrView B, CsrView Cm, uint32_t m, uint32_t n,
                                               const void* __restrict__ W, const void* __restrict__ tail,
                                               void* __restrict__ Az, void* __restrict__ Bz, void* __restrict__ Cz) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  CsrView M = blockIdx.y == 0 ? A : (blockIdx.y == 1 ? B : Cm);
  void* out = blockIdx.y == 0 ? Az : (blockIdx.y == 1 ? Bz : Cz);
  Fp<F> acc = Fp<F>::zero();
  uint32_t beg = M.rowptr[row], end = M.rowptr[row + 1];
  for (uint32_t k = beg; k < end; k++) {
    Fp<F> v = Fp<F>::load_nc(reinterpret_cast<const char*>(M.val) + (size_t)k * 32);
    Fp<F> z = load_z<F>(W, tail, n, __ldg(M.col + k));
    acc = fp_add(acc, coeff_mul(v, z));
  }
  acc.store(reinterpret_cast<char*>(out) + (size_t)row * 32);
}

// Row classes by non-zeros over A+B+C.  Every non-zero is a dependent chain col -> z[col] (two L2 gathers),
// so a thread that owns a 30-term Poseidon row is latency-bound for ~60 us: give such rows several lanes.
constexpr uint32_t R1CS_SHORT_ROW = 6;   // <= 6 : one thread   (bit / copy rows: ~3 non-zeros)
constexpr uint32_t R1CS_LONG_ROW = 64;   // 7..64: 8 lanes ; > 64: a whole warp (240-term packing rows)

// v * z with the two overwhelmingly common coefficients (1 and -1: bit / copy / C rows) short-cut
template <class F>
VIMZ_DI Fp<F> coeff_mul(const Fp<F>& v, const Fp<F>& z) {
  bool is_one = true, is_m1 = true;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    is_one &= v.v[i] == F::one(i);
    // -1 in Montgomery form = p - R mod p
  }
  if (is_one) return z;
  Fp<F> m1 = fp_neg(Fp<F>::one());
  is_m1 = (v == m1);
  if (is_m1) return fp_neg(z);
  return fp_mul_noinline<F>(v, z);
}

template <class F>
VIMZ_DI void row_dot2(const CsrView& M, uint32_t beg, uint32_t end, uint32_t stride, uint32_t n,
                      const void* __restrict__ W1, const void* __restrict__ t1,
                      const void* __restrict__ W2, const void* __restrict__ t2, Fp<F>& d1, Fp<F>& d2) {
  d1 = Fp<F>::zero();
  d2 = Fp<F>::zero();
  for (uint32_t k = beg; k < end; k += stride) {
    Fp<F> v = Fp<F>::load_nc(reinterpret_cast<const char*>(M.val) + (size_t)k * 32);
    uint32_t c = __ldg(M.col + k);
    d1 = fp_add(d1, coeff_mul(v, load_z<F>(W1, t1, n, c)));
    d2 = fp_add(d2, coeff_mul(v, load_z<F>(W2, t2, n, c)));
  }
}

template <class F>
VIMZ_DI Fp<F> cross_term_row(const Fp<F>& a1, const Fp<F>& a2, const Fp<F>& b1, const Fp<F>& b2, const Fp<F>& c1, const Fp<F>& c2,
                             const Fp<F>& u1) {
  Fp<F> t = fp_add(fp_mul_noinline<F>(a1, b2), fp_mul_noinline<F>(a2, b1));
  t = fp_sub(t, fp_mul_noinline<F>(u1, c2));
  return fp_sub(t, c1);
}

// T[row] = cross term of the six row products; optional digit histogram for the commit(T) that follows.
// Out of line (with the field product): the cross-term kernels are latency-bound and were stalling on
// instruction fetch with ~185 KB of inlined code; one shared copy keeps them inside the instruction caches.
#ifndef VIMZ_CROSS_AGG
#define VIMZ_CROSS_AGG true  // warp-aggregated histogram of T's digits: late in a proof ~10^5 rows share the top window's digit
#endif
template <class F>
__device__ __noinline__ void cross_term_finish(Fp<F> a1, Fp<F> a2, Fp<F> b1, Fp<F> b2, Fp<F> c1, Fp<F> c2, Fp<F> u1, void* T, uint32_t row,
                                               DigitCount dc) {
  Fp<F> t = cross_term_row<F>(a1, a2, b1, b2, c1, c2, u1);
  t.store(reinterpret_cast<char*>(T) + (size_t)row * 32);
  if (dc.digits) recode_scalar<F, VIMZ_CROSS_AGG>(t, dc.c, dc.nwin, dc.counts, dc.digits, dc.stride, row);
}

// sum over the GROUP lanes of a row group (GROUP = 8 or 32, groups are aligned inside the warp)
template <class F, int GROUP>
__device__ __noinline__ Fp<F> group_sum_fp(Fp<F> v) {
#pragma unroll
  for (int o = GROUP / 2; o > 0; o >>= 1) {
    Fp<F> other;
#pragma unroll
    for (int k = 0; k < 8; k++) other.v[k] = __shfl_down_sync(0xffffffffu, v.v[k], o);
    v = fp_add(v, other);
  }
  return v;
}

struct CrossArgs {
  CsrView A, B, Cm;
  uint32_t m, n;
  const void *W1, *tail1, *W2, *tail2;
  void* T;
  DigitCount dc;
};

// one thread per short row (<= R1CS_SHORT_ROW non-zeros over A+B+C); rows owned by the group roles are skipped
template <class F>
VIMZ_DI void cross_term_short(const CrossArgs& a, uint32_t row) {
  if (row >= a.m) return;
  uint32_t ab = a.A.rowptr[row], ae = a.A.rowptr[row + 1], bb = a.B.rowptr[row], be = a.B.rowptr[row + 1];
  uint32_t cb = a.Cm.rowptr[row], ce = a.Cm.rowptr[row + 1];
  if ((ae - ab) + (be - bb) + (ce - cb) > R1CS_SHORT_ROW) return;
  Fp<F> a1, a2, b1, b2, c1, c2;
  row_dot2<F>(a.A, ab, ae, 1, a.n, a.W1, a.tail1, a.W2, a.tail2, a1, a2);
  row_dot2<F>(a.B, bb, be, 1, a.n, a.W1, a.tail1, a.W2, a.tail2, b1, b2);
  row_dot2<F>(a.Cm, cb, ce, 1, a.n, a.W1, a.tail1, a.W2, a.tail2, c1, c2);
  cross_term_finish<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(a.tail1), a.T, row, a.dc);
}

// GROUP lanes per row (8 for Poseidon-like rows, 32 for the 240-term Num2Bits packing rows):
// lanes stride the non-zeros, partial dot products are folded with shuffles (group-collective: invalid groups join).
template <class F, int GROUP>
VIMZ_DI void cross_term_grouped_row(const CrossArgs& a, uint32_t row, bool valid) {
  const uint32_t lane = threadIdx.x % GROUP;
  Fp<F> a1, a2, b1, b2, c1, c2;
  const uint32_t none = 0;
  row_dot2<F>(a.A, valid ? a.A.rowptr[row] + lane : none, valid ? a.A.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, a1, a2);
  row_dot2<F>(a.B, valid ? a.B.rowptr[row] + lane : none, valid ? a.B.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, b1, b2);
  row_dot2<F>(a.Cm, valid ? a.Cm.rowptr[row] + lane : none, valid ? a.Cm.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, c1, c2);
  a1 = group_sum_fp<F, GROUP>(a1); a2 = group_sum_fp<F, GROUP>(a2);
  b1 = group_sum_fp<F, GROUP>(b1); b2 = group_sum_fp<F, GROUP>(b2);
  c1 = group_sum_fp<F, GROUP>(c1); c2 = group_sum_fp<F, GROUP>(c2);
  if (lane == 0 && valid) {
    cross_term_finish<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(a.tail1), a.T, row, a.dc);
  }
}
template <class F, int GROUP>
VIMZ_DI void cross_term_grouped(const CrossArgs& a, const uint32_t* __restrict__ rows, uint32_t n_rows, uint32_t g) {
  const bool valid = g < n_rows;  // whole groups are valid or not
  cross_term_grouped_row<F, GROUP>(a, rows[valid ? g : 0], valid);
}

// T[i] = Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1   (u2 = 1; u1 = tail1[0]) for ALL rows in one launch of 128-thread
// blocks; the block index selects the row class (the classes with the longest per-row chains are scheduled first):
//   [0, nb_long)          rows with > R1CS_LONG_ROW non-zeros: a warp each
//   [nb_long, +nb_mid)    rows with R1CS_SHORT_ROW+1 .. R1CS_LONG_ROW non-zeros: 8 lanes each
//   remaining blocks      every other row: a thread each
template <class F>
__global__ void __launch_bounds__(128) k_cross_term(CrossArgs a, const uint32_t* __restrict__ long_rows, uint32_t n_long, uint32_t nb_long,
                                                    const uint32_t* __restrict__ mid_rows, uint32_t n_mid, uint32_t nb_mid) {
  if (blockIdx.x < nb_long) {
    cross_term_grouped<F, 32>(a, long_rows, n_long, (blockIdx.x * blockDim.x + threadIdx.x) / 32);
  } else if (blockIdx.x < nb_long + nb_mid) {
    cross_term_grouped<F, 8>(a, mid_rows, n_mid, ((blockIdx.x - nb_long) * blockDim.x + threadIdx.x) / 8);
  } else {
    cross_term_short<F>(a, (blockIdx.x - nb_long - nb_mid) * blockDim.x + threadIdx.x);
  }
}

// ---- streamed cross term (default) -------------------------------------------------------------------------
// The row-class kernel above is bound by dependent loads (rowptr -> col/val -> z) of threads that own whole rows.
// Here a 256-thread block owns a CHUNK of consecutive rows holding <= CROSS_CHUNK_NNZ non-zeros over A+B+C:
//   phase 1: the threads stride the chunk's non-zeros of A, then B, then C -- coalesced (col, value-index) reads,
//            independent z1/z2 gathers (L2), products v*z1, v*z2 written to shared memory;
//   phase 2: one thread per row sums its products out of shared memory and forms T (rows above CROSS_ROW_COOP
//            non-zeros are summed by a warp with shuffles instead).
// Coefficients come from the shape's value dictionary (4 B per non-zero instead of 32 B; +1 / -1 need no product).
#ifndef VIMZ_CROSS_CHUNK_NNZ
#define VIMZ_CROSS_CHUNK_NNZ 1024
#endif
constexpr uint32_t CROSS_CHUNK_NNZ = VIMZ_CROSS_CHUNK_NNZ;  // products per chunk: 64 B of shared memory each
constexpr uint32_t CROSS_CHUNK_ROWS = 256;  // = block size
constexpr uint32_t CROSS_ROW_MAX = CROSS_CHUNK_NNZ < 512 ? CROSS_CHUNK_NNZ : 512;     // longer rows are chunks of their own (a warp walks them in global memory)
constexpr uint32_t CROSS_ROW_COOP = 32;     // rows above this are summed by a warp in phase 2

struct CrossStreamArgs {
  CrossArgs a;
  const uint32_t* vidx[3];
  const void* dict;
  const uint32_t* chunk_start;
  // CACHED variant (the resident accumulator): the products with the running z1 are linear in the fold,
  //   A (z1 + r z2) = A z1 + r A z2,
  // so the accumulator keeps (Az1, Bz1, Cz1) = cache1[3][m] resident and folds them in step_end with the
  // (Az2, Bz2, Cz2) = cache2[3][m] written here: no z1 gather, no z1 product, half the shared memory.
  const void* cache1;
  void* cache2;
};

#ifndef VIMZ_CROSS_MINB
#define VIMZ_CROSS_MINB 2
#endif
template <class F, bool CACHED>
__global__ void __launch_bounds__(256, CACHED ? VIMZ_CROSS_MINB : 2) k_cross_term_stream(CrossStreamArgs s) {
  extern __shared__ __align__(32) unsigned char cross_smem[];
  char* P1 = reinterpret_cast<char*>(cross_smem);
  char* P2 = CACHED ? P1 : P1 + (size_t)CROSS_CHUNK_NNZ * 32;  // CACHED: only the z2 products are staged
  __shared__ uint32_t big_rows[32];
  __shared__ uint32_t nbig;
  const CrossArgs& a = s.a;
  const uint32_t cs = s.chunk_start[blockIdx.x], ce = s.chunk_start[blockIdx.x + 1];
  const uint32_t r0 = cs & 0x7fffffffu, r1 = ce & 0x7fffffffu;
  const size_t m32 = (size_t)a.m * 32;
  auto keep2 = [&](uint32_t row_, const Fp<F>& a2_, const Fp<F>& b2_, const Fp<F>& c2_) {  // (Az2, Bz2, Cz2)[row] for step_end
    char* c2p = reinterpret_cast<char*>(s.cache2) + (size_t)row_ * 32;
    a2_.store(c2p); b2_.store(c2p + m32); c2_.store(c2p + 2 * m32);
  };
  if (cs >> 31) {  // a single row longer than CROSS_ROW_MAX: a warp walks it in global memory (both z: rare)
    if (threadIdx.x < 32) {
      if (!CACHED) {
        cross_term_grouped_row<F, 32>(a, r0, true);
      } else {
        const uint32_t lane = threadIdx.x;
        Fp<F> a1, a2, b1, b2, c1, c2;
        row_dot2<F>(a.A, a.A.rowptr[r0] + lane, a.A.rowptr[r0 + 1], 32, a.n, a.W1, a.tail1, a.W2, a.tail2, a1, a2);
        row_dot2<F>(a.B, a.B.rowptr[r0] + lane, a.B.rowptr[r0 + 1], 32, a.n, a.W1, a.tail1, a.W2, a.tail2, b1, b2);
        row_dot2<F>(a.Cm, a.Cm.rowptr[r0] + lane, a.Cm.rowptr[r0 + 1], 32, a.n, a.W1, a.tail1, a.W2, a.tail2, c1, c2);
        a1 = group_sum_fp<F, 32>(a1); a2 = group_sum_fp<F, 32>(a2);
        b1 = group_sum_fp<F, 32>(b1); b2 = group_sum_fp<F, 32>(b2);
        c1 = group_sum_fp<F, 32>(c1); c2 = group_sum_fp<F, 32>(c2);
        if (lane == 0) {
          cross_term_finish<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(a.tail1), a.T, r0, a.dc);
          keep2(r0, a2, b2, c2);
        }
      }
    }
    return;
  }
  if (threadIdx.x == 0) nbig = 0;
  const uint32_t begA = a.A.rowptr[r0], begB = a.B.rowptr[r0], begC = a.Cm.rowptr[r0];
  const uint32_t nA = a.A.rowptr[r1] - begA, nB = a.B.rowptr[r1] - begB, nC = a.Cm.rowptr[r1] - begC;
  const uint32_t total = nA + nB + nC;
  // this thread's row bounds for phase 2, fetched now so their latency hides behind phase 1
  const uint32_t row = r0 + threadIdx.x;
  const bool have_row = row < r1;
  uint32_t ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0, rc0 = 0, rc1 = 0;
  if (have_row) {
    ra0 = a.A.rowptr[row]; ra1 = a.A.rowptr[row + 1];
    rb0 = a.B.rowptr[row]; rb1 = a.B.rowptr[row + 1];
    rc0 = a.Cm.rowptr[row]; rc1 = a.Cm.rowptr[row + 1];
  }
  const Fp<F> u1 = Fp<F>::load(a.tail1);
  auto entry = [&](uint32_t k, uint32_t& col, uint32_t& vi) {
    if (k < nA) { col = __ldg(a.A.col + begA + k); vi = __ldg(s.vidx[0] + begA + k); }
    else if (k < nA + nB) { col = __ldg(a.B.col + begB + (k - nA)); vi = __ldg(s.vidx[1] + begB + (k - nA)); }
    else { col = __ldg(a.Cm.col + begC + (k - nA - nB)); vi = __ldg(s.vidx[2] + begC + (k - nA - nB)); }
  };
  auto product = [&](uint32_t k, uint32_t vi, Fp<F> z1, Fp<F> z2) {
    if (vi >= 2) {
      Fp<F> v = Fp<F>::load_nc(reinterpret_cast<const char*>(s.dict) + (size_t)vi * 32);
      if (!CACHED) z1 = fp_mul_noinline<F>(v, z1);
      z2 = fp_mul_noinline<F>(v, z2);
    } else if (vi == 1) {
      if (!CACHED) z1 = fp_neg(z1);
      z2 = fp_neg(z2);
    }
    if (!CACHED) z1.store(P1 + (size_t)k * 32);
    z2.store(P2 + (size_t)k * 32);
  };
  for (uint32_t k = threadIdx.x; k < total; k += 512) {  // two entries in flight per thread
    const uint32_t k2 = k + 256;
    const bool two = k2 < total;
    uint32_t c0, v0, c1 = 0, v1 = 0;
    entry(k, c0, v0);
    if (two) entry(k2, c1, v1);
    Fp<F> x1 = Fp<F>::zero(), y1 = Fp<F>::zero(), y2 = Fp<F>::zero();
    if (!CACHED) x1 = load_z<F>(a.W1, a.tail1, a.n, c0);
    Fp<F> x2 = load_z<F>(a.W2, a.tail2, a.n, c0);
    if (two) {
      if (!CACHED) y1 = load_z<F>(a.W1, a.tail1, a.n, c1);
      y2 = load_z<F>(a.W2, a.tail2, a.n, c1);
    }
    product(k, v0, x1, x2);
    if (two) product(k2, v1, y1, y2);
  }
  __syncthreads();
  auto sum2 = [&](uint32_t first, uint32_t count, uint32_t start, uint32_t stride, Fp<F>& d1, Fp<F>& d2) {
    d1 = Fp<F>::zero();
    d2 = Fp<F>::zero();
    for (uint32_t j = start; j < count; j += stride) {
      if (!CACHED) d1 = fp_add(d1, Fp<F>::load(P1 + (size_t)(first + j) * 32));
      d2 = fp_add(d2, Fp<F>::load(P2 + (size_t)(first + j) * 32));
    }
  };
  auto cached1 = [&](uint32_t row_, Fp<F>& a1_, Fp<F>& b1_, Fp<F>& c1_) {  // (Az1, Bz1, Cz1)[row] kept by the accumulator
    const char* c1p = reinterpret_cast<const char*>(s.cache1) + (size_t)row_ * 32;
    a1_ = Fp<F>::load(c1p); b1_ = Fp<F>::load(c1p + m32); c1_ = Fp<F>::load(c1p + 2 * m32);
  };
  if (have_row) {
    const uint32_t cnt = (ra1 - ra0) + (rb1 - rb0) + (rc1 - rc0);
    if (cnt > CROSS_ROW_COOP) {
      big_rows[atomicAdd(&nbig, 1u)] = threadIdx.x;  // at most CROSS_CHUNK_NNZ / (CROSS_ROW_COOP + 1) = 31 per chunk
    } else {
      Fp<F> a1, a2, b1, b2, c1, c2;
      sum2(ra0 - begA, ra1 - ra0, 0, 1, a1, a2);
      sum2(nA + (rb0 - begB), rb1 - rb0, 0, 1, b1, b2);
      sum2(nA + nB + (rc0 - begC), rc1 - rc0, 0, 1, c1, c2);
      if (CACHED) cached1(row, a1, b1, c1);
      cross_term_finish<F>(a1, a2, b1, b2, c1, c2, u1, a.T, row, a.dc);
      if (CACHED) keep2(row, a2, b2, c2);
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x >> 5; i < nbig; i += 8) {  // warp-uniform
    const uint32_t br = r0 + big_rows[i];
    const uint32_t qa0 = a.A.rowptr[br], qa1 = a.A.rowptr[br + 1], qb0 = a.B.rowptr[br], qb1 = a.B.rowptr[br + 1];
    const uint32_t qc0 = a.Cm.rowptr[br], qc1 = a.Cm.rowptr[br + 1];
    Fp<F> a1, a2, b1, b2, c1, c2;
    sum2(qa0 - begA, qa1 - qa0, lane, 32, a1, a2);
    sum2(nA + (qb0 - begB), qb1 - qb0, lane, 32, b1, b2);
    sum2(nA + nB + (qc0 - begC), qc1 - qc0, lane, 32, c1, c2);
    if (!CACHED) { a1 = group_sum_fp<F, 32>(a1); b1 = group_sum_fp<F, 32>(b1); c1 = group_sum_fp<F, 32>(c1); }
    a2 = group_sum_fp<F, 32>(a2);
    b2 = group_sum_fp<F, 32>(b2);
    c2 = group_sum_fp<F, 32>(c2);
    if (lane == 0) {
      if (CACHED) cached1(br, a1, b1, c1);
      cross_term_finish<F>(a1, a2, b1, b2, c1, c2, u1, a.T, br, a.dc);
      if (CACHED) keep2(br, a2, b2, c2);
    }
  }
}

// out[i] = a[i] + r * b[i]
template <class F>
__global__ void __launch_bounds__(256) k_axpy(const void* __restrict__ a, const void* __restrict__ b, Fp<F> r, size_t len,
                                              void* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
    Fp<F> x = Fp<F>::load(reinterpret_cast<const char*>(a) + i * 32);
    Fp<F> y = Fp<F>::load(reinterpret_cast<const char*>(b) + i * 32);
    fp_add(x, fp_mul(r, y)).store(reinterpret_cast<char*>(out) + i * 32);
  }
}

// RelaxedR1CSWitness::fold + the scalar half of RelaxedR1CSInstance::fold in one launch:
// W1 += r*W2, E1 += r*T, (u1, X1) += r*(1, X2) -- three in-place segments sharing r.
struct AxpySeg {
  void* a;
  const void* b;
  size_t len;
};
constexpr int AXPY_MAX_SEGS = 6;
struct AxpySegs {
  AxpySeg s[AXPY_MAX_SEGS];
  size_t end[AXPY_MAX_SEGS];  // running end offsets
  int count;
};
template <class F>
__global__ void __launch_bounds__(256) k_axpy3(AxpySegs segs, Fp<F> r) {
  const size_t total = segs.end[segs.count - 1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int k = 0;
#pragma unroll
    for (int q = 0; q < AXPY_MAX_SEGS - 1; q++)
      if (q < segs.count - 1 && i >= segs.end[q]) k = q + 1;
    const size_t j = i - (k ? segs.end[k - 1] : 0);
    char* pa = reinterpret_cast<char*>(segs.s[k].a) + j * 32;
    Fp<F> x = Fp<F>::load(pa);
    Fp<F> y = Fp<F>::load(reinterpret_cast<const char*>(segs.s[k].b) + j * 32);
    fp_add(x, fp_mul(r, y)).store(pa);
  }
}

// element-wise field op for the parity tests of fp.cuh (op 0 mul, 1 add, 2 sub)
template <class F>
__global__ void k_field_op(int op, const void* __restrict__ a, const void* __restrict__ b, size_t len, void* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  Fp<F> x = Fp<F>::load(reinterpret_cast<const char*>(a) + i * 32);
  Fp<F> y = Fp<F>::load(reinterpret_cast<const char*>(b) + i * 32);
  Fp<F> r = op == 0 ? fp_mul(x, y) : (op == 1 ? fp_add(x, y) : fp_sub(x, y));
  r.store(reinterpret_cast<char*>(out) + i * 32);
}

}  // namespace vimz
