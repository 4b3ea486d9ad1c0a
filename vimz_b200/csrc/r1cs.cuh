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
// shifted = true (a "booleanity" row of a shape whose accumulator keeps K_S, see k_cross_finish): the digits recoded for the
// commit are those of T + Az1, the stored T is the true one.
template <class F>
__device__ __noinline__ void cross_term_finish(Fp<F> a1, Fp<F> a2, Fp<F> b1, Fp<F> b2, Fp<F> c1, Fp<F> c2, Fp<F> u1, void* T, uint32_t row,
                                               DigitCount dc, bool shifted = false) {
  Fp<F> t = cross_term_row<F>(a1, a2, b1, b2, c1, c2, u1);
  t.store(reinterpret_cast<char*>(T) + (size_t)row * 32);
  if (shifted) t = fp_add(t, a1);
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

// ---- streamed mat-vec triple + element-wise cross term (default) -------------------------------------------
// The row-class kernel above is bound by dependent loads (rowptr -> col/val -> z) of threads that own whole rows, and a
// single fused kernel (mat-vecs + T + digit recoding, 122 registers, two blocks per SM) spent its time in three waves of
// latency-bound blocks (round 1: 67 us for 130 k rows, 11 % of the HBM roofline).  Now two launches:
//
//   k_matvec_stream   one 256-thread block per CHUNK of consecutive rows (<= CROSS_CHUNK_NNZ non-zeros over A+B+C,
//                     <= CROSS_CHUNK_ROWS rows).  The chunk's (column, coefficient-index) pairs were packed into ONE
//                     contiguous, 16-byte aligned stream when the shape was uploaded:
//                       1. one thread issues a TMA bulk copy (cp.async.bulk, mbarrier complete_tx) of that stream into shared
//                          memory -- no rowptr -> col dependent loads, no per-thread index traffic;
//                       2. every thread issues its z gathers as cp.async (LDGSTS) 2 x 16 B straight into the product slots:
//                          all of a chunk's gathers are in flight at once, without holding registers;
//                       3. coefficients: +1 / -1 need no product, the others come from the shape's value dictionary;
//                       4. one thread per row sums its slots (a warp for rows above CROSS_ROW_COOP non-zeros) and writes
//                          (Az, Bz, Cz)[row].
//                     ~60 registers: four to five blocks per SM instead of two.
//   k_cross_finish    one thread per row, perfectly coalesced: T = Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1, plus the signed-digit
//                     recoding and bucket histogram of T for the commit that follows.
// The resident accumulator keeps (Az1, Bz1, Cz1) folded (A (z1 + r z2) = A z1 + r A z2), so a step runs the mat-vec
// kernel once, on z2; the stand-alone commit_T runs it twice.
#ifndef VIMZ_CROSS_CHUNK_NNZ
#define VIMZ_CROSS_CHUNK_NNZ 1024
#endif
constexpr uint32_t CROSS_CHUNK_NNZ = VIMZ_CROSS_CHUNK_NNZ;  // products per chunk: 8 B of index stream + 32 B of product slot each
constexpr uint32_t CROSS_CHUNK_ROWS = 256;  // = block size
constexpr uint32_t CROSS_ROW_MAX = CROSS_CHUNK_NNZ < 512 ? CROSS_CHUNK_NNZ : 512;     // longer rows are chunks of their own (a warp walks them in global memory)
constexpr uint32_t CROSS_ROW_COOP = 32;     // rows above this are summed by a warp
constexpr uint32_t CHUNK_LONG_ROW = 0xffffffffu;

struct ChunkDesc {        // 32 bytes, one per chunk
  uint32_t off;           // first (col, vidx) pair of the chunk in the packed stream (even => 16-byte aligned); CHUNK_LONG_ROW: a single long row
  uint32_t nA, nB, nC;    // non-zeros of the chunk per matrix; the stream holds A's, then B's, then C's, in CSR order
  uint32_t r0, nrows;
  uint32_t pad0, pad1;
};

struct MatvecStreamArgs {
  CsrView A, B, Cm;       // row pointers (and, for long rows, columns / values)
  uint32_t m, n;
  const void *W, *tail;   // z = (W || tail)
  const uint2* stream;    // packed (col, vidx) pairs
  const ChunkDesc* desc;
  const void* dict;
  void* out;              // (Az, Bz, Cz)[3][m]
};

VIMZ_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class F>
__global__ void __launch_bounds__(256, 4) k_matvec_stream(MatvecStreamArgs s) {
  extern __shared__ __align__(128) unsigned char mv_smem[];
  uint2* idx = reinterpret_cast<uint2*>(mv_smem);                       // [CROSS_CHUNK_NNZ]
  char* P = reinterpret_cast<char*>(mv_smem) + (size_t)CROSS_CHUNK_NNZ * 8;  // [CROSS_CHUNK_NNZ][32]
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t big_rows[32];
  __shared__ uint32_t nbig;
  const uint4* dp = reinterpret_cast<const uint4*>(s.desc + blockIdx.x);
  const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
  const uint32_t off = d0.x, nA = d0.y, nB = d0.z, nC = d0.w, r0 = d1.x, nrows = d1.y;
  const size_t m32 = (size_t)s.m * 32;
  char* out = reinterpret_cast<char*>(s.out);
  if (off == CHUNK_LONG_ROW) {  // a single row longer than CROSS_ROW_MAX: a warp walks it in global memory (rare)
    if (threadIdx.x < 32) {
      const uint32_t lane = threadIdx.x;
      Fp<F> acc[3];
      const CsrView* M[3] = {&s.A, &s.B, &s.Cm};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        acc[k] = Fp<F>::zero();
        for (uint32_t e = M[k]->rowptr[r0] + lane; e < M[k]->rowptr[r0 + 1]; e += 32) {
          Fp<F> v = Fp<F>::load_nc(reinterpret_cast<const char*>(M[k]->val) + (size_t)e * 32);
          acc[k] = fp_add(acc[k], coeff_mul(v, load_z<F>(s.W, s.tail, s.n, __ldg(M[k]->col + e))));
        }
        acc[k] = group_sum_fp<F, 32>(acc[k]);
      }
      if (lane == 0) {
        acc[0].store(out + (size_t)r0 * 32); acc[1].store(out + m32 + (size_t)r0 * 32); acc[2].store(out + 2 * m32 + (size_t)r0 * 32);
      }
    }
    return;
  }
  const uint32_t total = nA + nB + nC;
  const uint32_t mb = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    nbig = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && total) {  // 1. TMA bulk copy of the chunk's index stream (16-byte granules; the stream is padded)
    const uint32_t bytes = (total * 8 + 15) & ~15u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(idx)),
                 "l"(s.stream + off), "r"(bytes), "r"(mb)
                 : "memory");
  }
  // this thread's row bounds for step 4, fetched now so their latency hides behind steps 1-3
  const uint32_t row = r0 + threadIdx.x;
  const bool have_row = threadIdx.x < nrows;
  const uint32_t begA = __ldg(s.A.rowptr + r0), begB = __ldg(s.B.rowptr + r0), begC = __ldg(s.Cm.rowptr + r0);
  uint32_t ra0 = 0, ra1 = 0, rb0 = 0, rb1 = 0, rc0 = 0, rc1 = 0;
  if (have_row) {
    ra0 = __ldg(s.A.rowptr + row); ra1 = __ldg(s.A.rowptr + row + 1);
    rb0 = __ldg(s.B.rowptr + row); rb1 = __ldg(s.B.rowptr + row + 1);
    rc0 = __ldg(s.Cm.rowptr + row); rc1 = __ldg(s.Cm.rowptr + row + 1);
  }
  if (total) {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb) : "memory");
    }
  }
  // 2. z gathers: cp.async straight into the product slots, all in flight at once
  for (uint32_t k = threadIdx.x; k < total; k += 256) {
    const uint32_t col = idx[k].x;
    const char* src = col < s.n ? reinterpret_cast<const char*>(s.W) + (size_t)col * 32 : reinterpret_cast<const char*>(s.tail) + (size_t)(col - s.n) * 32;
    const uint32_t dst = smem_u32(P + (size_t)k * 32);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 16) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // 3. coefficients (own slots only: no barrier needed before, one after)
  for (uint32_t k = threadIdx.x; k < total; k += 256) {
    const uint32_t vi = idx[k].y;
    if (vi == 0) continue;
    Fp<F> z = Fp<F>::load(P + (size_t)k * 32);
    if (vi == 1) z = fp_neg(z);
    else z = fp_mul_noinline<F>(Fp<F>::load_nc(reinterpret_cast<const char*>(s.dict) + (size_t)vi * 32), z);
    z.store(P + (size_t)k * 32);
  }
  __syncthreads();
  // 4. row sums
  auto sum = [&](uint32_t first, uint32_t count, uint32_t start, uint32_t stride) {
    Fp<F> d = Fp<F>::zero();
    for (uint32_t j = start; j < count; j += stride) d = fp_add(d, Fp<F>::load(P + (size_t)(first + j) * 32));
    return d;
  };
  if (have_row) {
    const uint32_t cnt = (ra1 - ra0) + (rb1 - rb0) + (rc1 - rc0);
    if (cnt > CROSS_ROW_COOP) {
      big_rows[atomicAdd(&nbig, 1u)] = threadIdx.x;  // at most CROSS_CHUNK_NNZ / (CROSS_ROW_COOP + 1) = 31 per chunk
    } else {
      sum(ra0 - begA, ra1 - ra0, 0, 1).store(out + (size_t)row * 32);
      sum(nA + (rb0 - begB), rb1 - rb0, 0, 1).store(out + m32 + (size_t)row * 32);
      sum(nA + nB + (rc0 - begC), rc1 - rc0, 0, 1).store(out + 2 * m32 + (size_t)row * 32);
    }
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x >> 5; i < nbig; i += 8) {  // warp-uniform
    const uint32_t br = r0 + big_rows[i];
    const uint32_t qa0 = s.A.rowptr[br], qa1 = s.A.rowptr[br + 1], qb0 = s.B.rowptr[br], qb1 = s.B.rowptr[br + 1];
    const uint32_t qc0 = s.Cm.rowptr[br], qc1 = s.Cm.rowptr[br + 1];
    Fp<F> a = group_sum_fp<F, 32>(sum(qa0 - begA, qa1 - qa0, lane, 32));
    Fp<F> b = group_sum_fp<F, 32>(sum(nA + (qb0 - begB), qb1 - qb0, lane, 32));
    Fp<F> c = group_sum_fp<F, 32>(sum(nA + nB + (qc0 - begC), qc1 - qc0, lane, 32));
    if (lane == 0) {
      a.store(out + (size_t)br * 32); b.store(out + m32 + (size_t)br * 32); c.store(out + 2 * m32 + (size_t)br * 32);
    }
  }
}

// T[row] = Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1 from the two product triples p1 = (Az1, Bz1, Cz1)[3][m], p2 = (Az2, Bz2, Cz2)[3][m];
// recodes / histograms T's digits for the commit that follows (dc.digits != nullptr).
// 128-thread blocks (10 K registers each): on the secondary curve this kernel must find room on SMs whose register files
// already hold three k_msm_direct blocks of the other lane -- with 256-thread blocks it waited ~60 us for them to drain.
constexpr int CROSS_FINISH_THREADS = 128;
//
// BOOLEANITY ROWS (rowflag != nullptr).  85-93 % of a pixel circuit's constraints are b * (b - 1) = 0 (A = {b}, B = {b, -one},
// C = {}).  For such a row T = W1[b] (W2[b] - 1) + W2[b] (W1[b] - u1), so T + Az1 = W2[b] (2 W1[b] - u1): ZERO whenever the fresh
// bit is 0, for any running instance.  The accumulator therefore keeps K_S = sum_{i in S} (A z1)_i ck_i over the set S of these
// rows -- linear in z1, so it folds like every other commitment: K_S += r * sum_{i in S} (A z2)_i ck_i, a plain sum of ~55 k points
// beside the step -- and commits T' = T + [i in S] Az1 instead of T: comm_T = commit(T') - K_S, the SAME group element, with
// half of the booleanity rows contributing no bucket insertion at all (grayscale HD: 1.32 M -> 0.77 M insertions per step).
// Exact for any W2 (a fresh wire that is not 0/1 just keeps its insertions); the true T is what is stored and folded into E.
template <class F>
__global__ void __launch_bounds__(CROSS_FINISH_THREADS) k_cross_finish(const void* __restrict__ p1, const void* __restrict__ p2, const void* __restrict__ tail1,
                                                      uint32_t m, void* __restrict__ T, DigitCount dc, const uint8_t* __restrict__ rowflag) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;  // (the warp-aggregated histogram matches the lanes that are still converged)
  const size_t m32 = (size_t)m * 32;
  const char* q1 = reinterpret_cast<const char*>(p1) + (size_t)row * 32;
  const char* q2 = reinterpret_cast<const char*>(p2) + (size_t)row * 32;
  Fp<F> a1 = Fp<F>::load(q1), b1 = Fp<F>::load(q1 + m32), c1 = Fp<F>::load(q1 + 2 * m32);
  Fp<F> a2 = Fp<F>::load(q2), b2 = Fp<F>::load(q2 + m32), c2 = Fp<F>::load(q2 + 2 * m32);
  cross_term_finish<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(tail1), T, row, dc, rowflag != nullptr && rowflag[row] != 0);
}

// out[i] = [rowflag[i]] * v[i]: the vector whose commitment updates K_S (v = A z2 of the step) or initialises it (v = A z1)
template <class F>
__global__ void __launch_bounds__(256) k_mask_rows(const void* __restrict__ v, const uint8_t* __restrict__ rowflag, uint32_t m, void* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  Fp<F> x = rowflag[i] ? Fp<F>::load(reinterpret_cast<const char*>(v) + (size_t)i * 32) : Fp<F>::zero();
  x.store(reinterpret_cast<char*>(out) + (size_t)i * 32);
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
  Fp<F> r = op == 0 ? fp_mul(x, y) : (op == 1 ? fp_add(x, y) : (op == 2 ? fp_sub(x, y) : fp_sqr(x)));  // 3: a^2 (b ignored)
  r.store(reinterpret_cast<char*>(out) + i * 32);
}

}  // namespace vimz
