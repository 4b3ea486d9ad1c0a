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

// optional fusion with the MSM that follows: histogram T's bucket digits here (counts == nullptr: off)
struct DigitCount {
  uint32_t* counts;
  int c, nwin;
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
  return fp_mul(v, z);
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
  Fp<F> t = fp_add(fp_mul(a1, b2), fp_mul(a2, b1));
  t = fp_sub(t, fp_mul(u1, c2));
  return fp_sub(t, c1);
}

// sum over the GROUP lanes of a row group (GROUP = 8 or 32, groups are aligned inside the warp)
template <class F, int GROUP>
VIMZ_DI Fp<F> group_sum_fp(Fp<F> v) {
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
  Fp<F> t = cross_term_row<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(a.tail1));
  t.store(reinterpret_cast<char*>(a.T) + (size_t)row * 32);
  if (a.dc.counts) count_scalar_digits(t, a.dc.c, a.dc.nwin, a.dc.counts);
}

// GROUP lanes per listed row (8 for Poseidon-like rows, 32 for the 240-term Num2Bits packing rows):
// lanes stride the non-zeros, partial dot products are folded with shuffles.  g = group index (warp-collective).
template <class F, int GROUP>
VIMZ_DI void cross_term_grouped(const CrossArgs& a, const uint32_t* __restrict__ rows, uint32_t n_rows, uint32_t g) {
  const uint32_t lane = threadIdx.x % GROUP;
  const bool valid = g < n_rows;  // whole groups are valid or not; invalid groups still join the shuffles
  const uint32_t row = rows[valid ? g : 0];
  Fp<F> a1, a2, b1, b2, c1, c2;
  const uint32_t none = 0;
  row_dot2<F>(a.A, valid ? a.A.rowptr[row] + lane : none, valid ? a.A.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, a1, a2);
  row_dot2<F>(a.B, valid ? a.B.rowptr[row] + lane : none, valid ? a.B.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, b1, b2);
  row_dot2<F>(a.Cm, valid ? a.Cm.rowptr[row] + lane : none, valid ? a.Cm.rowptr[row + 1] : none, GROUP, a.n, a.W1, a.tail1, a.W2, a.tail2, c1, c2);
  a1 = group_sum_fp<F, GROUP>(a1); a2 = group_sum_fp<F, GROUP>(a2);
  b1 = group_sum_fp<F, GROUP>(b1); b2 = group_sum_fp<F, GROUP>(b2);
  c1 = group_sum_fp<F, GROUP>(c1); c2 = group_sum_fp<F, GROUP>(c2);
  if (lane == 0 && valid) {
    Fp<F> t = cross_term_row<F>(a1, a2, b1, b2, c1, c2, Fp<F>::load(a.tail1));
    t.store(reinterpret_cast<char*>(a.T) + (size_t)row * 32);
    if (a.dc.counts) count_scalar_digits(t, a.dc.c, a.dc.nwin, a.dc.counts);
  }
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
template <class F>
__global__ void __launch_bounds__(256) k_axpy3(AxpySeg s0, AxpySeg s1, AxpySeg s2, Fp<F> r) {
  const size_t total = s0.len + s1.len + s2.len;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const bool in0 = i < s0.len, in1 = i < s0.len + s1.len;
    const AxpySeg& sg = in0 ? s0 : (in1 ? s1 : s2);
    const size_t j = in0 ? i : (in1 ? i - s0.len : i - s0.len - s1.len);
    char* pa = reinterpret_cast<char*>(sg.a) + j * 32;
    Fp<F> x = Fp<F>::load(pa);
    Fp<F> y = Fp<F>::load(reinterpret_cast<const char*>(sg.b) + j * 32);
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
