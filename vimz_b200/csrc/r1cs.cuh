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
    acc = fp_add(acc, fp_mul(v, z));
  }
  acc.store(reinterpret_cast<char*>(out) + (size_t)row * 32);
}

template <class F>
VIMZ_DI void row_dot2(const CsrView& M, uint32_t row, uint32_t n, const void* __restrict__ W1, const void* __restrict__ t1,
                      const void* __restrict__ W2, const void* __restrict__ t2, Fp<F>& d1, Fp<F>& d2) {
  d1 = Fp<F>::zero();
  d2 = Fp<F>::zero();
  uint32_t beg = M.rowptr[row], end = M.rowptr[row + 1];
  for (uint32_t k = beg; k < end; k++) {
    Fp<F> v = Fp<F>::load_nc(reinterpret_cast<const char*>(M.val) + (size_t)k * 32);
    uint32_t c = __ldg(M.col + k);
    d1 = fp_add(d1, fp_mul(v, load_z<F>(W1, t1, n, c)));
    d2 = fp_add(d2, fp_mul(v, load_z<F>(W2, t2, n, c)));
  }
}

// T[i] = Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1   (u2 = 1; u1 = tail1[0])
template <class F>
__global__ void __launch_bounds__(256) k_cross_term(CsrView A, CsrView B, CsrView Cm, uint32_t m, uint32_t n,
                                                    const void* __restrict__ W1, const void* __restrict__ tail1,
                                                    const void* __restrict__ W2, const void* __restrict__ tail2,
                                                    void* __restrict__ T) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  Fp<F> a1, a2, b1, b2, c1, c2;
  row_dot2<F>(A, row, n, W1, tail1, W2, tail2, a1, a2);
  row_dot2<F>(B, row, n, W1, tail1, W2, tail2, b1, b2);
  row_dot2<F>(Cm, row, n, W1, tail1, W2, tail2, c1, c2);
  Fp<F> u1 = Fp<F>::load(tail1);
  Fp<F> t = fp_add(fp_mul(a1, b2), fp_mul(a2, b1));
  t = fp_sub(t, fp_mul(u1, c2));
  t = fp_sub(t, c1);
  t.store(reinterpret_cast<char*>(T) + (size_t)row * 32);
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
