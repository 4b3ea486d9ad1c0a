// mul_latency.cu -- dependent-chain latency of one field multiplication for a warp that runs (almost) alone, the
// regime of the bucket-reduction / combine / scalar-multiplication kernels: cycles per product of fp_mul (throughput
// form, PTX carry chains) and fp_mul_lat (64-bit integer form, fp_lat.cuh) at 1, 2, 4 warps per scheduler, plus a
// bit-exactness check of the two on the same chain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Ivimz_b200/csrc -o tools/mul_latency tools/mul_latency.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "fp_lat.cuh"
using namespace vimz;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
#define MUL_ITERS 2048

template <class F, int V>
__device__ __noinline__ Fp<F> mul_call(Fp<F> a, Fp<F> b) {
  if (V == 0) return fp_mul(a, b);
  return fp_mul_lat(a, b);
}

template <class F, int V>
__global__ void k_chain(uint32_t* out, const uint32_t* in, long long* cyc) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  Fp<F> x = Fp<F>::load(in + 8 * (tid & 1023)), y = Fp<F>::load(in + 8 * ((tid + 7) & 1023));
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < MUL_ITERS; i++) {
    x = mul_call<F, V>(x, y);
    y = mul_call<F, V>(y, x);
  }
  long long t1 = clock64();
  fp_add(x, y).store(out + 8 * tid);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F>
int run(const char* fname, int sms, uint32_t* out, uint32_t* out2, uint32_t* in, long long* cyc, bool last) {
  long long* h = new long long[sms * 8];
  uint32_t* h0 = new uint32_t[(size_t)sms * 8 * 128 * 8];
  uint32_t* h1 = new uint32_t[(size_t)sms * 8 * 128 * 8];
  for (int w = 1; w <= 8; w *= 2) {
    int blocks = sms * w;
    double res[2];
    for (int v = 0; v < 2; v++) {
      uint32_t* o = v ? out2 : out;
      for (int r = 0; r < 2; r++) {
        if (v == 0) k_chain<F, 0><<<blocks, 128>>>(o, in, cyc);
        else k_chain<F, 1><<<blocks, 128>>>(o, in, cyc);
      }
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
      double avg = 0;
      for (int i = 0; i < blocks; i++) avg += h[i];
      res[v] = avg / blocks / (2.0 * MUL_ITERS);
    }
    size_t bytes = (size_t)blocks * 128 * 32;
    CK(cudaMemcpy(h0, out, bytes, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h1, out2, bytes, cudaMemcpyDeviceToHost));
    bool same = memcmp(h0, h1, bytes) == 0;
    printf("  {\"field\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_mul_fp_mul\": %.1f, \"cycles_per_mul_fp_mul_lat\": %.1f, \"identical\": %s}%s\n", fname,
           w, res[0], res[1], same ? "true" : "false", (last && w == 8) ? "" : ",");
  }
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  uint32_t *out, *out2, *in;
  long long* cyc;
  size_t n = (size_t)sms * 8 * 128;
  CK(cudaMalloc(&out, n * 32));
  CK(cudaMalloc(&out2, n * 32));
  CK(cudaMalloc(&in, 1024 * 32));
  uint32_t* hin = new uint32_t[1024 * 8];
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 1024 * 8; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    hin[i] = (uint32_t)(s >> 16);
    if ((i & 7) == 7) hin[i] &= 0x1fffffffu;  // < 2^253 < p for all four fields
  }
  CK(cudaMemcpy(in, hin, 1024 * 32, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&cyc, sms * 8 * sizeof(long long)));
  printf("{\"device\": \"%s\", \"tests\": [\n", prop.name);
  if (run<FieldPallasP>("pallas_base", sms, out, out2, in, cyc, false)) return 1;
  if (run<FieldVestaP>("vesta_base", sms, out, out2, in, cyc, false)) return 1;
  if (run<FieldBnP>("bn254_base", sms, out, out2, in, cyc, false)) return 1;
  if (run<FieldBnR>("bn254_scalar", sms, out, out2, in, cyc, true)) return 1;
  printf("]}\n");
  return 0;
}
