// msm.cuh -- Pippenger MSM over a resident window table, sm_100a.
//
// Replaces CommitmentEngine::commit -> Group::vartime_multiscalar_mul ([EXT nova-snark 0.23.0]
// src/provider/pedersen.rs, src/provider/mod.rs `cpu_best_multiexp`; pasta-msm for Pallas/Vesta),
// i.e. SURVEY.md rows a9/a10/a13.  Same group element, different schedule:
//
//   * The commitment key is fixed for the whole proof, so at upload we expand it ONCE into
//     table[j][i] = 2^(c*j) * ck_i (affine, 64 B).  A scalar s_i = sum_j d_ij 2^(c*j) with signed
//     digits then contributes d_ij * table[j][i]: every window shares ONE set of M = 2^(c-1)
//     buckets, there is no per-window bucket reduction and no final doubling chain.
//   * digits (k_count) -> counting sort by bucket (scan + k_scatter) -> bucket sums
//     (k_accumulate: one thread per bucket, buckets ordered by size so a warp's lanes run the same
//     trip count; oversized buckets are split into block tasks) -> sum_k (k+1) * B_k
//     (k_reduce_chunks / k_reduce_bits / k_reduce_scale / k_reduce_out).
//   * All arithmetic is exact; the result is the unique group element sum_i s_i * ck_i.
#pragma once
#include "common.cuh"
#include "ec.cuh"
#include "ecq.cuh"

namespace vimz {

constexpr int MSM_BIG_CHUNK = 256;    // entries per warp task for oversized buckets
constexpr int MSM_MAX_CLASSES = 4096; // bucket-size classes for the size ordering
constexpr int MSM_MAX_WINDOWS = 32;  // c >= 8 for 255-bit scalars

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

template <class C>
__global__ void k_msm_count(const uint32_t* __restrict__ scalars, uint32_t n, int c, int nwin,
                            uint32_t* __restrict__ counts) {
  using Fs = Fp<typename C::Fs>;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Fs s = fp_from_mont(Fs::load_nc(scalars + 8 * (size_t)i));
    if (fp_gt_half(s)) s = fp_neg(s);  // s*P = (q-s)*(-P): digits of the smaller magnitude
    for_each_digit(s.v, c, nwin, [&](int, uint32_t mag, bool) { atomicAdd(&counts[mag - 1], 1u); });
  }
}

template <class C>
__global__ void k_msm_scatter(const uint32_t* __restrict__ scalars, uint32_t n, int c, int nwin,
                              uint32_t table_stride, uint32_t first,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  using Fs = Fp<typename C::Fs>;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Fs s = fp_from_mont(Fs::load_nc(scalars + 8 * (size_t)i));
    // Nova's running vectors are dominated by values of small magnitude modulo q (sums of 128-bit
    // challenges and their negatives): recode q - s with the point negated, so a "negative small" scalar
    // costs as few bucket insertions as a positive one instead of all windows.
    bool flip = fp_gt_half(s);
    if (flip) s = fp_neg(s);
    for_each_digit(s.v, c, nwin, [&](int j, uint32_t mag, bool neg) {
      uint32_t pos = atomicAdd(&cursor[mag - 1], 1u);
      sorted[pos] = ((uint32_t)j * table_stride + first + i) | ((neg != flip) ? 0x80000000u : 0u);
    });
  }
}

// ---- exclusive scan over the bucket counts (3 small kernels) --------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread -> 2048 per block

static __global__ void k_scan_blocksum(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ blocksums) {
  __shared__ uint32_t sh[SCAN_THREADS / 32];
  uint32_t base = blockIdx.x * SCAN_THREADS * SCAN_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    uint32_t idx = base + k * SCAN_THREADS + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) t += sh[w];
    blocksums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of blocksums in place (nblocks <= 65536)
static __global__ void k_scan_top(uint32_t* __restrict__ blocksums, uint32_t nblocks, uint32_t* __restrict__ total) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    uint32_t idx = base + threadIdx.x;
    uint32_t v = idx < nblocks ? blocksums[idx] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = sh[threadIdx.x];
    if (idx < nblocks) blocksums[idx] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// per block: exclusive scan + block prefix; writes offsets[] and a copy into cursor[]
static __global__ void k_scan_apply(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ blocksums,
                             uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t sh[SCAN_THREADS];
  uint32_t base = blockIdx.x * SCAN_THREADS * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;  // blocked arrangement
  uint32_t v[SCAN_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < SCAN_THREADS; o <<= 1) {
    uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  uint32_t run = blocksums[blockIdx.x] + sh[threadIdx.x] - s;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) {
      offsets[base + k] = run;
      cursor[base + k] = run;
    }
    run += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) offsets[n] = run;
}

// ---- bucket schedule: order buckets by size (largest first), split oversized ones ------------
// cls layout (uint32): [0, NC) histogram, [NC, 2NC) class start, [2NC, 3NC) class cursor,
// then ctrl: [3NC+0] nbig, [3NC+1] ntasks, [3NC+2] task counter.
// The split threshold is decided on the device from the ACTUAL number of insertions E = offsets[M]
// (zeros and short scalars make it much smaller than n * windows): buckets above
// max(24, 2E/M + 8, 4e-5 E) entries are cut into warp tasks so no thread walks a long chain alone.
struct MsmSchedule {
  const uint32_t* total;  // &offsets[M]
  uint32_t M;
  uint32_t* hist;
  uint32_t* cstart;
  uint32_t* ccursor;
  uint32_t* ctrl;
  uint32_t* biglist;    // [maxbig] bucket ids
  uint32_t* taskstart;  // [maxbig + 1]
  uint32_t cap;         // buckets with count > cap are "big"
  uint32_t maxbig;
};

__device__ __forceinline__ uint32_t sched_cap(const MsmSchedule& sc) {
  uint32_t e = *sc.total;
  // a chain of k insertions costs ~3.5 us * k; the whole kernel needs ~e * 1.4e-4 us at full throughput,
  // so chains up to e * 4e-5 are free, and never split below twice the mean bucket
  uint32_t free_chain = (uint32_t)((uint64_t)e * 41u >> 20);
  return min(sc.cap, max(max(24u, free_chain), 2u * (e / sc.M) + 8u));
}

static __global__ void k_sched_hist(const uint32_t* __restrict__ counts, uint32_t M, MsmSchedule sc) {
  extern __shared__ uint32_t sh_hist[];  // cap + 1
  for (uint32_t k = threadIdx.x; k <= sc.cap; k += blockDim.x) sh_hist[k] = 0;
  __syncthreads();
  const uint32_t cap = sched_cap(sc);
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < M; b += gridDim.x * blockDim.x) {
    uint32_t cnt = counts[b];
    if (cnt > cap) {
      uint32_t e = atomicAdd(&sc.ctrl[0], 1u);
      if (e < sc.maxbig) sc.biglist[e] = b;
    } else {
      atomicAdd(&sh_hist[cnt], 1u);
    }
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k <= sc.cap; k += blockDim.x)
    if (sh_hist[k]) atomicAdd(&sc.hist[k], sh_hist[k]);
}

// single block: class starts (descending size) and big-bucket task starts
static __global__ void k_sched_scan(const uint32_t* __restrict__ counts, MsmSchedule sc) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  // classes in descending order: position p <-> class cap - p
  uint32_t nc = sc.cap + 1;
  for (uint32_t base = 0; base < nc; base += 1024) {
    uint32_t p = base + threadIdx.x;
    uint32_t v = p < nc ? sc.hist[sc.cap - p] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = sh[threadIdx.x];
    if (p < nc) {
      sc.cstart[sc.cap - p] = carry + incl - v;
      sc.ccursor[sc.cap - p] = carry + incl - v;
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  // big buckets: tasks of MSM_BIG_CHUNK entries
  uint32_t nbig = min(sc.ctrl[0], sc.maxbig);
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nbig; base += 1024) {
    uint32_t e = base + threadIdx.x;
    uint32_t v = e < nbig ? (counts[sc.biglist[e]] + MSM_BIG_CHUNK - 1) / MSM_BIG_CHUNK : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = sh[threadIdx.x];
    if (e < nbig) sc.taskstart[e] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sc.taskstart[nbig] = carry;
    sc.ctrl[1] = carry;
  }
}

static __global__ void k_sched_scatter(const uint32_t* __restrict__ counts, uint32_t M, MsmSchedule sc, uint32_t* __restrict__ order) {
  const uint32_t cap = sched_cap(sc);
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < M; b += gridDim.x * blockDim.x) {
    uint32_t cnt = counts[b];
    if (cnt <= cap) {
      uint32_t pos = atomicAdd(&sc.ccursor[cnt], 1u);
      order[pos] = b;
    }
  }
}

// ---- bucket accumulation ----------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128, 4) k_msm_accumulate(const uint32_t* __restrict__ order, const uint32_t* __restrict__ counts,
                                                        const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ sorted,
                                                        const void* __restrict__ table, uint32_t M, const uint32_t* __restrict__ ctrl,
                                                        void* __restrict__ buckets) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t nsmall = M - min(ctrl[0], M);
  if (j >= nsmall) return;
  uint32_t b = order[j];
  uint32_t cnt = counts[b];
  const uint32_t* ent = sorted + offsets[b];
  Xyzz<C> acc = Xyzz<C>::identity();
  if (cnt > 0) {
    uint32_t e = ent[0];
    Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(table) + (size_t)(e & 0x7fffffffu) * 64);
    for (uint32_t k = 0; k < cnt; k++) {
      Affine<C> cur = p;
      bool neg = (e >> 31) != 0;
      if (k + 1 < cnt) {  // prefetch the next base while this madd runs
        e = ent[k + 1];
        p = Affine<C>::load_nc(reinterpret_cast<const char*>(table) + (size_t)(e & 0x7fffffffu) * 64);
      }
      xyzz_madd<C>(acc, cur, neg);
    }
  }
  acc.store(reinterpret_cast<char*>(buckets) + (size_t)b * 128);
}

// value of lane (lane + o); lanes whose partner is out of range get the identity (NOT their own value:
// acc + acc would send them down the doubling path and the whole warp would pay for it)
template <class C>
__device__ __forceinline__ Xyzz<C> shfl_down_xyzz(const Xyzz<C>& acc, int o, int width = 32) {
  Xyzz<C> other;
  bool valid = (int)(threadIdx.x & 31) + o < width;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    other.x.v[k] = __shfl_down_sync(0xffffffffu, acc.x.v[k], o);
    other.y.v[k] = __shfl_down_sync(0xffffffffu, acc.y.v[k], o);
    uint32_t zz = __shfl_down_sync(0xffffffffu, acc.zz.v[k], o);
    other.zz.v[k] = valid ? zz : 0u;
    other.zzz.v[k] = __shfl_down_sync(0xffffffffu, acc.zzz.v[k], o);
  }
  return other;
}

// register-only warp tree (5 levels); result valid in lane 0
template <class C>
__device__ __forceinline__ Xyzz<C> warp_reduce_xyzz(Xyzz<C> acc) {
#pragma unroll 1  // one copy of the addition: these warps run alone and are instruction-fetch bound
  for (int o = 16; o > 0; o >>= 1) {
    Xyzz<C> other = shfl_down_xyzz<C>(acc, o);
    xyzz_add_call<C>(acc, other);
  }
  return acc;
}

// block of 128 threads: warp trees, then the 4 warp leaders are combined by warp 0; result valid in thread 0
template <class C>
__device__ __forceinline__ void block_reduce_xyzz_128(Xyzz<C>& acc, uint32_t* smem /* 4*32 words */) {
  acc = warp_reduce_xyzz<C>(acc);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) acc.store(smem + warp * 32);
  __syncthreads();
  if (warp == 0) {
    acc = lane < 4 ? Xyzz<C>::load(smem + lane * 32) : Xyzz<C>::identity();
#pragma unroll 1
    for (int o = 2; o > 0; o >>= 1) {
      Xyzz<C> other = shfl_down_xyzz<C>(acc, o, 2 * o);
      xyzz_add_call<C>(acc, other);
    }
  }
  __syncthreads();
}

// persistent warps pull (big bucket, chunk) tasks of MSM_BIG_CHUNK entries; each writes one XYZZ partial
template <class C>
__global__ void __launch_bounds__(128) k_msm_accumulate_big(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                            const uint32_t* __restrict__ sorted, const void* __restrict__ table,
                                                            MsmSchedule sc, void* __restrict__ partials, void* __restrict__ buckets) {
  uint32_t nbig = min(sc.ctrl[0], sc.maxbig);
  uint32_t ntasks = sc.ctrl[1];
  uint32_t lane = threadIdx.x & 31;
  for (;;) {
    uint32_t task = 0;
    if (lane == 0) task = atomicAdd(&sc.ctrl[2], 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= ntasks) break;
    // find e with taskstart[e] <= task < taskstart[e+1]
    uint32_t lo = 0, hi = nbig;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (sc.taskstart[mid] <= task) lo = mid; else hi = mid;
    }
    uint32_t b = sc.biglist[lo];
    uint32_t chunk = task - sc.taskstart[lo];
    uint32_t cnt = counts[b];
    uint32_t beg = chunk * MSM_BIG_CHUNK, end = min(cnt, beg + MSM_BIG_CHUNK);
    const uint32_t* ent = sorted + offsets[b];
    Xyzz<C> acc = Xyzz<C>::identity();
    for (uint32_t k = beg + lane; k < end; k += 32) {
      uint32_t e = ent[k];
      Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(table) + (size_t)(e & 0x7fffffffu) * 64);
      xyzz_madd_call<C>(acc, p, (e >> 31) != 0);
    }
    // 32 per-lane sums -> one point: each quad first folds its own four lanes' points, then a quad tree
    QPoint<C> qa = QPoint<C>::identity();
#pragma unroll 1
    for (int j = 0; j < 4; j++) qa = q_add<C>(qa, q_from_lane<C>(acc, j));
    qa = q_warp_reduce<C>(qa);
    // a bucket that fits one task is finished here; otherwise k_msm_big_combine adds the partials
    bool single = sc.taskstart[lo + 1] - sc.taskstart[lo] == 1;
    if (lane < 4) qa.store(single ? reinterpret_cast<char*>(buckets) + (size_t)b * 128 : reinterpret_cast<char*>(partials) + (size_t)task * 128);
  }
}

// 32 quads of a 128-thread block each hold one point: warp trees, then warp 0 folds the four warp results.
// Result valid in quad 0 of warp 0 (threads 0..3).  smem: 4 XYZZ records.
template <class C>
__device__ __forceinline__ QPoint<C> q_block_reduce_128(QPoint<C> acc, uint32_t* smem) {
  acc = q_warp_reduce<C>(acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) acc.store(smem + warp * 32);
  __syncthreads();
  if (warp == 0) {  // warp-uniform: all 32 lanes run the cooperative tree; quads >= 4 contribute the identity
    acc = lane < 16 ? QPoint<C>::load(smem + (lane >> 2) * 32) : QPoint<C>::identity();
    acc = q_warp_reduce<C>(acc);
  }
  __syncthreads();
  return acc;
}

// one 128-thread block per big bucket: its task partials are summed by 32 cooperating quads
template <class C>
__global__ void __launch_bounds__(128) k_msm_big_combine(MsmSchedule sc, const void* __restrict__ partials, void* __restrict__ buckets) {
  __shared__ __align__(16) uint32_t smem[4 * 32];
  uint32_t nbig = min(sc.ctrl[0], sc.maxbig);
  const uint32_t quad = threadIdx.x >> 2;
  for (uint32_t e = blockIdx.x; e < nbig; e += gridDim.x) {
    uint32_t t0 = sc.taskstart[e], t1 = sc.taskstart[e + 1];
    if (t1 - t0 <= 1) continue;  // written directly by its only task (block-uniform branch)
    QPoint<C> acc = QPoint<C>::identity();
    uint32_t iters = (t1 - t0 + 31) / 32;
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {  // warp-uniform trip count: every lane joins the shuffles
      uint32_t t = t0 + it * 32 + quad;
      QPoint<C> p = t < t1 ? QPoint<C>::load(reinterpret_cast<const char*>(partials) + (size_t)t * 128) : QPoint<C>::identity();
      acc = q_add<C>(acc, p);
    }
    acc = q_block_reduce_128<C>(acc, smem);
    if (threadIdx.x < 4) acc.store(reinterpret_cast<char*>(buckets) + (size_t)sc.biglist[e] * 128);
  }
}

// ---- bucket reduction: sum_{k=0}^{M-1} (k+1) * B_k ------------------------------------------
// Every step below is a chain of EC additions executed by warps that run alone, so it is organised for
// DEPTH and every addition is done by a quad of lanes (ecq.cuh):
//   level 1 (k_reduce_chunks): quad t owns K consecutive buckets: A_t = sum B, L_t = sum (j+1) B_{tK+j}   [2K adds deep]
//   level 2 (k_reduce_bits):   sum = S_L + K * sum_t t*A_t = S_L + sum_b 2^(b+logK) S_b with the plain sums
//                              S_b = sum_{t: bit b set} A_t, S_L = sum L_t                                  [trees]
//   level 3 (k_reduce_scale):  one warp per sum folds its G block partials, then applies the 2^(b+logK)
//                              doublings -- all sums in parallel instead of a serial Horner chain
//   level 4 (k_reduce_out):    one warp adds the <= 32 scaled sums and writes the Jacobian result.
template <class C>
__global__ void __launch_bounds__(128) k_reduce_chunks(const void* __restrict__ buckets, uint32_t T, int K,
                                                       void* __restrict__ chunkA, void* __restrict__ chunkL) {
  uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool valid = t < T;  // whole quads are valid or not; invalid quads still run the shuffles
  QPoint<C> running = QPoint<C>::identity(), acc = QPoint<C>::identity();
  const char* base = reinterpret_cast<const char*>(buckets) + (size_t)(valid ? t : 0) * K * 128;
#pragma unroll 1
  for (int j = K - 1; j >= 0; j--) {
    QPoint<C> b = valid ? QPoint<C>::load(base + (size_t)j * 128) : QPoint<C>::identity();
    running = q_add<C>(running, b);
    acc = q_add<C>(acc, running);
  }
  if (valid) {
    running.store(reinterpret_cast<char*>(chunkA) + (size_t)t * 128);
    acc.store(reinterpret_cast<char*>(chunkL) + (size_t)t * 128);
  }
}

// sum id s = blockIdx.y: s < nb -> sum of A_t over t with bit s set; s == nb -> sum of L_t.
template <class C>
__global__ void __launch_bounds__(128) k_reduce_bits(const void* __restrict__ chunkA, const void* __restrict__ chunkL, uint32_t T, int nb,
                                                     void* __restrict__ bitsums) {
  __shared__ __align__(16) uint32_t smem[4 * 32];
  const int s = blockIdx.y;
  // s < nb: enumerate exactly the indices with bit s set so every quad is busy; s == nb: all of chunkL
  const bool plain = (s == nb);
  const uint32_t count = plain ? T : (T >> 1), lowmask = plain ? 0u : ((1u << s) - 1);
  const char* src = reinterpret_cast<const char*>(plain ? chunkL : chunkA);
  const uint32_t quad = threadIdx.x >> 2, stride = gridDim.x * 32;
  const uint32_t iters = (count + stride - 1) / stride;
  QPoint<C> acc = QPoint<C>::identity();
#pragma unroll 1
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t u = it * stride + blockIdx.x * 32 + quad;
    uint32_t t = plain ? u : (((u & ~lowmask) << 1) | (1u << s) | (u & lowmask));
    QPoint<C> p = u < count ? QPoint<C>::load(src + (size_t)t * 128) : QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_block_reduce_128<C>(acc, smem);
  if (threadIdx.x < 4) acc.store(reinterpret_cast<char*>(bitsums) + ((size_t)s * gridDim.x + blockIdx.x) * 128);
}

// one warp (block) per sum s: fold the G block partials, then scale by 2^(s+logK) (s < nb) -> scaled[s]
template <class C>
__global__ void __launch_bounds__(32) k_reduce_scale(const void* __restrict__ bitsums, int nb, int G, int logK, void* __restrict__ scaled) {
  const int s = blockIdx.x, quad = threadIdx.x >> 2;
  QPoint<C> acc = QPoint<C>::identity();
  const int iters = (G + 7) / 8;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    int g = it * 8 + quad;
    QPoint<C> p = g < G ? QPoint<C>::load(reinterpret_cast<const char*>(bitsums) + ((size_t)s * G + g) * 128) : QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_warp_reduce<C>(acc);
  const int ndbl = s < nb ? s + logK : 0;  // block-uniform
#pragma unroll 1
  for (int k = 0; k < ndbl; k++) acc = q_dbl<C>(acc);
  if (threadIdx.x < 4) acc.store(reinterpret_cast<char*>(scaled) + (size_t)s * 128);
}

// one warp: add the nsums (<= 32) scaled sums, convert to Jacobian
template <class C>
__global__ void __launch_bounds__(32) k_reduce_out(const void* __restrict__ scaled, int nsums, void* __restrict__ out_jac) {
  const int quad = threadIdx.x >> 2;
  QPoint<C> acc = QPoint<C>::identity();
  const int iters = (nsums + 7) / 8;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    int g = it * 8 + quad;
    QPoint<C> p = g < nsums ? QPoint<C>::load(reinterpret_cast<const char*>(scaled) + (size_t)g * 128) : QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_warp_reduce<C>(acc);
  q_store_jacobian<C>(acc, out_jac, threadIdx.x < 4);
}

// ---- window-table expansion (once per commitment key) ---------------------------------------
// thread i: table[j][i] = 2^(c*j) * base_i for j < nwin, normalised to affine with one inversion.
template <class C>
__global__ void __launch_bounds__(128) k_precompute(const void* __restrict__ bases, uint32_t n, int c, int nwin, void* __restrict__ table) {
  using F = Fp<typename C::Fb>;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p = Affine<C>::load(reinterpret_cast<const char*>(bases) + (size_t)i * 64);
  p.store(reinterpret_cast<char*>(table) + (size_t)i * 64);
  if (nwin <= 1) return;
  Affine<C> zero;
  zero.x = F::zero(); zero.y = F::zero();
  if (p.is_identity()) {
    for (int j = 1; j < nwin; j++) zero.store(reinterpret_cast<char*>(table) + ((size_t)j * n + i) * 64);
    return;
  }
  Xyzz<C> pts[MSM_MAX_WINDOWS];
  F prefix[MSM_MAX_WINDOWS];
  Xyzz<C> cur = Xyzz<C>::from_affine(p);
  F run = F::one();
  for (int j = 1; j < nwin; j++) {
    for (int k = 0; k < c; k++) cur = xyzz_dbl<C>(cur);
    pts[j] = cur;          // prime-order group: never the identity
    prefix[j] = run;       // product of zzz_1 .. zzz_{j-1}
    run = fp_mul(run, cur.zzz);
  }
  F inv = fp_inv(run);
  for (int j = nwin - 1; j >= 1; j--) {
    F zi = fp_mul(inv, prefix[j]);      // 1 / zzz_j
    inv = fp_mul(inv, pts[j].zzz);
    F zzi = fp_sqr(fp_mul(pts[j].zz, zi));  // 1 / zz_j
    Affine<C> a;
    a.x = fp_mul(pts[j].x, zzi);
    a.y = fp_mul(pts[j].y, zi);
    a.store(reinterpret_cast<char*>(table) + ((size_t)j * n + i) * 64);
  }
}

// ---- small single-thread group kernels ---------------------------------------------------------
template <class C>
__global__ void k_point_sum(const void* __restrict__ pts, uint32_t k, void* __restrict__ out) {
  using F = Fp<typename C::Fb>;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Xyzz<C> acc = Xyzz<C>::identity();
  for (uint32_t i = 0; i < k; i++) {
    const char* p = reinterpret_cast<const char*>(pts) + (size_t)i * 96;
    Xyzz<C> q = xyzz_from_jacobian<C, MulCall>(F::load(p), F::load(p + 32), F::load(p + 64));
    xyzz_add_call<C>(acc, q);
  }
  F X, Y, Z;
  xyzz_to_jacobian<C, MulCall>(acc, X, Y, Z);
  char* o = reinterpret_cast<char*>(out);
  X.store(o); Y.store(o + 32); Z.store(o + 64);
}

template <class C>
__global__ void k_point_to_affine(const void* __restrict__ pt, void* __restrict__ out) {
  using F = Fp<typename C::Fb>;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const char* p = reinterpret_cast<const char*>(pt);
  Xyzz<C> q = xyzz_from_jacobian<C, MulCall>(F::load(p), F::load(p + 32), F::load(p + 64));
  Affine<C> a = xyzz_to_affine<C>(q);
  a.store(out);
}

// out[t] = a[t] + r * b[t] for count <= 8 independent pairs, one QUAD each (one warp in total); r is a
// Montgomery scalar shared by all pairs (RelaxedR1CSInstance::fold uses the same r for comm_W and comm_E).
template <class C>
__global__ void __launch_bounds__(32) k_point_scale_add(const void* __restrict__ a, const void* __restrict__ r_mont, const void* __restrict__ b,
                                                        void* __restrict__ out, int count) {
  using Fs = Fp<typename C::Fs>;
  const int t = threadIdx.x >> 2;
  const bool valid = t < count;
  Fs r = fp_from_mont(Fs::load(r_mont));
  const int tt = valid ? t : 0;
  QPoint<C> base = q_load_jacobian<C>(reinterpret_cast<const char*>(b) + (size_t)tt * 96);
  QPoint<C> pa = q_load_jacobian<C>(reinterpret_cast<const char*>(a) + (size_t)tt * 96);
  QPoint<C> acc = QPoint<C>::identity();
  int top = 255;
  while (top >= 0 && !((r.v[top >> 5] >> (top & 31)) & 1)) top--;
#pragma unroll 1
  for (int bit = top; bit >= 0; bit--) {  // r is uniform across the warp, so the branch is too
    acc = q_dbl<C>(acc);
    if ((r.v[bit >> 5] >> (bit & 31)) & 1) acc = q_add<C>(acc, base);
  }
  acc = q_add<C>(acc, pa);
  // each quad writes its own result (q_store_jacobian stores from lanes 0..2 of the quad)
  q_store_jacobian<C>(acc, reinterpret_cast<char*>(out) + (size_t)tt * 96, valid);
}

// bases[i] = (k0 + i*dk) * G ; each thread walks a run of GEN_RUN consecutive multiples.
constexpr int GEN_RUN = 16;
template <class C>
__global__ void __launch_bounds__(128) k_gen_bases(uint64_t k0, uint64_t dk, uint32_t n, void* __restrict__ out) {
  using F = Fp<typename C::Fb>;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t i0 = t * GEN_RUN;
  if (i0 >= n) return;
  Affine<C> g;
#pragma unroll
  for (int k = 0; k < 8; k++) { g.x.v[k] = C::gx_mont(k); g.y.v[k] = C::gy_mont(k); }
  auto smul = [&](uint64_t s) {
    Xyzz<C> acc = Xyzz<C>::identity();
    for (int bit = 63; bit >= 0; bit--) {
      xyzz_dbl_call<C>(acc);
      if ((s >> bit) & 1) xyzz_madd_call<C>(acc, g, false);
    }
    return acc;
  };
  Xyzz<C> step = smul(dk);
  Xyzz<C> cur = smul(k0 + (uint64_t)i0 * dk);
  Xyzz<C> pts[GEN_RUN];
  F prefix[GEN_RUN];
  F run = F::one();
  int cnt = min((uint32_t)GEN_RUN, n - i0);
  for (int j = 0; j < cnt; j++) {
    pts[j] = cur;
    prefix[j] = run;
    if (!cur.is_identity()) run = fp_mul(run, cur.zzz);
    xyzz_add_call<C>(cur, step);
  }
  F inv = fp_inv(run);
  for (int j = cnt - 1; j >= 0; j--) {
    Affine<C> a;
    if (pts[j].is_identity()) {
      a.x = F::zero(); a.y = F::zero();
    } else {
      F zi = fp_mul(inv, prefix[j]);
      inv = fp_mul(inv, pts[j].zzz);
      F zzi = fp_sqr(fp_mul(pts[j].zz, zi));
      a.x = fp_mul(pts[j].x, zzi);
      a.y = fp_mul(pts[j].y, zi);
    }
    a.store(reinterpret_cast<char*>(out) + (size_t)(i0 + j) * 64);
  }
}

}  // namespace vimz
