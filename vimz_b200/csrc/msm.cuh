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
//     (k_accumulate: every thread walks an equal segment of the sorted insertions, k_combine adds the
//     pieces of buckets cut by segment boundaries) -> sum_k (k+1) * B_k
//     (k_reduce_chunks / k_reduce_bits / k_reduce_scale / k_reduce_out).
//   * All arithmetic is exact; the result is the unique group element sum_i s_i * ck_i.
#pragma once
#include "common.cuh"
#include "digits.cuh"
#include "ec.cuh"
#include "ecq.cuh"

namespace vimz {

#ifndef VIMZ_ACC_MINBLOCKS
#define VIMZ_ACC_MINBLOCKS 4  // resident blocks per SM the accumulation kernel is compiled for (128 registers at 4)
#endif
#ifndef VIMZ_ACC_MUL
#define VIMZ_ACC_MUL MulInline  // multiplication policy of the accumulation hot loop (A/B: -DVIMZ_ACC_MUL=MulCall)
#endif

constexpr int MSM_MAX_WINDOWS = 32;  // c >= 8 for 255-bit scalars

constexpr uint32_t SEG_MIN_DEFAULT = 8; // shortest segment (entries per thread); option "msm_seg_min"
#ifndef VIMZ_SCATTER_AGG
#define VIMZ_SCATTER_AGG 1
#endif
#ifndef VIMZ_COMBINE_SPAN
#define VIMZ_COMBINE_SPAN 8
#endif
#ifndef VIMZ_COMBINE_MID
#define VIMZ_COMBINE_MID 256
#endif
#ifndef VIMZ_GIANT_CHUNK
#define VIMZ_GIANT_CHUNK 256
#endif
constexpr uint32_t COMBINE_SPAN = VIMZ_COMBINE_SPAN;  // a quad adds at most this many partials serially
constexpr uint32_t COMBINE_MID = VIMZ_COMBINE_MID;    // up to this many partials: one warp (8 cooperating quads) per bucket
constexpr uint32_t GIANT_CHUNK = VIMZ_GIANT_CHUNK;    // pieces of a giant bucket summed by one block (32 quads x GIANT_CHUNK/32)
static_assert(GIANT_CHUNK % 32 == 0 && GIANT_CHUNK >= 32, "GIANT_CHUNK: whole rounds of the block's 32 quads");

__device__ __forceinline__ uint32_t seg_len(uint32_t E, uint32_t nthreads, uint32_t seg_min) {
  uint32_t L = (E + nthreads - 1) / nthreads;
  return L < seg_min ? seg_min : L;
}

// Control block of one MSM (zeroed by a single memset before the launch sequence).
constexpr uint32_t CTRL_NGIANT = 0, CTRL_NCHUNK = 1, CTRL_NMID = 2;
constexpr uint32_t CTRL_REDUCE = 8;      // [nb + 2] arrival counters of the reduction tail (nb + 1 sums, then the final one)
constexpr uint32_t CTRL_GIANT_DONE = 64; // [max_giants] chunks finished per giant bucket
struct MsmCombine {
  uint32_t* ctrl;       // see CTRL_*
  uint32_t* mids;       // [M] ids of buckets cut into COMBINE_SPAN+1 .. COMBINE_MID pieces
  uint32_t* giants;     // [max_giants][3]: bucket id, first chunk, number of chunks
  uint32_t* chunk_rec;  // [max_chunks][2]: giant index, chunk index inside the giant
  void* chunk_sums;     // [max_chunks] XYZZ
  uint32_t max_giants, max_chunks;
};

// File bucket b = [s, e) of the sorted list by the number of segments (length L) it is cut into: 2..COMBINE_SPAN
// pieces need no record (a quad adds them), more go to the mid list (a warp each) or the giant list (blocks of
// GIANT_CHUNK pieces).  Depends only on the scan, so it runs BEFORE the accumulation, off its critical path.
// Warp-collective (ballots): call with all 32 lanes; list appends are one atomic per warp.
__device__ __forceinline__ void classify_bucket(bool valid, uint32_t b, uint32_t s, uint32_t e, uint32_t L, const MsmCombine& cb) {
  uint32_t pieces = 0;
  if (valid && e > s) pieces = (e - 1) / L - s / L + 1;
  const bool mid = pieces > COMBINE_SPAN && pieces <= COMBINE_MID;
  const unsigned lane = threadIdx.x & 31;
  const unsigned mm = __ballot_sync(0xffffffffu, mid);
  if (mm) {
    const int leader = __ffs(mm) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&cb.ctrl[CTRL_NMID], (uint32_t)__popc(mm));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mid) cb.mids[base + __popc(mm & ((1u << lane) - 1u))] = b;
  }
  if (pieces > COMBINE_MID) {  // giant (rare: the 0/1 bucket of a witness vector, all-equal scalars)
    uint32_t nch = (pieces + GIANT_CHUNK - 1) / GIANT_CHUNK;
    uint32_t g = atomicAdd(&cb.ctrl[CTRL_NGIANT], 1u);
    uint32_t base = atomicAdd(&cb.ctrl[CTRL_NCHUNK], nch);
    if (g < cb.max_giants && base + nch <= cb.max_chunks) {
      cb.giants[3 * g] = b; cb.giants[3 * g + 1] = base; cb.giants[3 * g + 2] = nch;
      for (uint32_t j = 0; j < nch; j++) { cb.chunk_rec[2 * (base + j)] = g; cb.chunk_rec[2 * (base + j) + 1] = j; }
    }
  }
}

template <class C>
__global__ void k_msm_digits(const uint32_t* __restrict__ scalars, uint32_t n, int c, int nwin,
                             uint32_t* __restrict__ counts, uint32_t* __restrict__ digits) {
  using Fs = Fp<typename C::Fs>;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    recode_scalar<typename C::Fs, true>(Fs::load_nc(scalars + 8 * (size_t)i), c, nwin, counts, digits, n, i);
  }
}

// One thread per (window, scalar) entry of the digit array: a non-zero digit takes the next slot of its bucket.
// sorted[] holds table indices (window row * table_stride + point) with the sign in bit 31.
static __global__ void __launch_bounds__(256) k_msm_scatter(const uint32_t* __restrict__ digits, uint32_t n, int nwin,
                                                            uint32_t table_stride, uint32_t first,
                                                            uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted,
                                                            const uint32_t* __restrict__ offsets, uint32_t M, uint32_t nthreads,
                                                            uint32_t seg_min, MsmCombine cb) {
  {  // bucket classification for k_msm_combine_all rides along (it only needs the finished scan): warp-collective
    const uint32_t L = seg_len(offsets[M], nthreads, seg_min);
    const uint32_t lane = threadIdx.x & 31, nwarps = (gridDim.x * gridDim.y * blockDim.x) >> 5;
    const uint32_t block = blockIdx.y * gridDim.x + blockIdx.x;
    for (uint32_t base = ((block * blockDim.x + threadIdx.x) >> 5) * 32; base < M; base += nwarps * 32) {
      const uint32_t b = base + lane;
      const bool valid = b < M;
      classify_bucket(valid, b, valid ? offsets[b] : 0u, valid ? offsets[b + 1] : 0u, L, cb);
    }
  }
  const uint32_t j = blockIdx.y;  // window
  const uint32_t* row = digits + (size_t)j * n;
#if VIMZ_SCATTER_AGG
  // Warp-aggregated cursors.  Late in a proof the top window of T holds one or two bits, so ~10^5 entries of that window
  // carry the SAME digit (and a witness vector puts half its scalars into bucket 1): one returning atomic per entry on a
  // single address serialises in L2 (+0.15 ms per step, measured).  The lanes of a warp that target the same bucket are
  // matched, one of them claims the whole run, the others take consecutive slots.  Every warp runs the same number of
  // iterations, so the collectives see converged lanes.
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
    const uint32_t i = base + lane;
    const uint32_t d = i < n ? __ldg(row + i) : 0u;
    const uint32_t b = (d & 0x7fffffffu) - 1;
    // only small digits are shared by many lanes (see recode_scalar): they claim their slots per group of equal digits,
    // everything else takes its slot with a plain returning atomic (uniform digits: nothing to merge, and the match unit
    // was this kernel's top stall)
    const bool small = d != 0 && b < AGG_MAX_DIGIT;
    const unsigned smask = __ballot_sync(0xffffffffu, small);
    if (small) {
      const unsigned peers = __match_any_sync(smask, b);
      const int leader = __ffs(peers) - 1;
      uint32_t slot = 0;
      if ((int)lane == leader) slot = atomicAdd(&cursor[b], (uint32_t)__popc(peers));
      slot = __shfl_sync(peers, slot, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      sorted[slot] = (j * table_stride + first + i) | (d & 0x80000000u);
    } else if (d != 0) {
      const uint32_t slot = atomicAdd(&cursor[b], 1u);
      sorted[slot] = (j * table_stride + first + i) | (d & 0x80000000u);
    }
  }
#else
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t d = __ldg(row + i);
    if (d == 0) continue;
    const uint32_t pos = atomicAdd(&cursor[(d & 0x7fffffffu) - 1], 1u);
    sorted[pos] = (j * table_stride + first + i) | (d & 0x80000000u);
  }
#endif
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

// ---- bucket accumulation: equal segments of the sorted entry list -----------------------------
// After the counting sort the E insertions are grouped by bucket.  Thread t takes entries
// [t*L, (t+1)*L) -- the same number for every thread, whatever the bucket sizes (uniform digits, the
// 55k-entry 0/1 bucket of a witness vector, the short scalars of T) -- and walks them with one XYZZ
// accumulator, flushing whenever the bucket changes:
//   * a run that covers its bucket completely is stored straight into buckets[b];
//   * the first / last run of a segment may be cut by the segment boundary: stored as partial 0 / 1;
// k_msm_combine then adds the <= few partials of every cut bucket (one thread per bucket; buckets cut
// into more than COMBINE_SPAN pieces go to k_msm_combine_big, one block of cooperating quads each).
// Small bucket counts (M <= SCAN1_MAX_M, every fold-step MSM): ONE block scans the counts and writes the offsets and
// the scatter cursors -- one launch instead of three on a latency-bound chain.
#ifndef VIMZ_SCAN1_THREADS
#define VIMZ_SCAN1_THREADS 1024
#endif
constexpr uint32_t SCAN1_THREADS = VIMZ_SCAN1_THREADS;  // a multiple of 32, at most 1024
constexpr uint32_t SCAN1_MAX_M = 32768;
static __global__ void __launch_bounds__(SCAN1_THREADS) k_scan_single(const uint32_t* __restrict__ counts, uint32_t M,
                                                                      uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t ipt = ((M + SCAN1_THREADS - 1) / SCAN1_THREADS + 3) & ~3u;  // consecutive counts per thread, multiple of 4
  const uint32_t base = threadIdx.x * ipt;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t sum = 0;
  if (base + ipt <= M) {
    for (uint32_t k = 0; k < ipt; k += 4) {
      uint4 v = *reinterpret_cast<const uint4*>(counts + base + k);
      sum += v.x + v.y + v.z + v.w;
    }
  } else {
    for (uint32_t k = 0; k < ipt; k++) sum += (base + k < M) ? counts[base + k] : 0u;
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN1_THREADS / 32 ? warp_tot[lane] : 0u, wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if ((int)lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;  // exclusive prefix of the warp totals
    if (lane == 31) offsets[M] = wi;
  }
  __syncthreads();
  uint32_t run = warp_tot[warp] + incl - sum;  // exclusive prefix of this thread's first count
  if (base + ipt <= M) {
    for (uint32_t k = 0; k < ipt; k += 4) {
      uint4 v = *reinterpret_cast<const uint4*>(counts + base + k);
      uint4 o;
      o.x = run; o.y = o.x + v.x; o.z = o.y + v.y; o.w = o.z + v.z;
      run = o.w + v.w;
      *reinterpret_cast<uint4*>(offsets + base + k) = o;
      *reinterpret_cast<uint4*>(cursor + base + k) = o;
    }
  } else {
    for (uint32_t k = 0; k < ipt; k++) {
      if (base + k < M) {
        offsets[base + k] = run;
        cursor[base + k] = run;
        run += counts[base + k];
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(128, VIMZ_ACC_MINBLOCKS) k_msm_accumulate(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ sorted,
                                                           const void* __restrict__ table, uint32_t M, uint32_t nthreads, uint32_t seg_min,
                                                           void* __restrict__ buckets, void* __restrict__ partials) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  const uint32_t E = offsets[M];
  const uint32_t L = seg_len(E, nthreads, seg_min);
  const uint64_t beg64 = (uint64_t)t * L;
  if (beg64 >= E) return;
  const uint32_t beg = (uint32_t)beg64, end = min(E, beg + L);
  // bucket of the first entry: the largest b with offsets[b] <= beg (then offsets[b+1] > beg)
  uint32_t lo = 0, hi = M;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= beg) lo = mid; else hi = mid;
  }
  uint32_t b = lo, bstart = offsets[b], bend = offsets[b + 1];
  uint32_t bnext = offsets[min(b + 2, M)];  // one boundary ahead, so a flush never waits on this load
  uint32_t run_start = beg;
  Xyzz<C> acc = Xyzz<C>::identity();
  uint32_t e = sorted[beg];
  Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(table) + (size_t)(e & 0x7fffffffu) * 64);
  for (uint32_t k = beg; k < end; k++) {
    if (k == bend) {  // bucket boundary: the finished run is complete unless it began at the segment start mid-bucket
      bool complete = run_start == bstart;
      char* dst = complete ? reinterpret_cast<char*>(buckets) + (size_t)b * 128 : reinterpret_cast<char*>(partials) + (size_t)(2 * t) * 128;
      acc.store(dst);
      acc = Xyzz<C>::identity();
      run_start = k;
      do { b++; bstart = bend; bend = bnext; bnext = offsets[min(b + 2, M)]; } while (bend == k);  // skip empty buckets
    }
    Affine<C> cur = p;
    bool neg = (e >> 31) != 0;
    if (k + 1 < end) {  // prefetch the next base while this madd runs
      e = sorted[k + 1];
      p = Affine<C>::load_nc(reinterpret_cast<const char*>(table) + (size_t)(e & 0x7fffffffu) * 64);
    }
    xyzz_madd<C, VIMZ_ACC_MUL>(acc, cur, neg);
  }
  // last run: complete only if it started at its bucket's start and the segment ends exactly at the bucket end
  bool complete = (run_start == bstart) && (end == bend);
  char* dst = complete ? reinterpret_cast<char*>(buckets) + (size_t)b * 128
                       : reinterpret_cast<char*>(partials) + (size_t)(2 * t + (run_start == beg ? 0 : 1)) * 128;
  acc.store(dst);
}

// which partial slot of segment t holds the piece of the bucket that starts at entry s
__device__ __forceinline__ size_t seg_partial_index(uint32_t t, uint32_t t0, uint32_t s, uint32_t L) {
  // in its first segment the bucket is the LAST run (slot 1) unless it starts exactly at the segment start
  return (size_t)2 * t + ((t == t0 && s != t0 * L) ? 1 : 0);
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
    acc = q_warp_reduce<C>(acc, 2);  // four points: two levels
  }
  __syncthreads();
  return acc;
}

// Last stage of the bucket reduction, shared by its two kinds of producers (k_reduce_tail): the nb + 1 scaled bit-plane
// sums and the deferred giant buckets each announce themselves on one counter; whoever arrives last adds them all and
// writes the Jacobian result.  Called by warp 0 of a block (all 32 lanes).
struct TailFinal {
  void* scaled;          // [nsums] XYZZ
  int nsums;
  void* deferred;        // [ngiant] XYZZ, already multiplied by the bucket weight
  const uint32_t* ngiant_p;
  uint32_t max_giants;
  uint32_t* cnt;         // arrival counter
  void* out_jac;
  void* out_host;        // optional second copy of the result in mapped page-locked host memory (nullptr: none)
  const void* sub_jac;   // optional Jacobian point SUBTRACTED from the sum before it is written (the accumulator's K_S, r1cs.cuh)
};
template <class C>
__device__ __forceinline__ void reduce_arrive_final(const TailFinal& f) {
  const uint32_t ngiant = f.deferred ? min(*f.ngiant_p, f.max_giants) : 0u;
  uint32_t done = 0;
  if ((threadIdx.x & 31) == 0) done = atomicAdd(f.cnt, 1u) + 1;
  done = __shfl_sync(0xffffffffu, done, 0);
  if (done != (uint32_t)f.nsums + ngiant) return;
  __threadfence();
  const uint32_t q8 = (threadIdx.x & 31) >> 2, total = (uint32_t)f.nsums + ngiant;
  QPoint<C> acc = QPoint<C>::identity();
#pragma unroll 1
  for (uint32_t it = 0; it < (total + 7) / 8; it++) {
    const uint32_t g = it * 8 + q8;
    const char* src = g < (uint32_t)f.nsums ? reinterpret_cast<const char*>(f.scaled) + (size_t)g * 128
                                            : reinterpret_cast<const char*>(f.deferred) + (size_t)(g - f.nsums) * 128;
    QPoint<C> p = g < total ? QPoint<C>::load_cg(src) : QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_warp_reduce<C>(acc);
  if (f.sub_jac) {  // warp-uniform
    QPoint<C> k = q_load_jacobian<C>(f.sub_jac);
    if ((threadIdx.x & 3) == 1) k.c = fp_neg(k.c);  // -(X, Y, ZZ, ZZZ) = (X, -Y, ZZ, ZZZ)
    acc = q_add<C>(acc, k);
  }
  q_store_jacobian<C>(acc, f.out_jac, (threadIdx.x & 31) < 4, f.out_host);
}

// Giant buckets (cut into more than COMBINE_MID pieces: the 0/1 bucket of a witness vector, the one- or two-bit top window of
// the cross term late in a proof).  Blocks `block`, `block + nblocks`, ... each sum one chunk of GIANT_CHUNK pieces (32 quads);
// the block that finishes the last chunk of a giant also folds its chunk sums.  Two uses:
//   * inside k_msm_combine_all (deferred == nullptr): the giant's sum goes to buckets[b] like every other bucket;
//   * inside k_reduce_tail (deferred != nullptr): the giant was left OUT of the bucket array (identity there), its sum is
//     scaled by its weight b + 1 here and handed to the final stage -- a ~19-addition chain that then runs beside the
//     bucket reduction instead of in front of it.
template <class C>
__device__ __forceinline__ void giant_role(const uint32_t* __restrict__ offsets, uint32_t L, uint32_t block, uint32_t nblocks,
                                           const void* __restrict__ partials, const MsmCombine& cb, uint32_t* smem, uint32_t* last_flag_p,
                                           void* __restrict__ buckets, const TailFinal* fin) {
  void* deferred = fin ? fin->deferred : nullptr;
  const char* parts = reinterpret_cast<const char*>(partials);
  const uint32_t nchunks = min(cb.ctrl[CTRL_NCHUNK], cb.max_chunks);
  const uint32_t quad = threadIdx.x >> 2;
  for (uint32_t j = block; j < nchunks; j += nblocks) {
    const uint32_t g = cb.chunk_rec[2 * j], idx = cb.chunk_rec[2 * j + 1];
    const uint32_t b = cb.giants[3 * g], gbase = cb.giants[3 * g + 1], nch = cb.giants[3 * g + 2];
    const uint32_t s = offsets[b], e = offsets[b + 1];
    const uint32_t t0 = s / L, t1 = (e - 1) / L;
    const uint32_t first = t0 + idx * GIANT_CHUNK, last = min(t1, first + GIANT_CHUNK - 1);
    QPoint<C> acc = QPoint<C>::identity();
#pragma unroll 1
    for (uint32_t it = 0; it < GIANT_CHUNK / 32; it++) {  // uniform trip count: every lane joins the shuffles
      uint32_t t = first + it * 32 + quad;
      QPoint<C> p = t <= last ? QPoint<C>::load(parts + seg_partial_index(t, t0, s, L) * 128) : QPoint<C>::identity();
      acc = q_add<C>(acc, p);
    }
    acc = q_block_reduce_128<C>(acc, smem);
    if (threadIdx.x < 4) {
      acc.store(reinterpret_cast<char*>(cb.chunk_sums) + (size_t)j * 128);
      __threadfence();  // publish the chunk sum before announcing it
    }
    __syncthreads();
    if (threadIdx.x == 0) *last_flag_p = (atomicAdd(&cb.ctrl[CTRL_GIANT_DONE + g], 1u) + 1 == nch) ? 1u : 0u;
    __syncthreads();
    if (*last_flag_p && threadIdx.x < 32) {  // last chunk of giant g: one warp folds all its chunk sums (read through L2)
      __threadfence();
      const uint32_t q8 = threadIdx.x >> 2;
      QPoint<C> tot = QPoint<C>::identity();
      const uint32_t iters = (nch + 7) / 8;
#pragma unroll 1
      for (uint32_t it = 0; it < iters; it++) {
        uint32_t jj = it * 8 + q8;
        QPoint<C> p = jj < nch ? QPoint<C>::load_cg(reinterpret_cast<const char*>(cb.chunk_sums) + (size_t)(gbase + jj) * 128)
                               : QPoint<C>::identity();
        tot = q_add<C>(tot, p);
      }
      tot = q_warp_reduce<C>(tot);
      if (!deferred) {
        if (threadIdx.x < 4) tot.store(reinterpret_cast<char*>(buckets) + (size_t)b * 128);
      } else {
        // weight of bucket b in sum_k (k+1) B_k: (b + 1) * tot by double-and-add (warp-uniform: b is)
        const uint32_t w = b + 1;
        QPoint<C> sc = QPoint<C>::identity();
#pragma unroll 1
        for (int bit = 31 - __clz(w); bit >= 0; bit--) {
          sc = q_dbl<C>(sc);
          if ((w >> bit) & 1) sc = q_add<C>(sc, tot);
        }
        if (threadIdx.x < 4) {
          sc.store(reinterpret_cast<char*>(deferred) + (size_t)g * 128);
          __threadfence();
        }
        __syncwarp();
        reduce_arrive_final<C>(*fin);  // the last arrival -- a bit-plane sum or a giant -- adds everything up
      }
    }
    __syncthreads();
  }
}

// All pieces of cut buckets are added in ONE launch after the accumulation; the block index selects the role
// (longest chains first so they start first):
//   blocks [0, nb_big)            giant buckets: a block (32 quads) per chunk of GIANT_CHUNK pieces; the block that
//                                 finishes the last chunk of a giant also folds the chunk sums into the bucket
//   blocks [nb_big, +nb_mid)      mid buckets: a warp each, its 8 quads stride over the pieces, then a quad tree
//   remaining blocks              a quad per bucket: 2..COMBINE_SPAN pieces summed serially; empty buckets get the
//                                 identity
// The lists were filled by classify_bucket before the accumulation started.
template <class C>
__global__ void __launch_bounds__(128) k_msm_combine_all(const uint32_t* __restrict__ offsets, uint32_t M, uint32_t nthreads, uint32_t seg_min,
                                                         uint32_t nb_big, uint32_t nb_mid, const void* __restrict__ partials,
                                                         void* __restrict__ buckets, MsmCombine cb, bool defer_giants) {
  __shared__ __align__(16) uint32_t smem[4 * 32];
  __shared__ uint32_t last_flag;
  const uint32_t E = offsets[M], L = seg_len(E, nthreads, seg_min);
  const char* parts = reinterpret_cast<const char*>(partials);
  if (blockIdx.x < nb_big) {
    giant_role<C>(offsets, L, blockIdx.x, nb_big, partials, cb, smem, &last_flag, buckets, nullptr);
    return;
  }
  if (blockIdx.x < nb_big + nb_mid) {
    const uint32_t nmid = cb.ctrl[CTRL_NMID];
    const uint32_t warp = ((blockIdx.x - nb_big) * blockDim.x + threadIdx.x) >> 5, nwarps = (nb_mid * blockDim.x) >> 5;
    const uint32_t quad = (threadIdx.x & 31) >> 2;
    for (uint32_t i = warp; i < nmid; i += nwarps) {
      const uint32_t b = cb.mids[i];
      const uint32_t s = offsets[b], e = offsets[b + 1];
      const uint32_t t0 = s / L, t1 = (e - 1) / L;
      QPoint<C> acc = QPoint<C>::identity();
      const uint32_t iters = (t1 - t0 + 1 + 7) / 8;
#pragma unroll 1
      for (uint32_t it = 0; it < iters; it++) {  // warp-uniform trip count
        uint32_t t = t0 + it * 8 + quad;
        QPoint<C> p = t <= t1 ? QPoint<C>::load(parts + seg_partial_index(t, t0, s, L) * 128) : QPoint<C>::identity();
        acc = q_add<C>(acc, p);
      }
      acc = q_warp_reduce<C>(acc);
      if ((threadIdx.x & 31) < 4) acc.store(reinterpret_cast<char*>(buckets) + (size_t)b * 128);
    }
    return;
  }
  const uint32_t b = ((blockIdx.x - nb_big - nb_mid) * blockDim.x + threadIdx.x) >> 2;
  const bool in_range = b < M;
  const uint32_t s = in_range ? offsets[b] : 0, e = in_range ? offsets[b + 1] : 0;
  const bool empty = in_range && e == s;
  uint32_t t0 = 0, pieces = 0;
  if (in_range && !empty) {
    t0 = s / L;
    pieces = (e - 1) / L - t0 + 1;  // 1: stored complete by k_msm_accumulate
  }
  const uint32_t mine = (pieces >= 2 && pieces <= COMBINE_SPAN) ? pieces : 0;  // pieces this quad adds itself
  const uint32_t trips = __reduce_max_sync(0xffffffffu, mine);               // warp-uniform trip count for the shuffles
  QPoint<C> acc = mine ? QPoint<C>::load(parts + seg_partial_index(t0, t0, s, L) * 128) : QPoint<C>::identity();
  QPoint<C> nxt = mine > 1 ? QPoint<C>::load(parts + seg_partial_index(t0 + 1, t0, s, L) * 128) : QPoint<C>::identity();
#pragma unroll 1
  for (uint32_t i = 1; i < trips; i++) {
    QPoint<C> q = nxt;  // the next piece is fetched while this addition runs
    nxt = i + 1 < mine ? QPoint<C>::load(parts + seg_partial_index(t0 + i + 1, t0, s, L) * 128) : QPoint<C>::identity();
    acc = q_add<C>(acc, q);
  }
  // a deferred giant stays out of the bucket array (identity here): k_reduce_tail adds its weighted sum at the very end
  if (empty || mine || (defer_giants && pieces > COMBINE_MID)) acc.store(reinterpret_cast<char*>(buckets) + (size_t)b * 128);
}

// ---- bucket reduction: sum_{k=0}^{M-1} (k+1) * B_k ------------------------------------------
// Every step below is a chain of EC additions executed by warps that run alone, so it is organised for
// DEPTH and every addition is done by a quad of lanes (ecq.cuh):
//   level 1 (k_reduce_chunks): quad t owns K consecutive buckets: A_t = sum B, L_t = sum (j+1) B_{tK+j}   [2K adds deep]
//   level 2 (k_reduce_tail):   sum = S_L + K * sum_t t*A_t = S_L + sum_b 2^(b+logK) S_b with the plain sums
//                              S_b = sum_{t: bit b set} A_t, S_L = sum L_t                                  [trees]
//   level 3 (same launch):     one warp per sum folds its G block partials, then applies the 2^(b+logK)
//                              doublings -- all sums in parallel instead of a serial Horner chain
//   level 4 (same launch):     one warp adds the <= 32 scaled sums and writes the Jacobian result.
template <class C>
__global__ void __launch_bounds__(128) k_reduce_chunks(const void* __restrict__ buckets, uint32_t T, int K,
                                                       void* __restrict__ chunkA, void* __restrict__ chunkL) {
  uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  const bool valid = t < T;  // whole quads are valid or not; invalid quads still run the shuffles
  const char* base = reinterpret_cast<const char*>(buckets) + (size_t)(valid ? t : 0) * K * 128;
  auto bucket = [&](int j) { return valid ? QPoint<C>::load(base + (size_t)j * 128) : QPoint<C>::identity(); };
  // running_j = running_{j+1} + B_j, acc_j = acc_{j+1} + running_j, from running_{K-1} = acc_{K-1} = B_{K-1} (no addition).  The
  // weighted sum lags one step behind the running sum, so each step's two additions are independent and share their product
  // rounds (q_add2): K - 2 paired steps + 2 single ones instead of 2 K sequential additions.
  QPoint<C> running = bucket(K - 1), acc = running;  // running_{K-1}, acc_{K-1}
  if (K >= 2) {
    running = q_add<C>(running, bucket(K - 2));      // running_{K-2}
#pragma unroll 1
    for (int j = K - 3; j >= 0; j--) {
      // running_j = running_{j+1} + B_j   ||   acc_{j+1} = acc_{j+2} + running_{j+1}
      QPair<C> r = q_add2<C>(running, bucket(j), acc, running);
      running = r.a;
      acc = r.b;
    }
    acc = q_add<C>(acc, running);                    // acc_0 = acc_1 + running_0
  }
  if (valid) {
    running.store(reinterpret_cast<char*>(chunkA) + (size_t)t * 128);
    acc.store(reinterpret_cast<char*>(chunkL) + (size_t)t * 128);
  }
}

// Levels 2-4 in ONE launch (the chain is latency-bound, so launch boundaries are pure loss): block (g, s) sums its
// share of sum s (s < nb: the A_t with bit s of t set; s == nb: all L_t); the block that completes sum s folds the
// G block partials and applies the 2^(s+logK) doublings; the block that completes the last sum adds the nb+1
// scaled sums and writes the Jacobian result.  Hand-offs: __threadfence + arrival counters (cnt[0..nb] per sum,
// cnt[nb+1] for the final), zeroed with the MSM's control block; later stages read through L2 (ld.cg).
template <class C>
__global__ void __launch_bounds__(128) k_reduce_tail(const void* __restrict__ chunkA, const void* __restrict__ chunkL, uint32_t T, int nb, int logK,
                                                     void* __restrict__ bitsums, void* __restrict__ scaled, uint32_t* __restrict__ cnt,
                                                     void* __restrict__ out_jac, void* __restrict__ out_host, const void* __restrict__ sub_jac,
                                                     // deferred giants (deferred == nullptr: none): row nb + 1 of the grid sums them
                                                     const uint32_t* __restrict__ offsets, uint32_t M, uint32_t nthreads, uint32_t seg_min,
                                                     const void* __restrict__ partials, MsmCombine cb, void* __restrict__ deferred) {
  __shared__ __align__(16) uint32_t smem[4 * 32];
  __shared__ uint32_t flag;
  const int s = blockIdx.y, G = gridDim.x;
  TailFinal fin;
  fin.scaled = scaled; fin.nsums = nb + 1; fin.deferred = deferred; fin.ngiant_p = cb.ctrl + CTRL_NGIANT; fin.max_giants = cb.max_giants;
  fin.cnt = cnt + nb + 1; fin.out_jac = out_jac; fin.out_host = out_host; fin.sub_jac = sub_jac;
  if (s == nb + 1) {  // giant role (only launched when giants are deferred)
    if (cb.ctrl[CTRL_NGIANT] == 0) return;
    const uint32_t L = seg_len(offsets[M], nthreads, seg_min);
    giant_role<C>(offsets, L, blockIdx.x, (uint32_t)G, partials, cb, smem, &flag, nullptr, &fin);
    return;
  }
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
  if (threadIdx.x < 4) {
    acc.store(reinterpret_cast<char*>(bitsums) + ((size_t)s * G + blockIdx.x) * 128);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) flag = (atomicAdd(&cnt[s], 1u) + 1 == (uint32_t)G) ? 1u : 0u;
  __syncthreads();
  if (!flag || threadIdx.x >= 32) return;
  // ---- this warp completes sum s: fold the G block partials, scale by 2^(s+logK)
  __threadfence();
  const int q8 = threadIdx.x >> 2;
  acc = QPoint<C>::identity();
#pragma unroll 1
  for (int it = 0; it < (G + 7) / 8; it++) {
    int g = it * 8 + q8;
    QPoint<C> p = g < G ? QPoint<C>::load_cg(reinterpret_cast<const char*>(bitsums) + ((size_t)s * G + g) * 128) : QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_warp_reduce<C>(acc);
  const int ndbl = s < nb ? s + logK : 0;  // warp-uniform
#pragma unroll 1
  for (int k = 0; k < ndbl; k++) acc = q_dbl<C>(acc);
  if (threadIdx.x < 4) {
    acc.store(reinterpret_cast<char*>(scaled) + (size_t)s * 128);
    __threadfence();
  }
  __syncwarp();
  // ---- the last arrival (a sum or a deferred giant) adds the nb + 1 scaled sums and the giants, converts to Jacobian
  reduce_arrive_final<C>(fin);
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

// ---- direct-table MSM for short commitment keys ------------------------------------------------------------
// The fold step's secondary curve (and the hash / redact step circuits) commit ~10^4-point vectors: there the
// bucket pipeline is pure latency -- sort, accumulate, combine and a ~40-addition-deep bucket reduction for a few
// hundred thousand insertions.  With 180 GB of HBM the key can instead hold EVERY digit multiple:
//   dtable[((j * n + i) << (c-1)) + (k-1)] = k * 2^(c*j) * ck_i,   k = 1 .. 2^(c-1)   (c = 8: 256 KB per point, c = 10: 832 KB)
// so a signed digit is one gather + one mixed addition into a per-thread accumulator and the MSM is a plain sum:
// no buckets, no sort, no bucket reduction -- digits kernel + ONE launch (thread sums -> quad sums -> block sum ->
// last-arriving block of each group of DIRECT_GROUP blocks -> last-arriving group writes the Jacobian result).
#ifndef VIMZ_DIRECT_MUL
#define VIMZ_DIRECT_MUL MulCall  // few additions per thread: a small kernel that stays in the instruction caches
#endif
constexpr int DIRECT_C = 8;
constexpr int DIRECT_C_SHORT = 10;  // keys of <= 16 384 points: 26 windows instead of 32 (832 KB per point)
constexpr uint32_t DIRECT_GROUP = 32;
constexpr uint32_t DIRECT_CTRL_FINAL = 32;  // ctrl[0..31]: per-group arrival counters, ctrl[32]: groups finished
constexpr uint32_t DIRECT_MAX_BLOCKS = DIRECT_GROUP * 32;
constexpr int DIRECT_BATCH = 4;             // multiples normalised per inversion in k_precompute_direct

// one thread per (window j, point i): the 2^(c-1) multiples of table[j][i], affine, DIRECT_BATCH per inversion
template <class C>
__global__ void __launch_bounds__(128) k_precompute_direct(const void* __restrict__ table, uint32_t entries, int cshift,
                                                           void* __restrict__ dtable) {
  using F = Fp<typename C::Fb>;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= entries) return;
  const Affine<C> base = Affine<C>::load(reinterpret_cast<const char*>(table) + (size_t)e * 64);
  char* dst = reinterpret_cast<char*>(dtable) + (((size_t)e) << cshift) * 64;
  const uint32_t K = 1u << cshift;
  if (base.is_identity()) {
    Affine<C> zero;
    zero.x = F::zero(); zero.y = F::zero();
    for (uint32_t k = 0; k < K; k++) zero.store(dst + (size_t)k * 64);
    return;
  }
  base.store(dst);
  Xyzz<C> cur = Xyzz<C>::from_affine(base);
  for (uint32_t k0 = 1; k0 < K; k0 += DIRECT_BATCH) {  // multiples k0+1 .. k0+DIRECT_BATCH
    Xyzz<C> pts[DIRECT_BATCH];
    F prefix[DIRECT_BATCH];
    F run = F::one();
    const int cnt = (int)min((uint32_t)DIRECT_BATCH, K - k0);
    for (int b = 0; b < cnt; b++) {
      xyzz_madd_call<C>(cur, base, false);  // prime-order group: k * P is never the identity for k <= 2^(c-1)
      pts[b] = cur;
      prefix[b] = run;
      run = fp_mul_noinline<typename C::Fb>(run, cur.zzz);
    }
    F inv = fp_inv(run);
    for (int b = cnt - 1; b >= 0; b--) {
      F zi = fp_mul_noinline<typename C::Fb>(inv, prefix[b]);       // 1 / zzz_b
      inv = fp_mul_noinline<typename C::Fb>(inv, pts[b].zzz);
      F t = fp_mul_noinline<typename C::Fb>(pts[b].zz, zi);
      F zzi = fp_mul_noinline<typename C::Fb>(t, t);                // 1 / zz_b
      Affine<C> a;
      a.x = fp_mul_noinline<typename C::Fb>(pts[b].x, zzi);
      a.y = fp_mul_noinline<typename C::Fb>(pts[b].y, zi);
      a.store(dst + (size_t)(k0 + b) * 64);
    }
  }
}

// Tree shared by k_msm_direct and k_masked_base_sum: every thread holds an XYZZ sum; thread sums -> one point per quad -> block
// sum -> the last-arriving block of each group of DIRECT_GROUP blocks adds the group's block sums -> the last-arriving group adds
// the group sums and writes the Jacobian result.  `self_reset`: the final block zeroes the arrival counters for the next launch.
template <class C>
__device__ __forceinline__ void direct_tree(const Xyzz<C>& acc, uint32_t* smem, uint32_t* flag_p, void* __restrict__ partials,
                                            uint32_t* __restrict__ ctrl, void* __restrict__ out_jac, void* __restrict__ out_host, bool self_reset) {
  QPoint<C> q = q_from_lane<C>(acc, 0);
#pragma unroll 1
  for (int j = 1; j < 4; j++) q = q_add<C>(q, q_from_lane<C>(acc, j));
  q = q_block_reduce_128<C>(q, smem);
  char* parts = reinterpret_cast<char*>(partials);
  const uint32_t nblocks = gridDim.x, ngroups = (nblocks + DIRECT_GROUP - 1) / DIRECT_GROUP;
  const uint32_t g = blockIdx.x / DIRECT_GROUP;
  const uint32_t gsize = min(DIRECT_GROUP, nblocks - g * DIRECT_GROUP);
  if (threadIdx.x < 4) {
    q.store(parts + (size_t)blockIdx.x * 128);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) *flag_p = (atomicAdd(&ctrl[g], 1u) + 1 == gsize) ? 1u : 0u;
  __syncthreads();
  if (!*flag_p) return;
  // last block of group g: its 32 quads fetch the group's block sums (through L2) and add them
  __threadfence();
  const uint32_t quad = threadIdx.x >> 2;
  q = quad < gsize ? QPoint<C>::load_cg(parts + (size_t)(g * DIRECT_GROUP + quad) * 128) : QPoint<C>::identity();
  q = q_block_reduce_128<C>(q, smem);
  if (threadIdx.x < 4) {
    q.store(parts + (size_t)(nblocks + g) * 128);
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (self_reset) ctrl[g] = 0;
    *flag_p = (atomicAdd(&ctrl[DIRECT_CTRL_FINAL], 1u) + 1 == ngroups) ? 1u : 0u;
  }
  __syncthreads();
  if (!*flag_p) return;
  // last group: add the <= 32 group sums, write the Jacobian result
  __threadfence();
  q = quad < ngroups ? QPoint<C>::load_cg(parts + (size_t)(nblocks + quad) * 128) : QPoint<C>::identity();
  q = q_block_reduce_128<C>(q, smem);
  if (threadIdx.x < 32) q_store_jacobian<C>(q, out_jac, threadIdx.x < 4, out_host);
  if (self_reset && threadIdx.x == 0) ctrl[DIRECT_CTRL_FINAL] = 0;
}

template <class C>
__global__ void __launch_bounds__(128, 4) k_msm_direct(const uint32_t* __restrict__ digits, uint32_t n, int nwin, uint32_t ck_n, uint32_t first,
                                                       int cshift, const void* __restrict__ dtable, void* __restrict__ partials,
                                                       uint32_t* __restrict__ ctrl, void* __restrict__ out_jac, void* __restrict__ out_host) {
  __shared__ __align__(16) uint32_t smem[4 * 32];
  __shared__ uint32_t flag;
  const uint32_t nthreads = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = n * (uint32_t)nwin;
  const char* tab = reinterpret_cast<const char*>(dtable);
  auto address = [&](uint32_t e, uint32_t d) {
    const uint32_t j = e / n, i = e - j * n;
    return tab + ((((size_t)j * ck_n + first + i) << cshift) + ((d & 0x7fffffffu) - 1)) * 64;
  };
  // entry e = j * n + i of the recoded digit array; thread t takes e = t, t + nthreads, ... (coalesced digit reads)
  Xyzz<C> acc = Xyzz<C>::identity();
  uint32_t e = t;
  uint32_t d = e < total ? __ldg(digits + e) : 0u;
  while (e < total) {
    const uint32_t en = e + nthreads;
    const uint32_t dn = en < total ? __ldg(digits + en) : 0u;  // next digit in flight during this addition
    if (d != 0) {
      Affine<C> p = Affine<C>::load_nc(address(e, d));
      xyzz_madd<C, VIMZ_DIRECT_MUL>(acc, p, (d >> 31) != 0);
    }
    e = en;
    d = dn;
  }
  direct_tree<C>(acc, smem, &flag, partials, ctrl, out_jac, out_host, true);
}

// sum over the booleanity rows i (bitcol[i] != ~0) of vals[bitcol[i]] * bases[i] for a vector `vals` whose entries there are
// (almost always) 0 or 1: the fresh witness W2 of a step -- on such a row (A z2)_i IS the wire W2[bitcol[i]] (r1cs.cuh,
// k_cross_finish; a satisfying witness cannot hold anything else there), so this sum needs no mat-vec and starts with the step.
// An entry equal to one is ONE mixed addition; any other non-zero value -- only an unsatisfying witness has them -- is multiplied
// out bit by bit so that the result stays exact.  Same tree as k_msm_direct; the arrival counters reset themselves.
template <class C>
__global__ void __launch_bounds__(128, 2) k_masked_base_sum(const void* __restrict__ vals, const uint32_t* __restrict__ bitcol, uint32_t m,
                                                            const void* __restrict__ bases, void* __restrict__ partials,
                                                            uint32_t* __restrict__ ctrl, void* __restrict__ out_jac) {
  using Fs = Fp<typename C::Fs>;
  __shared__ __align__(16) uint32_t smem[4 * 32];
  __shared__ uint32_t flag;
  // Rows whose wire is 1 are ~43 % of a chunk, scattered: added where they are found, every lane of a warp would sit through an
  // addition for every chunk.  Each warp queues the row numbers of its hits instead (ballot + prefix count) and adds 32 at a time,
  // one per lane: ~6 dense additions per thread instead of ~14 sparse ones.
  __shared__ uint32_t queue[4][64];
  const uint32_t nthreads = gridDim.x * blockDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Xyzz<C> acc = Xyzz<C>::identity();
  uint32_t qn = 0;  // warp-uniform
  const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
#pragma unroll 1
  for (uint32_t base = first; base < m; base += nthreads) {  // warp-uniform trip count
    const uint32_t i = base + lane;
    bool hit = false;
    if (i < m) {
      const uint32_t col = __ldg(bitcol + i);
      if (col != 0xffffffffu) {
        const Fs v = Fs::load(reinterpret_cast<const char*>(vals) + (size_t)col * 32);
        if (v == Fs::one()) {
          hit = true;
        } else if (!v.is_zero()) {  // rare, slow, exact: v * p by double-and-add
          const Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(bases) + (size_t)i * 64);
          const Fs raw = fp_from_mont(v);
          Xyzz<C> t = Xyzz<C>::identity();
          for (int bit = 255; bit >= 0; bit--) {
            xyzz_dbl_call<C>(t);
            if ((raw.v[bit >> 5] >> (bit & 31)) & 1) xyzz_madd_call<C>(t, p, false);
          }
          xyzz_add_call<C>(acc, t);
        }
      }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, hit);
    if (hit) queue[warp][qn + __popc(mask & ((1u << lane) - 1))] = i;
    qn += __popc(mask);
    __syncwarp();
    if (qn >= 32) {
      const uint32_t idx = queue[warp][lane];
      const uint32_t spill = lane < qn - 32 ? queue[warp][32 + lane] : 0;
      __syncwarp();
      if (lane < qn - 32) queue[warp][lane] = spill;
      qn -= 32;
      __syncwarp();
      const Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(bases) + (size_t)idx * 64);
      xyzz_madd<C, VIMZ_DIRECT_MUL>(acc, p, false);
    }
  }
  if (lane < qn) {
    const Affine<C> p = Affine<C>::load_nc(reinterpret_cast<const char*>(bases) + (size_t)queue[warp][lane] * 64);
    xyzz_madd_call<C>(acc, p, false);
  }
  direct_tree<C>(acc, smem, &flag, partials, ctrl, out_jac, nullptr, true);
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

// out[s] = sum_r pts[r * sets + s] for `sets` independent sums of k Jacobian points each (combining the per-rank
// partial commitments of a sharded fold step): block s, one warp, its 8 quads stride over the k points.
template <class C>
__global__ void __launch_bounds__(32) k_point_sum_batch(const void* __restrict__ pts, uint32_t k, uint32_t sets, void* __restrict__ out) {
  const uint32_t s = blockIdx.x, quad = threadIdx.x >> 2;
  QPoint<C> acc = QPoint<C>::identity();
#pragma unroll 1
  for (uint32_t it = 0; it < (k + 7) / 8; it++) {  // warp-uniform trip count
    const uint32_t r = it * 8 + quad;
    QPoint<C> p = q_load_jacobian<C>(reinterpret_cast<const char*>(pts) + ((size_t)(r < k ? r : 0) * sets + s) * 96);
    if (r >= k) p = QPoint<C>::identity();
    acc = q_add<C>(acc, p);
  }
  acc = q_warp_reduce<C>(acc);
  q_store_jacobian<C>(acc, reinterpret_cast<char*>(out) + (size_t)s * 96, threadIdx.x < 4);
}

// out = a + b for two Jacobian points (the early and the in-step share of comm_W2), second copy to mapped host memory
template <class C>
__global__ void __launch_bounds__(32) k_point_add2(const void* __restrict__ a, const void* __restrict__ b, void* __restrict__ out, void* __restrict__ out_host) {
  QPoint<C> acc = q_add<C>(q_load_jacobian<C>(a), q_load_jacobian<C>(b));
  q_store_jacobian<C>(acc, out, threadIdx.x < 4, out_host);
}

// out[t] = a[t] + r * b[t] for count <= 8 independent pairs, one QUAD each (one warp in total); r is a
// Montgomery scalar shared by all pairs (RelaxedR1CSInstance::fold uses the same r for comm_W and comm_E).
template <class C>
__device__ __forceinline__ void point_scale_add_body(const void* __restrict__ a, Fp<typename C::Fs> r_mont, const void* __restrict__ b,
                                                     void* __restrict__ out, int count) {
  using Fs = Fp<typename C::Fs>;
  const int t = threadIdx.x >> 2;
  const bool valid = t < count;
  Fs r = fp_from_mont(r_mont);
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
template <class C>
__global__ void __launch_bounds__(32) k_point_scale_add(const void* __restrict__ a, const void* __restrict__ r_mont, const void* __restrict__ b,
                                                        void* __restrict__ out, int count) {
  point_scale_add_body<C>(a, Fp<typename C::Fs>::load(r_mont), b, out, count);
}
// r by value (the fold step: no staging copy, no cross-stream event for the challenge)
template <class C>
__global__ void __launch_bounds__(32) k_point_scale_add_val(const void* __restrict__ a, Fp<typename C::Fs> r_mont, const void* __restrict__ b,
                                                            void* __restrict__ out, int count) {
  point_scale_add_body<C>(a, r_mont, b, out, count);
}

// parts[j] = 2^(PART_BITS j) * P, j = 0 .. SCALE_PARTS-1, as XYZZ records (128 B each), from a Jacobian P: one quad walks the doublings.
// Runs inside step_begin beside the commitments (P = the step's P_S is known ~0.15 ms into the step, r only after it), so that the
// scalar multiplication r * P_S in step_end is SCALE_PARTS independent PART_BITS-bit pieces instead of one 128-bit chain.
constexpr int SCALE_PARTS = 2, PART_BITS = 64;
template <class C>
__global__ void __launch_bounds__(32) k_point_pow2_parts(const void* __restrict__ jac, void* __restrict__ parts) {
  QPoint<C> p = q_load_jacobian<C>(jac);
  if (threadIdx.x < 4) p.store(parts);
#pragma unroll 1
  for (int j = 1; j < SCALE_PARTS; j++) {
#pragma unroll 1
    for (int b = 0; b < PART_BITS; b++) p = q_dbl<C>(p);
    if (threadIdx.x < 4) p.store(reinterpret_cast<char*>(parts) + (size_t)j * 128);
  }
}

// out = a + r * P given parts[j] = 2^(PART_BITS j) P (k_point_pow2_parts): warp j multiplies parts[j] by bits [PART_BITS j, PART_BITS (j + 1))
// of r, warp 0 adds the pieces and `a`.  A 128-bit r (Nova's challenge) is 64 doublings + ~32 additions deep instead of 128 + ~64; a
// wider r takes the plain chain on warp 0 (exact for any field element).  One block of SCALE_PARTS warps; every quad of a warp does
// the same work.
template <class C>
__global__ void __launch_bounds__(32 * SCALE_PARTS) k_point_scale_add_parts(const void* __restrict__ a, Fp<typename C::Fs> r_mont,
                                                                           const void* __restrict__ parts, void* __restrict__ out) {
  using Fs = Fp<typename C::Fs>;
  static_assert(SCALE_PARTS * PART_BITS == 128 && PART_BITS % 32 == 0 && SCALE_PARTS <= 8, "pieces of a 128-bit challenge");
  __shared__ __align__(16) uint32_t smem[SCALE_PARTS * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs r = fp_from_mont(r_mont);
  const bool wide = (r.v[4] | r.v[5] | r.v[6] | r.v[7]) != 0;  // block-uniform
  QPoint<C> acc = QPoint<C>::identity();
  if (!wide || warp == 0) {
    const QPoint<C> base = QPoint<C>::load(reinterpret_cast<const char*>(parts) + (size_t)warp * 128);
    int top = wide ? 255 : PART_BITS * (warp + 1) - 1;
    const int low = wide ? 0 : PART_BITS * warp;
    while (top >= low && !((r.v[top >> 5] >> (top & 31)) & 1)) top--;
#pragma unroll 1
    for (int bit = top; bit >= low; bit--) {  // warp-uniform trip count and branch
      acc = q_dbl<C>(acc);
      if ((r.v[bit >> 5] >> (bit & 31)) & 1) acc = q_add<C>(acc, base);
    }
  }
  if (lane < 4) acc.store(smem + warp * 32);
  __syncthreads();
  if (warp != 0) return;
  acc = lane < 4 * SCALE_PARTS ? QPoint<C>::load(smem + (lane >> 2) * 32) : QPoint<C>::identity();
  acc = q_warp_reduce<C>(acc, SCALE_PARTS > 4 ? 4 : (SCALE_PARTS > 2 ? 2 : 1));
  acc = q_add<C>(acc, q_load_jacobian<C>(a));
  q_store_jacobian<C>(acc, out, lane < 4);
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
