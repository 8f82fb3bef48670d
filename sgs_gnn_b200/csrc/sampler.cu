// K2: exponential-race ("Gumbel") top-q edge sampler.
//   key_e = s_e / noise_e  (IEEE fp32, no FMA contraction), MSD radix select over the key bit
//   patterns (11 + 11 + 9 bits below the sign bit), ties at the threshold broken by lowest edge id,
//   ballot/scan compaction in ascending edge id.  HBM-bound: 128-bit streaming loads throughout.
#include "common.cuh"

#include <stdlib.h>

namespace sgs {

constexpr int kBins = SGS_TOPQ_BINS;  // 2048
// keys are >= 0, so bit 31 is always clear: digits are [30:20], [19:9], [8:0]
__host__ __device__ __forceinline__ int level_shift(int level) { return level == 0 ? 20 : (level == 1 ? 9 : 0); }
__host__ __device__ __forceinline__ int level_bins(int level) { return level == 2 ? 512 : 2048; }

__device__ __forceinline__ uint32_t make_key(float p, float prob, float noise, float S_eff, float c_p,
                                             float c_prob, int mode, bool& bad) {
  float s;
  if (mode == SGS_SAMPLE_RAW) {
    s = p;
  } else {
    s = __fdiv_rn(p, S_eff);
    if (mode == SGS_SAMPLE_TRAIN) s = __fadd_rn(__fmul_rn(c_p, s), __fmul_rn(c_prob, prob));
  }
  bad |= !(s >= 0.f) || isinf(s);
  const float key = __fdiv_rn(s, noise);
  return __float_as_uint(key) & 0x7fffffffu;
}

// warp-aggregated shared-memory histogram increment
__device__ __forceinline__ void hist_add(int* sh, int bin, bool valid) {
  const unsigned active = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const unsigned peers = __match_any_sync(active, bin);
  if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(sh + bin, __popc(peers));
}

__global__ void __launch_bounds__(512)
topq_keys_kernel(const float* __restrict__ p, const float* __restrict__ prob, const float* __restrict__ noise,
                 int64_t E, float c_p, float c_prob, int mode, const float* __restrict__ S,
                 uint32_t* __restrict__ keys, unsigned long long* __restrict__ hist,
                 long long* __restrict__ state) {
  __shared__ int sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float S_eff = (mode == SGS_SAMPLE_RAW) ? 1.f : __fadd_rn(S[0], 1e-12f);
  bool bad = false;
  const int64_t n4 = E >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const bool use_prob = (mode == SGS_SAMPLE_TRAIN);
  // n4 rounded up to a warp multiple so that every lane reaches the warp-collective hist_add
  const int64_t n4_round = (n4 + 31) & ~(int64_t)31;
  for (int64_t i = tid; i < n4_round; i += nthreads) {
    const bool valid = i < n4;
    uint4 k = make_uint4(0, 0, 0, 0);
    if (valid) {
      const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
      const float4 z = ld_stream_f4(reinterpret_cast<const float4*>(noise) + i);
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (use_prob) b = ld_stream_f4(reinterpret_cast<const float4*>(prob) + i);
      k.x = make_key(a.x, b.x, z.x, S_eff, c_p, c_prob, mode, bad);
      k.y = make_key(a.y, b.y, z.y, S_eff, c_p, c_prob, mode, bad);
      k.z = make_key(a.z, b.z, z.z, S_eff, c_p, c_prob, mode, bad);
      k.w = make_key(a.w, b.w, z.w, S_eff, c_p, c_prob, mode, bad);
      reinterpret_cast<uint4*>(keys)[i] = k;
    }
    hist_add(sh, k.x >> 20, valid);
    hist_add(sh, k.y >> 20, valid);
    hist_add(sh, k.z >> 20, valid);
    hist_add(sh, k.w >> 20, valid);
  }
  // tail (E % 4 elements), handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const bool valid = i < E;
    uint32_t k = 0;
    if (valid) {
      k = make_key(p[i], use_prob ? prob[i] : 0.f, noise[i], S_eff, c_p, c_prob, mode, bad);
      keys[i] = k;
    }
    hist_add(sh, k >> 20, valid);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
  if (bad) state[5] = 1;
}

// gate (may be NULL): the kernel runs only if *gate != 0 (fallback passes of the sampled-window fast path)
__global__ void __launch_bounds__(512)
topq_hist_kernel(const uint32_t* __restrict__ keys, int64_t E, unsigned long long* __restrict__ hist,
                 const long long* __restrict__ state, int level, const uint32_t* __restrict__ gate) {
  __shared__ int sh[kBins];
  if (gate && *gate == 0) return;
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int shift = level_shift(level);
  const int hi_shift = (level == 0) ? 31 : ((level == 1) ? 20 : 9);  // bits above this digit must match the prefix
  const uint32_t prefix_hi = (level == 0) ? 0u : ((uint32_t)state[0] >> hi_shift);
  const uint32_t mask = (uint32_t)level_bins(level) - 1;
  const int64_t n4 = E >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4_round = (n4 + 31) & ~(int64_t)31;
  for (int64_t i = tid; i < n4_round; i += nthreads) {
    const bool valid = i < n4;
    uint4 k = make_uint4(0, 0, 0, 0);
    if (valid) k = ld_stream_u4(reinterpret_cast<const uint4*>(keys) + i);
    hist_add(sh, (k.x >> shift) & mask, valid && (k.x >> hi_shift) == prefix_hi);
    hist_add(sh, (k.y >> shift) & mask, valid && (k.y >> hi_shift) == prefix_hi);
    hist_add(sh, (k.z >> shift) & mask, valid && (k.z >> hi_shift) == prefix_hi);
    hist_add(sh, (k.w >> shift) & mask, valid && (k.w >> hi_shift) == prefix_hi);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const bool valid = i < E;
    const uint32_t k = valid ? keys[i] : 0u;
    hist_add(sh, (k >> shift) & mask, valid && (k >> hi_shift) == prefix_hi);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// One block of 1024 threads: the bin holding the k-th largest key, by a block-wide suffix scan of the histogram
// (the bin b >= 1, scanning down from the top, with  #keys above b < k <= #keys above b + hist[b];  bin 0 if none).
__global__ void __launch_bounds__(1024)
topq_find_kernel(unsigned long long* __restrict__ hist, long long* __restrict__ state, long long k_total,
                 int level, const uint32_t* __restrict__ gate) {
  __shared__ long long warp_tot[32];
  __shared__ long long s_bin, s_rem, s_cnt;
  if (gate && *gate == 0) return;
  const int nb = level_bins(level);
  const int per = nb >= 1024 ? nb / 1024 : 1;          // bins per thread (2 or 1)
  const int b0 = threadIdx.x * per;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long k = (level == 0) ? k_total : state[1];
  long long c[2] = {0, 0};
  if (b0 < nb) {
    c[0] = (long long)hist[b0];
    if (per == 2) c[1] = (long long)hist[b0 + 1];
  }
  if (threadIdx.x == 0) s_bin = -1;
  // suffix sums over threads: above = number of keys in bins owned by higher threads
  long long v = c[0] + c[1];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  if (lane == 0) warp_tot[wid] = v;
  __syncthreads();
  long long above = v - (c[0] + c[1]);
  for (int w = wid + 1; w < 32; ++w) above += warp_tot[w];
  if (b0 < nb) {
    long long cum = above;
    for (int j = per - 1; j >= 0; --j) {
      const int b = b0 + j;
      const bool hit = (k <= 0) ? (b == nb - 1) : (b >= 1 && cum < k && cum + c[j] >= k);
      if (hit) {
        s_bin = b;
        s_rem = k - cum;   // how many keys to take from bin b (>= 1 when k >= 1)
        s_cnt = c[j];
      }
      cum += c[j];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_bin < 0) {       // fewer than k keys above bin 0: everything comes from bin 0 downwards
      s_bin = 0;
      s_rem = k - (above + c[1]);
      s_cnt = c[0];
    }
    const long long prefix = ((level == 0) ? 0 : state[0]) | (s_bin << level_shift(level));
    state[0] = prefix;
    state[1] = s_rem;
    if (level == 2) {
      state[2] = prefix;           // tau bit pattern
      state[3] = k_total - s_rem;  // # keys strictly greater than tau
      state[4] = s_rem;            // # threshold ties to take (lowest edge ids first)
      state[6] = s_cnt;            // # keys equal to tau
    }
  }
  __syncthreads();   // every thread has read its bins
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) hist[i] = 0ull;
}

// ---------------------------------------------------------------------------------------
// Sampled-window fast path of the first two radix levels (single-GPU sgs_sample_topq, large E).
// A full level-0 histogram costs one MATCH.ANY (or one shared-memory atomic) per key: ncu r01 shows the keys pass
// bound by the ADU pipe at 91 % with DRAM at 24-30 %.  Instead:
//   (1) a 1/64 strided sample of the keys is histogrammed and the level-0 bin holding the q-th largest key is
//       PREDICTED (window = that bin and its two neighbours),
//   (2) the full keys pass only counts the keys above the window (one compare per key, per-thread counters) and
//       histograms the 22-bit prefix (level-0 bin, level-1 digit) of the ~15 % of keys inside the window with plain
//       shared-memory atomics,
//   (3) the exact counts decide: if the q-th largest key lies inside the window the 22-bit prefix is read off the
//       window histogram (levels 0 and 1 done, bit-exact by construction); otherwise a device flag enables the
//       classic level-0 / level-1 passes over the stored keys (they are launched either way and exit on the flag).
// ---------------------------------------------------------------------------------------
constexpr int kWinBins = 3;                                    // level-0 bins in the window
constexpr int kWinHist = kWinBins * kBins;                     // 6144 (window bin, level-1 digit) counters
constexpr size_t kWinScratchBytes = (size_t)(kBins + kWinHist + 16) * sizeof(unsigned long long);
// scratch layout (unsigned long long): shist[kBins] | whist[kWinHist] | misc[16]: 0 = window low bin, 1 = miss flag
// (also read as uint32 gate), 2 = #keys above the window, 3 = #sampled keys

__global__ void __launch_bounds__(256)
topq_sample_kernel(const float* __restrict__ p, const float* __restrict__ prob, const float* __restrict__ noise,
                   int64_t E, float c_p, float c_prob, int mode, const float* __restrict__ S, int stride,
                   unsigned long long* __restrict__ shist, unsigned long long* __restrict__ misc) {
  __shared__ int sh[kBins];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float S_eff = (mode == SGS_SAMPLE_RAW) ? 1.f : __fadd_rn(S[0], 1e-12f);
  const bool use_prob = (mode == SGS_SAMPLE_TRAIN);
  const int64_t n4 = E >> 2;
  const int64_t ns = (n4 + stride - 1) / stride;               // sampled float4 groups
  bool bad = false;
  int cnt = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t * stride;
    const float4 a = reinterpret_cast<const float4*>(p)[i];
    const float4 z = reinterpret_cast<const float4*>(noise)[i];
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (use_prob) b = reinterpret_cast<const float4*>(prob)[i];
    atomicAdd(&sh[make_key(a.x, b.x, z.x, S_eff, c_p, c_prob, mode, bad) >> 20], 1);
    atomicAdd(&sh[make_key(a.y, b.y, z.y, S_eff, c_p, c_prob, mode, bad) >> 20], 1);
    atomicAdd(&sh[make_key(a.z, b.z, z.z, S_eff, c_p, c_prob, mode, bad) >> 20], 1);
    atomicAdd(&sh[make_key(a.w, b.w, z.w, S_eff, c_p, c_prob, mode, bad) >> 20], 1);
    cnt += 4;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(shist + i, (unsigned long long)sh[i]);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(misc + 3, (unsigned long long)cnt);
}

// one block of 1024 threads: window low bin from the sample histogram (force_miss: a window that cannot hold tau)
__global__ void __launch_bounds__(1024)
topq_predict_kernel(const unsigned long long* __restrict__ shist, unsigned long long* __restrict__ misc,
                    long long k_total, long long E, int force_miss) {
  __shared__ long long warp_tot[32];
  __shared__ int s_bin;
  const int b0 = threadIdx.x * 2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long c0 = (long long)shist[b0], c1 = (long long)shist[b0 + 1];
  const long long ns = (long long)misc[3];
  // rank of the threshold inside the sample (rounded to nearest, at least 1)
  long long ks = (long long)((double)k_total * (double)ns / (double)E + 0.5);
  if (ks < 1) ks = 1;
  if (threadIdx.x == 0) s_bin = 0;
  long long v = c0 + c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  if (lane == 0) warp_tot[wid] = v;
  __syncthreads();
  long long above = v - (c0 + c1);
  for (int w = wid + 1; w < 32; ++w) above += warp_tot[w];
  if (above + c1 < ks && above + c1 + c0 >= ks) s_bin = b0;   // unique (suffix sums are monotone)
  if (above < ks && above + c1 >= ks) s_bin = b0 + 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    int lo = s_bin - 1;
    if (lo < 0) lo = 0;
    if (lo > kBins - kWinBins) lo = kBins - kWinBins;
    if (force_miss) lo = 0;        // keys of valid inputs never reach the three lowest exponent bins' neighbourhood
    misc[0] = (unsigned long long)lo;
  }
}

__global__ void __launch_bounds__(512)
topq_keys_window_kernel(const float* __restrict__ p, const float* __restrict__ prob, const float* __restrict__ noise,
                        int64_t E, float c_p, float c_prob, int mode, const float* __restrict__ S,
                        uint32_t* __restrict__ keys, unsigned long long* __restrict__ whist,
                        unsigned long long* __restrict__ misc, long long* __restrict__ state) {
  __shared__ int sh[kWinHist];
  for (int i = threadIdx.x; i < kWinHist; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const float S_eff = (mode == SGS_SAMPLE_RAW) ? 1.f : __fadd_rn(S[0], 1e-12f);
  const uint32_t lo = (uint32_t)misc[0];
  bool bad = false;
  int above = 0;
  const int64_t n4 = E >> 2;
  const bool use_prob = (mode == SGS_SAMPLE_TRAIN);
  auto classify = [&](uint32_t k) {
    const uint32_t w = (k >> 20) - lo;            // window bin (unsigned: bins below the window wrap to huge values)
    above += (int)(w >= (uint32_t)kWinBins && (k >> 20) > lo);
    if (w < (uint32_t)kWinBins) atomicAdd(&sh[w * kBins + ((k >> 9) & (kBins - 1))], 1);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
    const float4 z = ld_stream_f4(reinterpret_cast<const float4*>(noise) + i);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (use_prob) b = ld_stream_f4(reinterpret_cast<const float4*>(prob) + i);
    uint4 k;
    k.x = make_key(a.x, b.x, z.x, S_eff, c_p, c_prob, mode, bad);
    k.y = make_key(a.y, b.y, z.y, S_eff, c_p, c_prob, mode, bad);
    k.z = make_key(a.z, b.z, z.z, S_eff, c_p, c_prob, mode, bad);
    k.w = make_key(a.w, b.w, z.w, S_eff, c_p, c_prob, mode, bad);
    reinterpret_cast<uint4*>(keys)[i] = k;
    classify(k.x);
    classify(k.y);
    classify(k.z);
    classify(k.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < 4) {      // tail (E % 4 elements)
    const int64_t i = (n4 << 2) + threadIdx.x;
    if (i < E) {
      const uint32_t k = make_key(p[i], use_prob ? prob[i] : 0.f, noise[i], S_eff, c_p, c_prob, mode, bad);
      keys[i] = k;
      classify(k);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kWinHist; i += blockDim.x)
    if (sh[i]) atomicAdd(whist + i, (unsigned long long)sh[i]);
  above = warp_sum(above);
  if ((threadIdx.x & 31) == 0 && above) atomicAdd(misc + 2, (unsigned long long)above);
  if (bad) state[5] = 1;
}

// one block of 1024 threads, 6 window counters per thread: either the 22-bit prefix (levels 0 and 1 done) or the
// miss flag that enables the classic passes
__global__ void __launch_bounds__(1024)
topq_find_window_kernel(const unsigned long long* __restrict__ whist, unsigned long long* __restrict__ misc,
                        long long* __restrict__ state, long long k_total) {
  constexpr int PER = kWinHist / 1024;
  __shared__ long long warp_tot[32];
  __shared__ long long s_idx, s_rem;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int b0 = threadIdx.x * PER;
  long long c[PER], tsum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    c[j] = (long long)whist[b0 + j];
    tsum += c[j];
  }
  if (threadIdx.x == 0) s_idx = -1;
  long long v = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  if (lane == 0) warp_tot[wid] = v;
  __syncthreads();
  long long above = v - tsum;
  for (int w = wid + 1; w < 32; ++w) above += warp_tot[w];
  const long long k = k_total - (long long)misc[2];     // rank inside the window (from the top)
  long long cum = above;
#pragma unroll
  for (int j = PER - 1; j >= 0; --j) {
    if (k >= 1 && cum < k && cum + c[j] >= k) {
      s_idx = b0 + j;
      s_rem = k - cum;
    }
    cum += c[j];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_idx < 0) {
      misc[1] = 1ull;                                   // miss: tau is not inside the window
    } else {
      const long long lo = (long long)misc[0];
      state[0] = ((lo + s_idx / kBins) << 20) | ((s_idx % kBins) << 9);
      state[1] = s_rem;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Compaction: count per 8192-key chunk, scan the chunk counts, then write.
// ---------------------------------------------------------------------------------------
constexpr int kChunkThreads = 1024;
constexpr int kPerThread = 8;
constexpr int kChunk = kChunkThreads * kPerThread;

__device__ __forceinline__ void load8(const uint32_t* __restrict__ keys, int64_t base, int64_t E, uint32_t k[8]) {
  if (base + 8 <= E && ((uintptr_t)(keys + base) & 15) == 0) {
    const uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(keys + base));
    const uint4 b = ld_stream_u4(reinterpret_cast<const uint4*>(keys + base) + 1);
    k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
    k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = (base + j < E) ? keys[base + j] : 0xffffffffu;  // never matches
  }
}

__global__ void __launch_bounds__(kChunkThreads)
topq_count_kernel(const uint32_t* __restrict__ keys, int64_t E, const long long* __restrict__ state,
                  int* __restrict__ blk_gt, int* __restrict__ blk_eq) {
  __shared__ int s_gt, s_eq;
  if (threadIdx.x == 0) { s_gt = 0; s_eq = 0; }
  __syncthreads();
  const uint32_t tau = (uint32_t)state[2];
  const int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kPerThread;
  uint32_t k[8];
  load8(keys, base, E, k);
  int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool in = base + j < E;
    gt += in && (k[j] > tau);
    eq += in && (k[j] == tau);
  }
  gt = warp_sum(gt);
  eq = warp_sum(eq);
  if ((threadIdx.x & 31) == 0) {
    if (gt) atomicAdd(&s_gt, gt);
    if (eq) atomicAdd(&s_eq, eq);
  }
  __syncthreads();
  if (threadIdx.x == 0) { blk_gt[blockIdx.x] = s_gt; blk_eq[blockIdx.x] = s_eq; }
}

// single block exclusive scan of both count arrays (in place)
__global__ void __launch_bounds__(1024)
topq_scan_kernel(int* __restrict__ blk_gt, int* __restrict__ blk_eq, int nb) {
  __shared__ int sg[1024], se[1024];
  __shared__ int carry_g, carry_e;
  if (threadIdx.x == 0) { carry_g = 0; carry_e = 0; }
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const int g = i < nb ? blk_gt[i] : 0;
    const int e = i < nb ? blk_eq[i] : 0;
    sg[threadIdx.x] = g;
    se[threadIdx.x] = e;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      int vg = 0, ve = 0;
      if (threadIdx.x >= o) { vg = sg[threadIdx.x - o]; ve = se[threadIdx.x - o]; }
      __syncthreads();
      sg[threadIdx.x] += vg;
      se[threadIdx.x] += ve;
      __syncthreads();
    }
    if (i < nb) {
      blk_gt[i] = carry_g + sg[threadIdx.x] - g;
      blk_eq[i] = carry_e + se[threadIdx.x] - e;
    }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_g += sg[1023]; carry_e += se[1023]; }
    __syncthreads();
  }
}

// The selected ids of a chunk are first compacted into shared memory (block-local output order) and then copied out
// with coalesced 4-byte stores: ncu r01 showed the direct form -- every thread storing its ~1.6 selected ids one
// predicated STG at a time -- at 19 % of DRAM bandwidth, 0.32 ms per draw.
__global__ void __launch_bounds__(kChunkThreads)
topq_write_kernel(const uint32_t* __restrict__ keys, int64_t E, const long long* __restrict__ state,
                  long long tie_skip, const int* __restrict__ blk_gt, const int* __restrict__ blk_eq,
                  int32_t* __restrict__ sel, int64_t q_cap, uint8_t* __restrict__ mask,
                  long long* __restrict__ n_sel_out) {
  __shared__ int wg[32], we[32];
  __shared__ int tot_g, tot_e;
  __shared__ int32_t stage[kChunk];
  const uint32_t tau = (uint32_t)state[2];
  long long avail = state[4] - tie_skip;  // ties this shard may still take
  if (avail < 0) avail = 0;
  const int64_t base = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kPerThread;
  uint32_t k[8];
  load8(keys, base, E, k);
  int gt = 0, eq = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool in = base + j < E;
    gt += in && (k[j] > tau);
    eq += in && (k[j] == tau);
  }
  // block-wide exclusive scan of (gt, eq) in thread order
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int ig = gt, ie = eq;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int vg = __shfl_up_sync(0xffffffffu, ig, o);
    const int ve = __shfl_up_sync(0xffffffffu, ie, o);
    if (lane >= o) { ig += vg; ie += ve; }
  }
  if (lane == 31) { wg[wid] = ig; we[wid] = ie; }
  __syncthreads();
  if (wid == 0) {
    int a = wg[lane], b = we[lane];
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int va = __shfl_up_sync(0xffffffffu, ia, o);
      const int vb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += va; ib += vb; }
    }
    wg[lane] = ia - a;
    we[lane] = ib - b;
    if (lane == 31) { tot_g = ia; tot_e = ib; }
  }
  __syncthreads();
  const long long bg = blk_gt[blockIdx.x], be = blk_eq[blockIdx.x];
  const long long be_taken = be < avail ? be : avail;        // ties taken before this chunk
  const long long pos_block0 = bg + be_taken;                // output position of the chunk's first selected id
  uint64_t mbits = 0;
  if (tot_e == 0) {
    // no threshold tie in this chunk (ties are a handful per 10^8 keys): block-local positions in 32-bit arithmetic
    int loc = wg[wid] + (ig - gt);
    const int32_t id0 = (int32_t)base;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (base + j < E && k[j] > tau) {
        stage[loc++] = id0 + j;
        mbits |= (uint64_t)1 << (8 * j);
      }
    }
  } else {
    long long g_before = bg + wg[wid] + (ig - gt);
    long long e_before = be + we[wid] + (ie - eq);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool in = base + j < E;
      const bool is_gt = in && (k[j] > tau);
      const bool is_eq = in && (k[j] == tau);
      if (is_gt || (is_eq && e_before < avail)) {
        const long long pos = g_before + (e_before < avail ? e_before : avail);
        stage[(int)(pos - pos_block0)] = (int32_t)(base + j);
        mbits |= (uint64_t)1 << (8 * j);
      }
      g_before += is_gt;
      e_before += is_eq;
    }
  }
  if (mask) {
    if (base + 8 <= E && ((uintptr_t)(mask + base) & 7) == 0) {
      *reinterpret_cast<uint64_t*>(mask + base) = mbits;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j < E) mask[base + j] = (uint8_t)((mbits >> (8 * j)) & 1);
    }
  }
  __syncthreads();
  const long long e_end = be + tot_e;
  const int n_out = tot_g + (int)((e_end < avail ? e_end : avail) - be_taken);
  if (sel) {
    for (int i = threadIdx.x; i < n_out; i += kChunkThreads)
      if (pos_block0 + i < q_cap) sel[pos_block0 + i] = stage[i];
  }
  if (n_sel_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_sel_out = pos_block0 + n_out;
}

// ---------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------
constexpr int kRedBlocks = 1024;

__global__ void __launch_bounds__(256) sum_partial_kernel(const float* __restrict__ p, int64_t n,
                                                          double* __restrict__ partial) {
  __shared__ double sh[8];
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nt) {
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(p) + i);
    acc += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
  if (tid < (n & 3)) acc += (double)p[(n4 << 2) + tid];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void sum_final_kernel(const double* __restrict__ partial, int nb, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nb; ++i) t += partial[i];
    out[0] = (float)t;
  }
}

__global__ void __launch_bounds__(256) max_partial_kernel(const float* __restrict__ x, int64_t n,
                                                          float* __restrict__ partial) {
  __shared__ float sh[8];
  float m = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
    partial[blockIdx.x] = fmaxf(m, sh[0]);
  }
}
__global__ void max_final_kernel(const float* __restrict__ partial, int nb, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float m = -INFINITY;
    for (int i = 0; i < nb; ++i) m = fmaxf(m, partial[i]);
    out[0] = m;
  }
}
__global__ void __launch_bounds__(256) expsum_partial_kernel(const float* __restrict__ x, int64_t n,
                                                             const float* __restrict__ mx,
                                                             double* __restrict__ partial) {
  __shared__ double sh[8];
  const float m = mx[0];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += (double)expf(x[i] - m);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void softmax_write_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ mx,
                                     const float* __restrict__ sum, float* __restrict__ out) {
  const float m = mx[0], s = sum[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(expf(x[i] - m), s);
}

__global__ void exponential_kernel(float* __restrict__ out, int64_t n, uint64_t seed) {
  const int64_t n2 = (n + 1) >> 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t)i));
    const float u0 = ((float)((uint32_t)r >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float u1 = ((float)((uint32_t)(r >> 32) >> 9) + 0.5f) * (1.0f / 8388608.0f);
    out[2 * i] = -logf(u0);
    if (2 * i + 1 < n) out[2 * i + 1] = -logf(u1);
  }
}

// noise for the edges with GLOBAL ids gid[i]: the value exponential_kernel gives element gid[i] of a contiguous draw,
// so a destination-sharded edge list sees exactly the noise of the unsharded one (identical keys for any rank count)
__global__ void exponential_ids_kernel(float* __restrict__ out, const int64_t* __restrict__ gid, int64_t n,
                                       uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t g = (uint64_t)gid[i];
    const uint64_t r = splitmix64(seed ^ splitmix64(g >> 1));
    const uint32_t w = (g & 1) ? (uint32_t)(r >> 32) : (uint32_t)r;
    out[i] = -logf(((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f));
  }
}

__global__ void gather_selected_kernel(const float* __restrict__ p, const float* __restrict__ prob,
                                       const int32_t* __restrict__ sel, int64_t q, float c_p, float c_prob,
                                       int mode, const float* __restrict__ S, float* __restrict__ p_sel,
                                       float* __restrict__ w_st) {
  const float S_eff = (w_st && mode != SGS_SAMPLE_RAW) ? __fadd_rn(S[0], 1e-12f) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = sel[i];
    const float pe = p[e];
    if (p_sel) p_sel[i] = pe;
    if (w_st) {
      float s = pe;
      if (mode != SGS_SAMPLE_RAW) {
        s = __fdiv_rn(pe, S_eff);
        if (mode == SGS_SAMPLE_TRAIN) s = __fadd_rn(__fmul_rn(c_p, s), __fmul_rn(c_prob, prob[e]));
      }
      const float st = __fadd_rn(__fsub_rn(1.0f, s), s);  // (one_hot - s) + s, sampling.py:137
      w_st[i] = fminf(fmaxf(__fmul_rn(pe, st), 0.f), 1.f);
    }
  }
}

__global__ void scatter_selected_kernel(const float* __restrict__ src, const int32_t* __restrict__ sel,
                                        int64_t q, float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q; i += (int64_t)gridDim.x * blockDim.x)
    dst[sel[i]] += src[i];
}

static inline int stream_grid(int64_t n_items, int block, int per_sm = 4) {
  int64_t g = ceil_div(n_items, block);
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sgs

using namespace sgs;

extern "C" {

int32_t sgs_sum_f32(const float* p, int64_t n, float* S_out, void* ws, size_t ws_bytes, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0 && p && S_out && ws, "bad arguments");
  SGS_CHECK_ARG(((uintptr_t)p & 15) == 0, "p must be 16-byte aligned");
  if (ws_bytes < kRedBlocks * sizeof(double)) { set_error("sgs_sum_f32: workspace too small"); return SGS_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  sum_partial_kernel<<<kRedBlocks, 256, 0, st>>>(p, n, (double*)ws);
  SGS_LAUNCH_CHECK();
  sum_final_kernel<<<1, 32, 0, st>>>((const double*)ws, kRedBlocks, S_out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_softmax_f32(const float* in, int64_t n, float* out, void* ws, size_t ws_bytes, sgs_stream_t stream) {
  SGS_CHECK_ARG(n > 0 && in && out && ws, "bad arguments");
  if (ws_bytes < 16384) { set_error("sgs_softmax_f32: workspace too small"); return SGS_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  // ws layout: double partial[1024] | float fpartial[1024] | float mx | float sum
  double* dpart = (double*)ws;
  float* fpart = (float*)(dpart + kRedBlocks);
  float* mx = fpart + kRedBlocks;
  float* sm = mx + 1;
  max_partial_kernel<<<kRedBlocks, 256, 0, st>>>(in, n, fpart);
  SGS_LAUNCH_CHECK();
  max_final_kernel<<<1, 32, 0, st>>>(fpart, kRedBlocks, mx);
  SGS_LAUNCH_CHECK();
  expsum_partial_kernel<<<kRedBlocks, 256, 0, st>>>(in, n, mx, dpart);
  SGS_LAUNCH_CHECK();
  sum_final_kernel<<<1, 32, 0, st>>>(dpart, kRedBlocks, sm);
  SGS_LAUNCH_CHECK();
  softmax_write_kernel<<<stream_grid(n, 256, 8), 256, 0, st>>>(in, n, mx, sm, out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_exponential_f32(float* noise, int64_t n, uint64_t seed, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0 && (n == 0 || noise), "bad arguments");
  if (n == 0) return SGS_OK;
  exponential_kernel<<<stream_grid((n + 1) / 2, 256, 8), 256, 0, as_stream(stream)>>>(noise, n, seed);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_exponential_ids_f32(float* noise, const int64_t* gid, int64_t n, uint64_t seed, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0 && (n == 0 || (noise && gid)), "bad arguments");
  if (n == 0) return SGS_OK;
  exponential_ids_kernel<<<stream_grid(n, 256, 8), 256, 0, as_stream(stream)>>>(noise, gid, n, seed);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_topq_keys(const float* p, const float* prob, const float* noise, int64_t E, float one_minus_coef,
                      float coef, int32_t mode, const float* S, uint32_t* keys, int64_t* hist, int64_t* state,
                      sgs_stream_t stream) {
  SGS_CHECK_ARG(E > 0 && E < (1ll << 31), "E out of range");
  SGS_CHECK_ARG(p && noise && keys && hist && state, "null pointer");
  SGS_CHECK_ARG(mode == SGS_SAMPLE_RAW || S, "S required");
  SGS_CHECK_ARG(mode != SGS_SAMPLE_TRAIN || prob, "prob required in train mode");
  SGS_CHECK_ARG((((uintptr_t)p | (uintptr_t)noise | (uintptr_t)keys | (uintptr_t)prob) & 15) == 0,
                "p/prob/noise/keys must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(hist, 0, kBins * sizeof(int64_t), st));
  SGS_CUDA(cudaMemsetAsync(state, 0, 8 * sizeof(int64_t), st));
  topq_keys_kernel<<<stream_grid(E / 4 + 1, 512, 4), 512, 0, st>>>(p, prob, noise, E, one_minus_coef, coef, mode, S,
                                                                   keys, (unsigned long long*)hist,
                                                                   (long long*)state);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_topq_find(int64_t* hist, int64_t* state, int64_t k_total, int32_t level, sgs_stream_t stream) {
  SGS_CHECK_ARG(hist && state && level >= 0 && level <= 2 && k_total >= 1, "bad arguments");
  topq_find_kernel<<<1, 1024, 0, as_stream(stream)>>>((unsigned long long*)hist, (long long*)state, k_total, level,
                                                        nullptr);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_topq_hist(const uint32_t* keys, int64_t E, int64_t* hist, const int64_t* state, int32_t level,
                      sgs_stream_t stream) {
  SGS_CHECK_ARG(keys && hist && state && (level == 1 || level == 2) && E > 0, "bad arguments");
  topq_hist_kernel<<<stream_grid(E / 4 + 1, 512, 4), 512, 0, as_stream(stream)>>>(
      keys, E, (unsigned long long*)hist, (const long long*)state, level, nullptr);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

static size_t compact_bytes(int64_t E) {
  const int64_t nb = ceil_div(E > 0 ? E : 1, kChunk);
  return (size_t)((2 * nb * sizeof(int32_t) + 255) & ~(size_t)255) + 256;
}
size_t sgs_topq_workspace_bytes(int64_t E) { return compact_bytes(E) + kWinScratchBytes; }

int32_t sgs_topq_compact(const uint32_t* keys, int64_t E, const int64_t* state, int64_t tie_skip, int32_t* sel,
                         int64_t q_cap, uint8_t* mask, int64_t* n_sel_out, void* ws, size_t ws_bytes,
                         sgs_stream_t stream) {
  SGS_CHECK_ARG(keys && state && ws && E > 0 && tie_skip >= 0, "bad arguments");
  if (ws_bytes < sgs_topq_workspace_bytes(E)) { set_error("sgs_topq_compact: workspace too small"); return SGS_E_WORKSPACE; }
  cudaStream_t st = as_stream(stream);
  const int nb = (int)ceil_div(E, kChunk);
  int* blk_gt = (int*)ws;
  int* blk_eq = blk_gt + nb;
  topq_count_kernel<<<nb, kChunkThreads, 0, st>>>(keys, E, (const long long*)state, blk_gt, blk_eq);
  SGS_LAUNCH_CHECK();
  topq_scan_kernel<<<1, 1024, 0, st>>>(blk_gt, blk_eq, nb);
  SGS_LAUNCH_CHECK();
  topq_write_kernel<<<nb, kChunkThreads, 0, st>>>(keys, E, (const long long*)state, tie_skip, blk_gt, blk_eq, sel,
                                                  q_cap, mask, (long long*)n_sel_out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_sample_topq(const float* p, const float* prob, const float* noise, int64_t E, int64_t q,
                        float one_minus_coef, float coef, int32_t mode, const float* S, uint32_t* keys,
                        int32_t* sel, uint8_t* mask, int64_t* state, void* ws, size_t ws_bytes,
                        sgs_stream_t stream) {
  SGS_CHECK_ARG(q >= 1 && q <= E, "cannot sample q > E (or q < 1) edges without replacement");
  SGS_CHECK_ARG(ws && ws_bytes >= sgs_topq_workspace_bytes(E) + kBins * sizeof(int64_t), "workspace too small");
  int64_t* hist = (int64_t*)ws;
  void* ws2 = (char*)ws + kBins * sizeof(int64_t);
  size_t ws2_bytes = ws_bytes - kBins * sizeof(int64_t);
  int32_t rc;
  // debug / test knobs (read on every call): SGS_TOPQ_FAST_MIN_E moves the size threshold of the fast path,
  // SGS_TOPQ_FORCE_MISS=1 makes the predicted window miss so that the fallback passes run
  const char* env_min = getenv("SGS_TOPQ_FAST_MIN_E");
  const int64_t fast_min_e = env_min ? (int64_t)atoll(env_min) : (int64_t)1 << 20;
  if (E >= fast_min_e && E >= 4096) {
    // ---- sampled-window fast path (see above); scratch = the tail of the workspace ----
    SGS_CHECK_ARG(p && noise && keys && state, "null pointer");
    SGS_CHECK_ARG(mode == SGS_SAMPLE_RAW || S, "S required");
    SGS_CHECK_ARG(mode != SGS_SAMPLE_TRAIN || prob, "prob required in train mode");
    SGS_CHECK_ARG((((uintptr_t)p | (uintptr_t)noise | (uintptr_t)keys | (uintptr_t)prob) & 15) == 0,
                  "p/prob/noise/keys must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    unsigned long long* shist = (unsigned long long*)((char*)ws2 + compact_bytes(E));
    unsigned long long* whist = shist + kBins;
    unsigned long long* misc = whist + kWinHist;
    const uint32_t* gate = (const uint32_t*)(misc + 1);
    const char* env_miss = getenv("SGS_TOPQ_FORCE_MISS");
    const int force_miss = env_miss ? atoi(env_miss) : 0;
    SGS_CUDA(cudaMemsetAsync(hist, 0, kBins * sizeof(int64_t), st));
    SGS_CUDA(cudaMemsetAsync(state, 0, 8 * sizeof(int64_t), st));
    SGS_CUDA(cudaMemsetAsync(shist, 0, kWinScratchBytes, st));
    const int stride = 64;
    topq_sample_kernel<<<stream_grid(E / 4 / stride + 1, 256, 4), 256, 0, st>>>(p, prob, noise, E, one_minus_coef,
                                                                                coef, mode, S, stride, shist, misc);
    SGS_LAUNCH_CHECK();
    topq_predict_kernel<<<1, 1024, 0, st>>>(shist, misc, q, E, force_miss);
    SGS_LAUNCH_CHECK();
    topq_keys_window_kernel<<<stream_grid(E / 4 + 1, 512, 4), 512, 0, st>>>(p, prob, noise, E, one_minus_coef, coef,
                                                                            mode, S, keys, whist, misc,
                                                                            (long long*)state);
    SGS_LAUNCH_CHECK();
    topq_find_window_kernel<<<1, 1024, 0, st>>>(whist, misc, (long long*)state, q);
    SGS_LAUNCH_CHECK();
    // classic level-0 / level-1 passes over the stored keys: exit immediately unless the window missed tau
    const int hg = stream_grid(E / 4 + 1, 512, 4);
    topq_hist_kernel<<<hg, 512, 0, st>>>(keys, E, (unsigned long long*)hist, (const long long*)state, 0, gate);
    SGS_LAUNCH_CHECK();
    topq_find_kernel<<<1, 1024, 0, st>>>((unsigned long long*)hist, (long long*)state, q, 0, gate);
    SGS_LAUNCH_CHECK();
    topq_hist_kernel<<<hg, 512, 0, st>>>(keys, E, (unsigned long long*)hist, (const long long*)state, 1, gate);
    SGS_LAUNCH_CHECK();
    topq_find_kernel<<<1, 1024, 0, st>>>((unsigned long long*)hist, (long long*)state, q, 1, gate);
    SGS_LAUNCH_CHECK();
  } else {
    if ((rc = sgs_topq_keys(p, prob, noise, E, one_minus_coef, coef, mode, S, keys, hist, state, stream))) return rc;
    if ((rc = sgs_topq_find(hist, state, q, 0, stream))) return rc;
    if ((rc = sgs_topq_hist(keys, E, hist, state, 1, stream))) return rc;
    if ((rc = sgs_topq_find(hist, state, q, 1, stream))) return rc;
  }
  if ((rc = sgs_topq_hist(keys, E, hist, state, 2, stream))) return rc;
  if ((rc = sgs_topq_find(hist, state, q, 2, stream))) return rc;
  return sgs_topq_compact(keys, E, state, 0, sel, q, mask, (int64_t*)state + 7, ws2, ws2_bytes, stream);
}

int32_t sgs_gather_selected(const float* p, const float* prob, const int32_t* sel, int64_t q, float one_minus_coef,
                            float coef, int32_t mode, const float* S, float* p_sel, float* w_st,
                            sgs_stream_t stream) {
  SGS_CHECK_ARG(q >= 0, "negative q");
  if (q == 0) return SGS_OK;
  SGS_CHECK_ARG(p && sel && (p_sel || w_st), "null pointer");
  SGS_CHECK_ARG(!w_st || mode == SGS_SAMPLE_RAW || S, "S required for straight-through weights");
  SGS_CHECK_ARG(!w_st || mode != SGS_SAMPLE_TRAIN || prob, "prob required");
  gather_selected_kernel<<<stream_grid(q, 256, 8), 256, 0, as_stream(stream)>>>(p, prob, sel, q, one_minus_coef,
                                                                               coef, mode, S, p_sel, w_st);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_scatter_selected(const float* src, const int32_t* sel, int64_t q, float* dst, sgs_stream_t stream) {
  SGS_CHECK_ARG(q >= 0, "negative q");
  if (q == 0) return SGS_OK;
  SGS_CHECK_ARG(src && sel && dst, "null pointer");
  scatter_selected_kernel<<<stream_grid(q, 256, 8), 256, 0, as_stream(stream)>>>(src, sel, q, dst);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}
}
