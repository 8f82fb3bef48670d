// K3: gcn_norm, CSR SpMM (forward, and backward via the by-source CSR), activation backward,
// bias gradient and the edge-weight gradient (SDDMM + per-node sums).
// All segment reductions are atomic-free: one warp owns one CSR row.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"

namespace sgs {

constexpr int kWarpsPerBlock = 8;
constexpr int kBlock = kWarpsPerBlock * 32;

// Sharded SpMM (one graph split by destination range over the GPUs, peer.cu): only the rows [row_lo, row_hi) this
// rank owns are computed, and every finished row is also stored into the same [N, D] buffer of every peer
// (bases[g] + off floats: symmetric arena over NVLink) -- the slab all-gather happens inside the epilogue.
struct SpmmPeers {
  const uint64_t* bases;   // device array of `world` arena base addresses, or nullptr (no peer stores)
  int world, rank;
  int64_t off;             // element offset of the destination buffer inside every arena
  int64_t row_lo, row_hi;  // owned row range
  int phases;              // bit 0: block-cooperative hub rows, bit 1: one warp per remaining row (3 = both)
  float* out2;             // fp16-table form, pair mode: columns [split, D) go to out2 (both outputs [N, split])
  int split;               // 0 = single output [N, D]
};
__device__ __forceinline__ void peer_store4(const SpmmPeers& pe, int64_t idx, const float4& v) {
  for (int g = 0; g < pe.world; ++g)
    if (g != pe.rank) reinterpret_cast<float4*>(pe.bases[g] + (uint64_t)(pe.off + idx) * sizeof(float))[0] = v;
}
__device__ __forceinline__ void peer_store1(const SpmmPeers& pe, int64_t idx, float v) {
  for (int g = 0; g < pe.world; ++g)
    if (g != pe.rank) reinterpret_cast<float*>(pe.bases[g] + (uint64_t)(pe.off + idx) * sizeof(float))[0] = v;
}

// ---------------------------------------------------------------------------------------
// gcn_norm phase 1: weighted in-degree with "remaining" self loops.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock)
gcn_degree_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                  const int32_t* __restrict__ nbr, const float* __restrict__ w, int64_t N,
                  float* __restrict__ deg, float* __restrict__ dis, float* __restrict__ loopw) {
  int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  int64_t nrows_step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; row < N; row += nrows_step) {
    int beg = rowptr[row], end = rowptr[row + 1];
    float acc = 0.f;
    int loop_e = -1;  // highest edge id among input self loops (CPU "last write wins")
    for (int i = beg + lane; i < end; i += 32) {
      int s = nbr[i];
      int e = perm[i];
      if (s == (int)row) {
        loop_e = max(loop_e, e);
      } else {
        acc += w ? w[e] : 1.0f;
      }
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loop_e = max(loop_e, __shfl_xor_sync(0xffffffffu, loop_e, o));
    if (lane == 0) {
      float lw = (loop_e >= 0 && w) ? w[loop_e] : 1.0f;
      float d = acc + lw;
      float r = 1.0f / sqrtf(d);
      if (isinf(r)) r = 0.f;
      deg[row] = d;
      dis[row] = r;
      loopw[row] = lw;
    }
  }
}

// phase 2: what[i] = (dis[nbr] * w) * dis[row]   (0 for input self loops)
__global__ void __launch_bounds__(kBlock)
gcn_what_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                const int32_t* __restrict__ nbr, const float* __restrict__ w,
                const float* __restrict__ dis, int64_t N, float* __restrict__ what) {
  int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; row < N; row += step) {
    int beg = rowptr[row], end = rowptr[row + 1];
    float dr = dis[row];
    // four 32-edge groups per iteration: the dependent gathers dis[nbr], w[perm] of all four are in flight together
    int i = beg + lane;
    for (; i + 96 < end; i += 128) {
      int s[4], e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = nbr[i + 32 * u];
        e[u] = w ? perm[i + 32 * u] : 0;
      }
      float ds[4], we[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ds[u] = dis[s[u]];
        we[u] = w ? w[e[u]] : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) what[i + 32 * u] = (s[u] == (int)row) ? 0.f : (ds[u] * we[u]) * dr;
    }
    for (; i < end; i += 32) {
      int s = nbr[i];
      float we = w ? w[perm[i]] : 1.0f;
      what[i] = (s == (int)row) ? 0.f : (dis[s] * we) * dr;
    }
  }
}

// ---------------------------------------------------------------------------------------
// SpMM: out[r, cols] = act(sum_i what[i] * h[nbr[i], cols] + selfw * h[r, cols] + bias)
// A warp owns a row; each lane owns K chunks of VEC consecutive columns:
//   col(k, lane) = col0 + (k * 32 + lane) * VEC.
// (nbr, what) pairs are fetched 32 at a time (coalesced) and broadcast with shuffles so the
// feature-row loads of consecutive neighbours are independent and stay in flight together.
// ---------------------------------------------------------------------------------------
// Rows with more than kHeavyDeg edges (power-law hubs; listed first in `order`, their number at order[N])
// are processed by a whole block -- the 8 warps take interleaved 32-edge groups and their partial rows
// are summed through shared memory in a fixed order -- so that one hub cannot become the kernel's tail.
template <int VEC, int K>
struct SpmmRow {
  float acc[K][VEC];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
  }

  // accumulate the edges [beg, end) taking 32-edge groups base = beg + 32*first, stride 32*nstep
  __device__ __forceinline__ void gather(const int32_t* __restrict__ nbr, const float* __restrict__ what,
                                         const float* __restrict__ h, int D, int col0, int lane, int beg, int end,
                                         int first, int nstep) {
    for (int base = beg + 32 * first; base < end; base += 32 * nstep) {
      int my_n = 0;
      float my_w = 0.f;
      if (base + lane < end) {
        my_n = nbr[base + lane];
        my_w = what[base + lane];
      }
      const int cnt = min(32, end - base);
#pragma unroll 8
      for (int j = 0; j < cnt; ++j) {
        const int n = __shfl_sync(0xffffffffu, my_n, j);
        const float wv = __shfl_sync(0xffffffffu, my_w, j);
        const float* hp = h + (int64_t)n * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = col0 + (k * 32 + lane) * VEC;
          if (c < D) {
            if (VEC == 4) {
              const float4 x = *reinterpret_cast<const float4*>(hp + c);
              acc[k][0] = fmaf(wv, x.x, acc[k][0]);
              acc[k][1] = fmaf(wv, x.y, acc[k][1]);
              acc[k][2] = fmaf(wv, x.z, acc[k][2]);
              acc[k][3] = fmaf(wv, x.w, acc[k][3]);
            } else {
              acc[k][0] = fmaf(wv, hp[c], acc[k][0]);
            }
          }
        }
      }
    }
  }

  // self loop, bias, activation, store
  __device__ __forceinline__ void finish(int64_t row, const float* __restrict__ dis, const float* __restrict__ loopw,
                                         const float* __restrict__ h, int D, int col0, int lane,
                                         const float* __restrict__ bias, float* __restrict__ out, int flags,
                                         float scale, uint32_t thr, uint64_t seed, const SpmmPeers& pe) {
    float selfw = 0.f;
    if (dis) {
      const float d = dis[row];
      selfw = d * d * (loopw ? loopw[row] : 1.0f);
    }
    const float* hr = h + row * D;
    float* orow = out + row * D;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = col0 + (k * 32 + lane) * VEC;
      if (c < D) {
        float v[VEC];
#pragma unroll
        for (int t = 0; t < VEC; ++t) {
          v[t] = acc[k][t];
          if (dis) v[t] = fmaf(selfw, hr[c + t], v[t]);
          if (flags & SGS_SPMM_ADD_ROOT) v[t] += orow[c + t];   // root term (SAGEConv), inside the activation
          if (bias) v[t] += bias[c + t];
          if (flags & SGS_SPMM_RELU) v[t] = fmaxf(v[t], 0.f);
        }
        if (flags & SGS_SPMM_DROPOUT) {
          const uint64_t bits = dropout_bits(seed, (uint64_t)row, (uint32_t)(c >> 2));
#pragma unroll
          for (int t = 0; t < VEC; ++t)
            v[t] = dropout_keep(bits, (c + t) & 3, thr) ? v[t] * scale : 0.f;
        }
        if (flags & SGS_SPMM_ACCUM) {
#pragma unroll
          for (int t = 0; t < VEC; ++t) v[t] += orow[c + t];
        }
        if (VEC == 4) {
          const float4 o = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(orow + c) = o;
          if (pe.bases) peer_store4(pe, row * D + c, o);
        } else {
          orow[c] = v[0];
          if (pe.bases) peer_store1(pe, row * D + c, v[0]);
        }
      }
    }
  }
};

template <int VEC, int K>
__global__ void __launch_bounds__(kBlock)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
            const float* __restrict__ what, const float* __restrict__ dis,
            const float* __restrict__ loopw, const float* __restrict__ h, int64_t N, int D,
            const float* __restrict__ bias, float* __restrict__ out, int flags, float p_drop,
            uint64_t seed, const int32_t* __restrict__ order, const SpmmPeers pe) {
  __shared__ float red[kWarpsPerBlock - 1][32 * VEC * K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = blockIdx.y * (32 * VEC * K);
  const uint32_t thr = dropout_threshold(p_drop);
  const float scale = (flags & SGS_SPMM_DROPOUT) ? 1.0f / (1.0f - p_drop) : 1.0f;
  const int n_heavy = order ? order[N] : 0;
  SpmmRow<VEC, K> r;

  // phase 1: hub rows, one block per row
  for (int hidx = blockIdx.x; hidx < n_heavy; hidx += gridDim.x) {
    const int64_t row = order[hidx];
    if (row < pe.row_lo || row >= pe.row_hi) continue;   // block-uniform
    r.clear();
    r.gather(nbr, what, h, D, col0, lane, rowptr[row], rowptr[row + 1], warp, kWarpsPerBlock);
    if (warp > 0) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) red[warp - 1][(k * 32 + lane) * VEC + v] = r.acc[k][v];
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 0; w < kWarpsPerBlock - 1; ++w)
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int v = 0; v < VEC; ++v) r.acc[k][v] += red[w][(k * 32 + lane) * VEC + v];
      r.finish(row, dis, loopw, h, D, col0, lane, bias, out, flags, scale, thr, seed, pe);
    }
    __syncthreads();
  }

  // phase 2: one warp per row; `order` lists the rows heaviest-first, dealing them round-robin to the warps
  // balances what is left of the power-law tail
  int64_t idx = (int64_t)n_heavy + (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; idx < N; idx += step) {
    const int64_t row = order ? order[idx] : idx;
    if (row < pe.row_lo || row >= pe.row_hi) continue;
    r.clear();
    r.gather(nbr, what, h, D, col0, lane, rowptr[row], rowptr[row + 1], 0, 1);
    r.finish(row, dis, loopw, h, D, col0, lane, bias, out, flags, scale, thr, seed, pe);
  }
}


// ---------------------------------------------------------------------------------------
// 16-bit gather tables.  The D = 256 SpMM / SDDMM are bound by the row gathers (ncu r01: 1 KB per edge from a 239 MB
// fp32 table, L2 hit rate 50 %, DRAM at 58 %).  With the table in fp16 -- 119 MB at Reddit scale, i.e. L2-resident --
// a gathered row is 512 B and (nearly) never leaves L2.  Accumulation stays fp32; the table carries a power-of-two
// scale (gradient tables: max |v| -> [8192, 16384)) that is divided out in the epilogue.
// ---------------------------------------------------------------------------------------
__global__ void table_absmax_kernel(const float* __restrict__ x, int64_t n4, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(x) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
// scale[0] = S (power of two), scale[1] = 1 / S; S = 1 when no absmax is given (activations: |v| << 65504)
__global__ void table_scale_kernel(const float* __restrict__ absmax, float* __restrict__ scale) {
  float S = 1.0f;
  if (absmax) {
    const float m = absmax[0];
    if (m > 0.f && !isinf(m) && !isnan(m)) S = exp2f(floorf(log2f(16384.0f / m)));
  }
  scale[0] = S;
  scale[1] = 1.0f / S;
}
__global__ void table_convert_kernel(const float* __restrict__ in, int64_t n8, const float* __restrict__ scale,
                                     uint4* __restrict__ outp) {
  const float S = scale[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = ld_stream_f4(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = ld_stream_f4(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    auto pk = [&](float x, float y) {
      const __half2 v = __floats2half2_rn(fminf(fmaxf(x * S, -65504.f), 65504.f), fminf(fmaxf(y * S, -65504.f), 65504.f));
      return *reinterpret_cast<const uint32_t*>(&v);
    };
    outp[i] = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(b.x, b.y), pk(b.z, b.w));
  }
}

__device__ __forceinline__ void h8_to_float(const uint4& u, float* f) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// SpMM over an fp16 table [N, D], D % 8 == 0, D <= 256 * K: lane owns the 8 columns 8 * (k * 32 + lane).
template <int K>
struct SpmmRowH {
  float acc[K][8];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < 8; ++v) acc[k][v] = 0.f;
  }
  __device__ __forceinline__ void gather(const int32_t* __restrict__ nbr, const float* __restrict__ what,
                                         const __half* __restrict__ h, int D, int lane, int beg, int end, int first,
                                         int nstep) {
    for (int base = beg + 32 * first; base < end; base += 32 * nstep) {
      int my_n = 0;
      float my_w = 0.f;
      if (base + lane < end) {
        my_n = nbr[base + lane];
        my_w = what[base + lane];
      }
      const int cnt = min(32, end - base);
#pragma unroll 8
      for (int j = 0; j < cnt; ++j) {
        const int n = __shfl_sync(0xffffffffu, my_n, j);
        const float wv = __shfl_sync(0xffffffffu, my_w, j);
        const __half* hp = h + (int64_t)n * D;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = (k * 32 + lane) * 8;
          if (c < D) {
            float x[8];
            h8_to_float(*reinterpret_cast<const uint4*>(hp + c), x);
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[k][t] = fmaf(wv, x[t], acc[k][t]);
          }
        }
      }
    }
  }
  // K == 1 (D <= 256): software-pipelined form.  A ring of RING row loads stays in flight -- the load of edge
  // j + RING is issued before the FMAs of edge j, so there is no drain bubble between 8-edge batches -- and the
  // (nbr, what) pairs of the NEXT 32-edge group are fetched while this group is processed.  (The plain form above pays
  // one exposed L2 latency per batch and per group: ~20 per average row of ~100 edges.)
  __device__ __forceinline__ void gather_pipelined(const int32_t* __restrict__ nbr, const float* __restrict__ what,
                                                   const __half* __restrict__ h, int D, int lane, int beg, int end,
                                                   int first, int nstep) {
    static_assert(K == 1, "pipelined gather: one 8-column chunk per lane");
    constexpr int RING = 8;
    const int c = lane * 8;
    const bool act = c < D;
    int base = beg + 32 * first;
    if (base >= end) return;
    int nn = 0;
    float nw = 0.f;
    if (base + lane < end) {
      nn = nbr[base + lane];
      nw = what[base + lane];
    }
    while (base < end) {
      const int my_n = nn;
      const float my_w = nw;
      const int nbase = base + 32 * nstep;
      nn = 0;
      nw = 0.f;
      if (nbase + lane < end) {   // in flight during this group
        nn = nbr[nbase + lane];
        nw = what[nbase + lane];
      }
      const int cnt = min(32, end - base);
      uint4 ring[RING];
      // (loads are never conditional on the edge count: a predicated refill compiles to "load into a temporary,
      // then select", and the select waits for the load -- ncu r02a: long-scoreboard stalls on exactly those MOVs.
      // Slots past the group's end re-load its last row -- an L1 hit -- and meet a zero weight.)
      const int last = cnt - 1;
#pragma unroll
      for (int j = 0; j < RING; ++j) {
        const int n = __shfl_sync(0xffffffffu, my_n, min(j, last));
        ring[j] = *reinterpret_cast<const uint4*>(h + (int64_t)n * D + (act ? c : 0));
      }
#pragma unroll
      for (int jj = 0; jj < 32; jj += RING) {
        if (jj < cnt) {   // warp-uniform
#pragma unroll
          for (int r = 0; r < RING; ++r) {
            const int j = jj + r;
            const uint4 v = ring[r];
            const float wv = __shfl_sync(0xffffffffu, my_w, j);                  // 0 for j >= cnt
            const int n2 = __shfl_sync(0xffffffffu, my_n, min(j + RING, last));
            ring[r] = *reinterpret_cast<const uint4*>(h + (int64_t)n2 * D + (act ? c : 0));
            float x[8];
            h8_to_float(v, x);
#pragma unroll
            for (int t = 0; t < 8; ++t) acc[0][t] = fmaf(wv, x[t], acc[0][t]);
          }
        }
      }
      base = nbase;
    }
  }
  __device__ __forceinline__ void finish(int64_t row, const float* __restrict__ dis, const float* __restrict__ loopw,
                                         const __half* __restrict__ h, int D, int lane, float inv_scale,
                                         const float* __restrict__ bias, float* __restrict__ out, int flags,
                                         float scale, uint32_t thr, uint64_t seed, const SpmmPeers& pe) {
    float selfw = 0.f;
    if (dis) {
      const float d = dis[row];
      selfw = d * d * (loopw ? loopw[row] : 1.0f);
    }
    const __half* hr = h + row * D;
    float* orow = out + row * D;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        if (pe.split) {   // pair mode: two [N, split] outputs (no ADD_ROOT / ACCUM / peer stores in this mode)
          orow = (c < pe.split ? out + row * pe.split : pe.out2 + row * pe.split - pe.split);
        }
        float v[8], self[8];
        if (dis) h8_to_float(*reinterpret_cast<const uint4*>(hr + c), self);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          v[t] = acc[k][t];
          if (dis) v[t] = fmaf(selfw, self[t], v[t]);
          v[t] *= inv_scale;
          if (flags & SGS_SPMM_ADD_ROOT) v[t] += orow[c + t];
          if (bias) v[t] += bias[c + t];
          if (flags & SGS_SPMM_RELU) v[t] = fmaxf(v[t], 0.f);
        }
        if (flags & SGS_SPMM_DROPOUT) {
#pragma unroll
          for (int g4 = 0; g4 < 2; ++g4) {
            const uint64_t bits = dropout_bits(seed, (uint64_t)row, (uint32_t)((c >> 2) + g4));
#pragma unroll
            for (int t = 0; t < 4; ++t) v[4 * g4 + t] = dropout_keep(bits, t, thr) ? v[4 * g4 + t] * scale : 0.f;
          }
        }
        if (flags & SGS_SPMM_ACCUM) {
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] += orow[c + t];
        }
        const float4 o0 = make_float4(v[0], v[1], v[2], v[3]), o1 = make_float4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<float4*>(orow + c) = o0;
        *reinterpret_cast<float4*>(orow + c + 4) = o1;
        if (pe.bases) {
          peer_store4(pe, row * D + c, o0);
          peer_store4(pe, row * D + c + 4, o1);
        }
      }
    }
  }
};

template <int K, bool PIPE>
__global__ void __launch_bounds__(kBlock, PIPE ? 2 : 4)
spmm_h16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr, const float* __restrict__ what,
                const float* __restrict__ dis, const float* __restrict__ loopw, const __half* __restrict__ h,
                const float* __restrict__ tscale, int64_t N, int D, const float* __restrict__ bias,
                float* __restrict__ out, int flags, float p_drop, uint64_t seed, const int32_t* __restrict__ order,
                const SpmmPeers pe) {
  __shared__ float red[kWarpsPerBlock - 1][32 * 8 * K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t thr = dropout_threshold(p_drop);
  const float scale = (flags & SGS_SPMM_DROPOUT) ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float inv_scale = tscale[1];
  const int n_heavy = order ? order[N] : 0;
  SpmmRowH<K> r;
  auto do_gather = [&](int beg, int end, int first, int nstep) {
    if constexpr (K == 1 && PIPE) r.gather_pipelined(nbr, what, h, D, lane, beg, end, first, nstep);
    else r.gather(nbr, what, h, D, lane, beg, end, first, nstep);
  };
  for (int hidx = blockIdx.x; hidx < n_heavy && (pe.phases & 1); hidx += gridDim.x) {
    const int64_t row = order[hidx];
    if (row < pe.row_lo || row >= pe.row_hi) continue;   // block-uniform
    r.clear();
    do_gather(rowptr[row], rowptr[row + 1], warp, kWarpsPerBlock);
    if (warp > 0) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int v = 0; v < 8; ++v) red[warp - 1][(k * 32 + lane) * 8 + v] = r.acc[k][v];
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 0; w < kWarpsPerBlock - 1; ++w)
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int v = 0; v < 8; ++v) r.acc[k][v] += red[w][(k * 32 + lane) * 8 + v];
      r.finish(row, dis, loopw, h, D, lane, inv_scale, bias, out, flags, scale, thr, seed, pe);
    }
    __syncthreads();
  }
  // one warp per row, rows dealt heaviest-first.  The row id and the extent of the NEXT row are fetched while the
  // current row is gathered (two dependent loads -- order[], rowptr[] -- off the critical path of every row).
  if (!(pe.phases & 2)) return;
  const int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  int64_t idx = (int64_t)n_heavy + (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  auto row_of = [&](int64_t i) -> int64_t { return i < N ? (order ? (int64_t)order[i] : i) : -1; };
  int64_t row = row_of(idx), row1 = row_of(idx + step);
  int beg = 0, end = 0;
  if (row >= 0) {
    beg = rowptr[row];
    end = rowptr[row + 1];
  }
  for (; idx < N; idx += step) {
    const int64_t row2 = row_of(idx + 2 * step);
    int beg1 = 0, end1 = 0;
    if (row1 >= 0) {
      beg1 = rowptr[row1];
      end1 = rowptr[row1 + 1];
    }
    if (row >= pe.row_lo && row < pe.row_hi) {
      r.clear();
      do_gather(beg, end, 0, 1);
      r.finish(row, dis, loopw, h, D, lane, inv_scale, bias, out, flags, scale, thr, seed, pe);
    }
    row = row1;
    row1 = row2;
    beg = beg1;
    end = end1;
  }
}

__global__ void act_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ out, int64_t n,
                               float scale, float* __restrict__ gin) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) gin[i] = out[i] > 0.f ? gout[i] * scale : 0.f;
}

__global__ void colsum_kernel(const float* __restrict__ G, int64_t N, int D, float* __restrict__ cs) {
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    for (int64_t r = blockIdx.x; r < N; r += gridDim.x) acc += G[r * D + c];
    atomicAdd(cs + c, acc);
  }
}

// ---------------------------------------------------------------------------------------
// Edge-weight gradient.  Phase A: SDDMM over by-dst rows + dst-side sums.
// ---------------------------------------------------------------------------------------
template <int VEC, int K>
__global__ void __launch_bounds__(kBlock)
edge_grad_sddmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                       const int32_t* __restrict__ nbr, const float* __restrict__ what,
                       const float* __restrict__ G, const float* __restrict__ h,
                       const float* __restrict__ dis, const float* __restrict__ loopw, int64_t N, int D,
                       float* __restrict__ tmp_g, float* __restrict__ tmp_t, float* __restrict__ tmp_a,
                       const int32_t* __restrict__ order) {
  __shared__ float red[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_heavy = order ? order[N] : 0;
  float g_row[K][VEC];

  auto load_row = [&](int64_t row) {
    const float* gr = G + row * D;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * 32 + lane) * VEC;
#pragma unroll
      for (int t = 0; t < VEC; ++t) g_row[k][t] = (c + t < D) ? gr[c + t] : 0.f;
    }
  };
  auto partial_dot = [&](const float* hp) {
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * 32 + lane) * VEC;
      if (c < D) {
        if (VEC == 4) {
          const float4 x = *reinterpret_cast<const float4*>(hp + c);
          d = fmaf(g_row[k][0], x.x, d);
          d = fmaf(g_row[k][1], x.y, d);
          d = fmaf(g_row[k][2], x.z, d);
          d = fmaf(g_row[k][3], x.w, d);
        } else {
          d = fmaf(g_row[k][0], hp[c], d);
        }
      }
    }
    return d;
  };
  auto dot_with = [&](const float* hp) { return warp_sum(partial_dot(hp)); };
  // g_e, t_e of the edges [beg, end) of the current row; returns sum t_e (valid in every lane)
  auto edges = [&](int beg, int end) {
    float tsum = 0.f;
    // eight neighbours per iteration: their row loads and shuffle reductions are independent (the kernel is bound by
    // the latency of the row gathers)
    int i = beg;
    for (; i + 8 <= end; i += 8) {
      const float* hp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) hp[u] = h + (int64_t)nbr[i + u] * D;
      float d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[u] = partial_dot(hp[u]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
      }
      if (lane < 8) {
        float g = d[0];
#pragma unroll
        for (int u = 1; u < 8; ++u) g = lane == u ? d[u] : g;
        const int e = perm[i + lane];
        const float t = g * what[i + lane];
        tmp_g[e] = g;
        tmp_t[e] = t;
        tsum += t;
      }
    }
    for (; i < end; ++i) {
      const float g = dot_with(h + (int64_t)nbr[i] * D);
      if (lane == 0) {
        const int e = perm[i];
        const float t = g * what[i];
        tmp_g[e] = g;
        tmp_t[e] = t;
        tsum += t;
      }
    }
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
    return __shfl_sync(0xffffffffu, tsum, 0);
  };
  auto finish = [&](int64_t row, float tsum) {
    const float gl = dot_with(h + row * D);
    if (lane == 0) {
      const float d = dis[row];
      const float tl = gl * d * d * loopw[row];
      tmp_a[row] = tsum + 2.0f * tl;
    }
  };

  // phase 1: hub rows (first n_heavy entries of `order`), the block's warps take contiguous edge ranges
  for (int hidx = blockIdx.x; hidx < n_heavy; hidx += gridDim.x) {
    const int64_t row = order[hidx];
    load_row(row);
    const int beg = rowptr[row], end = rowptr[row + 1];
    const int per = (((end - beg) + kWarpsPerBlock - 1) / kWarpsPerBlock + 7) & ~7;
    const int b = min(end, beg + warp * per), e = min(end, b + per);
    const float part = edges(b, e);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (warp == 0) {
      float tsum = 0.f;
      for (int w = 0; w < kWarpsPerBlock; ++w) tsum += red[w];
      finish(row, tsum);
    }
    __syncthreads();
  }
  // phase 2: one warp per row
  int64_t idx = (int64_t)n_heavy + (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; idx < N; idx += step) {
    const int64_t row = order ? order[idx] : idx;
    load_row(row);
    finish(row, edges(rowptr[row], rowptr[row + 1]));
  }
}

// SDDMM over an fp16 h table (same role as edge_grad_sddmm_kernel; D % 8 == 0, D <= 256 * K)
template <int K>
__global__ void __launch_bounds__(kBlock)
edge_grad_sddmm_h16_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ perm,
                           const int32_t* __restrict__ nbr, const float* __restrict__ what,
                           const float* __restrict__ G, const __half* __restrict__ h,
                           const float* __restrict__ tscale, const float* __restrict__ dis,
                           const float* __restrict__ loopw, int64_t N, int D, float* __restrict__ tmp_g,
                           float* __restrict__ tmp_t, float* __restrict__ tmp_a, const int32_t* __restrict__ order) {
  __shared__ float red[kWarpsPerBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_heavy = order ? order[N] : 0;
  const float inv_scale = tscale[1];
  float g_row[K][8];

  auto load_row = [&](int64_t row) {
    const float* gr = G + row * D;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        const float4 a = *reinterpret_cast<const float4*>(gr + c);
        const float4 b = *reinterpret_cast<const float4*>(gr + c + 4);
        g_row[k][0] = a.x * inv_scale; g_row[k][1] = a.y * inv_scale; g_row[k][2] = a.z * inv_scale;
        g_row[k][3] = a.w * inv_scale; g_row[k][4] = b.x * inv_scale; g_row[k][5] = b.y * inv_scale;
        g_row[k][6] = b.z * inv_scale; g_row[k][7] = b.w * inv_scale;
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) g_row[k][t] = 0.f;
      }
    }
  };
  auto partial_dot = [&](const __half* hp) {
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = (k * 32 + lane) * 8;
      if (c < D) {
        float x[8];
        h8_to_float(*reinterpret_cast<const uint4*>(hp + c), x);
#pragma unroll
        for (int t = 0; t < 8; ++t) d = fmaf(g_row[k][t], x[t], d);
      }
    }
    return d;
  };
  auto dot_with = [&](const __half* hp) { return warp_sum(partial_dot(hp)); };
  auto edges = [&](int beg, int end) {
    float tsum = 0.f;
    // eight neighbours per iteration: their row loads are all issued before the first dot product is reduced
    // (the kernel is bound by the latency of these gathers, ncu r02a: long-scoreboard stalls), then eight
    // interleaved shuffle reductions
    int i = beg;
    for (; i + 8 <= end; i += 8) {
      const __half* hp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) hp[u] = h + (int64_t)nbr[i + u] * D;
      float d[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) d[u] = partial_dot(hp[u]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] += __shfl_xor_sync(0xffffffffu, d[u], o);
      }
      if (lane < 8) {
        float g = d[0];
#pragma unroll
        for (int u = 1; u < 8; ++u) g = lane == u ? d[u] : g;
        const int e = perm[i + lane];
        const float t = g * what[i + lane];
        tmp_g[e] = g;
        tmp_t[e] = t;
        tsum += t;
      }
    }
    for (; i < end; ++i) {
      const float g = dot_with(h + (int64_t)nbr[i] * D);
      if (lane == 0) {
        const int e = perm[i];
        const float t = g * what[i];
        tmp_g[e] = g;
        tmp_t[e] = t;
        tsum += t;
      }
    }
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
    tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
    return __shfl_sync(0xffffffffu, tsum, 0);
  };
  auto finish = [&](int64_t row, float tsum) {
    const float gl = dot_with(h + row * D);
    if (lane == 0) {
      const float d = dis[row];
      const float tl = gl * d * d * loopw[row];
      tmp_a[row] = tsum + 2.0f * tl;
    }
  };
  for (int hidx = blockIdx.x; hidx < n_heavy; hidx += gridDim.x) {
    const int64_t row = order[hidx];
    load_row(row);
    const int beg = rowptr[row], end = rowptr[row + 1];
    const int per = (((end - beg) + kWarpsPerBlock - 1) / kWarpsPerBlock + 7) & ~7;
    const int b = min(end, beg + warp * per), e = min(end, b + per);
    const float part = edges(b, e);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (warp == 0) {
      float tsum = 0.f;
      for (int w = 0; w < kWarpsPerBlock; ++w) tsum += red[w];
      finish(row, tsum);
    }
    __syncthreads();
  }
  int64_t idx = (int64_t)n_heavy + (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; idx < N; idx += step) {
    const int64_t row = order ? order[idx] : idx;
    load_row(row);
    finish(row, edges(rowptr[row], rowptr[row + 1]));
  }
}

// Phase B: add the source-side sums  A[r] += sum_{e: src_e = r} t_e
__global__ void __launch_bounds__(kBlock)
edge_grad_srcsum_kernel(const int32_t* __restrict__ rowptr_src, const int32_t* __restrict__ perm_src,
                        const float* __restrict__ tmp_t, int64_t N, float* __restrict__ tmp_a) {
  const int lane = threadIdx.x & 31;
  int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * kWarpsPerBlock;
  for (; row < N; row += step) {
    const int beg = rowptr_src[row], end = rowptr_src[row + 1];
    float acc = 0.f;
    for (int i = beg + lane; i < end; i += 32) acc += tmp_t[perm_src[i]];
    acc = warp_sum(acc);
    if (lane == 0) tmp_a[row] += acc;
  }
}

// Phase C: dL/dw_e = g_e dis[r] dis[c] - A_c / (2 deg_c)
__global__ void edge_grad_final_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                       const float* __restrict__ tmp_g, const float* __restrict__ tmp_a,
                                       const float* __restrict__ dis, const float* __restrict__ deg,
                                       int64_t M, float* __restrict__ dw, int accumulate) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < M; e += stride) {
    const int r = src[e], c = dst[e];
    float v = 0.f;
    if (r != c) v = tmp_g[e] * dis[r] * dis[c] - tmp_a[c] / (2.0f * deg[c]);
    dw[e] = accumulate ? dw[e] + v : v;
  }
}

static inline int row_grid(int64_t N) {
  int64_t g = ceil_div(N, kWarpsPerBlock);
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sgs

using namespace sgs;

extern "C" {

int32_t sgs_gcn_norm(const int32_t* rowptr, const int32_t* perm, const int32_t* nbr, const float* w,
                     int64_t M, int64_t N, float* deg, float* dis, float* loopw, float* what,
                     sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M >= 0, "bad sizes");
  SGS_CHECK_ARG(rowptr && deg && dis && loopw && (M == 0 || (perm && nbr && what)), "null pointer");
  cudaStream_t st = as_stream(stream);
  gcn_degree_kernel<<<row_grid(N), kBlock, 0, st>>>(rowptr, perm, nbr, w, N, deg, dis, loopw);
  SGS_LAUNCH_CHECK();
  if (M > 0) {
    gcn_what_kernel<<<row_grid(N), kBlock, 0, st>>>(rowptr, perm, nbr, w, dis, N, what);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t sgs_gcn_norm_apply(const int32_t* rowptr, const int32_t* perm, const int32_t* nbr, const float* w,
                           const float* dis, int64_t M, int64_t N, float* what, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M >= 0, "bad sizes");
  if (M == 0) return SGS_OK;
  SGS_CHECK_ARG(rowptr && perm && nbr && dis && what, "null pointer");
  gcn_what_kernel<<<row_grid(N), kBlock, 0, as_stream(stream)>>>(rowptr, perm, nbr, w, dis, N, what);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

static int32_t spmm_impl(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                         const float* dis, const float* loopw, const float* h, const void* h16, const float* tscale,
                         int64_t N, int64_t D, const float* bias, float* out, int32_t flags, float p_drop,
                         uint64_t seed, const SpmmPeers& pe, cudaStream_t st) {
  if ((flags & SGS_SPMM_DROPOUT) && p_drop == 0.f) flags &= ~SGS_SPMM_DROPOUT;
  if (h16) {
    const __half* hh = reinterpret_cast<const __half*>(h16);
    // software-pipelined gather loop (r02: 2.11 -> 1.49 ms per D = 256 launch); SGS_SPMM_PIPE=0 selects the plain
    // loop for A/B measurements
    static int pipe = -1;
    if (pipe < 0) {
      const char* e = getenv("SGS_SPMM_PIPE");
      pipe = (e && e[0] == '0') ? 0 : 1;
    }
    if (D <= 256 && pipe)
      spmm_h16_kernel<1, true><<<row_grid(N), kBlock, 0, st>>>(rowptr, nbr, what, dis, loopw, hh, tscale, N, (int)D,
                                                               bias, out, flags, p_drop, seed, order, pe);
    else if (D <= 256)
      spmm_h16_kernel<1, false><<<row_grid(N), kBlock, 0, st>>>(rowptr, nbr, what, dis, loopw, hh, tscale, N, (int)D,
                                                                bias, out, flags, p_drop, seed, order, pe);
    else
      spmm_h16_kernel<2, false><<<row_grid(N), kBlock, 0, st>>>(rowptr, nbr, what, dis, loopw, hh, tscale, N, (int)D,
                                                                bias, out, flags, p_drop, seed, order, pe);
    SGS_LAUNCH_CHECK();
    return SGS_OK;
  }
  const bool vec4 = (D % 4 == 0) && (((uintptr_t)h | (uintptr_t)out) % 16 == 0) && (pe.off % 4 == 0);
  dim3 block(kBlock);
#define SGS_SPMM_LAUNCH(VEC, K)                                                                        \
  do {                                                                                                 \
    dim3 grid(row_grid(N), (unsigned)ceil_div(D, 32 * VEC * K));                                       \
    spmm_kernel<VEC, K><<<grid, block, 0, st>>>(rowptr, nbr, what, dis, loopw, h, N, (int)D, bias, out, \
                                                flags, p_drop, seed, order, pe);                       \
  } while (0)
  if (vec4) {
    if (D <= 128) SGS_SPMM_LAUNCH(4, 1);
    else if (D <= 256) SGS_SPMM_LAUNCH(4, 2);
    else SGS_SPMM_LAUNCH(4, 4);
  } else {
    if (D <= 32) SGS_SPMM_LAUNCH(1, 1);
    else if (D <= 64) SGS_SPMM_LAUNCH(1, 2);
    else SGS_SPMM_LAUNCH(1, 4);
  }
#undef SGS_SPMM_LAUNCH
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_spmm(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order, const float* dis,
                 const float* loopw, const float* h, int64_t N, int64_t D, const float* bias, float* out,
                 int32_t flags, float p_drop, uint64_t seed, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && D > 0 && D < (1 << 20), "bad sizes");
  SGS_CHECK_ARG(rowptr && h && out, "null pointer");
  SGS_CHECK_ARG(!(flags & SGS_SPMM_DROPOUT) || (p_drop >= 0.f && p_drop < 1.f), "p_drop must be in [0,1)");
  const SpmmPeers pe = {nullptr, 1, 0, 0, 0, N, 3, nullptr, 0};
  return spmm_impl(rowptr, nbr, what, order, dis, loopw, h, nullptr, nullptr, N, D, bias, out, flags, p_drop, seed, pe,
                   as_stream(stream));
}

int32_t sgs_spmm_sharded(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                         const float* dis, const float* loopw, const float* h, const void* h16, const float* tscale,
                         int64_t N, int64_t D, const float* bias, float* out, int32_t flags, float p_drop,
                         uint64_t seed, int64_t row_lo, int64_t row_hi, const uint64_t* peer_bases, int32_t world,
                         int32_t rank, int64_t elem_off, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && D > 0 && D < (1 << 20), "bad sizes");
  SGS_CHECK_ARG(rowptr && out && (h || (h16 && tscale)), "null pointer");
  SGS_CHECK_ARG(!h16 || (D % 8 == 0 && D <= 512 && (((uintptr_t)h16 | (uintptr_t)out) & 15) == 0),
                "the fp16-table SpMM needs D % 8 == 0, D <= 512 and 16-byte alignment");
  SGS_CHECK_ARG(0 <= row_lo && row_lo <= row_hi && row_hi <= N, "bad row range");
  SGS_CHECK_ARG(!peer_bases || (world >= 1 && rank >= 0 && rank < world && elem_off >= 0), "bad peer arguments");
  SGS_CHECK_ARG(!peer_bases || !h16 || elem_off % 4 == 0, "peer buffer offset must be a multiple of 4 floats");
  SGS_CHECK_ARG(!(flags & SGS_SPMM_DROPOUT) || (p_drop >= 0.f && p_drop < 1.f), "p_drop must be in [0,1)");
  const SpmmPeers pe = {peer_bases, world, rank, elem_off, row_lo, row_hi, 3, nullptr, 0};
  return spmm_impl(rowptr, nbr, what, order, dis, loopw, h, h16, tscale, N, D, bias, out, flags, p_drop, seed, pe,
                   as_stream(stream));
}

int32_t sgs_table_f16(const float* in, int64_t N, int64_t D, int32_t scaled, void* out16, float* tscale,
                      sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && D > 0 && (N * D) % 8 == 0, "N * D must be a positive multiple of 8");
  SGS_CHECK_ARG(in && out16 && tscale, "null pointer");
  SGS_CHECK_ARG((((uintptr_t)in | (uintptr_t)out16) & 15) == 0, "in / out16 must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int64_t n8 = N * D / 8;
  int64_t g = ceil_div(n8, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  if (g > cap) g = cap;
  if (scaled) {
    // tscale[2] is scratch for the absmax
    SGS_CUDA(cudaMemsetAsync(tscale + 2, 0, sizeof(float), st));
    table_absmax_kernel<<<(unsigned)g, 256, 0, st>>>(in, N * D / 4, tscale + 2);
    SGS_LAUNCH_CHECK();
  }
  table_scale_kernel<<<1, 1, 0, st>>>(scaled ? tscale + 2 : nullptr, tscale);
  SGS_LAUNCH_CHECK();
  table_convert_kernel<<<(unsigned)g, 256, 0, st>>>(in, n8, tscale, reinterpret_cast<uint4*>(out16));
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_spmm_h16(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                     const float* dis, const float* loopw, const void* h16, const float* tscale, int64_t N, int64_t D,
                     const float* bias, float* out, int32_t flags, float p_drop, uint64_t seed, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && D > 0 && D % 8 == 0 && D <= 512, "the fp16-table SpMM needs D % 8 == 0 and D <= 512");
  SGS_CHECK_ARG(rowptr && h16 && tscale && out, "null pointer");
  SGS_CHECK_ARG((((uintptr_t)h16 | (uintptr_t)out) & 15) == 0, "h16 / out must be 16-byte aligned");
  SGS_CHECK_ARG(!(flags & SGS_SPMM_DROPOUT) || (p_drop >= 0.f && p_drop < 1.f), "p_drop must be in [0,1)");
  const SpmmPeers pe = {nullptr, 1, 0, 0, 0, N, 3, nullptr, 0};
  return spmm_impl(rowptr, nbr, what, order, dis, loopw, nullptr, h16, tscale, N, D, bias, out, flags, p_drop, seed,
                   pe, as_stream(stream));
}

int32_t sgs_spmm_h16_pair(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                          const float* dis, const float* loopw, const void* h16, const float* tscale, int64_t N,
                          int64_t D, const float* bias, float* out_a, float* out_b, int32_t flags, float p_drop,
                          uint64_t seed, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && D > 0 && D % 16 == 0 && D <= 512, "the pair SpMM needs D % 16 == 0 and D <= 512");
  SGS_CHECK_ARG(rowptr && h16 && tscale && out_a && out_b, "null pointer");
  SGS_CHECK_ARG((((uintptr_t)h16 | (uintptr_t)out_a | (uintptr_t)out_b) & 15) == 0, "16-byte alignment required");
  SGS_CHECK_ARG(!(flags & (SGS_SPMM_ACCUM | SGS_SPMM_ADD_ROOT)), "ACCUM / ADD_ROOT are not available in pair mode");
  SGS_CHECK_ARG(!(flags & SGS_SPMM_DROPOUT) || (p_drop >= 0.f && p_drop < 1.f), "p_drop must be in [0,1)");
  const SpmmPeers pe = {nullptr, 1, 0, 0, 0, N, 3, out_b, (int)(D / 2)};
  return spmm_impl(rowptr, nbr, what, order, dis, loopw, nullptr, h16, tscale, N, D, bias, out_a, flags, p_drop, seed,
                   pe, as_stream(stream));
}

int32_t sgs_act_bwd(const float* gout, const float* out, int64_t n, float scale, float* gin,
                    sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0, "negative size");
  if (n == 0) return SGS_OK;
  SGS_CHECK_ARG(gout && out && gin, "null pointer");
  int64_t g = ceil_div(n, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  act_bwd_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, as_stream(stream)>>>(gout, out, n, scale, gin);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_colsum(const float* G, int64_t N, int64_t D, float* colsum, sgs_stream_t stream) {
  SGS_CHECK_ARG(N >= 0 && D > 0, "bad sizes");
  SGS_CHECK_ARG(G && colsum, "null pointer");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(colsum, 0, D * sizeof(float), st));
  if (N == 0) return SGS_OK;
  int64_t g = N < (int64_t)sm_count() * 4 ? N : (int64_t)sm_count() * 4;
  colsum_kernel<<<(unsigned)g, 256, 0, st>>>(G, N, (int)D, colsum);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

// phases of sgs_gcn_edge_grad: 1 = SDDMM + per-node sums (tmp_g, tmp_t, tmp_a), 2 = final formula (dw).
// A multi-GPU caller whose edges are sharded by destination all-reduces tmp_a[N] between the two.
static int32_t edge_grad_impl(int phases, const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                              const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                              const int32_t* perm_src, const int32_t* src, const int32_t* dst, const float* G,
                              const float* h, const float* dis, const float* deg, const float* loopw, int64_t M,
                              int64_t N, int64_t D, float* tmp_g, float* tmp_t, float* tmp_a, float* dw,
                              int32_t accumulate, cudaStream_t st, const void* h16 = nullptr,
                              const float* tscale = nullptr) {
  if ((phases & 1) && h16) {
    // fp16 gather table of h (the forward's): half the gathered bytes, table L2-resident
    if (D % 8 != 0 || D > 512) {
      set_error("sgs_gcn_edge_grad: the fp16-table form needs D %% 8 == 0 and D <= 512");
      return SGS_E_UNSUPPORTED;
    }
    const __half* hh = reinterpret_cast<const __half*>(h16);
    if (D <= 256)
      edge_grad_sddmm_h16_kernel<1><<<row_grid(N), kBlock, 0, st>>>(rowptr_dst, perm_dst, nbr_dst, what_dst, G, hh,
                                                                    tscale, dis, loopw, N, (int)D, tmp_g, tmp_t, tmp_a,
                                                                    order_dst);
    else
      edge_grad_sddmm_h16_kernel<2><<<row_grid(N), kBlock, 0, st>>>(rowptr_dst, perm_dst, nbr_dst, what_dst, G, hh,
                                                                    tscale, dis, loopw, N, (int)D, tmp_g, tmp_t, tmp_a,
                                                                    order_dst);
    SGS_LAUNCH_CHECK();
    edge_grad_srcsum_kernel<<<row_grid(N), kBlock, 0, st>>>(rowptr_src, perm_src, tmp_t, N, tmp_a);
    SGS_LAUNCH_CHECK();
  } else if (phases & 1) {
    const bool vec4 = (D % 4 == 0) && (((uintptr_t)h | (uintptr_t)G) % 16 == 0);
#define SGS_SDDMM_LAUNCH(VEC, K)                                                                         \
  edge_grad_sddmm_kernel<VEC, K><<<row_grid(N), kBlock, 0, st>>>(rowptr_dst, perm_dst, nbr_dst, what_dst, \
                                                                 G, h, dis, loopw, N, (int)D, tmp_g, tmp_t, \
                                                                 tmp_a, order_dst)
    if (vec4 && D <= 128) SGS_SDDMM_LAUNCH(4, 1);
    else if (vec4 && D <= 256) SGS_SDDMM_LAUNCH(4, 2);
    else if (vec4 && D <= 512) SGS_SDDMM_LAUNCH(4, 4);
    else if (vec4 && D <= 1024) SGS_SDDMM_LAUNCH(4, 8);
    else if (D <= 32) SGS_SDDMM_LAUNCH(1, 1);
    else if (D <= 64) SGS_SDDMM_LAUNCH(1, 2);
    else if (D <= 128) SGS_SDDMM_LAUNCH(1, 4);
    else if (D <= 256) SGS_SDDMM_LAUNCH(1, 8);
    else {
      set_error("sgs_gcn_edge_grad: unsupported width %lld", (long long)D);
      return SGS_E_UNSUPPORTED;
    }
#undef SGS_SDDMM_LAUNCH
    SGS_LAUNCH_CHECK();
    edge_grad_srcsum_kernel<<<row_grid(N), kBlock, 0, st>>>(rowptr_src, perm_src, tmp_t, N, tmp_a);
    SGS_LAUNCH_CHECK();
  }
  if (phases & 2) {
    int64_t g = ceil_div(M, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    edge_grad_final_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(src, dst, tmp_g, tmp_a, dis, deg, M, dw,
                                                                          accumulate);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t sgs_gcn_edge_grad(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                          const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                          const int32_t* perm_src,
                          const int32_t* src, const int32_t* dst, const float* G, const float* h,
                          const float* dis, const float* deg, const float* loopw, int64_t M, int64_t N,
                          int64_t D, float* tmp_g, float* tmp_t, float* tmp_a, float* dw, int32_t accumulate,
                          sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M >= 0 && D > 0, "bad sizes");
  if (M == 0) return SGS_OK;
  SGS_CHECK_ARG(rowptr_dst && perm_dst && nbr_dst && what_dst && rowptr_src && perm_src && src && dst && G &&
                    h && dis && deg && loopw && tmp_g && tmp_t && tmp_a && dw,
                "null pointer");
  return edge_grad_impl(3, rowptr_dst, perm_dst, nbr_dst, what_dst, order_dst, rowptr_src, perm_src, src, dst, G, h,
                        dis, deg, loopw, M, N, D, tmp_g, tmp_t, tmp_a, dw, accumulate, as_stream(stream));
}

int32_t sgs_gcn_edge_grad_h16(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                              const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                              const int32_t* perm_src, const int32_t* src, const int32_t* dst, const float* G,
                              const void* h16, const float* tscale, const float* dis, const float* deg,
                              const float* loopw, int64_t M, int64_t N, int64_t D, float* tmp_g, float* tmp_t,
                              float* tmp_a, float* dw, int32_t accumulate, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M >= 0 && D > 0, "bad sizes");
  if (M == 0) return SGS_OK;
  SGS_CHECK_ARG(rowptr_dst && perm_dst && nbr_dst && what_dst && rowptr_src && perm_src && src && dst && G && h16 &&
                    tscale && dis && deg && loopw && tmp_g && tmp_t && tmp_a && dw,
                "null pointer");
  SGS_CHECK_ARG((((uintptr_t)h16 | (uintptr_t)G) & 15) == 0, "G / h16 must be 16-byte aligned");
  return edge_grad_impl(3, rowptr_dst, perm_dst, nbr_dst, what_dst, order_dst, rowptr_src, perm_src, src, dst, G,
                        nullptr, dis, deg, loopw, M, N, D, tmp_g, tmp_t, tmp_a, dw, accumulate, as_stream(stream), h16,
                        tscale);
}

int32_t sgs_gcn_edge_grad_partial(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                                  const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                                  const int32_t* perm_src, const float* G, const float* h, const float* dis,
                                  const float* loopw, int64_t M, int64_t N, int64_t D, float* tmp_g, float* tmp_t,
                                  float* tmp_a, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M > 0 && D > 0, "bad sizes");
  SGS_CHECK_ARG(rowptr_dst && perm_dst && nbr_dst && what_dst && rowptr_src && perm_src && G && h && dis && loopw &&
                    tmp_g && tmp_t && tmp_a,
                "null pointer");
  return edge_grad_impl(1, rowptr_dst, perm_dst, nbr_dst, what_dst, order_dst, rowptr_src, perm_src, nullptr, nullptr,
                        G, h, dis, nullptr, loopw, M, N, D, tmp_g, tmp_t, tmp_a, nullptr, 0, as_stream(stream));
}

int32_t sgs_gcn_edge_grad_partial_h16(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                                      const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                                      const int32_t* perm_src, const float* G, const void* h16, const float* tscale,
                                      const float* dis, const float* loopw, int64_t M, int64_t N, int64_t D,
                                      float* tmp_g, float* tmp_t, float* tmp_a, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && M > 0 && D > 0, "bad sizes");
  SGS_CHECK_ARG(rowptr_dst && perm_dst && nbr_dst && what_dst && rowptr_src && perm_src && G && h16 && tscale && dis &&
                    loopw && tmp_g && tmp_t && tmp_a,
                "null pointer");
  SGS_CHECK_ARG((((uintptr_t)h16 | (uintptr_t)G) & 15) == 0, "G / h16 must be 16-byte aligned");
  return edge_grad_impl(1, rowptr_dst, perm_dst, nbr_dst, what_dst, order_dst, rowptr_src, perm_src, nullptr, nullptr,
                        G, nullptr, dis, nullptr, loopw, M, N, D, tmp_g, tmp_t, tmp_a, nullptr, 0, as_stream(stream),
                        h16, tscale);
}

int32_t sgs_gcn_edge_grad_final(const int32_t* src, const int32_t* dst, const float* tmp_g, const float* tmp_a,
                                const float* dis, const float* deg, int64_t M, float* dw, int32_t accumulate,
                                sgs_stream_t stream) {
  SGS_CHECK_ARG(M > 0, "bad sizes");
  SGS_CHECK_ARG(src && dst && tmp_g && tmp_a && dis && deg && dw, "null pointer");
  return edge_grad_impl(2, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, src, dst, nullptr, nullptr,
                        dis, deg, nullptr, M, 0, 0, const_cast<float*>(tmp_g), nullptr, const_cast<float*>(tmp_a), dw,
                        accumulate, as_stream(stream));
}
}
