// Producer side of the fused edge-scorer kernels (forward K1 and backward BA): gathers the 16-bit rows
// out[src], out[dst] of a 128-edge tile with 128-bit loads and writes the [x*y | x-y] feature blocks into the
// SWIZZLE_128B K-major stage ring consumed by tcgen05.mma.
//
// 256 producer threads; thread pt owns 16-byte chunk c = pt & 7 (8 columns) of the 4 CONSECUTIVE rows
// 4 * (pt >> 3) + i, i < 4: consecutive edges share their source even in a sampled / bucketed edge list (runs of
// ~10 edges), so the source row is loaded once for the four.
// Software pipeline (everything that can miss in L2 is issued at least one stage before it is consumed):
//   * the edge endpoints (src/dst ids) of tile t+1 are loaded while tile t is being built,
//   * the row chunks of stage g+1 (possibly the first stage of the next tile) are in flight while stage g is
//     converted and stored.
// CONTRACT: the consumer must execute fence_proxy_async_smem() after waiting on the stage's full barrier and before
// its tcgen05.mma reads the stage.
#pragma once
#include "tc.cuh"

namespace sgs {

// PAIR: the CTA is one half of a cta_group::2 pair -- the "stage full" barrier lives in the leader CTA (full0 is then
// a shared::cluster address and arrivals are released at cluster scope).
template <typename T, int H, int NSTAGE, int STAGE_BYTES, int TILE_M, bool PAIR = false>
struct FeatureProducer {
  static constexpr int NSP = H / 64;

  struct Rows {
    int32_t s[4], d[4];
  };

  __device__ static __forceinline__ Rows load_rows(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                   const int32_t* __restrict__ ids, int64_t n, int64_t t,
                                                   int row_base) {
    Rows r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t e = t * TILE_M + 4 * row_base + i;
      if (e >= n) e = n - 1;
      if (ids) e = ids[e];
      r.s[i] = src[e];
      r.d[i] = dst[e];
    }
    return r;
  }

  // Edge ids ascend by source, so the 4 rows of a thread nearly always share their source row: load it once.
  __device__ static __forceinline__ void issue(const T* __restrict__ tab, const Rows& r, int col, uint4* x, uint4* y) {
    const bool same_src = (r.s[0] == r.s[1]) & (r.s[1] == r.s[2]) & (r.s[2] == r.s[3]);
    x[0] = *reinterpret_cast<const uint4*>(tab + (int64_t)r.s[0] * H + col);
    if (same_src) {
      x[1] = x[2] = x[3] = x[0];
    } else {
#pragma unroll
      for (int i = 1; i < 4; ++i) x[i] = *reinterpret_cast<const uint4*>(tab + (int64_t)r.s[i] * H + col);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) y[i] = *reinterpret_cast<const uint4*>(tab + (int64_t)r.d[i] * H + col);
  }

  __device__ static void run(const T* __restrict__ tab, const int32_t* __restrict__ src,
                             const int32_t* __restrict__ dst, const int32_t* __restrict__ ids, int64_t n,
                             int64_t tile0, int64_t tstep, int64_t ntiles, uint8_t* stages, uint32_t full0,
                             uint32_t empty0, int pt) {
    using namespace tc;
    const int c = pt & 7;
    const int row_base = pt >> 3;
    int64_t t = tile0;
    if (t >= ntiles) return;
    Rows cur = load_rows(src, dst, ids, n, t, row_base);
    int64_t t_next = t + tstep;
    bool has_next = t_next < ntiles;
    Rows nxt = cur;
    if (has_next) nxt = load_rows(src, dst, ids, n, t_next, row_base);
    uint4 cx[4], cy[4], nx[4], ny[4];
    issue(tab, cur, c * 8, cx, cy);
    uint32_t it = 0;
    while (true) {
#pragma unroll
      for (int sp = 0; sp < NSP; ++sp, ++it) {
        if (sp + 1 < NSP) issue(tab, cur, (sp + 1) * 64 + c * 8, nx, ny);
        else if (has_next) issue(tab, nxt, c * 8, nx, ny);
        const uint32_t slot = it % NSTAGE;
        mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
        uint8_t* stage = stages + slot * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = sw128_offset(4 * row_base + i, c);
          *reinterpret_cast<uint4*>(stage + off) =
              make_uint4(Cvt<T>::mul2(cx[i].x, cy[i].x), Cvt<T>::mul2(cx[i].y, cy[i].y),
                         Cvt<T>::mul2(cx[i].z, cy[i].z), Cvt<T>::mul2(cx[i].w, cy[i].w));
          *reinterpret_cast<uint4*>(stage + TILE_M * 128 + off) =
              make_uint4(Cvt<T>::sub2(cx[i].x, cy[i].x), Cvt<T>::sub2(cx[i].y, cy[i].y),
                         Cvt<T>::sub2(cx[i].z, cy[i].z), Cvt<T>::sub2(cx[i].w, cy[i].w));
        }
        // No proxy fence here: fence.proxy.async lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC and the MEMBAR waits
        // for this thread's in-flight prefetch loads, i.e. it would expose the full gather latency once per stage
        // (ncu r01: ~22% of the stall samples).  The CTA-scope release of the arrive orders the stores before the
        // consumer's acquire; the consumer (one thread, nothing in flight) issues the proxy fence before its MMAs.
        if (PAIR) mbar_arrive_cluster(full0 + 8 * slot);
        else mbar_arrive(full0 + 8 * slot);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cx[i] = nx[i];
          cy[i] = ny[i];
        }
      }
      if (!has_next) break;
      cur = nxt;
      t = t_next;
      t_next = t + tstep;
      has_next = t_next < ntiles;
      if (has_next) nxt = load_rows(src, dst, ids, n, t_next, row_base);
    }
  }
};

}  // namespace sgs
