// K5: fused cross-entropy + accuracy (conditional gate) + assortative BCE (reg1) + consistency MSE
// (reg2), forward and backward.  One pass over the train rows and one over the sampled edges.
#include <stdlib.h>

#include "common.cuh"

namespace sgs {

constexpr int kThreads = 256;
constexpr float kCosEps = 1e-8f;

__device__ __forceinline__ void block_add(double* acc, const double* vals, int nvals) {
  // vals: per-thread partials; reduce within warp then one atomic per warp
  for (int k = 0; k < nvals; ++k) {
    double v = warp_sum(vals[k]);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(acc + k, v);
  }
}

// warp per node: CE + argmax accuracy over train rows
__global__ void __launch_bounds__(kThreads)
loss_nodes_fwd_kernel(const float* __restrict__ logits, int64_t N, int C, const int64_t* __restrict__ y,
                      const uint8_t* __restrict__ train_mask, double* __restrict__ acc, double q_edges) {
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) acc[7] = q_edges;
  int64_t n = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * (kThreads / 32);
  double part[3] = {0.0, 0.0, 0.0};
  for (; n < N; n += step) {
    if (!train_mask[n]) continue;
    const float* row = logits + n * C;
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > m) { m = v; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(row[c] - m);
    se = warp_sum(se);
    if (lane == 0) {
      const int64_t t = y[n];
      const float lse = m + logf(se);
      part[0] += (double)(lse - row[t]);
      part[1] += 1.0;
      part[2] += (am == (int)t) ? 1.0 : 0.0;
    }
  }
  block_add(acc, part, 3);
}

// 16 lanes per sampled edge
__global__ void __launch_bounds__(kThreads)
loss_edges_fwd_kernel(const float* __restrict__ logits, int C, const int64_t* __restrict__ y,
                      const uint8_t* __restrict__ train_mask, const int32_t* __restrict__ s_src,
                      const int32_t* __restrict__ s_dst, const float* __restrict__ p_s, int64_t q,
                      double* __restrict__ acc) {
  const int sl = threadIdx.x & 15;
  int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 4;
  const int64_t step = ((int64_t)gridDim.x * kThreads) >> 4;
  double part[4] = {0.0, 0.0, 0.0, 0.0};  // bce, n_valid, sum_label, mse
  const int64_t q_round = (q + 1) & ~(int64_t)1;  // both half-warps of a warp stay in the loop together
  for (; i < q_round; i += step) {
    const bool live = i < q;
    const int s = live ? s_src[i] : 0, d = live ? s_dst[i] : 0;
    const float* a = logits + (int64_t)s * C;
    const float* b = logits + (int64_t)d * C;
    float dot = 0.f, na = 0.f, nb = 0.f;
    if (live)
      for (int c = sl; c < C; c += 16) {
        const float av = a[c], bv = b[c];
        dot = fmaf(av, bv, dot);
        na = fmaf(av, av, na);
        nb = fmaf(bv, bv, nb);
      }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      na += __shfl_xor_sync(0xffffffffu, na, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (live && sl == 0) {
      const float cosv = dot / (fmaxf(sqrtf(na), kCosEps) * fmaxf(sqrtf(nb), kCosEps));
      const float p = p_s[i];
      const float diff = p - cosv;
      part[3] += (double)(diff * diff);
      if (train_mask[s] && train_mask[d]) {
        const bool same = y[s] == y[d];
        const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.0f - p), -100.f);
        part[0] += (double)(same ? -lp : -l1p);
        part[1] += 1.0;
        part[2] += same ? 1.0 : 0.0;
      }
    }
  }
  block_add(acc + 3, part, 4);
}

__global__ void loss_finish_kernel(const double* __restrict__ acc, float c0, float c1, float c2, int reg1,
                                   int reg2, float* __restrict__ loss_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float loss = (c0 != 0.f) ? c0 * (float)(acc[0] / acc[1]) : 0.f;
    if (reg1 && acc[5] > 1.0) loss += c1 * (float)(acc[3] / acc[4]);
    if (reg2) loss += c2 * (float)(acc[6] / acc[7]);
    loss_out[0] = loss;
  }
}

// dlogits rows: train rows get g*(softmax - onehot)/n_train, the rest 0 (every row is written)
__global__ void __launch_bounds__(kThreads)
loss_nodes_bwd_kernel(const float* __restrict__ logits, int64_t N, int C, const int64_t* __restrict__ y,
                      const uint8_t* __restrict__ train_mask, const double* __restrict__ acc, float c0,
                      const float* __restrict__ gscale, float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * (kThreads / 32);
  const float g = (c0 != 0.f) ? c0 * gscale[0] / (float)acc[1] : 0.f;
  for (; n < N; n += step) {
    const float* row = logits + n * C;
    float* drow = dlogits + n * C;
    if (c0 == 0.f || !train_mask[n]) {
      for (int c = lane; c < C; c += 32) drow[c] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(row[c] - m);
    se = warp_sum(se);
    const int t = (int)y[n];
    for (int c = lane; c < C; c += 32) drow[c] = g * (expf(row[c] - m) / se - (c == t ? 1.f : 0.f));
  }
}

__global__ void __launch_bounds__(kThreads)
loss_edges_bwd_kernel(const float* __restrict__ logits, int C, const int64_t* __restrict__ y,
                      const uint8_t* __restrict__ train_mask, const int32_t* __restrict__ s_src,
                      const int32_t* __restrict__ s_dst, const float* __restrict__ p_s, int64_t q,
                      const double* __restrict__ acc, float c1, float c2, int reg1, int reg2,
                      const float* __restrict__ gscale, float* __restrict__ dlogits, float* __restrict__ dp_s) {
  const int sl = threadIdx.x & 15;
  int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 4;
  const int64_t step = ((int64_t)gridDim.x * kThreads) >> 4;
  const float g = gscale[0];
  const bool reg1_on = reg1 && acc[5] > 1.0;
  const float g1 = reg1_on ? g * c1 / (float)acc[4] : 0.f;
  const float g2 = reg2 ? g * c2 * 2.0f / (float)acc[7] : 0.f;
  const int64_t q_round = (q + 1) & ~(int64_t)1;
  for (; i < q_round; i += step) {
    const bool live = i < q;
    const int s = live ? s_src[i] : 0, d = live ? s_dst[i] : 0;
    const float* a = logits + (int64_t)s * C;
    const float* b = logits + (int64_t)d * C;
    float dot = 0.f, na = 0.f, nb = 0.f;
    if (live && reg2)
      for (int c = sl; c < C; c += 16) {
        const float av = a[c], bv = b[c];
        dot = fmaf(av, bv, dot);
        na = fmaf(av, av, na);
        nb = fmaf(bv, bv, nb);
      }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      na += __shfl_xor_sync(0xffffffffu, na, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (!live) continue;
    const float p = p_s[i];
    const float sna = sqrtf(na), snb = sqrtf(nb);
    const float nae = fmaxf(sna, kCosEps), nbe = fmaxf(snb, kCosEps);
    const float cosv = reg2 ? dot / (nae * nbe) : 0.f;
    if (sl == 0) {
      float dp = g2 * (p - cosv);
      if (reg1_on && train_mask[s] && train_mask[d]) {
        const float lab = (y[s] == y[d]) ? 1.f : 0.f;
        dp += g1 * (p - lab) / fmaxf(p * (1.0f - p), 1e-12f);
      }
      dp_s[i] = dp;
    }
    if (reg2) {
      const float r = g2 * (cosv - p);  // dL/dcos
      const float inv = r / (nae * nbe);
      const float ka = (sna > kCosEps) ? r * cosv / na : 0.f;
      const float kb = (snb > kCosEps) ? r * cosv / nb : 0.f;
      float* da = dlogits + (int64_t)s * C;
      float* db = dlogits + (int64_t)d * C;
      for (int c = sl; c < C; c += 16) {
        const float av = a[c], bv = b[c];
        atomicAdd(da + c, inv * bv - ka * av);
        atomicAdd(db + c, inv * av - kb * bv);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fused edge pass: loss sums AND the (unscaled) gradients of the edge terms in ONE sweep over the sampled edges.
// The upstream scale g, the coefficients and the global counts only enter the backward as two scalars
//   g1 = g c1 / n_valid [sum_label > 1],   g2 = 2 g c2 / q
// so the sweep stores  u1[i] = (p - label) / max(p (1 - p), 1e-12)  (0 for edges outside the train mask),
// u2[i] = p - cos  and accumulates  E[n, :] = sum over incident sampled edges of d cos / d logits[n] * (cos - p);
// the backward is then two streaming passes:  dp = g1 u1 + g2 u2,  dlogits = CE part + g2 E  (no second gather,
// no second round of atomics).
// A 16-lane group walks RUN consecutive edges: edge ids ascend by source, so the source row, its norm and its
// gradient accumulator stay in registers over the run (one flush of C atomics per run instead of per edge); the
// destination rows of 4 edges are in flight at a time.
// ---------------------------------------------------------------------------------------------------------
__global__ void node_code_kernel(const int64_t* __restrict__ y, const uint8_t* __restrict__ train_mask, int64_t N,
                                 int32_t* __restrict__ code) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; n < N; n += step) code[n] = train_mask[n] ? (int32_t)y[n] : -1;
}

template <int KC, int UN>   // KC columns per lane (C <= 16 * KC); UN destination rows in flight
__global__ void __launch_bounds__(kThreads)
loss_edges_fused_kernel(const float* __restrict__ logits, int C, const int32_t* __restrict__ code,
                        const int32_t* __restrict__ s_src, const int32_t* __restrict__ s_dst,
                        const float* __restrict__ p_s, int64_t q, double* __restrict__ acc,
                        float* __restrict__ dlog_e, float* __restrict__ u1, float* __restrict__ u2) {
  constexpr int RUN = 16;   // consecutive edges per 16-lane group and iteration
  const int sl = threadIdx.x & 15;
  const uint32_t hm = 0xFFFFu << (threadIdx.x & 16);   // this half-warp
  int64_t grp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 4;
  const int64_t gstep = ((int64_t)gridDim.x * kThreads) >> 4;
  const int64_t nruns = (q + RUN - 1) / RUN;
  double part[4] = {0.0, 0.0, 0.0, 0.0};  // bce, n_valid, sum_label, mse
  for (; grp < nruns; grp += gstep) {
    const int64_t i0 = grp * RUN;
    const int cnt = (int)(q - i0 < RUN ? q - i0 : RUN);
    // lane k of the group holds edge i0 + k
    int my_s = 0, my_d = 0;
    float my_p = 0.5f;
    if (sl < cnt) {
      my_s = s_src[i0 + sl];
      my_d = s_dst[i0 + sl];
      my_p = p_s[i0 + sl];
    }
    const int my_cs = code[my_s], my_cd = code[my_d];
    int cur_s = -1;
    float av[KC], da[KC], na = 0.f, sna = 0.f, nae = 1.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) av[k] = da[k] = 0.f;
    for (int e0 = 0; e0 < cnt; e0 += UN) {
      int d[UN];
      float bv[UN][KC];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        d[u] = __shfl_sync(hm, my_d, (e0 + u) & 15, 16);
        const float* b = logits + (int64_t)d[u] * C;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          const int c = sl + 16 * k;
          bv[u][k] = (e0 + u < cnt && c < C) ? __ldg(b + c) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        if (e0 + u >= cnt) break;   // uniform in the group
        const int s = __shfl_sync(hm, my_s, e0 + u, 16);
        const float p = __shfl_sync(hm, my_p, e0 + u, 16);
        const int cs = __shfl_sync(hm, my_cs, e0 + u, 16), cd = __shfl_sync(hm, my_cd, e0 + u, 16);
        if (s != cur_s) {           // uniform: a new run of equal sources
          if (cur_s >= 0) {
            float* o = dlog_e + (int64_t)cur_s * C;
#pragma unroll
            for (int k = 0; k < KC; ++k)
              if (sl + 16 * k < C) atomicAdd(o + sl + 16 * k, da[k]);
          }
          cur_s = s;
          const float* a = logits + (int64_t)s * C;
          na = 0.f;
#pragma unroll
          for (int k = 0; k < KC; ++k) {
            const int c = sl + 16 * k;
            av[k] = c < C ? __ldg(a + c) : 0.f;
            da[k] = 0.f;
            na = fmaf(av[k], av[k], na);
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) na += __shfl_xor_sync(hm, na, o, 16);
          sna = sqrtf(na);
          nae = fmaxf(sna, kCosEps);
        }
        float dot = 0.f, nb = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          dot = fmaf(av[k], bv[u][k], dot);
          nb = fmaf(bv[u][k], bv[u][k], nb);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          dot += __shfl_xor_sync(hm, dot, o, 16);
          nb += __shfl_xor_sync(hm, nb, o, 16);
        }
        const float snb = sqrtf(nb), nbe = fmaxf(snb, kCosEps);
        const float cosv = dot / (nae * nbe);
        const float diff = p - cosv;
        // d/dlogits of the consistency term, without its scalar 2 g c2 / q
        const float r = -diff;   // cos - p
        const float inv = r / (nae * nbe);
        const float ka = (sna > kCosEps) ? r * cosv / na : 0.f;
        const float kb = (snb > kCosEps) ? r * cosv / nb : 0.f;
        float* ob = dlog_e + (int64_t)d[u] * C;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          da[k] += inv * bv[u][k] - ka * av[k];
          if (sl + 16 * k < C) atomicAdd(ob + sl + 16 * k, inv * av[k] - kb * bv[u][k]);
        }
        if (sl == 0) {
          part[3] += (double)(diff * diff);
          float g1u = 0.f;
          if (cs >= 0 && cd >= 0) {
            const bool same = cs == cd;
            const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.0f - p), -100.f);
            part[0] += (double)(same ? -lp : -l1p);
            part[1] += 1.0;
            part[2] += same ? 1.0 : 0.0;
            g1u = (p - (same ? 1.f : 0.f)) / fmaxf(p * (1.0f - p), 1e-12f);
          }
          u1[i0 + e0 + u] = g1u;
          u2[i0 + e0 + u] = diff;
        }
      }
    }
    if (cur_s >= 0) {
      float* o = dlog_e + (int64_t)cur_s * C;
#pragma unroll
      for (int k = 0; k < KC; ++k)
        if (sl + 16 * k < C) atomicAdd(o + sl + 16 * k, da[k]);
    }
  }
  block_add(acc + 3, part, 4);
}

// dlogits = CE part (train rows) + g2 * E      (every row is written)
__global__ void __launch_bounds__(kThreads)
loss_nodes_bwd_fused_kernel(const float* __restrict__ logits, int64_t N, int C, const int64_t* __restrict__ y,
                            const uint8_t* __restrict__ row_mask, const double* __restrict__ acc, float c0, float c2,
                            int reg2, const float* __restrict__ gscale, const float* __restrict__ dlog_e,
                            float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  int64_t n = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * (kThreads / 32);
  const float gs = gscale[0];
  const float g = (c0 != 0.f) ? c0 * gs / (float)acc[1] : 0.f;
  const float g2 = reg2 ? gs * c2 * 2.0f / (float)acc[7] : 0.f;
  for (; n < N; n += step) {
    const float* row = logits + n * C;
    const float* erow = dlog_e + n * C;
    float* drow = dlogits + n * C;
    if (c0 == 0.f || !row_mask[n]) {
      for (int c = lane; c < C; c += 32) drow[c] = g2 * erow[c];
      continue;
    }
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(row[c] - m);
    se = warp_sum(se);
    const int t = (int)y[n];
    for (int c = lane; c < C; c += 32)
      drow[c] = g * (expf(row[c] - m) / se - (c == t ? 1.f : 0.f)) + g2 * erow[c];
  }
}

__global__ void __launch_bounds__(kThreads)
loss_dp_kernel(const float* __restrict__ u1, const float* __restrict__ u2, int64_t q, const double* __restrict__ acc,
               float c1, float c2, int reg1, int reg2, const float* __restrict__ gscale, float* __restrict__ dp_s) {
  const float g = gscale[0];
  const float g1 = (reg1 && acc[5] > 1.0) ? g * c1 / (float)acc[4] : 0.f;
  const float g2 = reg2 ? g * c2 * 2.0f / (float)acc[7] : 0.f;
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * kThreads;
  for (; i < q; i += step) dp_s[i] = g1 * u1[i] + g2 * u2[i];
}

static inline int lgrid(int64_t items_per_block_unit, int64_t n) {
  int64_t g = ceil_div(n, items_per_block_unit);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sgs

using namespace sgs;

extern "C" {

int32_t sgs_loss_fwd(const float* logits, int64_t N, int64_t C, const int64_t* y, const uint8_t* train_mask,
                     const uint8_t* row_mask, const int32_t* s_src, const int32_t* s_dst, const float* p_s, int64_t q, int32_t with_edges,
                     double* acc, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && C > 0 && q >= 0, "bad sizes");
  SGS_CHECK_ARG(logits && y && train_mask && acc, "null pointer");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(acc, 0, 8 * sizeof(double), st));
  loss_nodes_fwd_kernel<<<lgrid(kThreads / 32, N), kThreads, 0, st>>>(logits, N, (int)C, y,
                                                                     row_mask ? row_mask : train_mask, acc, (double)q);
  SGS_LAUNCH_CHECK();
  if (with_edges && q > 0) {
    SGS_CHECK_ARG(s_src && s_dst && p_s, "null pointer (edges)");
    loss_edges_fwd_kernel<<<lgrid(kThreads / 16, q), kThreads, 0, st>>>(logits, (int)C, y, train_mask, s_src, s_dst,
                                                                       p_s, q, acc);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t sgs_loss_finish(const double* acc, float c0, float c1, float c2, int32_t reg1, int32_t reg2,
                        float* loss_out, sgs_stream_t stream) {
  SGS_CHECK_ARG(acc && loss_out, "null pointer");
  loss_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(acc, c0, c1, c2, reg1, reg2, loss_out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_loss_bwd(const float* logits, int64_t N, int64_t C, const int64_t* y, const uint8_t* train_mask,
                     const uint8_t* row_mask, const int32_t* s_src, const int32_t* s_dst, const float* p_s, int64_t q, int32_t with_edges,
                     const double* acc, float c0, float c1, float c2, int32_t reg1, int32_t reg2,
                     const float* gscale, float* dlogits, float* dp_s, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && C > 0 && q >= 0, "bad sizes");
  SGS_CHECK_ARG(logits && y && train_mask && acc && gscale && dlogits, "null pointer");
  cudaStream_t st = as_stream(stream);
  loss_nodes_bwd_kernel<<<lgrid(kThreads / 32, N), kThreads, 0, st>>>(logits, N, (int)C, y,
                                                                     row_mask ? row_mask : train_mask, acc, c0, gscale,
                                                                     dlogits);
  SGS_LAUNCH_CHECK();
  if (with_edges && q > 0) {
    SGS_CHECK_ARG(s_src && s_dst && p_s && dp_s, "null pointer (edges)");
    loss_edges_bwd_kernel<<<lgrid(kThreads / 16, q), kThreads, 0, st>>>(logits, (int)C, y, train_mask, s_src, s_dst,
                                                                       p_s, q, acc, c1, c2, reg1, reg2, gscale,
                                                                       dlogits, dp_s);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t sgs_loss_fwd_fused(const float* logits, int64_t N, int64_t C, const int64_t* y, const uint8_t* train_mask,
                           const uint8_t* row_mask, const int32_t* s_src, const int32_t* s_dst, const float* p_s,
                           int64_t q, double* acc, float* dlog_e, float* u_reg1, float* u_reg2, int32_t* node_code,
                           sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && C > 0 && q > 0, "bad sizes");
  SGS_CHECK_ARG(C <= 64, "the fused edge pass supports at most 64 classes");
  SGS_CHECK_ARG(logits && y && train_mask && acc && s_src && s_dst && p_s && dlog_e && u_reg1 && u_reg2 && node_code,
                "null pointer");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(acc, 0, 8 * sizeof(double), st));
  SGS_CUDA(cudaMemsetAsync(dlog_e, 0, (size_t)N * C * sizeof(float), st));
  loss_nodes_fwd_kernel<<<lgrid(kThreads / 32, N), kThreads, 0, st>>>(logits, N, (int)C, y,
                                                                     row_mask ? row_mask : train_mask, acc, (double)q);
  SGS_LAUNCH_CHECK();
  node_code_kernel<<<lgrid(kThreads, N), kThreads, 0, st>>>(y, train_mask, N, node_code);
  SGS_LAUNCH_CHECK();
  const int grid = lgrid(kThreads, q);   // 16 groups of 16 edges per block and iteration
  // four destination rows in flight per 16-lane group (eight measured slower: 4.0 -> 5.9 ms, registers / occupancy)
#define SGS_LEF(KC)                                                                                            \
  loss_edges_fused_kernel<KC, 4><<<grid, kThreads, 0, st>>>(logits, (int)C, node_code, s_src, s_dst, p_s, q, acc, \
                                                            dlog_e, u_reg1, u_reg2)
  if (C <= 16) SGS_LEF(1);
  else if (C <= 32) SGS_LEF(2);
  else if (C <= 48) SGS_LEF(3);
  else SGS_LEF(4);
#undef SGS_LEF
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_loss_bwd_fused(const float* logits, int64_t N, int64_t C, const int64_t* y, const uint8_t* train_mask,
                           const uint8_t* row_mask, int64_t q, const double* acc, float c0, float c1, float c2,
                           int32_t reg1, int32_t reg2, const float* gscale, const float* dlog_e, const float* u_reg1,
                           const float* u_reg2, float* dlogits, float* dp_s, sgs_stream_t stream) {
  SGS_CHECK_ARG(N > 0 && C > 0 && q > 0, "bad sizes");
  SGS_CHECK_ARG(logits && y && train_mask && acc && gscale && dlog_e && u_reg1 && u_reg2 && dlogits && dp_s,
                "null pointer");
  cudaStream_t st = as_stream(stream);
  loss_nodes_bwd_fused_kernel<<<lgrid(kThreads / 32, N), kThreads, 0, st>>>(
      logits, N, (int)C, y, row_mask ? row_mask : train_mask, acc, c0, c2, reg2, gscale, dlog_e, dlogits);
  SGS_LAUNCH_CHECK();
  loss_dp_kernel<<<lgrid(kThreads, q), kThreads, 0, st>>>(u_reg1, u_reg2, q, acc, c1, c2, reg1, reg2, gscale, dp_s);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}
}
