// K5: fused cross-entropy + accuracy (conditional gate) + assortative BCE (reg1) + consistency MSE
// (reg2), forward and backward.  One pass over the train rows and one over the sampled edges.
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
}
