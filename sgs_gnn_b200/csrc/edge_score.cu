// K1 / K1b edge scorer, fp32 parity mode (CUDA cores) + dispatcher.
//   p_e = sigmoid(w2 . dropout(relu(W1 . [x*y | x-y] + b1)) + b2),  x = out[src_e], y = out[dst_e]
// The parity mode works in bounded edge chunks: features -> fp32 GEMM -> fused epilogue, and for
// the backward: recompute -> dA -> dF = dA.W1, dW1 += dA^T.F -> scatter into d_out.
// The tensor-core (tcgen05) forward lives in edge_score_tc.cu.
// precision == SGS_PREC_TF32 ("tensor-core parity mode"): the same chunked pipeline with its three contractions on
// the tcgen05 kind::tf32 GEMM (gemm_tc.cu) and chunks small enough that F and Z stay L2-resident between kernels.
#include "common.cuh"

namespace sgs {

int32_t edge_score_fwd_tc(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                          const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                          const float* b2, float p_drop, uint64_t seed, float* p, void* ws, size_t ws_bytes,
                          int32_t precision, cudaStream_t st);
size_t edge_score_tc_workspace_bytes(int64_t n, int64_t N, int64_t H);
size_t edge_score_bwd_tc_workspace_bytes(int64_t n, int64_t N, int64_t H);
int32_t edge_score_bwd_tc(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                          const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                          float p_drop, uint64_t seed, const float* p_fwd, const float* dp, float* d_out, float* dW1,
                          float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int32_t precision,
                          cudaStream_t st);
static inline bool tc_bwd_supported(int32_t precision, int64_t H) {
  return (precision == SGS_PREC_BF16 || precision == SGS_PREC_FP16) && (H == 128 || H == 256);
}

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int64_t kMaxChunk = 131072;
constexpr int64_t kMaxChunkTf32 = 32768;   // F (64 MB at H = 256) + Z (32 MB) of a chunk fit the 126 MB L2
static inline bool is_16bit(int32_t precision) { return precision == SGS_PREC_BF16 || precision == SGS_PREC_FP16; }

// Wr = rna_tf32(W), Wt[c, r] = rna_tf32(W[r, c])   (W [R, C] row-major; tiny: H x 2H)
__global__ void round_transpose_small_kernel(const float* __restrict__ W, int R, int C, float* __restrict__ Wr,
                                             float* __restrict__ Wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R * C) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(W[i]));
    Wr[i] = __uint_as_float(r);
    if (Wt) Wt[(int64_t)(i % C) * R + i / C] = __uint_as_float(r);
  }
}

__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// F[i, c] = x*y, F[i, H+c] = x-y   (one warp per edge, float4 columns); RND: values rounded to the nearest tf32
// (operands of the truncating kind::tf32 MMA, see gemm.cu)
template <bool RND>
__global__ void __launch_bounds__(kThreads)
edge_feat_kernel(const float* __restrict__ out, int H, const int32_t* __restrict__ src,
                 const int32_t* __restrict__ dst, const int32_t* __restrict__ ids, int64_t e0, int64_t n,
                 float* __restrict__ F) {
  const int lane = threadIdx.x & 31;
  int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * kWarps;
  for (; i < n; i += step) {
    const int64_t e = ids ? ids[e0 + i] : e0 + i;
    const float* x = out + (int64_t)src[e] * H;
    const float* y = out + (int64_t)dst[e] * H;
    float* f = F + i * 2 * H;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(x + c);
      const float4 b = *reinterpret_cast<const float4*>(y + c);
      float4 pr = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
      float4 df = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
      if (RND) {
        pr = make_float4(rna_tf32(pr.x), rna_tf32(pr.y), rna_tf32(pr.z), rna_tf32(pr.w));
        df = make_float4(rna_tf32(df.x), rna_tf32(df.y), rna_tf32(df.z), rna_tf32(df.w));
      }
      *reinterpret_cast<float4*>(f + c) = pr;
      *reinterpret_cast<float4*>(f + H + c) = df;
    }
  }
}

// hidden = dropout(relu(Z + b1)); z = w2.hidden + b2; p = sigmoid(z)
// BWD: also dA = dz * w2 * keep_scale * [Z + b1 > 0] written over Z, and the small parameter
// gradients reduced per block.
template <bool BWD, bool RND = false>
__global__ void __launch_bounds__(kThreads)
edge_hidden_kernel(float* __restrict__ Z, int H, const float* __restrict__ b1, const float* __restrict__ w2,
                   const float* __restrict__ b2, const int32_t* __restrict__ ids, int64_t e0, int64_t n,
                   float p_drop, uint64_t seed, float* __restrict__ p_out, const float* __restrict__ dp,
                   float* __restrict__ dw2, float* __restrict__ db1, float* __restrict__ db2) {
  extern __shared__ float sh[];  // BWD: [2*H + 1] block accumulators
  const int lane = threadIdx.x & 31;
  const uint32_t thr = dropout_threshold(p_drop);
  const bool drop = p_drop > 0.f;
  const float scale = drop ? 1.0f / (1.0f - p_drop) : 1.0f;
  const float bias2 = b2[0];
  if (BWD) {
    for (int c = threadIdx.x; c < 2 * H + 1; c += blockDim.x) sh[c] = 0.f;
    __syncthreads();
  }
  float acc_db2 = 0.f;
  float acc_w2[4][4], acc_b1[4][4];  // per-lane partial sums for columns lane*4 + 128*k (H <= 512)
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int t = 0; t < 4; ++t) { acc_w2[k][t] = 0.f; acc_b1[k][t] = 0.f; }
  int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * kWarps;
  for (; i < n; i += step) {
    const int64_t e = ids ? ids[e0 + i] : e0 + i;
    float* zr = Z + i * H;
    float z = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(zr + c);
      const float4 bb = *reinterpret_cast<const float4*>(b1 + c);
      const float4 ww = *reinterpret_cast<const float4*>(w2 + c);
      float hv[4] = {fmaxf(a.x + bb.x, 0.f), fmaxf(a.y + bb.y, 0.f), fmaxf(a.z + bb.z, 0.f),
                     fmaxf(a.w + bb.w, 0.f)};
      if (drop) {
        const uint64_t bits = dropout_bits(seed, (uint64_t)e, (uint32_t)(c >> 2));
#pragma unroll
        for (int t = 0; t < 4; ++t) hv[t] = dropout_keep(bits, t, thr) ? hv[t] * scale : 0.f;
      }
      z = fmaf(ww.x, hv[0], z);
      z = fmaf(ww.y, hv[1], z);
      z = fmaf(ww.z, hv[2], z);
      z = fmaf(ww.w, hv[3], z);
    }
    z = warp_sum(z) + bias2;
    const float pe = 1.0f / (1.0f + expf(-z));
    if (!BWD) {
      if (lane == 0) p_out[i] = pe;
    } else {
      const float dz = dp[i] * pe * (1.0f - pe);
      if (lane == 0) acc_db2 += dz;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = lane * 4 + 128 * k;
        if (c < H) {
          const float4 a = *reinterpret_cast<const float4*>(zr + c);
          const float4 bb = *reinterpret_cast<const float4*>(b1 + c);
          const float4 ww = *reinterpret_cast<const float4*>(w2 + c);
          const float pre[4] = {a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w};
          const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
          uint64_t bits = 0;
          if (drop) bits = dropout_bits(seed, (uint64_t)e, (uint32_t)(c >> 2));
          float da[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const bool on = (!drop || dropout_keep(bits, t, thr)) && pre[t] > 0.f;
            da[t] = on ? dz * wv[t] * scale : 0.f;
            acc_w2[k][t] += on ? dz * (pre[t] * scale) : 0.f;
            acc_b1[k][t] += da[t];
          }
          if (RND) {   // dA is the operand of two more kind::tf32 GEMMs
#pragma unroll
            for (int t = 0; t < 4; ++t) da[t] = rna_tf32(da[t]);
          }
          *reinterpret_cast<float4*>(zr + c) = make_float4(da[0], da[1], da[2], da[3]);
        }
      }
    }
  }
  if (BWD) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane * 4 + 128 * k;
      if (c < H) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          atomicAdd(&sh[c + t], acc_w2[k][t]);
          atomicAdd(&sh[H + c + t], acc_b1[k][t]);
        }
      }
    }
    if (lane == 0 && acc_db2 != 0.f) atomicAdd(&sh[2 * H], acc_db2);
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
      atomicAdd(dw2 + c, sh[c]);
      atomicAdd(db1 + c, sh[H + c]);
    }
    if (threadIdx.x == 0) atomicAdd(db2, sh[2 * H]);
  }
}

// d_out[src] += dF1*y + dF2 ; d_out[dst] += dF1*x - dF2
__global__ void __launch_bounds__(kThreads)
edge_feat_bwd_kernel(const float* __restrict__ out, int H, const int32_t* __restrict__ src,
                     const int32_t* __restrict__ dst, const int32_t* __restrict__ ids, int64_t e0, int64_t n,
                     const float* __restrict__ dF, float* __restrict__ d_out) {
  const int lane = threadIdx.x & 31;
  int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int64_t step = (int64_t)gridDim.x * kWarps;
  for (; i < n; i += step) {
    const int64_t e = ids ? ids[e0 + i] : e0 + i;
    const int64_t s = src[e], d = dst[e];
    const float* x = out + s * H;
    const float* y = out + d * H;
    const float* g = dF + i * 2 * H;
    float* dx = d_out + s * H;
    float* dy = d_out + d * H;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(x + c);
      const float4 b = *reinterpret_cast<const float4*>(y + c);
      const float4 g1 = *reinterpret_cast<const float4*>(g + c);
      const float4 g2 = *reinterpret_cast<const float4*>(g + H + c);
      atomicAdd(dx + c + 0, fmaf(g1.x, b.x, g2.x));
      atomicAdd(dx + c + 1, fmaf(g1.y, b.y, g2.y));
      atomicAdd(dx + c + 2, fmaf(g1.z, b.z, g2.z));
      atomicAdd(dx + c + 3, fmaf(g1.w, b.w, g2.w));
      atomicAdd(dy + c + 0, fmaf(g1.x, a.x, -g2.x));
      atomicAdd(dy + c + 1, fmaf(g1.y, a.y, -g2.y));
      atomicAdd(dy + c + 2, fmaf(g1.z, a.z, -g2.z));
      atomicAdd(dy + c + 3, fmaf(g1.w, a.w, -g2.w));
    }
  }
}

static inline int edge_grid(int64_t n) {
  int64_t g = ceil_div(n, kWarps);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static inline size_t per_edge_bytes(int64_t H, int backward) {
  return (size_t)(backward ? 5 : 3) * (size_t)H * sizeof(float);
}

}  // namespace sgs

using namespace sgs;

extern "C" {

size_t sgs_edge_score_workspace_bytes(int64_t n, int64_t N, int64_t H, int32_t precision, int32_t backward) {
  if (n <= 0 || H <= 0) return 256;
  if (is_16bit(precision) && !backward) return edge_score_tc_workspace_bytes(n, N, H);
  if (backward && tc_bwd_supported(precision, H)) return edge_score_bwd_tc_workspace_bytes(n, N, H);
  const int64_t cap = precision == SGS_PREC_TF32 ? kMaxChunkTf32 : kMaxChunk;
  const int64_t chunk = n < cap ? n : cap;
  return (size_t)chunk * per_edge_bytes(H, backward) + 256 + 2 * ((size_t)2 * H * H * sizeof(float) + 256);
}

int32_t sgs_edge_score_fwd(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                           const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                           const float* b2, float p_drop, uint64_t seed, float* p, void* ws, size_t ws_bytes,
                           int32_t precision, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0 && N > 0 && H > 0 && H % 4 == 0, "bad sizes (H must be a multiple of 4)");
  if (n == 0) return SGS_OK;
  SGS_CHECK_ARG(out && src && dst && W1 && b1 && w2 && b2 && p && ws, "null pointer");
  SGS_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "p_drop must be in [0,1)");
  cudaStream_t st = as_stream(stream);
  if (is_16bit(precision))
    return edge_score_fwd_tc(out, N, H, src, dst, ids, n, W1, b1, w2, b2, p_drop, seed, p, ws, ws_bytes,
                             precision, st);
  SGS_CHECK_ARG(precision == SGS_PREC_FP32 || precision == SGS_PREC_TF32, "unknown precision");
  const int32_t gprec = precision == SGS_PREC_TF32 ? SGS_PREC_TF32 : SGS_PREC_FP32;
  const size_t pe = per_edge_bytes(H, 0);
  int64_t chunk = (int64_t)((ws_bytes > 256 ? ws_bytes - 256 : 0) / pe);
  if (chunk > (gprec == SGS_PREC_TF32 ? kMaxChunkTf32 : kMaxChunk)) chunk = gprec == SGS_PREC_TF32 ? kMaxChunkTf32 : kMaxChunk;
  if (chunk > n) chunk = n;
  if (chunk < 1) { set_error("sgs_edge_score_fwd: workspace too small"); return SGS_E_WORKSPACE; }
  const size_t wt_bytes = (size_t)2 * H * H * sizeof(float) + 256;
  if (gprec == SGS_PREC_TF32 && ws_bytes < 256 + pe + wt_bytes) {
    set_error("sgs_edge_score_fwd: workspace too small");
    return SGS_E_WORKSPACE;
  }
  if (gprec == SGS_PREC_TF32) {
    const int64_t room = (int64_t)((ws_bytes - 256 - wt_bytes) / pe);
    if (chunk > room) chunk = room;
  }
  float* F = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* Z = F + chunk * 2 * H;
  const float* W1g = W1;
  if (gprec == SGS_PREC_TF32) {   // weights rounded to the nearest tf32 once per call
    float* W1r = (float*)(((uintptr_t)(Z + chunk * H) + 255) & ~(uintptr_t)255);
    round_transpose_small_kernel<<<(unsigned)ceil_div(2 * H * H, 256), 256, 0, st>>>(W1, (int)H, (int)(2 * H), W1r,
                                                                                    nullptr);
    SGS_LAUNCH_CHECK();
    W1g = W1r;
  }
  for (int64_t e0 = 0; e0 < n; e0 += chunk) {
    const int64_t m = (n - e0 < chunk) ? n - e0 : chunk;
    if (gprec == SGS_PREC_TF32) edge_feat_kernel<true><<<edge_grid(m), kThreads, 0, st>>>(out, (int)H, src, dst, ids, e0, m, F);
    else edge_feat_kernel<false><<<edge_grid(m), kThreads, 0, st>>>(out, (int)H, src, dst, ids, e0, m, F);
    SGS_LAUNCH_CHECK();
    int32_t rc = sgs_gemm(F, 2 * H, 1, W1g, 2 * H, 1, Z, H, m, H, 2 * H, 0, gprec, stream);
    if (rc) return rc;
    edge_hidden_kernel<false><<<edge_grid(m), kThreads, 0, st>>>(Z, (int)H, b1, w2, b2, ids, e0, m, p_drop, seed,
                                                                 p + e0, nullptr, nullptr, nullptr, nullptr);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t sgs_edge_score_bwd(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                           const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                           const float* b2, float p_drop, uint64_t seed, const float* p_fwd, const float* dp,
                           float* d_out, float* dW1, float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes,
                           int32_t precision, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0 && N > 0 && H > 0 && H % 4 == 0, "bad sizes (H must be a multiple of 4)");
  if (n == 0) return SGS_OK;
  SGS_CHECK_ARG(out && src && dst && W1 && b1 && w2 && b2 && dp && d_out && dW1 && db1 && dw2 && db2 && ws,
                "null pointer");
  SGS_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f, "p_drop must be in [0,1)");
  SGS_CHECK_ARG(H <= 512, "backward supports H <= 512");
  cudaStream_t st = as_stream(stream);
  // tensor-core modes: fused tcgen05 kernels for H in {128, 256}; other widths keep the fp32 CUDA-core path
  if (tc_bwd_supported(precision, H)) {
    SGS_CHECK_ARG(p_fwd != nullptr, "tensor-core backward needs the forward probabilities p_fwd");
    return edge_score_bwd_tc(out, N, H, src, dst, ids, n, W1, b1, w2, p_drop, seed, p_fwd, dp, d_out, dW1, db1, dw2,
                             db2, ws, ws_bytes, precision, st);
  }
  SGS_CHECK_ARG(precision == SGS_PREC_FP32 || precision == SGS_PREC_TF32 || is_16bit(precision), "unknown precision");
  // kind::tf32 needs 16-byte aligned rows; 16-bit modes at widths the fused kernels do not cover use fp32
  const bool tf32 = precision == SGS_PREC_TF32;
  const int32_t gprec = tf32 ? SGS_PREC_TF32 : SGS_PREC_FP32;
  const size_t pe = per_edge_bytes(H, 1);
  const size_t wt_bytes = 2 * ((size_t)2 * H * H * sizeof(float) + 256);
  int64_t chunk = (int64_t)((ws_bytes > 256 + wt_bytes ? ws_bytes - 256 - wt_bytes : 0) / pe);
  if (chunk > (tf32 ? kMaxChunkTf32 : kMaxChunk)) chunk = tf32 ? kMaxChunkTf32 : kMaxChunk;
  if (chunk > n) chunk = n;
  if (chunk < 1) { set_error("sgs_edge_score_bwd: workspace too small"); return SGS_E_WORKSPACE; }
  float* F = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* Z = F + chunk * 2 * H;   // becomes dA
  float* dF = Z + chunk * H;
  float* W1t = (float*)(((uintptr_t)(dF + chunk * 2 * H) + 255) & ~(uintptr_t)255);   // [2H, H]
  float* W1r = (float*)(((uintptr_t)(W1t + 2 * H * H) + 255) & ~(uintptr_t)255);      // [H, 2H]
  const float* W1g = W1;
  if (tf32) {
    round_transpose_small_kernel<<<(unsigned)ceil_div(2 * H * H, 256), 256, 0, st>>>(W1, (int)H, (int)(2 * H), W1r,
                                                                                    W1t);
    SGS_LAUNCH_CHECK();
    W1g = W1r;
  }
  const size_t shmem = (size_t)(2 * H + 1) * sizeof(float);
  for (int64_t e0 = 0; e0 < n; e0 += chunk) {
    const int64_t m = (n - e0 < chunk) ? n - e0 : chunk;
    if (tf32) edge_feat_kernel<true><<<edge_grid(m), kThreads, 0, st>>>(out, (int)H, src, dst, ids, e0, m, F);
    else edge_feat_kernel<false><<<edge_grid(m), kThreads, 0, st>>>(out, (int)H, src, dst, ids, e0, m, F);
    SGS_LAUNCH_CHECK();
    int32_t rc = sgs_gemm(F, 2 * H, 1, W1g, 2 * H, 1, Z, H, m, H, 2 * H, 0, gprec, stream);
    if (rc) return rc;
    if (tf32)
      edge_hidden_kernel<true, true><<<edge_grid(m), kThreads, shmem, st>>>(Z, (int)H, b1, w2, b2, ids, e0, m, p_drop,
                                                                            seed, nullptr, dp + e0, dw2, db1, db2);
    else
      edge_hidden_kernel<true><<<edge_grid(m), kThreads, shmem, st>>>(Z, (int)H, b1, w2, b2, ids, e0, m, p_drop, seed,
                                                                      nullptr, dp + e0, dw2, db1, db2);
    SGS_LAUNCH_CHECK();
    // dF[m,2H] = dA[m,H] . W1[H,2H]
    if (tf32) rc = sgs_gemm(Z, H, 1, W1t, H, 1, dF, 2 * H, m, 2 * H, H, 0, SGS_PREC_TF32, stream);   // NT vs W1^T
    else rc = sgs_gemm(Z, H, 1, W1, 1, 2 * H, dF, 2 * H, m, 2 * H, H, 0, SGS_PREC_FP32, stream);
    if (rc) return rc;
    // dW1[H,2H] += dA^T[H,m] . F[m,2H]   (TN form: both operands unit-stride along M / N)
    rc = sgs_gemm(Z, 1, H, F, 1, 2 * H, dW1, 2 * H, H, 2 * H, m, 1, gprec, stream);
    if (rc) return rc;
    edge_feat_bwd_kernel<<<edge_grid(m), kThreads, 0, st>>>(out, (int)H, src, dst, ids, e0, m, dF, d_out);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}
}
