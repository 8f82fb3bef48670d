// K4 dense contraction C[M,N] (+)= A[M,K] . B[N,K]^T with arbitrary element strides.
// This file holds the CUDA-core fp32 path (parity mode) and the dispatcher; the tcgen05/TMA
// tensor-core path lives in gemm_tc.cu.
#include "common.cuh"

namespace sgs {

int32_t gemm_tc_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                   int64_t N, int64_t K, int32_t accumulate, int32_t precision, cudaStream_t st);
int32_t gemm_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                int64_t N, int64_t K, int32_t accumulate, int32_t precision, cudaStream_t st);

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <bool SPLITK>
__global__ void __launch_bounds__(256)
gemm_fp32_kernel(const float* __restrict__ A, int64_t a_sm, int64_t a_sk, const float* __restrict__ B,
                 int64_t b_sn, int64_t b_sk, float* __restrict__ C, int64_t ldc, int M, int N, int64_t K,
                 int64_t k_per_split, int accumulate) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 micro-tile each
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const bool a_kfast = (a_sk == 1), b_kfast = (b_sk == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int m, k;
      if (a_kfast) { k = idx & (BK - 1); m = idx >> 4; } else { m = idx & (BM - 1); k = idx >> 6; }
      const int gm = m0 + m;
      const int64_t gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? A[gm * a_sm + gk * a_sk] : 0.f;
      int n, kb;
      if (b_kfast) { kb = idx & (BK - 1); n = idx >> 4; } else { n = idx & (BN - 1); kb = idx >> 6; }
      const int gn = n0 + n;
      const int64_t gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < kend) ? B[gn * b_sn + gkb * b_sk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      if (SPLITK) atomicAdd(c, acc[i][j]);
      else *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

__global__ void zero_rows_kernel(float* __restrict__ C, int64_t ldc, int M, int N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)M * N;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) C[(i / N) * ldc + (i % N)] = 0.f;
}

// tcgen05 kind::tf32 reads the upper 19 bits of each fp32 operand, i.e. it TRUNCATES the mantissa: a biased error of
// up to 2^-10 per product that does not average out over K.  Rounding the operands to the nearest tf32 value first
// (cvt.rna) makes the error unbiased, so it shrinks with sqrt(K) -- what keeps the tensor-core path inside the fp32
// 1e-4 parity bar (measured r02: 1.8e-4 -> see tests/test_gpu_tc.py).
__global__ void round_tf32_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  if ((((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
    for (int64_t k = i; k < n4; k += stride) {
      float4 v = reinterpret_cast<const float4*>(in)[k];
      uint32_t a, b, c, d;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(v.x));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(b) : "f"(v.y));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(c) : "f"(v.z));
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(d) : "f"(v.w));
      reinterpret_cast<uint4*>(out)[k] = make_uint4(a, b, c, d);
    }
    i += n4 * 4;
  }
  for (; i < n; i += stride) {
    uint32_t a;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(in[i]));
    out[i] = __uint_as_float(a);
  }
}

}  // namespace sgs

using namespace sgs;

extern "C" int32_t sgs_round_tf32(const float* in, int64_t n, float* out, sgs_stream_t stream) {
  SGS_CHECK_ARG(n >= 0, "negative size");
  if (n == 0) return SGS_OK;
  SGS_CHECK_ARG(in && out, "null pointer");
  int64_t g = ceil_div(n / 4 + 1, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  round_tf32_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, as_stream(stream)>>>(in, n, out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

extern "C" int32_t sgs_gemm(const float* A, int64_t a_sm, int64_t a_sk, const float* B, int64_t b_sn,
                            int64_t b_sk, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                            int32_t accumulate, int32_t precision, sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && M < (1ll << 31) && N < (1ll << 31), "bad sizes");
  SGS_CHECK_ARG(ldc >= N, "ldc < N");
  if (M == 0 || N == 0) return SGS_OK;
  SGS_CHECK_ARG(A && B && C, "null pointer");
  cudaStream_t st = as_stream(stream);
  if (precision != SGS_PREC_FP32) {
    if (a_sk == 1 && b_sk == 1 && K > 0) return gemm_tc(A, a_sm, B, b_sn, C, ldc, M, N, K, accumulate, precision, st);
    // TN form: both operands have unit stride along M / N (reduction over their rows)
    if (a_sm == 1 && b_sn == 1 && K > 0)
      return gemm_tc_tn(A, a_sk, B, b_sk, C, ldc, M, N, K, accumulate, precision, st);
    set_error("sgs_gemm: tensor-core modes need unit stride along K (NT) or along M and N (TN) for both operands");
    return SGS_E_UNSUPPORTED;
  }
  const int64_t tiles = ceil_div(M, BM) * ceil_div(N, BN);
  int64_t splits = 1;
  const int64_t target = 2ll * sm_count();
  if (tiles < target && K >= 4 * BK * 8) {
    splits = ceil_div(target, tiles);
    const int64_t max_splits = K / (BK * 8);
    if (splits > max_splits) splits = max_splits;
    if (splits > 512) splits = 512;
    if (splits < 1) splits = 1;
  }
  int64_t k_per_split = ceil_div(ceil_div(K, splits), BK) * BK;
  if (k_per_split == 0) k_per_split = BK;
  splits = K > 0 ? ceil_div(K, k_per_split) : 1;
  dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), (unsigned)splits);
  if (splits > 1) {
    if (!accumulate) {
      int64_t total = M * N;
      int64_t g = ceil_div(total, 256);
      int64_t cap = (int64_t)sm_count() * 8;
      zero_rows_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(C, ldc, (int)M, (int)N);
      SGS_LAUNCH_CHECK();
    }
    gemm_fp32_kernel<true><<<grid, 256, 0, st>>>(A, a_sm, a_sk, B, b_sn, b_sk, C, ldc, (int)M, (int)N, K,
                                                 k_per_split, 1);
  } else {
    gemm_fp32_kernel<false><<<grid, 256, 0, st>>>(A, a_sm, a_sk, B, b_sn, b_sk, C, ldc, (int)M, (int)N, K,
                                                  k_per_split, accumulate);
  }
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}
