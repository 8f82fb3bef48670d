// K4 tensor-core path: C[M,N] (+)= A[M,K] . B[N,K]^T on tcgen05 (kind::tf32 on the fp32 operands as they
// lie in HBM), operands streamed by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 6-stage mbarrier ring,
// fp32 accumulators double-buffered in TMEM, persistent CTAs over 128x128 output tiles.
//   warp 0: TMA producer (one elected thread)      warp 1: MMA issuer (one thread) + TMEM allocation
//   warps 2-5: epilogue (TMEM -> registers -> global, one warp per TMEM lane quarter)
// K and N tails come for free from TMA's out-of-bounds zero fill; M/N tails are masked in the epilogue.
// Requirements (checked, SGS_E_UNSUPPORTED otherwise): unit stride along K, row strides multiples of 4
// elements (16 B) and 16-byte aligned bases -- the Python wrapper pads odd strides once (e.g. F = 602).
//
// TN form (template TN = true): C[M,N] (+)= A^T . B with A stored [K, M] and B stored [K, N] (unit stride along
// M / N) -- the weight gradients dW = dh^T . x, a reduction over the N ~ 2.3e5 node rows with a tiny output.  Both
// operands are MN-major: TMA boxes of [BK rows x 32 floats] land as SWIZZLE_128B blocks and are consumed through
// MN-major UMMA descriptors (no transposed copy in HBM).  The K range is split over the CTAs (work item =
// output tile x K slice) and the epilogue reduces into C with red.global.add.
#include <cuda.h>
#include <cstdio>

#include "common.cuh"
#include "tc.cuh"
#include "tma.cuh"

namespace sgs {

namespace k4 {
constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 = 128 B rows
constexpr int STAGE_BYTES = (BM + BN) * BK * 4;     // 32 KB
constexpr int NSTAGE = 6;
constexpr int THREADS = 6 * 32;
}  // namespace k4

// 2-D fp32 tensor [rows, inner] with row stride ld (elements); box = [box_rows x 32 floats], 128-byte swizzle
static bool make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t K, int64_t ld,
                     int box_rows = k4::BM, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)k4::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool TN>
__global__ void __launch_bounds__(k4::THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 float* __restrict__ C, int64_t ldc, int M, int N, int K, int accumulate, int nsplit, int kb_per) {
  using namespace k4;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + NSTAGE * STAGE_BYTES);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * NSTAGE, tfull0 = empty0 + 8 * NSTAGE,
                 tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int ntiles = tiles_m * tiles_n;
  const int nkb_all = (K + BK - 1) / BK;
  const int nwork = ntiles * nsplit;   // work item w: output tile w % ntiles, K slice w / ntiles (TN only: nsplit > 1)
  auto kb_range = [&](int w, int& kb0, int& kb1) {
    const int ks = w / ntiles;
    kb0 = ks * kb_per;
    kb1 = kb0 + kb_per < nkb_all ? kb0 + kb_per : nkb_all;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4 * 32);
    }
    fence_mbar_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_s), 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        // consecutive CTAs walk down M inside one N tile: the B tile stays hot in L2
        const int t = w % ntiles;
        const int tm = t % tiles_m, tn = t / tiles_m;
        int kb0, kb1;
        kb_range(w, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
          mbar_expect_tx(full0 + 8 * slot, STAGE_BYTES);
          const uint32_t dst = sm_addr + slot * STAGE_BYTES;
          if (TN) {
            // [BK rows (k) x 32 floats (m or n)] boxes, one per 32-wide block of the tile
#pragma unroll
            for (int j = 0; j < BM / 32; ++j)
              tma_load_2d(dst + j * (BK * 128), &map_a, full0 + 8 * slot, tm * BM + 32 * j, kb * BK);
#pragma unroll
            for (int j = 0; j < BN / 32; ++j)
              tma_load_2d(dst + BM * BK * 4 + j * (BK * 128), &map_b, full0 + 8 * slot, tn * BN + 32 * j, kb * BK);
          } else {
            tma_load_2d(dst, &map_a, full0 + 8 * slot, kb * BK, tm * BM);
            tma_load_2d(dst + BM * BK * 4, &map_b, full0 + 8 * slot, kb * BK, tn * BN);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // TN: both operands MN-major (bits 15 / 16)
      const uint32_t idesc = umma_idesc(2 /* TF32 */, BM, BN) | (TN ? ((1u << 15) | (1u << 16)) : 0u);
      uint32_t it = 0, lt = 0;
      for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++lt) {
        const uint32_t acc = lt & 1;
        int kb0, kb1;
        kb_range(w, kb0, kb1);
        mbar_wait(tempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a_addr = sm_addr + slot * STAGE_BYTES;
          const uint32_t b_addr = a_addr + BM * BK * 4;
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            if (TN)   // K = 8 rows = two 4-row swizzle groups per MMA; 32-float blocks along M / N are BK*128 B apart
              umma_tf32(tmem_base + acc * BN, umma_desc_mn_sw128_32b(a_addr + k8 * 1024, BK * 128),
                        umma_desc_mn_sw128_32b(b_addr + k8 * 1024, BK * 128), idesc, (kb > kb0 || k8 != 0) ? 1u : 0u);
            else
              umma_tf32(tmem_base + acc * BN, umma_desc_k_sw128(a_addr + k8 * 32), umma_desc_k_sw128(b_addr + k8 * 32),
                        idesc, (kb > kb0 || k8 != 0) ? 1u : 0u);
          }
          umma_commit(empty0 + 8 * slot);
        }
        umma_commit(tfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    const int lg = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t lt = 0;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x, ++lt) {
      const uint32_t acc = lt & 1;
      const int t = w % ntiles;
      const int tm = t % tiles_m, tn = t / tiles_m;
      const int row = tm * BM + lg * 32 + lane;
      mbar_wait(tfull0 + 8 * acc, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {
          tc_fence_before();
          mbar_arrive(tempty0 + 8 * acc);
        }
        const int col = tn * BN + c0;
        if (row < M && col < N) {
          float* cp = C + (int64_t)row * ldc + col;
          if (TN && nsplit > 1) {   // K slices of one output tile meet in C (zeroed / pre-loaded by the launcher)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col + j < N) atomicAdd(cp + j, __uint_as_float(v[j]));
          } else if (col + 32 <= N && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (accumulate) {
                const float4 p = *reinterpret_cast<const float4*>(cp + j);
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
              }
              *reinterpret_cast<float4*>(cp + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col + j < N) cp[j] = accumulate ? cp[j] + __uint_as_float(v[j]) : __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

int32_t gemm_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                int64_t N, int64_t K, int32_t accumulate, int32_t precision, cudaStream_t st) {
  if (precision != SGS_PREC_TF32) {
    set_error("sgs_gemm: the tensor-core GEMM runs kind::tf32 on fp32 operands (precision SGS_PREC_TF32)");
    return SGS_E_UNSUPPORTED;
  }
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) {
    set_error("sgs_gemm: TMA needs 16-byte aligned operand bases and row strides (lda, ldb multiples of 4)");
    return SGS_E_UNSUPPORTED;
  }
  CUtensorMap ma, mb;
  if (!make_map(&ma, A, M, K, lda) || !make_map(&mb, B, N, K, ldb)) {
    set_error("sgs_gemm: cuTensorMapEncodeTiled failed");
    return SGS_E_CUDA;
  }
  const size_t smem = (size_t)k4::NSTAGE * k4::STAGE_BYTES + 256 + 1024;
  SGS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = ceil_div(M, k4::BM) * ceil_div(N, k4::BN);
  const int64_t grid = ntiles < sm_count() ? ntiles : sm_count();
  const int nkb = (int)ceil_div(K, k4::BK);
  gemm_tf32_kernel<false><<<(unsigned)grid, k4::THREADS, smem, st>>>(ma, mb, C, ldc, (int)M, (int)N, (int)K,
                                                                      accumulate, 1, nkb);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

// C[M,N] (+)= A^T . B,  A stored [K, M] (row stride lda), B stored [K, N] (row stride ldb)
int32_t gemm_tc_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                   int64_t N, int64_t K, int32_t accumulate, int32_t precision, cudaStream_t st) {
  if (precision != SGS_PREC_TF32) {
    set_error("sgs_gemm: the tensor-core GEMM runs kind::tf32 on fp32 operands (precision SGS_PREC_TF32)");
    return SGS_E_UNSUPPORTED;
  }
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) {
    set_error("sgs_gemm: TMA needs 16-byte aligned operand bases and row strides (lda, ldb multiples of 4)");
    return SGS_E_UNSUPPORTED;
  }
  CUtensorMap ma, mb;
  // inner (contiguous) dimension = M / N, rows = K; boxes of [BK rows x 32 floats].  MN-major 32-bit operands
  // use the 128-byte swizzle with 32-byte atoms (16-byte atoms cannot be transposed at 32-bit granularity)
  if (!make_map(&ma, A, K, M, lda, k4::BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) ||
      !make_map(&mb, B, K, N, ldb, k4::BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) {
    set_error("sgs_gemm: cuTensorMapEncodeTiled failed");
    return SGS_E_CUDA;
  }
  const size_t smem = (size_t)k4::NSTAGE * k4::STAGE_BYTES + 256 + 1024;
  SGS_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = ceil_div(M, k4::BM) * ceil_div(N, k4::BN);
  const int nkb = (int)ceil_div(K, k4::BK);
  int64_t nsplit = sm_count() / ntiles;
  if (nsplit > nkb / 8) nsplit = nkb / 8;   // at least 8 K blocks (256 rows) per slice
  if (nsplit < 1) nsplit = 1;
  const int kb_per = (int)ceil_div(nkb, nsplit);
  nsplit = ceil_div(nkb, kb_per);
  if (nsplit > 1 && !accumulate)
    SGS_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  const int64_t nwork = ntiles * nsplit;
  const int64_t grid = nwork < sm_count() ? nwork : sm_count();
  gemm_tf32_kernel<true><<<(unsigned)grid, k4::THREADS, smem, st>>>(ma, mb, C, ldc, (int)M, (int)N, (int)K,
                                                                     accumulate, (int)nsplit, kb_per);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

}  // namespace sgs
