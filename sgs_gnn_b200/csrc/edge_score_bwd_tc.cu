// K1b tensor-core backward of the edge scorer: three fused tcgen05 kernels.  Nothing of size [q, 2H] is ever
// written to HBM; the only intermediate is the 16-bit gate gradient G [q, H],
//       G[e, j] = S * dz_e / (1 - p_drop) * [Z_ej + b1_j > 0] * keep_ej ,        dz = dp * p * (1 - p),
// i.e. the hidden-layer gradient dA = G . diag(w2) WITHOUT its w2 factor.  That factor is folded into operands and
// final epilogues, so no per-tile epilogue reduces anything across rows:
//
//  BA  (recompute; H = 256: the CTA-pair forward kernel in MODE 1, edge_score_tc2.cu; H = 128: the kernel below)
//        MMA : Z  = F . W1^T ;   EPI : G -> HBM (16 bit), db2 += sum dz
//  BF  (per 128-edge tile, per block of 128 node-embedding columns; diag(w2) . W1[:, cols] resident in smem and read
//       as an MN-major operand -- the row-major bytes "transposed" by the descriptor):
//        MMA : dF = G . (diag(w2) W1)[:, cols]                 (K = H hidden units)
//        EPI : d_out[src] += dF1*y + dF2  (segment-reduced over equal-src rows of a warp, then one coalesced RED)
//              d_out[dst] += dF1*x - dF2  (128-bit vector RED)
//  BW  P[BN, 2H] = G^T . F : both operands are edge-major tiles read as MN-major; the [BN x 2H] fp32 accumulator
//        stays in TMEM (512 columns) for the whole kernel; the loaders also keep per-column sums g_j = sum_e G[e, j].
//        Final epilogue (once per CTA), all from P and g:
//              dW1[j, :] += w2_j * P[j, :] / S
//              db1[j]    += w2_j * g_j / S
//              dw2[j]    += ( W1[j, :] . P[j, :] + b1_j * g_j ) / S
//        The last line is  sum_e dz_e * keep * relu(Z_ej + b1_j) / (1 - p)  rewritten with Z = F . W1^T: the hidden
//        activations never have to meet dz in a per-tile reduction.
//
// S = 2^k (from max|dp|) keeps fp16 G in range; it is divided out again in the epilogues.
#include "common.cuh"
#include "tc.cuh"
#include "scorer_producer.cuh"
#include "tma.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <type_traits>

namespace sgs {

int32_t edge_score_bwd_gate_pair(const void* tab, int32_t is_bf16, const int32_t* src, const int32_t* dst,
                                 const int32_t* ids, int64_t n, const float* W1, const float* b1, float p_drop,
                                 uint64_t seed, const float* p_fwd, const float* dp, const float* dp_absmax,
                                 void* g_out, float* db2, const int32_t* key_ids, cudaStream_t st);


namespace kb {
constexpr int TILE_M = 128;
constexpr int STAGE_BYTES = TILE_M * 128 * 2;  // 32 KB: two [128 x 64] 16-bit blocks
constexpr int EPI_WARPS = 4;
constexpr int PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int PROD_WARP0 = EPI_WARPS + 1;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
// the dW1 kernel has no per-tile epilogue: 16 loader warps + 1 MMA warp
constexpr int BW_MMA_WARP = 8;
constexpr int BW_THREADS = 17 * 32;
}  // namespace kb

// in-warp transpose-reduce (recursive halving): every lane holds v[0..15] (one row, 16 columns); on return lanes L
// and L^16 both hold the sum over all 32 rows of column (L & 15).
__device__ __forceinline__ float warp_colsum16(float* v, int lane) {
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = upper ? v[j] : v[j + half];
      const float keep = upper ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

template <typename T>
__global__ void convert_rows_kernel_b(const float* __restrict__ in, int64_t n8, uint4* __restrict__ outp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
    const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    uint4 o;
    o.x = Cvt<T>::pack_table(a.x, a.y);
    o.y = Cvt<T>::pack_table(a.z, a.w);
    o.z = Cvt<T>::pack_table(b.x, b.y);
    o.w = Cvt<T>::pack_table(b.z, b.w);
    outp[i] = o;
  }
}

// =============================================================================================
// BA: recompute + hidden-layer gradient
// =============================================================================================
template <typename T, int BN, int H>
__global__ void __launch_bounds__(kb::THREADS, 1)
edge_score_bwd_da_kernel(const T* __restrict__ tab, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ W1,
                         const float* __restrict__ b1, float p_drop, uint64_t seed,
                         const float* __restrict__ p_fwd, const float* __restrict__ dp,
                         const float* __restrict__ dp_absmax, T* __restrict__ g_out, float* __restrict__ db2,
                         const int32_t* __restrict__ key_ids) {
  using namespace kb;
  using namespace tc;
  constexpr int NB = H / BN;
  constexpr int NSP = H / 64;
  constexpr int NSTAGE = 3;
  constexpr int B_BLOCK_BYTES = BN * 128;
  constexpr int B_BYTES = 2 * NSP * B_BLOCK_BYTES;
  constexpr int TMEM_COLS = 2 * BN;
  static_assert(H % 64 == 0 && BN * NB == H && BN % 64 == 0 && BN <= 128, "unsupported shape");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  constexpr uint32_t kUsed = B_BYTES + NSTAGE * STAGE_BYTES + BN * 8 + 16 * 8 + 16;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (pad + kUsed > dyn_size) __trap();
  }
  const uint32_t b_base = sm_addr;
  const uint32_t a_base = b_base + B_BYTES;
  float* b1s = reinterpret_cast<float*>(sm + B_BYTES + NSTAGE * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(b1s + 2 * BN);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full0 = bar0, empty0 = bar0 + 32, zfull0 = bar0 + 64, zempty0 = bar0 + 80;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb = blockIdx.x % NB;
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  const int64_t tile0 = blockIdx.x / NB;
  const int64_t tstep = gridDim.x / NB;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, PROD_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(zfull0 + 8 * s, 1);
      mbar_init(zempty0 + 8 * s, EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(smem_u32(tmem_ptr_s), TMEM_COLS);
    tmem_relinquish();
  }
  for (int idx = threadIdx.x; idx < BN * (2 * H / 8); idx += THREADS) {
    const int nrow = idx / (2 * H / 8);
    const int kc = idx % (2 * H / 8);
    const int k0 = kc * 8;
    const int half = k0 >= H;
    const int kk = half ? k0 - H : k0;
    const int sp = kk >> 6;
    const int c16 = (kk & 63) >> 3;
    const float* g = W1 + (int64_t)(nb * BN + nrow) * (2 * H) + k0;
    const float4 a = *reinterpret_cast<const float4*>(g);
    const float4 b = *reinterpret_cast<const float4*>(g + 4);
    uint4 o;
    o.x = Cvt<T>::pack(a.x, a.y);
    o.y = Cvt<T>::pack(a.z, a.w);
    o.z = Cvt<T>::pack(b.x, b.y);
    o.w = Cvt<T>::pack(b.z, b.w);
    *reinterpret_cast<uint4*>(sm + (2 * sp + half) * B_BLOCK_BYTES + sw128_offset(nrow, c16)) = o;
  }
  for (int j = threadIdx.x; j < BN; j += THREADS) b1s[j] = b1[nb * BN + j];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp >= PROD_WARP0) {
    // ------------------------------- producers (as the forward) -------------------------------
    FeatureProducer<T, H, NSTAGE, STAGE_BYTES, TILE_M>::run(tab, src, dst, ids, n, tile0, tstep, ntiles,
                                                             sm + B_BYTES, full0, empty0,
                                                             threadIdx.x - PROD_WARP0 * 32);
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, TILE_M, BN);
      uint32_t it = 0, lt = 0;
      for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
        const uint32_t acc = lt & 1;
        mbar_wait(zempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll 1
        for (int sp = 0; sp < NSP; ++sp, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          fence_proxy_async_smem();   // producers' generic-proxy stores -> async proxy (see scorer_producer.cuh)
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t ad = umma_desc_k_sw128(a_base + slot * STAGE_BYTES + half * (TILE_M * 128) + k16 * 32);
              const uint64_t bd = umma_desc_k_sw128(b_base + (2 * sp + half) * B_BLOCK_BYTES + k16 * 32);
              umma_f16(d_tmem, ad, bd, idesc, (sp | half | k16) != 0 ? 1u : 0u);
            }
          umma_commit(empty0 + 8 * slot);
        }
        umma_commit(zfull0 + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: Z -> G (16 bit), db2 -------------------------------
    const int lg = warp & 3;        // one epilogue warp per TMEM lane quarter handles all BN columns
    const int r = lg * 32 + lane;
    const uint32_t thr32 = dropout_threshold32(dropout_threshold(p_drop));
    const bool drop = p_drop > 0.f;
    const float gscale = grad_scale(dp_absmax[0]) * (drop ? 1.0f / (1.0f - p_drop) : 1.0f);
    float acc_b2 = 0.f;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    uint32_t lt = 0;
    for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
      const uint32_t acc = lt & 1;
      const int64_t i = t * TILE_M + r;
      const bool live = i < n;
      float dz = 0.f;
      uint32_t rowkey = 0;
      if (live) {
        dz = dp[i];   // p_fwd == nullptr: dp already holds dz = dp * p * (1 - p)
        if (p_fwd) {
          const float pe = p_fwd[i];
          dz *= pe * (1.0f - pe);
        }
        if (drop) rowkey = dropout_rowkey(seed, (uint64_t)(key_ids ? key_ids[i] : (ids ? ids[i] : i)));
      }
      acc_b2 += dz;
      const float g = dz * gscale;
      const uint32_t gg = Cvt<T>::pack(g, g);
      const uint32_t g_lo = gg & 0xFFFFu, g_hi = gg & 0xFFFF0000u;
      mbar_wait(zfull0 + 8 * acc, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + acc * BN + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {  // accumulator fully read
          tc_fence_before();
          mbar_arrive(zempty0 + 8 * acc);
        }
        uint32_t o[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(b1s + c0 + j4 * 4);
          bool m0 = __uint_as_float(v[j4 * 4 + 0]) + bb.x > 0.f;
          bool m1 = __uint_as_float(v[j4 * 4 + 1]) + bb.y > 0.f;
          bool m2 = __uint_as_float(v[j4 * 4 + 2]) + bb.z > 0.f;
          bool m3 = __uint_as_float(v[j4 * 4 + 3]) + bb.w > 0.f;
          if (drop) {
            const uint32_t cp = (uint32_t)(nb * BN + c0 + j4 * 4) >> 1;
            const uint32_t xa = rowkey ^ dropout_colmix(cp);
            const uint32_t xb = rowkey ^ dropout_colmix(cp + 1);
            m0 = m0 && (xa * kDropMulEven >= thr32);
            m1 = m1 && (xa * kDropMulOdd >= thr32);
            m2 = m2 && (xb * kDropMulEven >= thr32);
            m3 = m3 && (xb * kDropMulOdd >= thr32);
          }
          o[2 * j4] = (m0 ? g_lo : 0u) | (m1 ? g_hi : 0u);
          o[2 * j4 + 1] = (m2 ? g_lo : 0u) | (m3 ? g_hi : 0u);
        }
        if (live) {
          uint4* gp = reinterpret_cast<uint4*>(g_out + i * H + nb * BN + c0);
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) gp[c8] = make_uint4(o[4 * c8], o[4 * c8 + 1], o[4 * c8 + 2], o[4 * c8 + 3]);
        }
      }
    }
    if (nb == 0) {
      acc_b2 = warp_sum(acc_b2);
      if (lane == 0) atomicAdd(db2, acc_b2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =============================================================================================
// BF: dF^T = (diag(w2) W1)[:, column block]^T . G^T   and the scatter into d_out
// =============================================================================================
// The product is computed TRANSPOSED: M = 128 node-embedding columns (TMEM lanes), N = 128 edges (TMEM columns).
// An epilogue thread therefore owns ONE embedding column and walks the tile's edges along its registers:
//   * the source-side sum over a run of equal sources (edge ids ascend by source) is a sequential in-register
//     accumulation -- no shuffles, no cross-lane reduction; one RED per run and warp,
//   * every RED (source and destination side) is a warp-coalesced 128-byte row segment: 32 lanes = 32 consecutive
//     columns of one d_out row, i.e. whole 32-byte sectors at the L2 atomic units,
//   * the gathers of x = tab[src], y = tab[dst] are 64-byte warp-coalesced 16-bit loads (the source row hits L1).
// A = this kind's columns of diag(w2) W1, resident in smem ([H x 64] blocks, MN-major through the descriptor);
// B = the G tile [128 e x 128 j] streamed by TMA (SWIZZLE_128B boxes) through a 3-stage ring -- no loader warps.
// Roles: warp 0 TMA producer (lane 0) + endpoint staging (all lanes: the tile's (src | run-start flag, dst) pairs
// go to shared memory, so the epilogue reads them with one broadcast LDS.64 per edge instead of shuffles),
// warp 1 MMA issue + TMEM allocation, warps 2-17 epilogue: 4 groups of 4 warps (one warp per TMEM lane quarter);
// group g drains accumulator buffer g & 1 (every other tile of this CTA), edges [64 (g >> 1), +64) of the tile.
namespace kbf {
constexpr int TILE_E = 128;                    // edges per tile (N of the MMA)
constexpr int CB = 128;                        // node-embedding columns per CTA kind (M of the MMA)
constexpr int STAGE_BYTES = TILE_E * 128 * 2;  // [128 e x 128 j] 16-bit = two [128 x 64] boxes
constexpr int NSTAGE = 3;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_GROUPS = 4;
constexpr int THREADS = (EPI_WARP0 + 4 * EPI_GROUPS) * 32;   // 576
constexpr int CHUNK = 4;                       // edges per epilogue chunk
constexpr int AHEAD = 3;                       // chunks of x / y gathers in flight ahead of the chunk being reduced
}  // namespace kbf

template <typename T>
__device__ __forceinline__ float tab_to_float(unsigned short u);
template <>
__device__ __forceinline__ float tab_to_float<__half>(unsigned short u) { return __half2float(__ushort_as_half(u)); }
template <>
__device__ __forceinline__ float tab_to_float<__nv_bfloat16>(unsigned short u) {
  return __uint_as_float((uint32_t)u << 16);
}

// 32 lanes x 4 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

// 18 warps = 5 on two of the four SM sub-partitions (16 K registers each): at most 96 registers per thread
template <typename T, int H>
__global__ void __launch_bounds__(kbf::THREADS, 1)
edge_score_bwd_df_kernel(const __grid_constant__ CUtensorMap map_g, const T* __restrict__ tab,
                         const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ W1,
                         const float* __restrict__ w2, const float* __restrict__ dp_absmax,
                         float* __restrict__ d_out) {
  using namespace kbf;
  using namespace tc;
  constexpr int NKIND = H / CB;
  constexpr int NQ = CB / 64;                   // 64-column blocks per part (product / difference)
  constexpr int BLK = H * 128;                  // one [H rows (j) x 64 columns] 16-bit block
  constexpr int A_BYTES = 2 * NQ * BLK;         // resident operand
  constexpr int NJ = H / 128;                   // G sub-tiles (128 hidden units each) per edge tile
  static_assert(H % 128 == 0 && H <= 256, "unsupported shape");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  constexpr uint32_t kUsed = A_BYTES + NSTAGE * STAGE_BYTES + 16 * 8 + 16 + 2 * TILE_E * 8;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (pad + kUsed > dyn_size) __trap();
  }
  const uint32_t a_base = sm_addr;               // resident diag(w2) W1 blocks
  const uint32_t g_base = a_base + A_BYTES;      // G stage ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + A_BYTES + NSTAGE * STAGE_BYTES);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 16);
  int2* ep_s = reinterpret_cast<int2*>(sm + A_BYTES + NSTAGE * STAGE_BYTES + 16 * 8 + 16);   // [2][TILE_E]
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full0 = bar0, empty0 = bar0 + 32, dffull0 = bar0 + 64, dfempty0 = bar0 + 80, epfull0 = bar0 + 96;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kind = blockIdx.x % NKIND;
  const int64_t ntiles = (n + TILE_E - 1) / TILE_E;
  const int64_t tile0 = blockIdx.x / NKIND;
  const int64_t tstep = gridDim.x / NKIND;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, 1);       // the producer's arrive.expect_tx; TMA completes the bytes
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(dffull0 + 8 * s, 1);
      mbar_init(dfempty0 + 8 * s, (EPI_GROUPS / 2) * 4 * 32);
      mbar_init(epfull0 + 8 * s, 32);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr_s), 512);
    tmem_relinquish();
  }
  // resident operand: rows j = 0..H-1 of diag(w2) . W1, this kind's columns, as blocks [H x 64]:
  //   block qd      : product-part columns    kind*CB + 64*qd + [0,64)
  //   block NQ + qd : difference-part columns H + kind*CB + 64*qd + [0,64)
  for (int idx = threadIdx.x; idx < H * (2 * CB / 8); idx += THREADS) {
    const int j = idx / (2 * CB / 8);
    const int kc = idx % (2 * CB / 8);          // 16-byte chunk among this kind's 2*CB columns
    const int blk = kc >> 3;                    // 0 .. 2*NQ-1: blk = half * NQ + qd
    const int half = blk / NQ, qd = blk % NQ;
    const int c16 = kc & 7;
    const int kcol = (half ? H : 0) + kind * CB + qd * 64 + c16 * 8;
    const float* g = W1 + (int64_t)j * (2 * H) + kcol;
    const float wj = w2[j];
    const float4 a = *reinterpret_cast<const float4*>(g);
    const float4 b = *reinterpret_cast<const float4*>(g + 4);
    uint4 o;
    o.x = Cvt<T>::pack(wj * a.x, wj * a.y);
    o.y = Cvt<T>::pack(wj * a.z, wj * a.w);
    o.z = Cvt<T>::pack(wj * b.x, wj * b.y);
    o.w = Cvt<T>::pack(wj * b.z, wj * b.w);
    *reinterpret_cast<uint4*>(sm + blk * BLK + sw128_offset(j, c16)) = o;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // --------------- TMA producer (lane 0): G sub-tiles [128 e x 128 j];  endpoint staging (all lanes) ---------------
    // lane l stages edges 4l .. 4l+3 of the tile as (src | run-start flag << 31, dst): the flag marks an edge whose
    // source differs from its predecessor's inside the same 64-edge half (the first edge of a half is never
    // flagged: the epilogue starts a fresh run there).  Dead edges (past n) become (0, 0): their G rows are zero.
    uint32_t it = 0, lt = 0;
    for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
      if (lane == 0) {
#pragma unroll 1
        for (int js = 0; js < NJ; ++js, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
          mbar_expect_tx(full0 + 8 * slot, STAGE_BYTES);
          const uint32_t dst_s = g_base + slot * STAGE_BYTES;
          // rows past n are zero-filled by TMA: dead edges contribute exact zeros
          tma_load_2d(dst_s, &map_g, full0 + 8 * slot, js * 128, (int)(t * TILE_E));
          tma_load_2d(dst_s + TILE_E * 128, &map_g, full0 + 8 * slot, js * 128 + 64, (int)(t * TILE_E));
        }
      }
      __syncwarp();
      int sv[4], dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t i = t * TILE_E + 4 * lane + k;
        sv[k] = dv[k] = 0;
        if (i < n) {
          const int64_t e = ids ? ids[i] : i;
          sv[k] = src[e];
          dv[k] = dst[e];
        }
      }
      const int prev = __shfl_up_sync(0xffffffffu, sv[3], 1);
      const uint32_t tb = lt & 1;
      mbar_wait(dfempty0 + 8 * tb, ((lt >> 1) & 1) ^ 1);   // the epilogue of tile lt - 2 has read its endpoints
      int2* ep = ep_s + tb * TILE_E + 4 * lane;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int before = k == 0 ? prev : sv[k - 1];
        const bool start = (k != 0 || (lane & 15) != 0) && sv[k] != before;
        ep[k] = make_int2(sv[k] | (start ? (int)0x80000000u : 0), dv[k]);
      }
      mbar_arrive(epfull0 + 8 * tb);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, CB, TILE_E) | (1u << 15);  // A is MN-major
      uint32_t it = 0, lt = 0;
      for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
        const uint32_t tb = lt & 1;
        mbar_wait(dfempty0 + 8 * tb, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int js = 0; js < NJ; ++js, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t ad = umma_desc_mn_sw128(a_base + (half * NQ) * BLK + (js * 8 + kk) * 2048, BLK);
              const uint64_t bd =
                  umma_desc_k_sw128(g_base + slot * STAGE_BYTES + (kk >> 2) * (TILE_E * 128) + (kk & 3) * 32);
              umma_f16(tmem_base + tb * 256 + half * 128, ad, bd, idesc, (js | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8 * slot);
        }
        umma_commit(dffull0 + 8 * tb);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: dF^T -> d_out[src], d_out[dst] -------------------------------
    const int lg = warp & 3;                        // TMEM lane quarter this warp may read
    const int grp = (warp - EPI_WARP0) >> 2;
    const uint32_t tb = grp & 1;
    const int eh = grp >> 1;                        // which 64 edges of the tile
    const int c = kind * CB + lg * 32 + lane;       // this thread's node-embedding column
    const unsigned short* tabc = reinterpret_cast<const unsigned short*>(tab) + c;
    float* doc = d_out + c;
    const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + tb * 256 + eh * 64;
    const float invS = 1.0f / grad_scale(dp_absmax[0]);
    const int2* ep = ep_s + tb * TILE_E + eh * 64;
    constexpr uint64_t ROWB2 = (uint64_t)H * 2, ROWB4 = (uint64_t)H * 4;
    const char* tabb = reinterpret_cast<const char*>(tabc);
    char* docb = reinterpret_cast<char*>(doc);
    uint32_t lt = tb;
    for (int64_t t = tile0 + tb * tstep; t < ntiles; t += 2 * tstep, lt += 2) {
      mbar_wait(epfull0 + 8 * tb, (lt >> 1) & 1);
      // software pipeline over chunks of CHUNK edges: the x / y gathers of the next AHEAD chunks are in flight while
      // chunk ch is reduced.  The epilogue is bound by the latency of these 2-byte gathers (ncu r02a: long-scoreboard
      // 51 %; tile time = chunks x latency / lookahead), so the lookahead distance is what sets BF's duration: 4 chunks
      // of 4 edges (16 edges ahead, 32 loads in flight per thread) in the registers that 2 chunks of 8 used before.
      constexpr int NCH = 64 / CHUNK;
      constexpr int NBUF = AHEAD + 1;
      static_assert(CHUNK == 4, "tmem_ld4 below");
      unsigned short xv[NBUF][CHUNK], yv[NBUF][CHUNK];
      auto issue = [&](int ch, unsigned short* xc, unsigned short* yc) {
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
          const int2 sd = ep[ch * CHUNK + j];
          xc[j] = __ldg(reinterpret_cast<const unsigned short*>(tabb + (uint32_t)(sd.x & 0x7FFFFFFF) * ROWB2));
          yc[j] = __ldg(reinterpret_cast<const unsigned short*>(tabb + (uint32_t)sd.y * ROWB2));
        }
      };
#pragma unroll
      for (int a = 0; a < AHEAD; ++a) issue(a, xv[a], yv[a]);
      uint32_t cur_s = (uint32_t)ep[0].x;           // never flagged (first edge of the half)
      float xcur = tab_to_float<T>(xv[0][0]);
      float acc = 0.f;                              // S-scaled source-side run sum
      mbar_wait(dffull0 + 8 * tb, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int cb = ch % NBUF;
        uint32_t f1[CHUNK], f2[CHUNK];
        tmem_ld4(taddr0 + ch * CHUNK, f1);
        tmem_ld4(taddr0 + 128 + ch * CHUNK, f2);
        if (ch + AHEAD < NCH) issue(ch + AHEAD, xv[(ch + AHEAD) % NBUF], yv[(ch + AHEAD) % NBUF]);
        tmem_ld_wait();
        int2 sd[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) sd[j] = ep[ch * CHUNK + j];
        if (ch == NCH - 1) {  // this warp has read all of its part of the accumulator and of the endpoints
          tc_fence_before();
          mbar_arrive(dfempty0 + 8 * tb);
        }
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
          if (sd[j].x < 0) {            // warp-uniform: a new run of equal sources starts
            atomicAdd(reinterpret_cast<float*>(docb + cur_s * ROWB4), acc * invS);
            acc = 0.f;
            cur_s = (uint32_t)sd[j].x & 0x7FFFFFFFu;
            xcur = tab_to_float<T>(xv[cb][j]);
          }
          const float a = __uint_as_float(f1[j]), b = __uint_as_float(f2[j]);
          acc += fmaf(a, tab_to_float<T>(yv[cb][j]), b);
          atomicAdd(reinterpret_cast<float*>(docb + (uint32_t)sd[j].y * ROWB4), fmaf(a, xcur, -b) * invS);
        }
      }
      atomicAdd(reinterpret_cast<float*>(docb + cur_s * ROWB4), acc * invS);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================
// BW: P[BN, 2H] = G^T . F over all edges of this CTA's tiles, g = column sums of G;  dW1, db1, dw2 from P and g
// =============================================================================================
template <typename T, int BN, int H>
__global__ void __launch_bounds__(kb::BW_THREADS, 1)
edge_score_bwd_dw_kernel(const T* __restrict__ tab, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ ids, int64_t n, const T* __restrict__ dA,
                         const float* __restrict__ dp_absmax, const float* __restrict__ W1,
                         const float* __restrict__ b1, const float* __restrict__ w2, float* __restrict__ dW1,
                         float* __restrict__ db1, float* __restrict__ dw2) {
  using namespace kb;
  using namespace tc;
  constexpr int NB = H / BN;
  constexpr int SUB_M = 64;                       // edges per stage (K of the MMA)
  constexpr int F_BYTES = SUB_M * 2 * H * 2;      // [64 e x 2H] as 2H/64 blocks of [64 x 64]
  constexpr int DA_BYTES = SUB_M * BN * 2;
  constexpr int STAGE = F_BYTES + DA_BYTES;
  constexpr int NSTAGE = (2 * STAGE + 4096 <= 232448) ? 2 : 1;
  constexpr int NCOLS = 2 * H;                    // TMEM columns of the accumulator
  static_assert(NCOLS <= 512 && BN == 128, "unsupported shape");
  constexpr int TMEM_ALLOC = NCOLS <= 128 ? 128 : (NCOLS <= 256 ? 256 : 512);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (pad + NSTAGE * STAGE + 128 + BN * 4 > dyn_size) __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + NSTAGE * STAGE);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 8);
  float* gsum_s = reinterpret_cast<float*>(sm + NSTAGE * STAGE + 128);   // [BN] column sums of G
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 16, done_bar = full0 + 32;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb = blockIdx.x % NB;
  const int64_t nsub = (n + SUB_M - 1) / SUB_M;
  const int64_t sub0 = blockIdx.x / NB;
  const int64_t sstep = gridDim.x / NB;
  constexpr int LOAD_THREADS = BW_THREADS - 32;      // every warp but the MMA warp fills stages

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full0 + 8 * s, LOAD_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  for (int j = threadIdx.x; j < BN; j += BW_THREADS) gsum_s[j] = 0.f;
  if (warp == BW_MMA_WARP) {
    tmem_alloc(smem_u32(tmem_ptr_s), TMEM_ALLOC);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == BW_MMA_WARP) {
    if (lane == 0) {
      // A = dA^T [M = BN (j) x K = 64 (e)]  MN-major;  B = F^T [N = 2H (k) x K = 64 (e)]  MN-major
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, BN, NCOLS > 256 ? 256 : NCOLS) | (1u << 15) | (1u << 16);
      uint32_t it = 0;
      for (int64_t s = sub0; s < nsub; s += sstep, ++it) {
        const uint32_t slot = it % NSTAGE;
        mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t f_base = sm_addr + slot * STAGE;
        const uint32_t d_base = f_base + F_BYTES;
#pragma unroll
        for (int nh = 0; nh < (NCOLS + 255) / 256; ++nh) {
#pragma unroll
          for (int kk = 0; kk < SUB_M / 16; ++kk) {
            const uint64_t ad = umma_desc_mn_sw128(d_base + kk * 2048, SUB_M * 128);
            const uint64_t bd = umma_desc_mn_sw128(f_base + nh * 4 * (SUB_M * 128) + kk * 2048, SUB_M * 128);
            umma_f16(tmem_base + nh * 256, ad, bd, idesc, (it | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8 * slot);
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // ---- stage fillers: F sub-tile (gather + product/difference) and the dA sub-tile from HBM ----
    const int lt_id = threadIdx.x < BW_MMA_WARP * 32 ? threadIdx.x : threadIdx.x - 32;  // 0 .. LOAD_THREADS-1
    constexpr int CH = H / 8;                            // 16-byte chunks per node-embedding row
    constexpr int NIT = SUB_M * CH / LOAD_THREADS;       // F items per thread and stage
    constexpr int NDA = SUB_M * (BN / 8) / LOAD_THREADS; // dA items per thread and stage
    static_assert(SUB_M * CH % LOAD_THREADS == 0 && SUB_M * (BN / 8) % LOAD_THREADS == 0, "loader mapping");
    // every G item of a thread is the same 16-byte chunk (8 hidden units) of some row: LOAD_THREADS % (BN / 8) == 0
    static_assert(LOAD_THREADS % (BN / 8) == 0, "column-sum mapping");
    float gs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gs[k] = 0.f;
    // thread -> 16-byte chunk cc of the NIT consecutive rows NIT * (lt_id / CH) + k: consecutive edges mostly share
    // their source (runs of ~10 in bucket order), whose row chunk is then loaded once
    const int cc = lt_id % CH;
    const int row0 = NIT * (lt_id / CH);
    auto load_idx = [&](int64_t s, int* sn, int* dn) {
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int64_t i = s * SUB_M + row0 + k;
        sn[k] = -1;
        dn[k] = 0;
        if (i < n) {
          const int64_t e = ids ? ids[i] : i;
          sn[k] = src[e];
          dn[k] = dst[e];
        }
      }
    };
    int sn[NIT], dn[NIT], sn2[NIT], dn2[NIT];
    if (sub0 < nsub) load_idx(sub0, sn, dn);
    uint32_t it = 0;
    for (int64_t s = sub0; s < nsub; s += sstep, ++it) {
      // all global loads of the stage are issued before the stage slot is waited for; the endpoints of the next
      // stage are prefetched behind them
      uint4 xv[NIT], yv[NIT], av[NDA];
      bool same = true;
#pragma unroll
      for (int k = 1; k < NIT; ++k) same = same && (sn[k] == sn[0]);
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        xv[k] = yv[k] = make_uint4(0, 0, 0, 0);
        if (sn[k] >= 0) {
          if (k == 0 || !same) xv[k] = *reinterpret_cast<const uint4*>(tab + (int64_t)sn[k] * H + cc * 8);
          yv[k] = *reinterpret_cast<const uint4*>(tab + (int64_t)dn[k] * H + cc * 8);
        }
      }
      if (same) {
#pragma unroll
        for (int k = 1; k < NIT; ++k) xv[k] = xv[0];
      }
#pragma unroll
      for (int k = 0; k < NDA; ++k) {
        const int item = lt_id + k * LOAD_THREADS;
        const int64_t i = s * SUB_M + item / (BN / 8);
        av[k] = make_uint4(0, 0, 0, 0);
        if (i < n) av[k] = ld_stream_u4(reinterpret_cast<const uint4*>(dA + i * H + nb * BN + (item % (BN / 8)) * 8));
      }
      if (s + sstep < nsub) load_idx(s + sstep, sn2, dn2);
      const uint32_t slot = it % NSTAGE;
      mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
      uint8_t* fst = sm + slot * STAGE;
      uint8_t* dst_da = fst + F_BYTES;
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int row = row0 + k;
        const uint32_t off = sw128_offset(row, cc & 7);
        *reinterpret_cast<uint4*>(fst + (2 * (cc >> 3)) * (SUB_M * 128) + off) =
            make_uint4(Cvt<T>::mul2(xv[k].x, yv[k].x), Cvt<T>::mul2(xv[k].y, yv[k].y),
                       Cvt<T>::mul2(xv[k].z, yv[k].z), Cvt<T>::mul2(xv[k].w, yv[k].w));
        *reinterpret_cast<uint4*>(fst + (2 * (cc >> 3) + 1) * (SUB_M * 128) + off) =
            make_uint4(Cvt<T>::sub2(xv[k].x, yv[k].x), Cvt<T>::sub2(xv[k].y, yv[k].y),
                       Cvt<T>::sub2(xv[k].z, yv[k].z), Cvt<T>::sub2(xv[k].w, yv[k].w));
      }
#pragma unroll
      for (int k = 0; k < NDA; ++k) {
        const int item = lt_id + k * LOAD_THREADS;
        const int row = item / (BN / 8), cc = item % (BN / 8);
        *reinterpret_cast<uint4*>(dst_da + (cc >> 3) * (SUB_M * 128) + sw128_offset(row, cc & 7)) = av[k];
      }
      fence_proxy_async_smem();
      mbar_arrive(full0 + 8 * slot);
#pragma unroll
      for (int k = 0; k < NDA; ++k) {
        const uint32_t w[4] = {av[k].x, av[k].y, av[k].z, av[k].w};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const float2 f = Cvt<T>::unpack(w[m]);
          gs[2 * m] += f.x;
          gs[2 * m + 1] += f.y;
        }
      }
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        sn[k] = sn2[k];
        dn[k] = dn2[k];
      }
    }
    {
      const int gc = lt_id % (BN / 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(gsum_s + gc * 8 + k, gs[k]);
    }
    // all loader warps (everything but the MMA warp) have added their column sums
    asm volatile("bar.sync 1, %0;" ::"n"(LOAD_THREADS) : "memory");
    // ---- final epilogue: P, g -> dW1 (original column order), db1, dw2; the scale S divided out ----
    if (warp < 4) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const float invS = 1.0f / grad_scale(dp_absmax[0]);
      const int jrow = warp * 32 + lane;   // TMEM lane == hidden unit inside the block
      const int j = nb * BN + jrow;
      const float wj = w2[j] * invS;
      const float* w1row = W1 + (int64_t)j * (2 * H);
      float* wrow = dW1 + (int64_t)j * (2 * H);
      float dot = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < NCOLS; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        // accumulator column n: block b = n / 64 (even: product part, odd: difference part), group b / 2
        const int b = c0 >> 6;
        const int kcol = ((b & 1) ? H : 0) + (b >> 1) * 64 + (c0 & 63);
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const float4 w = *reinterpret_cast<const float4*>(w1row + kcol + k);
          const float p0 = __uint_as_float(v[k]), p1 = __uint_as_float(v[k + 1]);
          const float p2 = __uint_as_float(v[k + 2]), p3 = __uint_as_float(v[k + 3]);
          dot = fmaf(w.x, p0, fmaf(w.y, p1, fmaf(w.z, p2, fmaf(w.w, p3, dot))));
          red_add_v4(wrow + kcol + k, p0 * wj, p1 * wj, p2 * wj, p3 * wj);
        }
      }
      const float gj = gsum_s[jrow];
      atomicAdd(db1 + j, wj * gj);
      atomicAdd(dw2 + j, (dot + b1[j] * gj) * invS);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BW_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_ALLOC);
  }
}

// =============================================================================================
// Pre-pass: destination-range bucketing of the edge list
// =============================================================================================
// The backward touches, per edge, the rows tab[dst] (gather) and d_out[dst] (RED) of a random destination: with
// N x H x 6 bytes of such rows (358 MB at Reddit scale) against 126 MB of L2, ~3/4 of the RED sectors and most
// gathers missed (ncu r01g: BF 46 GB of DRAM traffic for 12 GB of useful G reads).  The edges are therefore
// processed grouped by destination RANGE (2^r rows whose tab + d_out rows fit in L2), stably -- inside a bucket
// edge ids still ascend by source, so the source-side run sums keep working (runs get shorter).  All three
// kernels just see a permuted id list ids_b plus dz in the same order; every output is an order-independent sum.
__global__ void bucket_key_kernel(const int32_t* __restrict__ dst, const int32_t* __restrict__ ids, int64_t n,
                                  int shift, uint8_t* __restrict__ key, int32_t* __restrict__ pos) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t e = ids ? ids[i] : i;
    key[i] = (uint8_t)(dst[e] >> shift);
    pos[i] = (int32_t)i;
  }
}
// in bucket order: ids_b[i'] = edge id (dropout mask key), its endpoints src_b / dst_b (so that no kernel chases
// ids -> src/dst on its critical path) and dz_b[i'] = dp * p * (1 - p)
__global__ void bucket_gather_kernel(const int32_t* __restrict__ pos, const int32_t* __restrict__ ids,
                                     const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                     const float* __restrict__ p_fwd, const float* __restrict__ dp, int64_t n,
                                     int32_t* __restrict__ ids_b, int32_t* __restrict__ src_b,
                                     int32_t* __restrict__ dst_b, float* __restrict__ dz_b) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int64_t o = pos ? pos[i] : i;
    const int64_t e = ids ? ids[o] : o;
    ids_b[i] = (int32_t)e;
    src_b[i] = src[e];
    dst_b[i] = dst[e];
    const float pe = p_fwd[o];
    dz_b[i] = dp[o] * pe * (1.0f - pe);
  }
}

static size_t bucket_temp_bound(int64_t n) { return (size_t)(32u << 20) + (size_t)(n / 8) * 4; }
// rows per bucket = 2^shift: tab (2 B) + d_out (4 B) rows of one bucket within ~64 MB; at most 256 buckets
static int bucket_shift(int64_t N, int64_t H) {
  int shift = 0;
  while (((int64_t)2 << shift) * H * 6 <= ((int64_t)64 << 20)) ++shift;
  while (((N - 1) >> shift) > 255) ++shift;
  return shift;
}

// =============================================================================================
// host side
// =============================================================================================
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
size_t edge_score_bwd_tc_workspace_bytes(int64_t n, int64_t N, int64_t H) {
  // absmax | tab [N,H] 16 bit | G [n,H] 16 bit | ids_b, src_b, dst_b, dz_b, pos_a, pos_b [n] 32 bit |
  // key_a, key_b [n] 8 bit | cub temp
  return 4096 + align256((size_t)N * H * 2) + align256((size_t)n * H * 2) + 6 * align256((size_t)n * 4) +
         2 * align256((size_t)n) + bucket_temp_bound(n);
}

template <typename T, int H>
static int32_t launch_bwd(const float* out, int64_t N, const int32_t* src, const int32_t* dst, const int32_t* ids,
                          int64_t n, const float* W1, const float* b1, const float* w2, float p_drop, uint64_t seed,
                          const float* p_fwd, const float* dp, float* d_out, float* dW1, float* db1, float* dw2,
                          float* db2, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int BN = 128;
  constexpr int NB = H / BN;
  if (ws_bytes < edge_score_bwd_tc_workspace_bytes(n, N, H)) {
    set_error("sgs_edge_score_bwd: workspace too small");
    return SGS_E_WORKSPACE;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  uint8_t* cur = base;
  auto take = [&](size_t bytes) { uint8_t* p = cur; cur += align256(bytes); return p; };
  float* absmax = reinterpret_cast<float*>(take(256));
  T* tab = reinterpret_cast<T*>(take((size_t)N * H * 2));
  T* dA = reinterpret_cast<T*>(take((size_t)n * H * 2));
  int32_t* ids_b = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* src_b = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* dst_b = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  float* dz_b = reinterpret_cast<float*>(take((size_t)n * 4));
  int32_t* pos_a = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  int32_t* pos_b = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  uint8_t* key_a = take((size_t)n);
  uint8_t* key_b = take((size_t)n);
  void* cub_temp = cur;
  const size_t cub_avail = ws_bytes - (size_t)(cur - reinterpret_cast<uint8_t*>(ws));
  SGS_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
  const int64_t cap = (int64_t)sm_count() * 16;
  int64_t g = ceil_div(n, 256);
  absmax_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(dp, n, absmax);
  SGS_LAUNCH_CHECK();
  const int64_t n8 = N * H / 8;
  g = ceil_div(n8, 256);
  convert_rows_kernel_b<T><<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(out, n8, reinterpret_cast<uint4*>(tab));
  SGS_LAUNCH_CHECK();
  // ---- destination-range bucketing: ids_b (bucket order, stable), dz_b ----
  {
    const int shift = bucket_shift(N, H);
    const int nbuckets = (int)(((N - 1) >> shift) + 1);
    const int32_t* pos = nullptr;
    g = ceil_div(n, 256);
    const unsigned eg = (unsigned)(g > cap ? cap : g);
    if (nbuckets > 1) {
      int bits = 1;
      while ((1 << bits) < nbuckets) ++bits;
      bucket_key_kernel<<<eg, 256, 0, st>>>(dst, ids, n, shift, key_a, pos_a);
      SGS_LAUNCH_CHECK();
      cub::DoubleBuffer<uint8_t> dk(key_a, key_b);
      cub::DoubleBuffer<int32_t> dv(pos_a, pos_b);
      size_t need = 0;
      SGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, bits, st));
      if (need > cub_avail) {
        set_error("sgs_edge_score_bwd: cub temp %zu > available %zu", need, cub_avail);
        return SGS_E_WORKSPACE;
      }
      SGS_CUDA(cub::DeviceRadixSort::SortPairs(cub_temp, need, dk, dv, (int)n, 0, bits, st));
      count_launch(2);
      pos = dv.Current();
    }
    bucket_gather_kernel<<<eg, 256, 0, st>>>(pos, ids, src, dst, p_fwd, dp, n, ids_b, src_b, dst_b, dz_b);
    SGS_LAUNCH_CHECK();
    // from here on the edge list is (src_b, dst_b) in bucket order, addressed directly
    src = src_b;
    dst = dst_b;
    ids = nullptr;
    dp = dz_b;
    p_fwd = nullptr;  // dp holds dz
  }
  const int64_t ntiles = ceil_div(n, kb::TILE_M);
  auto grid_for = [&](int kinds, int64_t items) {
    int64_t gr = (int64_t)(sm_count() / kinds) * kinds;
    if (gr > items * kinds) gr = items * kinds;
    return (unsigned)gr;
  };
  if (H == 256 && n >= 2 * kb::TILE_M) {
    // CTA pairs: every edge tile is gathered and built once (edge_score_tc2.cu, MODE 1)
    const int32_t rc = edge_score_bwd_gate_pair(tab, std::is_same<T, __nv_bfloat16>::value ? 1 : 0, src, dst, ids, n,
                                                W1, b1, p_drop, seed, p_fwd, dp, absmax, dA, db2, ids_b, st);
    if (rc != SGS_OK) return rc;
  } else {
    auto kern = edge_score_bwd_da_kernel<T, BN, H>;
    constexpr size_t used = (size_t)2 * (H / 64) * BN * 128 + 3 * kb::STAGE_BYTES + BN * 8 + 16 * 8 + 16;
    const size_t smem = used + 1024 > 232448 ? 232448 : used + 1024;
    SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for(NB, ntiles), kb::THREADS, smem, st>>>(tab, src, dst, ids, n, W1, b1, p_drop, seed, p_fwd, dp,
                                                          absmax, dA, db2, ids_b);
    SGS_LAUNCH_CHECK();
  }
  {
    auto kern = edge_score_bwd_df_kernel<T, H>;
    CUtensorMap map_g;
    if (!make_map_16bit(&map_g, dA, n, H, H, kbf::TILE_E, std::is_same<T, __nv_bfloat16>::value)) {
      set_error("sgs_edge_score_bwd: cuTensorMapEncodeTiled failed");
      return SGS_E_CUDA;
    }
    constexpr size_t used =
        (size_t)2 * 2 * H * 128 + kbf::NSTAGE * kbf::STAGE_BYTES + 16 * 8 + 16 + 2 * kbf::TILE_E * 8;
    static_assert(used <= 232448, "BF shared memory");
    const size_t smem = used + 1024 > 232448 ? 232448 : used + 1024;   // the kernel traps if its alignment pad does not fit
    SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for(H / 128, ntiles), kbf::THREADS, smem, st>>>(map_g, tab, src, dst, ids, n, W1, w2, absmax, d_out);
    SGS_LAUNCH_CHECK();
  }
  {
    auto kern = edge_score_bwd_dw_kernel<T, BN, H>;
    constexpr size_t stage = (size_t)64 * 2 * H * 2 + 64 * BN * 2;
    constexpr int nstage = (2 * stage + 4096 <= 232448) ? 2 : 1;
    const size_t smem = nstage * stage + 128 + BN * 4 + 1024;
    SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for(NB, ceil_div(n, 64)), kb::BW_THREADS, smem, st>>>(tab, src, dst, ids, n, dA, absmax, W1, b1, w2,
                                                                      dW1, db1, dw2);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t edge_score_bwd_tc(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                          const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                          float p_drop, uint64_t seed, const float* p_fwd, const float* dp, float* d_out, float* dW1,
                          float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes, int32_t precision,
                          cudaStream_t st) {
#define SGS_KB(T, HH)                                                                                          \
  return launch_bwd<T, HH>(out, N, src, dst, ids, n, W1, b1, w2, p_drop, seed, p_fwd, dp, d_out, dW1, db1, dw2, \
                           db2, ws, ws_bytes, st)
  if (precision == SGS_PREC_BF16) {
    if (H == 256) SGS_KB(__nv_bfloat16, 256);
    if (H == 128) SGS_KB(__nv_bfloat16, 128);
  } else if (precision == SGS_PREC_FP16) {
    if (H == 256) SGS_KB(__half, 256);
    if (H == 128) SGS_KB(__half, 128);
  }
#undef SGS_KB
  set_error("sgs_edge_score_bwd: tensor-core path supports bf16/fp16 and H in {128, 256}");
  return SGS_E_UNSUPPORTED;
}

}  // namespace sgs
