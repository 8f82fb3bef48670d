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

#include <type_traits>

namespace sgs {

int32_t edge_score_bwd_gate_pair(const void* tab, int32_t is_bf16, const int32_t* src, const int32_t* dst,
                                 const int32_t* ids, int64_t n, const float* W1, const float* b1, float p_drop,
                                 uint64_t seed, const float* p_fwd, const float* dp, const float* dp_absmax,
                                 void* g_out, float* db2, cudaStream_t st);


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
                         const float* __restrict__ dp_absmax, T* __restrict__ g_out, float* __restrict__ db2) {
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
        const float pe = p_fwd[i];
        dz = dp[i] * pe * (1.0f - pe);
        if (drop) rowkey = dropout_rowkey(seed, (uint64_t)(ids ? ids[i] : i));
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
// BF: dF = G . (diag(w2) W1)[:, column block]  and the scatter into d_out
// =============================================================================================
template <typename T, int H>
__global__ void __launch_bounds__(kb::THREADS, 1)
edge_score_bwd_df_kernel(const T* __restrict__ tab, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ W1,
                         const float* __restrict__ w2, const T* __restrict__ dA, const float* __restrict__ dp_absmax,
                         float* __restrict__ d_out) {
  using namespace kb;
  using namespace tc;
  constexpr int CB = 128;                       // node-embedding columns per CTA kind
  constexpr int NKIND = H / CB;
  constexpr int NQ = CB / 64;                   // 64-column groups (each: product block + difference block)
  constexpr int BLK = H * 128;                  // one [H rows (j) x 64 k] 16-bit block
  constexpr int B_BYTES = 2 * NQ * BLK;
  constexpr int NJ = H / 128;                   // dA sub-tiles (128 hidden units each) per edge tile
  constexpr int NSTAGE = 3;
  static_assert(H % 128 == 0 && H <= 256, "unsupported shape");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  constexpr uint32_t kUsed = B_BYTES + NSTAGE * STAGE_BYTES + 16 * 8 + 16;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (pad + kUsed > dyn_size) __trap();
  }
  const uint32_t b_base = sm_addr;
  const uint32_t a_base = b_base + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + B_BYTES + NSTAGE * STAGE_BYTES);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full0 = bar0, empty0 = bar0 + 32, dffull0 = bar0 + 64, dfempty0 = bar0 + 80;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kind = blockIdx.x % NKIND;
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  const int64_t tile0 = blockIdx.x / NKIND;
  const int64_t tstep = gridDim.x / NKIND;
  // Roles: the epilogue (gathers, TMEM reads, segment reduction, REDs) is the long pole of this kernel and the dA
  // loaders are plain streaming loads, so 8 of the 13 warps drain accumulators: group 0 (warps 0-3) takes this
  // CTA's even tiles / accumulator 0, group 1 (warps 9-12) the odd ones; warps 5-8 load dA, warp 4 issues MMAs.
  constexpr int BF_PROD_WARP0 = MMA_WARP + 1, BF_PROD_WARPS = 4, BF_PROD_THREADS = BF_PROD_WARPS * 32;
  constexpr int BF_EPI1_WARP0 = BF_PROD_WARP0 + BF_PROD_WARPS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, BF_PROD_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(dffull0 + 8 * s, 1);
      mbar_init(dfempty0 + 8 * s, EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(smem_u32(tmem_ptr_s), 512);
    tmem_relinquish();
  }
  // resident operand: rows j = 0..H-1 of diag(w2) . W1, this kind's columns, as blocks [H x 64]:
  //   block 2*qd   : product-part columns    kind*CB + 64*qd + [0,64)
  //   block 2*qd+1 : difference-part columns H + kind*CB + 64*qd + [0,64)
  for (int idx = threadIdx.x; idx < H * (2 * CB / 8); idx += THREADS) {
    const int j = idx / (2 * CB / 8);
    const int kc = idx % (2 * CB / 8);          // 16-byte chunk among this kind's 2*CB columns
    const int blk = kc >> 3;                    // 0 .. 2*NQ-1 in (qd, half) order: blk = 2*qd + half
    const int qd = blk >> 1, half = blk & 1;
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
  const float invS = 1.0f / grad_scale(dp_absmax[0]);

  if (warp >= BF_PROD_WARP0 && warp < BF_EPI1_WARP0) {
    // ------------------------------- dA loaders: [128 e x 128 j] sub-tiles -------------------------------
    const int pt = threadIdx.x - BF_PROD_WARP0 * 32;
    const int c = pt & 15;          // 16-byte chunk (8 hidden units) inside the 128-unit sub-tile
    const int row_base = pt >> 4;   // rows row_base + 8*i
    uint32_t it = 0;
    for (int64_t t = tile0; t < ntiles; t += tstep) {
#pragma unroll 1
      for (int js = 0; js < NJ; ++js, ++it) {
        uint4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int64_t row = t * TILE_M + row_base + 8 * i;
          v[i] = make_uint4(0, 0, 0, 0);
          if (row < n) v[i] = ld_stream_u4(reinterpret_cast<const uint4*>(dA + row * H + js * 128 + c * 8));
        }
        const uint32_t slot = it % NSTAGE;
        mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
        uint8_t* stage = sm + B_BYTES + slot * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          *reinterpret_cast<uint4*>(stage + (c >> 3) * (TILE_M * 128) + sw128_offset(row_base + 8 * i, c & 7)) = v[i];
        mbar_arrive(full0 + 8 * slot);   // the MMA thread issues the proxy fence (see scorer_producer.cuh)
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, TILE_M, 128) | (1u << 16);  // B is MN-major
      uint32_t it = 0, lt = 0;
      for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
        const uint32_t tb = lt & 1;
        mbar_wait(dfempty0 + 8 * tb, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int js = 0; js < NJ; ++js, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          fence_proxy_async_smem();   // loaders' generic-proxy stores -> async proxy
          tc_fence_after();
#pragma unroll
          for (int qd = 0; qd < NQ; ++qd) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t ad =
                  umma_desc_k_sw128(a_base + slot * STAGE_BYTES + (kk >> 2) * (TILE_M * 128) + (kk & 3) * 32);
              const uint64_t bd = umma_desc_mn_sw128(b_base + (2 * qd) * BLK + (js * 8 + kk) * 2048, BLK);
              umma_f16(tmem_base + tb * 256 + qd * 128, ad, bd, idesc, (js | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8 * slot);
        }
        umma_commit(dffull0 + 8 * tb);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue: dF -> d_out[src], d_out[dst] -------------------------------
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    const uint32_t grp = warp >= BF_EPI1_WARP0 ? 1u : 0u;
    for (int64_t t = tile0 + grp * tstep, lt = grp; t < ntiles; t += 2 * tstep, lt += 2) {
      const uint32_t tb = grp;   // == lt & 1
      const int64_t i = t * TILE_M + r;
      const bool live = i < n;
      int64_t e = live ? i : n - 1;
      if (ids) e = ids[e];
      const int s_node = src[e], d_node = dst[e];
      const T* xrow = tab + (int64_t)s_node * H;
      const T* yrow = tab + (int64_t)d_node * H;
      float* dyrow = d_out + (int64_t)d_node * H;
      // the embedding-row slices of slice `part + 1` are in flight while slice `part` is reduced; the first ones
      // are issued before the accumulator is waited for
      uint4 nx0, nx1, ny0, ny1;
      {
        const int c0 = kind * CB;
        nx0 = *reinterpret_cast<const uint4*>(xrow + c0);
        nx1 = *reinterpret_cast<const uint4*>(xrow + c0 + 8);
        ny0 = *reinterpret_cast<const uint4*>(yrow + c0);
        ny1 = *reinterpret_cast<const uint4*>(yrow + c0 + 8);
      }
      mbar_wait(dffull0 + 8 * tb, (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int part = 0; part < NQ * 4; ++part) {
        const int qd = part >> 2, q16 = part & 3;                     // 16-column slice q16 of column group qd
        const int col0 = kind * CB + qd * 64 + q16 * 16;              // 16 node-embedding columns
        const uint4 xv0 = nx0, xv1 = nx1, yv0 = ny0, yv1 = ny1;
        if (part + 1 < NQ * 4) {
          const int cn = kind * CB + ((part + 1) >> 2) * 64 + ((part + 1) & 3) * 16;
          nx0 = *reinterpret_cast<const uint4*>(xrow + cn);
          nx1 = *reinterpret_cast<const uint4*>(xrow + cn + 8);
          ny0 = *reinterpret_cast<const uint4*>(yrow + cn);
          ny1 = *reinterpret_cast<const uint4*>(yrow + cn + 8);
        }
        uint32_t f1[16], f2[16];
        tmem_ld16(tmem_base + lane_off + tb * 256 + qd * 128 + q16 * 16, f1);
        tmem_ld16(tmem_base + lane_off + tb * 256 + qd * 128 + 64 + q16 * 16, f2);
        tmem_ld_wait();
        if (part == NQ * 4 - 1) {  // all of this tile's accumulator has been read
          tc_fence_before();
          mbar_arrive(dfempty0 + 8 * tb);
        }
        const uint32_t xs[8] = {xv0.x, xv0.y, xv0.z, xv0.w, xv1.x, xv1.y, xv1.z, xv1.w};
        const uint32_t ys[8] = {yv0.x, yv0.y, yv0.z, yv0.w, yv1.x, yv1.y, yv1.z, yv1.w};
        float gx[16], gy[16];
        const float sc = live ? invS : 0.f;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float2 xa = Cvt<T>::unpack(xs[m]);
          const float2 ya = Cvt<T>::unpack(ys[m]);
          const float a0 = __uint_as_float(f1[2 * m]) * sc, a1 = __uint_as_float(f1[2 * m + 1]) * sc;
          const float d0 = __uint_as_float(f2[2 * m]) * sc, d1 = __uint_as_float(f2[2 * m + 1]) * sc;
          gx[2 * m] = fmaf(a0, ya.x, d0);
          gx[2 * m + 1] = fmaf(a1, ya.y, d1);
          gy[2 * m] = fmaf(a0, xa.x, -d0);
          gy[2 * m + 1] = fmaf(a1, xa.y, -d1);
        }
        // destination side: random rows -> 128-bit vector RED per row
        if (live) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            red_add_v4(dyrow + col0 + 4 * k, gy[4 * k], gy[4 * k + 1], gy[4 * k + 2], gy[4 * k + 3]);
        }
        // source side: rows of a warp mostly share their source (edge ids ascend by source): reduce each run
        // of equal sources across the warp first, then one coalesced RED from 16 lanes.
        uint32_t todo = __ballot_sync(0xffffffffu, live);
#pragma unroll 1
        for (int iter = 0; iter < 2 && todo; ++iter) {
          const int leader = __ffs(todo) - 1;
          const int s_lead = __shfl_sync(0xffffffffu, s_node, leader);
          const bool in_seg = live && s_node == s_lead && ((todo >> lane) & 1u);
          const uint32_t seg = __ballot_sync(0xffffffffu, in_seg);
          float v[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) v[k] = in_seg ? gx[k] : 0.f;
          const float tot = warp_colsum16(v, lane);
          if (lane < 16) atomicAdd(d_out + (int64_t)s_lead * H + col0 + lane, tot);
          todo &= ~seg;
        }
        if ((todo >> lane) & 1u) {
          float* dxrow = d_out + (int64_t)s_node * H;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            red_add_v4(dxrow + col0 + 4 * k, gx[4 * k], gx[4 * k + 1], gx[4 * k + 2], gx[4 * k + 3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
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
    uint32_t it = 0;
    for (int64_t s = sub0; s < nsub; s += sstep, ++it) {
      // all global loads of the stage are issued before the stage slot is waited for
      int sn[NIT], dn[NIT];
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int item = lt_id + k * LOAD_THREADS;
        const int64_t i = s * SUB_M + item / CH;
        sn[k] = -1;
        dn[k] = 0;
        if (i < n) {
          const int64_t e = ids ? ids[i] : i;
          sn[k] = src[e];
          dn[k] = dst[e];
        }
      }
      uint4 xv[NIT], yv[NIT], av[NDA];
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int cc = (lt_id + k * LOAD_THREADS) % CH;
        xv[k] = yv[k] = make_uint4(0, 0, 0, 0);
        if (sn[k] >= 0) {
          xv[k] = *reinterpret_cast<const uint4*>(tab + (int64_t)sn[k] * H + cc * 8);
          yv[k] = *reinterpret_cast<const uint4*>(tab + (int64_t)dn[k] * H + cc * 8);
        }
      }
#pragma unroll
      for (int k = 0; k < NDA; ++k) {
        const int item = lt_id + k * LOAD_THREADS;
        const int64_t i = s * SUB_M + item / (BN / 8);
        av[k] = make_uint4(0, 0, 0, 0);
        if (i < n) av[k] = ld_stream_u4(reinterpret_cast<const uint4*>(dA + i * H + nb * BN + (item % (BN / 8)) * 8));
      }
      const uint32_t slot = it % NSTAGE;
      mbar_wait(empty0 + 8 * slot, ((it / NSTAGE) & 1) ^ 1);
      uint8_t* fst = sm + slot * STAGE;
      uint8_t* dst_da = fst + F_BYTES;
#pragma unroll
      for (int k = 0; k < NIT; ++k) {
        const int item = lt_id + k * LOAD_THREADS;
        const int row = item / CH, cc = item % CH;
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
    }
    {
      const int cc = lt_id % (BN / 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(gsum_s + cc * 8 + k, gs[k]);
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
// host side
// =============================================================================================
size_t edge_score_bwd_tc_workspace_bytes(int64_t n, int64_t N, int64_t H) {
  return 2048 + (size_t)N * H * 2 + (size_t)n * H * 2;
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
  float* absmax = reinterpret_cast<float*>(base);
  T* tab = reinterpret_cast<T*>(base + 256);
  T* dA = reinterpret_cast<T*>(base + 256 + (((size_t)N * H * 2 + 255) & ~(size_t)255));
  SGS_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
  const int64_t cap = (int64_t)sm_count() * 16;
  int64_t g = ceil_div(n, 256);
  absmax_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(dp, n, absmax);
  SGS_LAUNCH_CHECK();
  const int64_t n8 = N * H / 8;
  g = ceil_div(n8, 256);
  convert_rows_kernel_b<T><<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(out, n8, reinterpret_cast<uint4*>(tab));
  SGS_LAUNCH_CHECK();
  const int64_t ntiles = ceil_div(n, kb::TILE_M);
  auto grid_for = [&](int kinds, int64_t items) {
    int64_t gr = (int64_t)(sm_count() / kinds) * kinds;
    if (gr > items * kinds) gr = items * kinds;
    return (unsigned)gr;
  };
  if (H == 256 && n >= 2 * kb::TILE_M) {
    // CTA pairs: every edge tile is gathered and built once (edge_score_tc2.cu, MODE 1)
    const int32_t rc = edge_score_bwd_gate_pair(tab, std::is_same<T, __nv_bfloat16>::value ? 1 : 0, src, dst, ids, n,
                                                W1, b1, p_drop, seed, p_fwd, dp, absmax, dA, db2, st);
    if (rc != SGS_OK) return rc;
  } else {
    auto kern = edge_score_bwd_da_kernel<T, BN, H>;
    constexpr size_t used = (size_t)2 * (H / 64) * BN * 128 + 3 * kb::STAGE_BYTES + BN * 8 + 16 * 8 + 16;
    const size_t smem = used + 1024 > 232448 ? 232448 : used + 1024;
    SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for(NB, ntiles), kb::THREADS, smem, st>>>(tab, src, dst, ids, n, W1, b1, p_drop, seed, p_fwd, dp,
                                                          absmax, dA, db2);
    SGS_LAUNCH_CHECK();
  }
  {
    auto kern = edge_score_bwd_df_kernel<T, H>;
    constexpr size_t used = (size_t)2 * 2 * H * 128 + 3 * kb::STAGE_BYTES + 16 * 8 + 16;
    const size_t smem = used + 1024 > 232448 ? 232448 : used + 1024;
    SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid_for(H / 128, ntiles), kb::THREADS, smem, st>>>(tab, src, dst, ids, n, W1, w2, dA, absmax, d_out);
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
