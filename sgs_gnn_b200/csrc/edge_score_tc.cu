// K1 tensor-core forward: fused endpoint gather -> [x*y | x-y] tile build in swizzled smem ->
// tcgen05.mma (bf16/fp16 operands, fp32 accumulators in TMEM) -> bias/ReLU/dropout/fc2-dot epilogue.
// The [E, 2H] feature tensor and the [E, H] hidden tensor never exist in HBM.
//
// One persistent CTA per SM.  A CTA owns one block of BN <= 128 hidden units (its slice of W1 stays
// resident in shared memory for the whole kernel) and walks 128-edge tiles:
//   warps 0-3   epilogue : TMEM -> registers, +b1, ReLU, dropout, dot with w2  (one warp per TMEM lane quarter)
//   warp  4     MMA      : one elected thread issues tcgen05.mma / tcgen05.commit
//   warps 5-12  producers: gather 16-bit rows of out[src], out[dst] (128-bit loads), form x*y and
//                          x-y in fp32, round once, store into the SWIZZLE_128B K-major A stage
// Pipelines: smem stage ring (full/empty mbarriers) and a double-buffered TMEM accumulator
// (tmem_full/tmem_empty), so gathers, MMAs and the epilogue of consecutive tiles overlap.
// K is ordered in pairs of 64-column blocks [product cols 64j.. | difference cols 64j..] so one
// gathered 16-byte chunk of x and y feeds both blocks of the same stage.
#include <cstdlib>

#include "common.cuh"
#include "tc.cuh"
#include "scorer_producer.cuh"

namespace sgs {

namespace k1 {
constexpr int TILE_M = 128;
constexpr int STAGE_BYTES = TILE_M * 128 * 2;  // two [128 x 64] 16-bit blocks = 32 KB
constexpr int NSTAGE = 3;
constexpr int EPI_WARPS = 4;
constexpr int PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;
constexpr int PROD_WARP0 = EPI_WARPS + 1;
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;  // 416: at most 4 warps per SM sub-partition
constexpr int PROD_THREADS = PROD_WARPS * 32;               // 256
}  // namespace k1


// out fp32 [rows, H] -> 16-bit table
template <typename T>
__global__ void convert_rows_kernel(const float* __restrict__ in, int64_t n8, uint4* __restrict__ outp) {
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

__global__ void edge_score_finalize_kernel(const float* __restrict__ zpart, int nb, int64_t n,
                                           const float* __restrict__ b2, float* __restrict__ p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float bias = b2[0];
  for (; i < n; i += stride) {
    float z = bias;
    for (int k = 0; k < nb; ++k) z += zpart[(int64_t)k * n + i];
    p[i] = 1.0f / (1.0f + expf(-z));
  }
}

template <typename T, int BN, int H>
__global__ void __launch_bounds__(k1::THREADS, 1)
edge_score_tc_kernel(const T* __restrict__ tab, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                     const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ W1,
                     const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                     float p_drop, uint64_t seed, float* __restrict__ p_out, float* __restrict__ zpart) {
  using namespace k1;
  using namespace tc;
  constexpr int NB = H / BN;              // hidden-unit blocks (CTA kinds)
  constexpr int NSP = H / 64;             // stage pairs (128 K columns each) per tile
  constexpr int B_BLOCK_BYTES = BN * 128; // one [BN x 64] 16-bit K block
  constexpr int B_BYTES = 2 * NSP * B_BLOCK_BYTES;
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static_assert(H % 64 == 0 && BN * NB == H && BN % 32 == 0 && BN <= 128, "unsupported shape");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    constexpr uint32_t kUsed = 2 * (H / 64) * BN * 128 + NSTAGE * STAGE_BYTES + BN * 8 + 2 * TILE_M * 4 +
                               (2 * NSTAGE + 4) * 8 + 16;
    if (pad + kUsed > dyn_size) __trap();
  }
  // layout: [B resident | A stages | b1 slice | w2 slice | zsh[2][128] | barriers | tmem ptr]
  const uint32_t b_base = sm_addr;
  const uint32_t a_base = b_base + B_BYTES;
  uint8_t* small = sm + B_BYTES + NSTAGE * STAGE_BYTES;
  float* b1s = reinterpret_cast<float*>(small);
  float* w2s = b1s + BN;
  float* zsh = w2s + BN;                  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(zsh + 2 * TILE_M);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = full0 + 8 * NSTAGE;
  const uint32_t tfull0 = empty0 + 8 * NSTAGE;
  const uint32_t tempty0 = tfull0 + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nb = blockIdx.x % NB;
  const int64_t ntiles = (n + TILE_M - 1) / TILE_M;
  const int64_t tile0 = blockIdx.x / NB;
  const int64_t tstep = gridDim.x / NB;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, PROD_THREADS);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(smem_u32(tmem_ptr_s), TMEM_COLS);
    tmem_relinquish();
  }
  // resident B: rows nb*BN .. +BN of W1 [H, 2H], fp32 -> 16 bit, K reordered in (product, difference) pairs
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
  {
    const float w2_scale = p_drop > 0.f ? 1.0f / (1.0f - p_drop) : 1.0f;  // inverted-dropout scale folded into w2
    for (int j = threadIdx.x; j < BN; j += THREADS) {
      b1s[j] = b1[nb * BN + j];
      w2s[j] = w2[nb * BN + j] * w2_scale;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp >= PROD_WARP0) {
    // =============================== producers ===============================
    FeatureProducer<T, H, NSTAGE, STAGE_BYTES, TILE_M>::run(tab, src, dst, ids, n, tile0, tstep, ntiles,
                                                             sm + B_BYTES, full0, empty0,
                                                             threadIdx.x - PROD_WARP0 * 32);
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, TILE_M, BN);
      uint32_t it = 0, lt = 0;
      for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
        const uint32_t acc = lt & 1;
        mbar_wait(tempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll 1
        for (int sp = 0; sp < NSP; ++sp, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          fence_proxy_async_smem();   // producers' generic-proxy stores -> async proxy (see scorer_producer.cuh)
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t ad = umma_desc_k_sw128(a_base + slot * STAGE_BYTES + half * (TILE_M * 128) + k16 * 32);
              const uint64_t bd = umma_desc_k_sw128(b_base + (2 * sp + half) * B_BLOCK_BYTES + k16 * 32);
              umma_f16(d_tmem, ad, bd, idesc, (sp | half | k16) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8 * slot);   // smem stage reusable once these MMAs have read it
        }
        umma_commit(tfull0 + 8 * acc);      // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    const int lg = warp & 3;       // TMEM lane quarter (warps 0-3)
    const int r = lg * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t thr = dropout_threshold(p_drop);
    const bool drop = p_drop > 0.f;
    const float bias2 = b2[0];
    uint32_t lt = 0;
    for (int64_t t = tile0; t < ntiles; t += tstep, ++lt) {
      const uint32_t acc = lt & 1;
      const int64_t i = t * TILE_M + r;
      uint32_t rowkey = 0;
      if (drop) {
        int64_t e = i < n ? i : n - 1;
        if (ids) e = ids[e];
        rowkey = dropout_rowkey(seed, (uint64_t)e);
      }
      mbar_wait(tfull0 + 8 * acc, (lt >> 1) & 1);
      tc_fence_after();
      float z = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= BN) {  // accumulator drained: the MMA warp may overwrite it
          tc_fence_before();
          mbar_arrive(tempty0 + 8 * acc);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int col = c0 + j4 * 4;  // column inside this CTA's hidden block
          const float4 bb = *reinterpret_cast<const float4*>(b1s + col);
          const float4 ww = *reinterpret_cast<const float4*>(w2s + col);  // already scaled by 1/(1-p)
          const float h0 = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bb.x, 0.f);
          const float h1 = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bb.y, 0.f);
          const float h2 = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bb.z, 0.f);
          const float h3 = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bb.w, 0.f);
          if (drop) {
            const uint32_t cp = (uint32_t)(nb * BN + col) >> 1;
            const uint32_t r0 = dropout_pair(rowkey, cp), r1 = dropout_pair(rowkey, cp + 1);
            if ((r0 & 0xFFFFu) >= thr) z = fmaf(ww.x, h0, z);
            if ((r0 >> 16) >= thr) z = fmaf(ww.y, h1, z);
            if ((r1 & 0xFFFFu) >= thr) z = fmaf(ww.z, h2, z);
            if ((r1 >> 16) >= thr) z = fmaf(ww.w, h3, z);
          } else {
            z = fmaf(ww.x, h0, z);
            z = fmaf(ww.y, h1, z);
            z = fmaf(ww.z, h2, z);
            z = fmaf(ww.w, h3, z);
          }
        }
      }
      if (i < n) {
        if (NB == 1) p_out[i] = 1.0f / (1.0f + expf(-(z + bias2)));
        else zpart[(int64_t)nb * n + i] = z;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static size_t k1_used_bytes(int BN, int H) {
  return (size_t)(2 * (H / 64)) * BN * 128 + (size_t)k1::NSTAGE * k1::STAGE_BYTES + (size_t)BN * 8 +
         2 * k1::TILE_M * 4 + (2 * k1::NSTAGE + 4) * 8 + 16;
}
// 1 KB of slack for the 1024-byte alignment of the swizzled tiles, capped at the 227 KB per-CTA limit
// (the kernel traps if the alignment pad it actually needs does not fit).
static size_t k1_smem_bytes(int BN, int H) {
  const size_t want = k1_used_bytes(BN, H) + 1024;
  return want > 232448 ? 232448 : want;
}

size_t edge_score_tc_workspace_bytes(int64_t n, int64_t N, int64_t H) {
  const int64_t nb = H > 128 ? H / 128 : 1;
  return 1024 + (size_t)N * H * 2 + (nb > 1 ? (size_t)nb * n * 4 : 0);
}

int32_t edge_score_fwd_pair(const void* tab, int32_t is_bf16, const int32_t* src, const int32_t* dst,
                            const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                            const float* b2, float p_drop, uint64_t seed, float* p, cudaStream_t st);

// SGS_K1_SINGLE_CTA=1 selects the single-CTA kernel for H = 256 (A/B measurements)
static bool use_cta_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGS_K1_SINGLE_CTA");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

template <typename T, int BN, int H>
static int32_t launch_k1(const float* out, int64_t N, const int32_t* src, const int32_t* dst, const int32_t* ids,
                         int64_t n, const float* W1, const float* b1, const float* w2, const float* b2,
                         float p_drop, uint64_t seed, float* p, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int NB = H / BN;
  if (ws_bytes < edge_score_tc_workspace_bytes(n, N, H)) {
    set_error("sgs_edge_score_fwd: workspace too small");
    return SGS_E_WORKSPACE;
  }
  T* tab = reinterpret_cast<T*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* zpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(tab) + (((size_t)N * H * 2 + 255) & ~(size_t)255));
  const int64_t n8 = N * H / 8;
  int64_t g = ceil_div(n8, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  convert_rows_kernel<T><<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(out, n8, reinterpret_cast<uint4*>(tab));
  SGS_LAUNCH_CHECK();
  if (H == 256 && use_cta_pairs()) {
    // CTA pairs (cta_group::2): every 128-edge tile is gathered and built once instead of once per W1 slice
    return edge_score_fwd_pair(tab, Cvt<T>::kFmt == 1, src, dst, ids, n, W1, b1, w2, b2, p_drop, seed, p, st);
  }
  const size_t smem = k1_smem_bytes(BN, H);
  auto kern = edge_score_tc_kernel<T, BN, H>;
  SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = ceil_div(n, k1::TILE_M);
  int64_t grid = (int64_t)(sm_count() / NB) * NB;
  if (grid > ntiles * NB) grid = ntiles * NB;
  kern<<<(unsigned)grid, k1::THREADS, smem, st>>>(tab, src, dst, ids, n, W1, b1, w2, b2, p_drop, seed, p, zpart);
  SGS_LAUNCH_CHECK();
  if (NB > 1) {
    int64_t g2 = ceil_div(n, 256);
    edge_score_finalize_kernel<<<(unsigned)(g2 > cap ? cap : g2), 256, 0, st>>>(zpart, NB, n, b2, p);
    SGS_LAUNCH_CHECK();
  }
  return SGS_OK;
}

int32_t edge_score_fwd_tc(const float* out, int64_t N, int64_t H, const int32_t* src, const int32_t* dst,
                          const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                          const float* b2, float p_drop, uint64_t seed, float* p, void* ws, size_t ws_bytes,
                          int32_t precision, cudaStream_t st) {
  if (precision != SGS_PREC_BF16 && precision != SGS_PREC_FP16) {
    set_error("sgs_edge_score_fwd: tensor-core scorer supports bf16 / fp16 operands");
    return SGS_E_UNSUPPORTED;
  }
#define SGS_K1(T, BN, HH) \
  return launch_k1<T, BN, HH>(out, N, src, dst, ids, n, W1, b1, w2, b2, p_drop, seed, p, ws, ws_bytes, st)
  if (precision == SGS_PREC_BF16) {
    if (H == 256) SGS_K1(__nv_bfloat16, 128, 256);
    if (H == 128) SGS_K1(__nv_bfloat16, 128, 128);
    if (H == 64) SGS_K1(__nv_bfloat16, 64, 64);
  } else {
    if (H == 256) SGS_K1(__half, 128, 256);
    if (H == 128) SGS_K1(__half, 128, 128);
    if (H == 64) SGS_K1(__half, 64, 64);
  }
#undef SGS_K1
  set_error("sgs_edge_score_fwd: tensor-core scorer supports H in {64, 128, 256}, got %lld", (long long)H);
  return SGS_E_UNSUPPORTED;
}

}  // namespace sgs
