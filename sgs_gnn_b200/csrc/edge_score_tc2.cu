// K1 tensor-core forward for H = 256 on CTA PAIRS (tcgen05 cta_group::2).
//
// The single-CTA kernel (edge_score_tc.cu) keeps a 128-hidden-unit slice of W1 resident per CTA, so every
// 128-edge tile is gathered and built twice (once per slice) -- and ncu shows that kernel bound by the latency of
// exactly those gathers (r01: tensor pipe 40%, DRAM 5%, producers stalled on their row loads).  Here the two CTAs
// of a cluster form one MMA of M = 256 edges x N = 256 hidden units:
//   * each CTA gathers and builds the [x*y | x-y] features of ITS 128 edges once,
//   * each CTA keeps ITS half of W1 (128 hidden units x 2H, 128 KB) resident; the pair's tensor cores read both,
//   * the accumulator of a CTA (its 128 edges x all 256 hidden units, fp32) lives in its own TMEM, double buffered
//     (2 x 256 columns = all 512), so the epilogue finishes sigma(w2 . relu(.) + b2) without a second pass.
// Roles per CTA: warps 0-3 and 13-16 epilogue (group g drains accumulator buffer g, i.e. every other tile: one warp
// per TMEM lane quarter has two tile periods for its 256 columns), warp 4 MMA issue (leader CTA only) + TMEM
// allocation, warps 5-12 producers.
// Barriers (same offsets in both CTAs): producers arrive on their OWN CTA's full[s] (a plain CTA-scope arrive: a
// cluster-scope release costs a MEMBAR per producer thread that also drains its prefetched row loads); the idle MMA
// warp of the peer CTA relays "peer stage s full" to the leader's pfull[s] with a single release.cluster arrive.
// empty[s] / tmem_full[a] are signalled in both CTAs by a multicast tcgen05.commit; tmem_empty[a] lives in the
// leader and counts the epilogue threads of both CTAs (the peer's arrive remotely, relaxed: their TMEM reads are
// already complete).
#include <cuda_runtime.h>

#include "common.cuh"
#include "scorer_producer.cuh"
#include "tc.cuh"

namespace sgs {

namespace k1p {
constexpr int H = 256;
constexpr int TILE_M = 128;                    // edges per CTA and tile (256 per pair)
constexpr int BN = 128;                        // hidden units whose W1 rows this CTA keeps resident
constexpr int STAGE_BYTES = TILE_M * 128 * 2;  // [128 x 64] product block + [128 x 64] difference block
constexpr int NSTAGE = 3;
constexpr int NSP = H / 64;                    // stages per tile
constexpr int B_BLOCK_BYTES = BN * 128;
constexpr int B_BYTES = 2 * NSP * B_BLOCK_BYTES;   // 128 KB
#ifndef SGS_K1_EPI_GROUPS
#define SGS_K1_EPI_GROUPS 1
#endif
constexpr int EPI_GROUPS = SGS_K1_EPI_GROUPS;  // 2: a second group of epilogue warps drains every other tile
constexpr int EPI_WARPS = 4;                   // per epilogue group
constexpr int PROD_WARPS = 8;
// Warp order matters: the SM sub-partition arbiter favours the highest warp id, so the epilogue warps (the
// longest dependent instruction chains) come LAST and are never starved by producers polling their barriers.
constexpr int PROD_WARP0 = 0;
constexpr int MMA_WARP = PROD_WARPS;
constexpr int EPI0_WARP0 = PROD_WARPS + 1;                   // first epilogue group: warps 9-12
constexpr int EPI1_WARP0 = EPI0_WARP0 + EPI_WARPS;           // second epilogue group: warps 13-16
constexpr int THREADS = (EPI_GROUPS * EPI_WARPS + 1 + PROD_WARPS) * 32;   // 416 or 544
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int SMALL_BYTES = 2 * H * 4 + (3 * NSTAGE + 4) * 8 + 16;
constexpr int USED_BYTES = B_BYTES + NSTAGE * STAGE_BYTES + SMALL_BYTES;
}  // namespace k1p

// MODE 0: forward, p_out[i] = sigma(w2 . dropout(relu(Z + b1)) + b2).
// MODE 1: first backward kernel ("BA"): the same recompute of Z, but the epilogue writes the 16-bit hidden-layer
//         gate gradient  G[i, j] = S * dz_i / (1 - p_drop) * [Z_ij + b1_j > 0] * keep_ij   (dz = dp * p * (1 - p),
//         S = grad_scale(max|dp|)) to g_out [n, H] and accumulates db2 += sum_i dz_i.  The factor w2_j of
//         dA = G . diag(w2) is folded into the operands / epilogues of the BF and BW kernels
//         (edge_score_bwd_tc.cu), which also derive db1 and dw2 from G^T F -- this epilogue has no reductions.
template <typename T, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k1p::THREADS, 1)
edge_score_tc2_kernel(const T* __restrict__ tab, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                      const int32_t* __restrict__ ids, int64_t n, const float* __restrict__ W1,
                      const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                      float p_drop, uint64_t seed, float* __restrict__ p_out, const float* __restrict__ p_fwd,
                      const float* __restrict__ dp, const float* __restrict__ dp_absmax, T* __restrict__ g_out,
                      float* __restrict__ db2, const int32_t* __restrict__ key_ids) {
  using namespace k1p;
  using namespace tc;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;   // identical in both CTAs of the pair
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  {
    uint32_t dyn_size;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_size));
    if (pad + USED_BYTES > dyn_size) __trap();
  }
  // layout: [B resident | A stages | b1 | w2 (scaled) | barriers | tmem ptr]
  const uint32_t b_base = sm_addr;
  const uint32_t a_base = b_base + B_BYTES;
  uint8_t* small = sm + B_BYTES + NSTAGE * STAGE_BYTES;
  float* b1s = reinterpret_cast<float*>(small);
  float* w2s = b1s + H;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w2s + H);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 3 * NSTAGE + 4);
  const uint32_t full0 = smem_u32(bars);
  const uint32_t empty0 = full0 + 8 * NSTAGE;
  const uint32_t tfull0 = empty0 + 8 * NSTAGE;
  const uint32_t tempty0 = tfull0 + 16;
  const uint32_t pfull0 = tempty0 + 16;   // leader only: "the peer's stage s is full" (one relayed arrival)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();         // 0 = leader
  const int64_t npairs = gridDim.x / 2;
  const int64_t pair = blockIdx.x / 2;
  const int64_t ndt = (n + 2 * TILE_M - 1) / (2 * TILE_M);   // 256-edge double tiles
  const int64_t ntiles_padded = 2 * ndt;                     // the pair always works in lock step
  const int64_t tile0 = 2 * pair + rank;
  const int64_t tstep = 2 * npairs;

  // ---- one-time setup ----
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full0 + 8 * s, PROD_THREADS);       // this CTA's producers
      mbar_init(empty0 + 8 * s, 1);
      mbar_init(pfull0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 2 * EPI_WARPS * 32);   // used in the leader: group-a epilogue threads of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc_2cta(smem_u32(tmem_ptr_s), 512);
    tmem_relinquish_2cta();
  }
  // resident B: rows rank*128 .. +128 of W1 [H, 2H], fp32 -> 16 bit, K reordered in (product, difference) pairs
  for (int idx = threadIdx.x; idx < BN * (2 * H / 8); idx += THREADS) {
    const int nrow = idx / (2 * H / 8);
    const int kc = idx % (2 * H / 8);
    const int k0 = kc * 8;
    const int half = k0 >= H;
    const int kk = half ? k0 - H : k0;
    const int sp = kk >> 6;
    const int c16 = (kk & 63) >> 3;
    const float* g = W1 + (int64_t)(rank * BN + nrow) * (2 * H) + k0;
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
    for (int j = threadIdx.x; j < H; j += THREADS) {
      b1s[j] = b1[j];
      w2s[j] = w2[j] * w2_scale;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();   // barriers initialised, W1 halves resident and TMEM allocated in BOTH CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp < MMA_WARP) {
    // =============================== producers ===============================
    FeatureProducer<T, H, NSTAGE, STAGE_BYTES, TILE_M>::run(tab, src, dst, ids, n, tile0, tstep, ntiles_padded,
                                                             sm + B_BYTES, full0, empty0,
                                                             threadIdx.x - PROD_WARP0 * 32);
  } else if (warp == MMA_WARP) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = umma_idesc(Cvt<T>::kFmt, 2 * TILE_M, H);
      uint32_t it = 0, lt = 0;
      for (int64_t t = tile0; t < ntiles_padded; t += tstep, ++lt) {
        const uint32_t acc = lt & 1;
        mbar_wait_cluster(tempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * H;
#pragma unroll 1
        for (int sp = 0; sp < NSP; ++sp, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);            // own producers
          fence_proxy_async_smem();   // producers' generic-proxy stores -> async proxy (see scorer_producer.cuh)
          mbar_wait_cluster(pfull0 + 8 * slot, (it / NSTAGE) & 1);   // the peer's, relayed
          tc_fence_after();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int k16 = 0; k16 < 4; ++k16) {
              const uint64_t ad = umma_desc_k_sw128(a_base + slot * STAGE_BYTES + half * (TILE_M * 128) + k16 * 32);
              const uint64_t bd = umma_desc_k_sw128(b_base + (2 * sp + half) * B_BLOCK_BYTES + k16 * 32);
              umma_f16_2cta(d_tmem, ad, bd, idesc, (sp | half | k16) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2cta(empty0 + 8 * slot, 3);   // both CTAs' stage `slot` reusable once these MMAs have read it
        }
        umma_commit_2cta(tfull0 + 8 * acc, 3);      // accumulators complete in both CTAs
      }
    } else if (rank == 1 && lane == 0) {
      // relay: forward "my producers filled stage s" to the leader (one cluster-scope release per stage)
      const uint32_t pfull_leader = mapa_shared(pfull0, 0);
      uint32_t it = 0;
      for (int64_t t = tile0; t < ntiles_padded; t += tstep) {
#pragma unroll 1
        for (int sp = 0; sp < NSP; ++sp, ++it) {
          const uint32_t slot = it % NSTAGE;
          mbar_wait(full0 + 8 * slot, (it / NSTAGE) & 1);
          fence_proxy_async_smem();   // this CTA's stage stores -> async proxy, before the leader may issue the MMA
          mbar_arrive_cluster(pfull_leader + 8 * slot);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue ===============================
    // group 0 (warps 0-3) handles this CTA's tiles lt = 0, 2, 4, ... (accumulator 0), group 1 (warps 13-16) the
    // odd ones (accumulator 1); a warp reads the TMEM lane quarter (warp & 3)
    const uint32_t grp = warp >= EPI1_WARP0 ? 1u : 0u;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;  // row of this CTA's tile == TMEM lane
    const uint32_t thr32 = dropout_threshold32(dropout_threshold(p_drop));
    const bool drop = p_drop > 0.f;
    const float bias2 = MODE == 0 ? b2[0] : 0.f;
    const float gscale = MODE == 1 ? grad_scale(dp_absmax[0]) * (drop ? 1.0f / (1.0f - p_drop) : 1.0f) : 0.f;
    float acc_b2 = 0.f;
    const uint32_t tempty_leader0 = mapa_shared(tempty0, 0);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16);
    // per-row inputs of the NEXT tile (dropout key id; MODE 1: dz, p) are loaded one tile ahead: their latency (an
    // HBM miss each) would otherwise sit on the epilogue's critical path once per tile
    int64_t kid_n = 0;
    float dz_n = 0.f, pe_n = 0.f;
    auto prefetch_row = [&](int64_t t) {
      const int64_t i = t * TILE_M + r;
      if (drop) {
        int64_t e = i < n ? i : n - 1;
        if (key_ids) e = key_ids[e];   // the edge list was re-ordered by the caller: original ids for the mask
        else if (ids) e = ids[e];
        kid_n = e;
      }
      if (MODE == 1) {
        dz_n = 0.f;
        if (i < n) {
          dz_n = dp[i];
          if (p_fwd) pe_n = p_fwd[i];
        }
      }
    };
    uint32_t lt = grp;
    if (tile0 + grp * tstep < ntiles_padded) prefetch_row(tile0 + grp * tstep);
    for (int64_t t = tile0 + grp * tstep; t < ntiles_padded; t += EPI_GROUPS * tstep, lt += EPI_GROUPS) {
      const uint32_t acc = lt & 1;
      const uint32_t tempty_leader = tempty_leader0 + 8 * acc;
      const uint32_t taddr = taddr0 + acc * H;
      const int64_t i = t * TILE_M + r;
      const int64_t kid = kid_n;
      float dz = dz_n;
      if (MODE == 1 && p_fwd) dz *= pe_n * (1.0f - pe_n);   // p_fwd == nullptr: dp already holds dz = dp * p * (1 - p)
      if (t + EPI_GROUPS * tstep < ntiles_padded) prefetch_row(t + EPI_GROUPS * tstep);
      const uint32_t rowkey = drop ? dropout_rowkey(seed, (uint64_t)kid) : 0u;
      uint32_t g_lo = 0, g_hi = 0;   // MODE 1: this row's gate gradient in the low / high 16 bits
      if (MODE == 1) {
        acc_b2 += dz;
        const float g = dz * gscale;
        const uint32_t gg = Cvt<T>::pack(g, g);
        g_lo = gg & 0xFFFFu;
        g_hi = gg & 0xFFFF0000u;
      }
      mbar_wait(tfull0 + 8 * acc, (lt >> 1) & 1);
      tc_fence_after();
      float z0 = 0.f, z1 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < H / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch == H / 32 - 1) {  // accumulator drained: the leader's MMA thread may overwrite it
          tc_fence_before();
          if (rank == 0) mbar_arrive(tempty0 + 8 * acc);
          else mbar_arrive_cluster_relaxed(tempty_leader);
        }
        const uint32_t rk = rowkey ^ ((uint32_t)ch * 0x9E3779B9u);   // dropout_colmix: chunk part of the pair constant
        const float* bp = b1s + ch * 32;
        const float* wp = w2s + ch * 32;
        if (MODE == 1) {
          uint32_t o[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bb = *reinterpret_cast<const float4*>(bp + j4 * 4);
            bool m0 = __uint_as_float(v[j4 * 4 + 0]) + bb.x > 0.f;
            bool m1 = __uint_as_float(v[j4 * 4 + 1]) + bb.y > 0.f;
            bool m2 = __uint_as_float(v[j4 * 4 + 2]) + bb.z > 0.f;
            bool m3 = __uint_as_float(v[j4 * 4 + 3]) + bb.w > 0.f;
            if (drop) {
              const uint32_t xa = rk ^ ((uint32_t)(2 * j4) * 0x7FEB352Du);
              const uint32_t xb = rk ^ ((uint32_t)(2 * j4 + 1) * 0x7FEB352Du);
              m0 = m0 && (xa * kDropMulEven >= thr32);
              m1 = m1 && (xa * kDropMulOdd >= thr32);
              m2 = m2 && (xb * kDropMulEven >= thr32);
              m3 = m3 && (xb * kDropMulOdd >= thr32);
            }
            o[2 * j4] = (m0 ? g_lo : 0u) | (m1 ? g_hi : 0u);
            o[2 * j4 + 1] = (m2 ? g_lo : 0u) | (m3 ? g_hi : 0u);
          }
          if (i < n) {
            uint4* gp = reinterpret_cast<uint4*>(g_out + i * H + ch * 32);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) gp[c8] = make_uint4(o[4 * c8], o[4 * c8 + 1], o[4 * c8 + 2], o[4 * c8 + 3]);
          }
          continue;
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(bp + j4 * 4);
          const float4 ww = *reinterpret_cast<const float4*>(wp + j4 * 4);  // already scaled by 1/(1-p)
          const float h0 = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bb.x, 0.f);
          const float h1 = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bb.y, 0.f);
          const float h2 = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bb.z, 0.f);
          const float h3 = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bb.w, 0.f);
          if (drop) {
            const uint32_t xa = rk ^ ((uint32_t)(2 * j4) * 0x7FEB352Du);
            const uint32_t xb = rk ^ ((uint32_t)(2 * j4 + 1) * 0x7FEB352Du);
            if (xa * kDropMulEven >= thr32) z0 = fmaf(ww.x, h0, z0);
            if (xa * kDropMulOdd >= thr32) z1 = fmaf(ww.y, h1, z1);
            if (xb * kDropMulEven >= thr32) z0 = fmaf(ww.z, h2, z0);
            if (xb * kDropMulOdd >= thr32) z1 = fmaf(ww.w, h3, z1);
          } else {
            z0 = fmaf(ww.x, h0, z0);
            z1 = fmaf(ww.y, h1, z1);
            z0 = fmaf(ww.z, h2, z0);
            z1 = fmaf(ww.w, h3, z1);
          }
        }
      }
      if (MODE == 0 && i < n) p_out[i] = 1.0f / (1.0f + expf(-(z0 + z1 + bias2)));
    }
    if (MODE == 1) {
      acc_b2 = warp_sum(acc_b2);
      if (lane == 0) atomicAdd(db2, acc_b2);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

size_t edge_score_tc2_workspace_bytes(int64_t N) { return 1024 + (size_t)N * k1p::H * 2; }

template <typename T, int MODE>
static int32_t launch_k1_pair(const T* tab, const int32_t* src, const int32_t* dst, const int32_t* ids, int64_t n,
                              const float* W1, const float* b1, const float* w2, const float* b2, float p_drop,
                              uint64_t seed, float* p, cudaStream_t st, const float* p_fwd = nullptr,
                              const float* dp = nullptr, const float* dp_absmax = nullptr, T* g_out = nullptr,
                              float* db2 = nullptr, const int32_t* key_ids = nullptr) {
  size_t smem = (size_t)k1p::USED_BYTES + 1024;
  if (smem > 232448) smem = 232448;
  auto kern = edge_score_tc2_kernel<T, MODE>;
  SGS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ndt = ceil_div(n, 2 * k1p::TILE_M);
  int64_t pairs = sm_count() / 2;
  if (pairs > ndt) pairs = ndt;
  kern<<<(unsigned)(2 * pairs), k1p::THREADS, smem, st>>>(tab, src, dst, ids, n, W1, b1, w2, b2, p_drop, seed, p,
                                                          p_fwd, dp, dp_absmax, g_out, db2, key_ids);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

// tab: the 16-bit node-embedding table [N, 256] (already converted)
int32_t edge_score_fwd_pair(const void* tab, int32_t is_bf16, const int32_t* src, const int32_t* dst,
                            const int32_t* ids, int64_t n, const float* W1, const float* b1, const float* w2,
                            const float* b2, float p_drop, uint64_t seed, float* p, cudaStream_t st) {
  if (is_bf16)
    return launch_k1_pair<__nv_bfloat16, 0>(reinterpret_cast<const __nv_bfloat16*>(tab), src, dst, ids, n, W1, b1, w2,
                                            b2, p_drop, seed, p, st);
  return launch_k1_pair<__half, 0>(reinterpret_cast<const __half*>(tab), src, dst, ids, n, W1, b1, w2, b2, p_drop,
                                   seed, p, st);
}

// BA on CTA pairs (H = 256): g_out [n, 256] 16-bit gate gradients, db2 += sum dz.  tab / g_out: __half or bf16.
int32_t edge_score_bwd_gate_pair(const void* tab, int32_t is_bf16, const int32_t* src, const int32_t* dst,
                                 const int32_t* ids, int64_t n, const float* W1, const float* b1, float p_drop,
                                 uint64_t seed, const float* p_fwd, const float* dp, const float* dp_absmax,
                                 void* g_out, float* db2, const int32_t* key_ids, cudaStream_t st) {
  if (is_bf16)
    return launch_k1_pair<__nv_bfloat16, 1>(reinterpret_cast<const __nv_bfloat16*>(tab), src, dst, ids, n, W1, b1, b1,
                                            b1, p_drop, seed, nullptr, st, p_fwd, dp, dp_absmax,
                                            reinterpret_cast<__nv_bfloat16*>(g_out), db2, key_ids);
  return launch_k1_pair<__half, 1>(reinterpret_cast<const __half*>(tab), src, dst, ids, n, W1, b1, b1, b1, p_drop,
                                   seed, nullptr, st, p_fwd, dp, dp_absmax, reinterpret_cast<__half*>(g_out), db2, key_ids);
}

}  // namespace sgs
