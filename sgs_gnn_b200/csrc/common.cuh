// Shared helpers for libsgs_b200 (sm_100a).  Internal -- the public surface is include/sgs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sgs_b200.h"

namespace sgs {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define SGS_CHECK_ARG(cond, msg)                           \
  do {                                                     \
    if (!(cond)) {                                         \
      sgs::set_error("%s: %s", __func__, msg);             \
      return SGS_E_INVALID;                                \
    }                                                      \
  } while (0)

#define SGS_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      sgs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return SGS_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define SGS_LAUNCH_CHECK()                 \
  do {                                     \
    sgs::count_launch();                   \
    SGS_CUDA(cudaPeekAtLastError());       \
  } while (0)

static inline cudaStream_t as_stream(sgs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------
// Counter-based hash RNG (splitmix64 finaliser).  Dropout keep decisions are a pure function
// of (seed, row, column) so that forward and backward kernels (and the numpy mirror in
// sgs_gnn_b200/rng.py used by the parity tests) regenerate the same mask.
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Dropout keep decisions: a 32-bit per-row key (one splitmix64 per (seed, row)), then per PAIR of columns one
// xor with a column constant and two multiplicative hashes whose top 16 bits are compared with the threshold.
// The hot epilogues test `x * M >= threshold << 16` directly (one IMAD + one ISETP per column, no extraction).
constexpr uint32_t kDropMulEven = 0x85EBCA6Bu, kDropMulOdd = 0xC2B2AE35u;
__host__ __device__ __forceinline__ uint32_t dropout_rowkey(uint64_t seed, uint64_t row) {
  return (uint32_t)(splitmix64(seed + row * 0xD6E8FEB86659FD93ull) >> 32);
}
// column-pair constant: separable in (colpair >> 4, colpair & 15) so that a 32-column chunk of an unrolled loop
// needs one runtime xor (chunk part) and otherwise only immediates
__host__ __device__ __forceinline__ uint32_t dropout_colmix(uint32_t colpair) {
  return ((colpair >> 4) * 0x9E3779B9u) ^ ((colpair & 15u) * 0x7FEB352Du);
}
// 32 random bits for columns (2*colpair, 2*colpair + 1): low half -> even column, high half -> odd column
__host__ __device__ __forceinline__ uint32_t dropout_pair(uint32_t rowkey, uint32_t colpair) {
  const uint32_t x = rowkey ^ dropout_colmix(colpair);
  return ((x * kDropMulEven) >> 16) | ((x * kDropMulOdd) & 0xFFFF0000u);
}
// threshold for the direct 32-bit compare: keep iff x * M >= dropout_threshold32(p)  (p < 1)
__host__ __device__ __forceinline__ uint32_t dropout_threshold32(uint32_t thr16) {
  return thr16 >= 65536u ? 0xFFFFFFFFu : (thr16 << 16);
}
// 64 bits = four 16-bit lanes for the 4 consecutive columns [4*col4, 4*col4 + 4)
__host__ __device__ __forceinline__ uint64_t dropout_bits_rk(uint32_t rowkey, uint32_t col4) {
  return (uint64_t)dropout_pair(rowkey, 2 * col4) | ((uint64_t)dropout_pair(rowkey, 2 * col4 + 1) << 32);
}
__host__ __device__ __forceinline__ uint64_t dropout_bits(uint64_t seed, uint64_t row, uint32_t col4) {
  return dropout_bits_rk(dropout_rowkey(seed, row), col4);
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
  float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)t);
}
// keep column (4*col4 + k) iff its 16-bit lane >= threshold
__host__ __device__ __forceinline__ bool dropout_keep(uint64_t bits, int k, uint32_t thr) {
  return ((uint32_t)(bits >> (16 * k)) & 0xFFFFu) >= thr;
}

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// streaming 128-bit load that does not pollute L1 (data read exactly once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
#endif

}  // namespace sgs
