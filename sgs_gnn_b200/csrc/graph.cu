// Graph preparation: int64 COO -> int32, column gathers, CSR (by dst / by src) construction.
#include <cub/device/device_radix_sort.cuh>

#include <math.h>

#include "common.cuh"

namespace sgs {

__global__ void edge_index_split_kernel(const int64_t* __restrict__ ei, int64_t M, int64_t N,
                                        int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                        int32_t* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < M; i += stride) {
    int64_t s = ei[i], d = ei[M + i];
    bad |= (s < 0) | (s >= N) | (d < 0) | (d >= N);
    src[i] = (int32_t)s;
    dst[i] = (int32_t)d;
  }
  if (bad) *err = 1;
}

__global__ void edge_index_gather_kernel(const int64_t* __restrict__ ei, int64_t M,
                                         const int32_t* __restrict__ ids, int64_t q,
                                         int64_t* __restrict__ out, int32_t* __restrict__ so,
                                         int32_t* __restrict__ dout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < q; i += stride) {
    int64_t e = ids[i];
    int64_t s = ei[e], d = ei[M + e];
    if (out) {
      out[i] = s;
      out[q + i] = d;
    }
    if (so) so[i] = (int32_t)s;
    if (dout) dout[i] = (int32_t)d;
  }
}

// the int32 form of an edge list (host batches narrowed by Batch.compact()): range check only, no copy
__global__ void edge_index_check32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t M,
                                          int64_t N, int32_t* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < M; i += stride) {
    const int64_t s = src[i], d = dst[i];
    bad |= (s < 0) | (s >= N) | (d < 0) | (d >= N);
  }
  if (bad) *err = 1;
}

__global__ void edge_gather32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                     const int32_t* __restrict__ ids, int64_t q, int64_t* __restrict__ out,
                                     int32_t* __restrict__ so, int32_t* __restrict__ dout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < q; i += stride) {
    const int64_t e = ids[i];
    const int32_t s = src[e], d = dst[e];
    if (out) {
      out[i] = s;
      out[q + i] = d;
    }
    if (so) so[i] = s;
    if (dout) dout[i] = d;
  }
}

// Degree prior of datasets.py:141-156 (before its softmax): counts by atomics, then, op for op in fp32 as the
// reference evaluates it,  score_e = E^-1/2 / ( 1/(1/colcount[row_e]) + 1/(1/rowcount[col_e]) + 1e-10 ).
__global__ void degree_count_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t M,
                                    int32_t* __restrict__ rowcount, int32_t* __restrict__ colcount) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < M; i += stride) {
    atomicAdd(rowcount + src[i], 1);
    atomicAdd(colcount + dst[i], 1);
  }
}
__global__ void degree_score_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t M,
                                    const int32_t* __restrict__ rowcount, const int32_t* __restrict__ colcount,
                                    float scale, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < M; i += stride) {
    const float deg_in = __fdiv_rn(1.0f, (float)colcount[src[i]]);    // deg_in[row]
    const float deg_out = __fdiv_rn(1.0f, (float)rowcount[dst[i]]);   // deg_out[col]
    const float p = __fadd_rn(__fdiv_rn(1.0f, deg_in), __fdiv_rn(1.0f, deg_out));
    out[i] = __fmul_rn(__fdiv_rn(1.0f, __fadd_rn(p, 1e-10f)), scale);
  }
}

__global__ void iota_copy_kernel(const int32_t* __restrict__ key, int64_t M, int32_t* __restrict__ kout,
                                 int32_t* __restrict__ vout) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < M; i += stride) {
    kout[i] = key[i];
    vout[i] = (int32_t)i;
  }
}

// flag[0] |= 1 if keys are not non-decreasing
__global__ void unsorted_flag_kernel(const int32_t* __restrict__ key, int64_t M, int32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i + 1 < M; i += stride) bad |= key[i] > key[i + 1];
  if (bad) *flag = 1;
}

// perm = identity, nbr = other  (CSR of an edge list whose keys are already sorted)
__global__ void iota_nbr_kernel(const int32_t* __restrict__ other, int64_t M, int32_t* __restrict__ perm,
                                int32_t* __restrict__ nbr) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < M; i += stride) {
    perm[i] = (int32_t)i;
    nbr[i] = other[i];
  }
}

// rowptr[k] = first position in the sorted key array with key >= k   (k in [0, N])
__global__ void rowptr_kernel(const int32_t* __restrict__ sorted, int64_t M, int64_t N,
                              int32_t* __restrict__ rowptr) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > N) return;
  int64_t lo = 0, hi = M;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < (int32_t)k) lo = mid + 1; else hi = mid;
  }
  rowptr[k] = (int32_t)lo;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ idx,
                                  int64_t M, int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < M; i += stride) out[i] = src[idx[i]];
}

static inline int grid_for(int64_t n, int block, int waves = 8) {
  int64_t g = ceil_div(n, block);
  int64_t cap = (int64_t)sm_count() * waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static size_t cub_temp_bound(int64_t M) { return (size_t)(32u << 20) + (size_t)(M / 8) * 4; }

// keys for the degree sort: ~degree so that an ascending radix sort yields heaviest rows first
__global__ void degree_keys_kernel(const int32_t* __restrict__ rowptr, int64_t N, uint32_t* __restrict__ keys,
                                   int32_t* __restrict__ ids) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    keys[i] = ~(uint32_t)(rowptr[i + 1] - rowptr[i]);
    ids[i] = (int32_t)i;
  }
}

// order[N] = number of rows with more than SGS_HEAVY_ROW_DEG edges (they come first in `order`)
__global__ void heavy_count_kernel(const int32_t* __restrict__ rowptr, int64_t N, int32_t* __restrict__ count) {
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    c += (rowptr[i + 1] - rowptr[i]) > SGS_HEAVY_ROW_DEG;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

}  // namespace sgs

using namespace sgs;

extern "C" {

int32_t sgs_edge_index_split(const int64_t* edge_index, int64_t M, int64_t N, int32_t* src,
                             int32_t* dst, int32_t* err_flag, sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N >= 0 && N < (1ll << 31) && M < (1ll << 31), "M, N must fit int32");
  SGS_CHECK_ARG(M == 0 || (edge_index && src && dst && err_flag), "null pointer");
  if (M == 0) return SGS_OK;
  edge_index_split_kernel<<<grid_for(M, 256), 256, 0, as_stream(stream)>>>(edge_index, M, N, src, dst,
                                                                           err_flag);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_edge_index_gather(const int64_t* edge_index, int64_t M, const int32_t* ids, int64_t q,
                              int64_t* out, int32_t* src_out, int32_t* dst_out, sgs_stream_t stream) {
  SGS_CHECK_ARG(q >= 0 && M >= 0, "negative size");
  if (q == 0) return SGS_OK;
  SGS_CHECK_ARG(edge_index && ids, "null pointer");
  edge_index_gather_kernel<<<grid_for(q, 256), 256, 0, as_stream(stream)>>>(edge_index, M, ids, q, out,
                                                                            src_out, dst_out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_edge_index_check32(const int32_t* src, const int32_t* dst, int64_t M, int64_t N, int32_t* err_flag,
                               sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N > 0, "bad sizes");
  if (M == 0) return SGS_OK;
  SGS_CHECK_ARG(src && dst && err_flag, "null pointer");
  edge_index_check32_kernel<<<grid_for(M, 256), 256, 0, as_stream(stream)>>>(src, dst, M, N, err_flag);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_edge_gather32(const int32_t* src, const int32_t* dst, const int32_t* ids, int64_t q, int64_t* out,
                          int32_t* src_out, int32_t* dst_out, sgs_stream_t stream) {
  SGS_CHECK_ARG(q >= 0, "negative size");
  if (q == 0) return SGS_OK;
  SGS_CHECK_ARG(src && dst && ids, "null pointer");
  edge_gather32_kernel<<<grid_for(q, 256), 256, 0, as_stream(stream)>>>(src, dst, ids, q, out, src_out, dst_out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_degree_scores(const int32_t* src, const int32_t* dst, int64_t M, int64_t N, int32_t* counts, float* out,
                          sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N > 0, "bad sizes");
  if (M == 0) return SGS_OK;
  SGS_CHECK_ARG(src && dst && counts && out, "null pointer");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(counts, 0, (size_t)2 * N * sizeof(int32_t), st));
  degree_count_kernel<<<grid_for(M, 256), 256, 0, st>>>(src, dst, M, counts, counts + N);
  SGS_LAUNCH_CHECK();
  // len(prob) ** -0.5 is evaluated by Python in double and meets the fp32 tensor as an fp32 scalar
  const float scale = (float)pow((double)M, -0.5);
  degree_score_kernel<<<grid_for(M, 256), 256, 0, st>>>(src, dst, M, counts, counts + N, scale, out);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

size_t sgs_csr_workspace_bytes(int64_t M, int64_t N) {
  size_t m = (size_t)(M > N ? M : N);
  if (m < 1) m = 1;
  return 4 * m * sizeof(int32_t) + 1024 + cub_temp_bound((int64_t)m);
}

static int32_t degree_order(const int32_t* rowptr, int64_t N, int32_t* order, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  uint32_t* kA = (uint32_t*)ws;
  uint32_t* kB = kA + N;
  int32_t* vA = (int32_t*)(kB + N);
  int32_t* vB = vA + N;
  char* temp = (char*)(((uintptr_t)(vB + N) + 255) & ~(uintptr_t)255);
  size_t temp_avail = ws_bytes - (size_t)(temp - (char*)ws);
  degree_keys_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(rowptr, N, kA, vA);
  SGS_LAUNCH_CHECK();
  cub::DoubleBuffer<uint32_t> dk(kA, kB);
  cub::DoubleBuffer<int32_t> dv(vA, vB);
  size_t need = 0;
  SGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)N, 0, 32, st));
  if (need > temp_avail) {
    set_error("sgs_csr_build: cub temp %zu > available %zu", need, temp_avail);
    return SGS_E_WORKSPACE;
  }
  SGS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, dk, dv, (int)N, 0, 32, st));
  count_launch(4);
  SGS_CUDA(cudaMemcpyAsync(order, dv.Current(), N * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  SGS_CUDA(cudaMemsetAsync(order + N, 0, sizeof(int32_t), st));
  heavy_count_kernel<<<(unsigned)std::min<int64_t>(ceil_div(N, 256), 1024), 256, 0, st>>>(rowptr, N, order + N);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_csr_build(const int32_t* key, const int32_t* other, int64_t M, int64_t N, int32_t* rowptr,
                      int32_t* perm, int32_t* nbr, int32_t* order, void* ws, size_t ws_bytes,
                      sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N > 0, "bad sizes");
  SGS_CHECK_ARG(rowptr != nullptr, "null rowptr");
  cudaStream_t st = as_stream(stream);
  if (ws_bytes < sgs_csr_workspace_bytes(M, N) || !ws) {
    set_error("sgs_csr_build: workspace too small");
    return SGS_E_WORKSPACE;
  }
  if (M == 0) {
    SGS_CUDA(cudaMemsetAsync(rowptr, 0, (N + 1) * sizeof(int32_t), st));
    if (order) return degree_order(rowptr, N, order, ws, ws_bytes, st);
    return SGS_OK;
  }
  SGS_CHECK_ARG(key && other && perm && nbr, "null pointer");
  int32_t* kA = (int32_t*)ws;
  int32_t* kB = kA + M;
  int32_t* vA = kB + M;
  int32_t* vB = vA + M;
  char* temp = (char*)(vB + M);
  temp = (char*)(((uintptr_t)temp + 255) & ~(uintptr_t)255);
  size_t temp_avail = ws_bytes - (size_t)(temp - (char*)ws);
  iota_copy_kernel<<<grid_for(M, 256), 256, 0, st>>>(key, M, kA, vA);
  SGS_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 31 && (1ll << end_bit) < N) ++end_bit;
  cub::DoubleBuffer<int32_t> dk(kA, kB);
  cub::DoubleBuffer<int32_t> dv(vA, vB);
  size_t need = 0;
  SGS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)M, 0, end_bit, st));
  if (need > temp_avail) {
    set_error("sgs_csr_build: cub temp %zu > available %zu", need, temp_avail);
    return SGS_E_WORKSPACE;
  }
  SGS_CUDA(cub::DeviceRadixSort::SortPairs(temp, need, dk, dv, (int)M, 0, end_bit, st));
  count_launch(4);
  rowptr_kernel<<<(unsigned)ceil_div(N + 1, 256), 256, 0, st>>>(dk.Current(), M, N, rowptr);
  SGS_LAUNCH_CHECK();
  SGS_CUDA(cudaMemcpyAsync(perm, dv.Current(), M * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  gather_i32_kernel<<<grid_for(M, 256), 256, 0, st>>>(other, dv.Current(), M, nbr);
  SGS_LAUNCH_CHECK();
  if (order) return degree_order(rowptr, N, order, ws, ws_bytes, st);
  return SGS_OK;
}

int32_t sgs_keys_unsorted(const int32_t* key, int64_t M, int32_t* flag, sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && flag, "bad arguments");
  cudaStream_t st = as_stream(stream);
  SGS_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), st));
  if (M < 2) return SGS_OK;
  SGS_CHECK_ARG(key != nullptr, "null pointer");
  unsorted_flag_kernel<<<grid_for(M, 256), 256, 0, st>>>(key, M, flag);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_csr_build_sorted(const int32_t* key, const int32_t* other, int64_t M, int64_t N, int32_t* rowptr,
                             int32_t* perm, int32_t* nbr, int32_t* order, void* ws, size_t ws_bytes,
                             sgs_stream_t stream) {
  SGS_CHECK_ARG(M >= 0 && N > 0, "bad sizes");
  SGS_CHECK_ARG(rowptr != nullptr, "null rowptr");
  cudaStream_t st = as_stream(stream);
  if (ws_bytes < sgs_csr_workspace_bytes(M, N) || !ws) {
    set_error("sgs_csr_build_sorted: workspace too small");
    return SGS_E_WORKSPACE;
  }
  if (M == 0) {
    SGS_CUDA(cudaMemsetAsync(rowptr, 0, (N + 1) * sizeof(int32_t), st));
  } else {
    SGS_CHECK_ARG(key && other && perm && nbr, "null pointer");
    rowptr_kernel<<<(unsigned)ceil_div(N + 1, 256), 256, 0, st>>>(key, M, N, rowptr);
    SGS_LAUNCH_CHECK();
    iota_nbr_kernel<<<grid_for(M, 256), 256, 0, st>>>(other, M, perm, nbr);
    SGS_LAUNCH_CHECK();
  }
  if (order) return degree_order(rowptr, N, order, ws, ws_bytes, st);
  return SGS_OK;
}
}
