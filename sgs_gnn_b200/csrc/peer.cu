// Row-slab exchange over NVLink peer memory (one graph sharded by destination range).
//
// Every rank owns one SYMMETRIC arena (torch.distributed._symmetric_memory: same size and layout on every GPU, all
// bases mapped into every process); `peer_bases` is the device array of the `world` arena base addresses.  A [N, D]
// buffer lives at the same element offset `elem_off` in every arena, so rank g's copy is peer_bases[g] + elem_off.
//
//  * all-gather of owned row slabs  (model.py:107-111 GCNConv outputs gathered by arbitrary sources):
//      the PRODUCER writes its finished rows straight into every peer's buffer -- the SpMM epilogue does it
//      (SpmmPeers in gcn.cu: the transfer overlaps the kernel's gathers row by row), dense-layer slabs go through
//      push_rows_kernel -- then one cross-GPU barrier, then each rank copies the foreign rows out of its own arena.
//  * reduce-scatter of partial sums  (the by-source SpMM / scorer backward produce partial rows for ALL nodes):
//      every rank leaves its partial [N, D] in its arena, one barrier, then reduce_rows_kernel sums the `world`
//      copies of the rows it owns with peer LOADS.
// Loads / stores on mapped peer pointers travel over NVLink; they bypass the local L2 (B300_MICROARCH "NVLink").
#include "common.cuh"

namespace sgs {

__global__ void push_rows_kernel(const float4* __restrict__ src, const uint64_t* __restrict__ peer_bases, int world,
                                 int rank, int64_t elem_off, int64_t n4, int include_self) {
  // grid.y = destination rank: every CTA streams the slab once per destination (the slab is L2-resident after the
  // first pass; peer stores are the bottleneck)
  const int g = blockIdx.y;
  if (g == rank && !include_self) return;
  float4* dstp = reinterpret_cast<float4*>(peer_bases[g] + (uint64_t)elem_off * sizeof(float));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    dstp[i] = src[i];
}

// scalar forms for slabs that are not 16-byte aligned / not a multiple of 4 floats (the [N, 41] logits, [N, 3] norm
// statistics: a few MB, latency-bound anyway)
__global__ void push_rows_scalar_kernel(const float* __restrict__ src, const uint64_t* __restrict__ peer_bases,
                                        int world, int rank, int64_t elem_off, int64_t n, int include_self) {
  const int g = blockIdx.y;
  if (g == rank && !include_self) return;
  float* dstp = reinterpret_cast<float*>(peer_bases[g] + (uint64_t)elem_off * sizeof(float));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dstp[i] = src[i];
}
__global__ void reduce_rows_scalar_kernel(const uint64_t* __restrict__ peer_bases, int world, int rank,
                                          int64_t elem_off, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int g = 0; g < world; ++g)
      acc += reinterpret_cast<const float*>(peer_bases[g] + (uint64_t)elem_off * sizeof(float))[i];
    out[i] = acc;
  }
}

__global__ void reduce_rows_kernel(const uint64_t* __restrict__ peer_bases, int world, int rank, int64_t elem_off,
                                   int64_t n4, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // fixed summation order (rank 0, 1, ...): every rank reduces ITS rows, and the result is reproducible
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < world; ++g) {
      const float4 v = reinterpret_cast<const float4*>(peer_bases[g] + (uint64_t)elem_off * sizeof(float))[i];
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    out[i] = acc;
  }
}

}  // namespace sgs

using namespace sgs;

extern "C" {

int32_t sgs_peer_push_rows(const float* src, const uint64_t* peer_bases, int32_t world, int32_t rank, int64_t elem_off,
                           int64_t row0, int64_t rows, int64_t D, int32_t include_self, sgs_stream_t stream) {
  SGS_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && rows >= 0 && D > 0 && row0 >= 0, "bad arguments");
  if (rows == 0) return SGS_OK;
  SGS_CHECK_ARG(src && peer_bases, "null pointer");
  const int64_t cap = (int64_t)sm_count() * 4;
  if ((rows * D) % 4 != 0 || ((elem_off + row0 * D) % 4) != 0 || ((uintptr_t)src & 15) != 0) {
    const int64_t n = rows * D;
    int64_t gs = ceil_div(n, 256);
    dim3 grid_s((unsigned)(gs > cap ? cap : gs), (unsigned)world);
    push_rows_scalar_kernel<<<grid_s, 256, 0, as_stream(stream)>>>(src, peer_bases, world, rank, elem_off + row0 * D, n,
                                                                   include_self);
    SGS_LAUNCH_CHECK();
    return SGS_OK;
  }
  const int64_t n4 = rows * D / 4;
  int64_t g = ceil_div(n4, 256);
  dim3 grid((unsigned)(g > cap ? cap : g), (unsigned)world);
  push_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(src), peer_bases, world, rank,
                                                        elem_off + row0 * D, n4, include_self);
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}

int32_t sgs_peer_reduce_rows(const uint64_t* peer_bases, int32_t world, int32_t rank, int64_t elem_off, int64_t row0,
                             int64_t rows, int64_t D, float* out, sgs_stream_t stream) {
  SGS_CHECK_ARG(world >= 1 && rows >= 0 && D > 0 && row0 >= 0, "bad arguments");
  if (rows == 0) return SGS_OK;
  SGS_CHECK_ARG(out && peer_bases, "null pointer");
  const int64_t cap = (int64_t)sm_count() * 8;
  if ((rows * D) % 4 != 0 || ((elem_off + row0 * D) % 4) != 0 || ((uintptr_t)out & 15) != 0) {
    const int64_t n = rows * D;
    int64_t gs = ceil_div(n, 256);
    reduce_rows_scalar_kernel<<<(unsigned)(gs > cap ? cap : gs), 256, 0, as_stream(stream)>>>(
        peer_bases, world, rank, elem_off + row0 * D, n, out);
    SGS_LAUNCH_CHECK();
    return SGS_OK;
  }
  const int64_t n4 = rows * D / 4;
  int64_t g = ceil_div(n4, 256);
  reduce_rows_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, as_stream(stream)>>>(
      peer_bases, world, rank, elem_off + row0 * D, n4, reinterpret_cast<float4*>(out));
  SGS_LAUNCH_CHECK();
  return SGS_OK;
}
}
