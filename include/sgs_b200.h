/*
 * sgs_b200.h -- C ABI of libsgs_b200.so: hand-written sm_100a CUDA kernels for the
 * SGS-GNN learned-sparsifier training step (edge scoring, exponential-race top-q edge
 * sampling, gcn_norm + GCNConv forward/backward, fused losses).
 *
 * The reference (anonymousauthors001/SGS-GNN) is pure Python on top of PyTorch /
 * PyTorch-Geometric and has NO FFI of its own.  Each entry
 * point below therefore cites the reference *Python* call site whose library-kernel
 * sequence it replaces; INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add (it is the one sgs_gnn_b200/_lib.py uses).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless the
 *     name ends in _host; all buffers (outputs, workspaces) are caller-allocated;
 *   - every function returns 0 (SGS_OK) or a negative SGS_E_* code; sgs_last_error() holds a
 *     thread-local message for the last failure;
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*); no hidden syncs
 *     unless documented; the library keeps no global mutable state;
 *   - edge ids / node ids are int32 inside the library (E, N < 2^31); the Python API's int64
 *     `edge_index[2,E]` is narrowed once by sgs_edge_index_split.
 */
#ifndef SGS_B200_H
#define SGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGS_OK 0
#define SGS_E_INVALID (-1)   /* bad argument (null pointer, negative size, unsupported width) */
#define SGS_E_CUDA (-2)      /* a CUDA runtime call or launch failed */
#define SGS_E_WORKSPACE (-3) /* workspace too small */
#define SGS_E_UNSUPPORTED (-4)

/* precision selector for the dense contractions (X.W^T, edge-scorer fc1) */
#define SGS_PREC_FP32 0 /* CUDA-core fp32 FFMA: parity mode (<= 1e-5 rel vs the fp32 reference) */
#define SGS_PREC_BF16 1 /* tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM            */
#define SGS_PREC_FP16 2 /* tcgen05 kind::f16, fp16 operands, fp32 accumulate in TMEM            */
#define SGS_PREC_TF32 3 /* tcgen05 kind::tf32 on fp32 operands                                  */

typedef void* sgs_stream_t;

const char* sgs_last_error(void);
int32_t sgs_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
int64_t sgs_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Graph preparation (replaces PyG's per-call COO handling inside GCNConv.propagate and the
 * `batch.edge_index[:, idx]` gathers at training_hybrid.py:48,83).
 * ---------------------------------------------------------------------------------------- */

/* edge_index int64 [2,M] (row 0 = src, row 1 = dst) -> int32 src[M], dst[M].
 * *err_flag (device int32, caller-zeroed) is set to 1 if any id is outside [0,N). */
int32_t sgs_edge_index_split(const int64_t* edge_index, int64_t M, int64_t N, int32_t* src,
                             int32_t* dst, int32_t* err_flag, sgs_stream_t stream);

/* out[2,q] = edge_index[:, ids]  (int64 out, int32 ids), also narrowed copies if non-null. */
int32_t sgs_edge_index_gather(const int64_t* edge_index, int64_t M, const int32_t* ids, int64_t q,
                              int64_t* out, int32_t* src_out, int32_t* dst_out, sgs_stream_t stream);

/* The same two steps for an edge list that is ALREADY int32 (host batches narrowed before the upload: node ids fit
 * 31 bits, int64 only doubles the PCIe bytes of training_hybrid.py:42's batch.to(device)): range check without a
 * copy, and the gather of (src, dst)[ids] (int64 out optional). */
int32_t sgs_edge_index_check32(const int32_t* src, const int32_t* dst, int64_t M, int64_t N, int32_t* err_flag,
                               sgs_stream_t stream);
int32_t sgs_edge_gather32(const int32_t* src, const int32_t* dst, const int32_t* ids, int64_t q, int64_t* out,
                          int32_t* src_out, int32_t* dst_out, sgs_stream_t stream);

/* The degree prior `data.prob` of datasets.py:141-156 BEFORE its softmax (finish with sgs_softmax_f32):
 * out[e] = len^-1/2 / (colcount[src_e] + rowcount[dst_e] + 1e-10), every fp32 operation as the reference performs it
 * (1/(1/count) included).  counts: int32 [2N] scratch (row counts | column counts). */
int32_t sgs_degree_scores(const int32_t* src, const int32_t* dst, int64_t M, int64_t N, int32_t* counts, float* out,
                          sgs_stream_t stream);

size_t sgs_csr_workspace_bytes(int64_t M, int64_t N);
/* Stable counting sort of the M edges by key (dst for the forward CSR, src for the backward
 * one): rowptr[N+1], perm[M] = edge ids in key order, nbr[M] = other[perm]; optional order[N+1]:
 * order[0..N) = row ids sorted by descending degree, order[N] = number of rows with more than
 * SGS_HEAVY_ROW_DEG edges.  The SpMM / SDDMM kernels give each of those hub rows to a whole
 * thread block and deal the rest to warps heaviest-first, so a power-law tail stays balanced. */
#define SGS_HEAVY_ROW_DEG 512
int32_t sgs_csr_build(const int32_t* key, const int32_t* other, int64_t M, int64_t N,
                      int32_t* rowptr, int32_t* perm, int32_t* nbr, int32_t* order, void* ws,
                      size_t ws_bytes, sgs_stream_t stream);
/* Same outputs for an edge list whose `key` column is ALREADY non-decreasing (e.g. the by-source view of
 * edges selected in ascending id from a (src,dst)-sorted edge_index): no sort -- rowptr by binary search,
 * perm = identity, nbr = other.  The caller vouches for the order; sgs_keys_unsorted checks it on the device
 * (flag[0] = 1 if some key[i] > key[i+1]). */
int32_t sgs_csr_build_sorted(const int32_t* key, const int32_t* other, int64_t M, int64_t N,
                             int32_t* rowptr, int32_t* perm, int32_t* nbr, int32_t* order, void* ws,
                             size_t ws_bytes, sgs_stream_t stream);
int32_t sgs_keys_unsorted(const int32_t* key, int64_t M, int32_t* flag, sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3a gcn_norm  (PyG gcn_norm + add_remaining_self_loops as called from model.py:107-111,
 * 159-161).  CSR is by destination.  w is the per-edge weight in ORIGINAL edge
 * order or NULL (all ones).  Outputs: deg[N] (weighted in-degree incl. self loop), dis[N] =
 * deg^-1/2 (0 where deg == 0), loopw[N] (self-loop weight: 1, or the weight of an input
 * self-loop edge), what[M] = dis[src]*w*dis[dst] in CSR order (0 for input self loops).
 * ---------------------------------------------------------------------------------------- */
int32_t sgs_gcn_norm(const int32_t* rowptr, const int32_t* perm, const int32_t* nbr, const float* w,
                     int64_t M, int64_t N, float* deg, float* dis, float* loopw, float* what,
                     sgs_stream_t stream);
/* what for a second CSR (by source) of the same edge set, from dis computed above. */
int32_t sgs_gcn_norm_apply(const int32_t* rowptr, const int32_t* perm, const int32_t* nbr,
                           const float* w, const float* dis, int64_t M, int64_t N, float* what,
                           sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3b SpMM  (PyG propagate: index_select + mul + scatter_add, model.py:107-111,159-161).
 * out[r,:] = act( sum_{i in [rowptr[r],rowptr[r+1])} what[i]*h[nbr[i],:] + selfw*h[r,:] + bias )
 * with selfw = dis[r]^2*loopw[r] (pass dis = NULL for no self term).  flags: bit0 = ReLU,
 * bit1 = dropout(p_drop, seed) after ReLU, bit2 = accumulate into out (after the activation), bit3 = out holds a
 * root term that joins the sum inside the activation.
 * Atomic-free segment reduction; one warp (or several for hub rows) per destination row.
 * ---------------------------------------------------------------------------------------- */
#define SGS_SPMM_RELU 1
#define SGS_SPMM_DROPOUT 2
#define SGS_SPMM_ACCUM 4
#define SGS_SPMM_ADD_ROOT 8 /* out holds a per-row term that is added BEFORE bias / ReLU / dropout (SAGEConv root weight) */
int32_t sgs_spmm(const int32_t* rowptr, const int32_t* nbr, const float* what,
                 const int32_t* order /* [N+1] from sgs_csr_build, may be NULL */, const float* dis, const float* loopw,
                 const float* h, int64_t N, int64_t D, const float* bias, float* out, int32_t flags,
                 float p_drop, uint64_t seed, sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3c backward helpers (autograd of GCNConv, training_hybrid.py:135)
 * ---------------------------------------------------------------------------------------- */
/* fp16 gather tables (fast mode of the D = 256 SpMM / SDDMM: half the gathered bytes, table L2-resident).
 * sgs_table_f16: out16[N*D] = fp16(in * S), S a power of two -- 1 when scaled == 0 (activations), else chosen from
 * max|in| so that it maps into [8192, 16384) (gradient tables).  tscale: device float[4] = {S, 1/S, scratch, -}.
 * sgs_spmm_h16 / sgs_gcn_edge_grad_h16: as sgs_spmm / sgs_gcn_edge_grad with h given as such a table; fp32
 * accumulation, 1/S applied in the epilogue.  D % 8 == 0, D <= 512. */
int32_t sgs_table_f16(const float* in, int64_t N, int64_t D, int32_t scaled, void* out16, float* tscale,
                      sgs_stream_t stream);
int32_t sgs_spmm_h16(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                     const float* dis, const float* loopw, const void* h16, const float* tscale, int64_t N, int64_t D,
                     const float* bias, float* out, int32_t flags, float p_drop, uint64_t seed, sgs_stream_t stream);
/* Sharded form (one graph split by destination range over the GPUs): only rows [row_lo, row_hi) are
 * computed; with peer_bases != NULL every finished row is ALSO stored into the same [N, D] buffer of every peer
 * (peer_bases[g] + elem_off floats; bases of the ranks' symmetric arenas, mapped over NVLink) -- the slab all-gather
 * of training_hybrid.py's GCN layers happens inside the SpMM epilogue.  h (fp32) or h16 + tscale (fp16 table). */
int32_t sgs_spmm_sharded(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                         const float* dis, const float* loopw, const float* h, const void* h16, const float* tscale,
                         int64_t N, int64_t D, const float* bias, float* out, int32_t flags, float p_drop,
                         uint64_t seed, int64_t row_lo, int64_t row_hi, const uint64_t* peer_bases, int32_t world,
                         int32_t rank, int64_t elem_off, sgs_stream_t stream);
/* Slab exchange over peer memory (csrc/peer.cu): push_rows stores the local slab src[rows, D] (rows row0 .. of an
 * [N, D] buffer) into every peer's buffer; reduce_rows sums the `world` copies of rows [row0, row0 + rows) with peer
 * loads, in rank order.  The caller separates writers and readers with a cross-GPU barrier. */
int32_t sgs_peer_push_rows(const float* src, const uint64_t* peer_bases, int32_t world, int32_t rank, int64_t elem_off,
                           int64_t row0, int64_t rows, int64_t D, int32_t include_self, sgs_stream_t stream);
int32_t sgs_peer_reduce_rows(const uint64_t* peer_bases, int32_t world, int32_t rank, int64_t elem_off, int64_t row0,
                             int64_t rows, int64_t D, float* out, sgs_stream_t stream);
/* Two GCN layers over the SAME graph in one gather pass: h16 is the fp16 table of [h_a | h_b] ([N, D], D = 2 x
 * width), out_a / out_b receive the two [N, D/2] results (bias [D] = [b_a | b_b], one fused ReLU / dropout epilogue).
 * The D = 256 SpMM is bound by the RATE of row gathers, not by their bytes (profiles/r02_notes.md), so the pair costs
 * about one pass: the scorer's gcn1 and the random baseline's gcn1 (model.py:107, :159 over the same random subgraph,
 * training_hybrid.py:45-48,93) share it. */
int32_t sgs_spmm_h16_pair(const int32_t* rowptr, const int32_t* nbr, const float* what, const int32_t* order,
                          const float* dis, const float* loopw, const void* h16, const float* tscale, int64_t N,
                          int64_t D, const float* bias, float* out_a, float* out_b, int32_t flags, float p_drop,
                          uint64_t seed, sgs_stream_t stream);
int32_t sgs_gcn_edge_grad_h16(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                              const float* what_dst, const int32_t* order_dst, const int32_t* rowptr_src,
                              const int32_t* perm_src, const int32_t* src, const int32_t* dst, const float* G,
                              const void* h16, const float* tscale, const float* dis, const float* deg,
                              const float* loopw, int64_t M, int64_t N, int64_t D, float* tmp_g, float* tmp_t,
                              float* tmp_a, float* dw, int32_t accumulate, sgs_stream_t stream);
/* gin = gout * (out > 0) * scale      (ReLU + inverted-dropout backward; out is the saved
 * post-activation output, scale = 1/(1-p)) */
int32_t sgs_act_bwd(const float* gout, const float* out, int64_t n, float scale, float* gin,
                    sgs_stream_t stream);
/* colsum[D] = sum_n G[n,:]  (bias gradient) */
int32_t sgs_colsum(const float* G, int64_t N, int64_t D, float* colsum, sgs_stream_t stream);
/* dL/dw_e for one GCNConv (A.3): SDDMM over the by-dst CSR + the two per-node sums.
 * tmp_g[M], tmp_t[M], tmp_a[N] are scratch.  dw[M] in original edge order; accumulate != 0
 * adds to dw (both GCN layers share one edge_weight). */
int32_t sgs_gcn_edge_grad(const int32_t* rowptr_dst, const int32_t* perm_dst, const int32_t* nbr_dst,
                          const float* what_dst, const int32_t* order_dst /* may be NULL */,
                          const int32_t* rowptr_src, const int32_t* perm_src,
                          const int32_t* src, const int32_t* dst, const float* G, const float* h,
                          const float* dis, const float* deg, const float* loopw, int64_t M,
                          int64_t N, int64_t D, float* tmp_g, float* tmp_t, float* tmp_a, float* dw,
                          int32_t accumulate, sgs_stream_t stream);
/* The same computation in two phases for a caller whose edges are sharded by destination across
 * GPUs: _partial runs the SDDMM and the per-node sums over the local edges (G must be
 * zero outside the owned rows), the caller all-reduces tmp_a[N], _final applies the formula. */
int32_t sgs_gcn_edge_grad_partial(const int32_t* rowptr_dst, const int32_t* perm_dst,
                                  const int32_t* nbr_dst, const float* what_dst,
                                  const int32_t* order_dst /* may be NULL */, const int32_t* rowptr_src,
                                  const int32_t* perm_src, const float* G, const float* h,
                                  const float* dis, const float* loopw, int64_t M, int64_t N, int64_t D,
                                  float* tmp_g, float* tmp_t, float* tmp_a, sgs_stream_t stream);
/* _partial with the rows of h gathered from its fp16 table (sgs_table_f16), as sgs_gcn_edge_grad_h16 does. */
int32_t sgs_gcn_edge_grad_partial_h16(const int32_t* rowptr_dst, const int32_t* perm_dst,
                                      const int32_t* nbr_dst, const float* what_dst,
                                      const int32_t* order_dst /* may be NULL */, const int32_t* rowptr_src,
                                      const int32_t* perm_src, const float* G, const void* h16,
                                      const float* tscale, const float* dis, const float* loopw, int64_t M,
                                      int64_t N, int64_t D, float* tmp_g, float* tmp_t, float* tmp_a,
                                      sgs_stream_t stream);
int32_t sgs_gcn_edge_grad_final(const int32_t* src, const int32_t* dst, const float* tmp_g,
                                const float* tmp_a, const float* dis, const float* deg, int64_t M,
                                float* dw, int32_t accumulate, sgs_stream_t stream);

/* out[i] = in[i] rounded to the nearest tf32 value (10 mantissa bits; cvt.rna).  kind::tf32 MMAs truncate their
 * fp32 operands; rounding them first makes the error unbiased (in == out allowed). */
int32_t sgs_round_tf32(const float* in, int64_t n, float* out, sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4 dense contraction  C[M,N] (+)= A[M,K] . B[N,K]^T with arbitrary element strides
 * (NT / NN / TN forms of x.W^T, dh^T.x, dh.W; model.py:107-108,159-161 `lin`, and their
 * autograd).  precision: SGS_PREC_FP32 (CUDA-core) or a tensor-core mode (tcgen05 + TMA,
 * requires unit stride along K for both operands: a_sk == b_sk == 1).
 * ---------------------------------------------------------------------------------------- */
int32_t sgs_gemm(const float* A, int64_t a_sm, int64_t a_sk, const float* B, int64_t b_sn,
                 int64_t b_sk, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                 int32_t accumulate, int32_t precision, sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1 / K1b edge scorer  (model.py:115-122 `_edge_score`; identical body at :29-34, :71-78):
 *   p_e = sigmoid( w2 . dropout(relu(W1 . [x*y | x-y] + b1)) + b2 ),  x = out[src_e], y = out[dst_e]
 * ids == NULL scores edges [0,n) of (src,dst); otherwise edges ids[0..n).  The dropout keep
 * mask is a counter-based hash of (seed, edge id, column) so the backward regenerates it.
 * W1 is [H,2H] row-major, w2 [H], b2 [1].  The [E,2H] feature tensor is never materialised
 * in HBM in the tensor-core modes; the fp32 parity mode works in bounded chunks of `ws`.
 * ---------------------------------------------------------------------------------------- */
size_t sgs_edge_score_workspace_bytes(int64_t n, int64_t N, int64_t H, int32_t precision,
                                      int32_t backward);
int32_t sgs_edge_score_fwd(const float* out, int64_t N, int64_t H, const int32_t* src,
                           const int32_t* dst, const int32_t* ids, int64_t n, const float* W1,
                           const float* b1, const float* w2, const float* b2, float p_drop,
                           uint64_t seed, float* p, void* ws, size_t ws_bytes, int32_t precision,
                           sgs_stream_t stream);
/* Backward over the same edge list given dp[n] (upstream dL/dp_e, compact: dp[i] belongs to
 * edge ids[i] or i).  Accumulates (+=) into d_out[N,H], dW1[H,2H], db1[H], dw2[H], db2[1].
 * p_fwd[n]: the forward probabilities of these edges (required by the tensor-core modes, which
 * use dz = dp*p*(1-p); the fp32 mode recomputes p and ignores it, may be NULL there). */
int32_t sgs_edge_score_bwd(const float* out, int64_t N, int64_t H, const int32_t* src,
                           const int32_t* dst, const int32_t* ids, int64_t n, const float* W1,
                           const float* b1, const float* w2, const float* b2, float p_drop,
                           uint64_t seed, const float* p_fwd, const float* dp, float* d_out, float* dW1,
                           float* db1, float* dw2, float* db2, void* ws, size_t ws_bytes,
                           int32_t precision, sgs_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * K2 sampler  (sampling.py:91-155 gumbel_softmax_sampling + the compaction at
 * training_hybrid.py:83,86 and the baseline draw at :45-48).
 *   key_e = s_e / noise_e,  s = p/(S+1e-12) [* (1-coef) + coef*prob when mode == TRAIN],
 * evaluated op-for-op in IEEE fp32 without FMA contraction; the q largest keys are selected
 * by an MSD radix select on the key bit patterns (keys >= 0: 11+11+9 bits below the sign); ties at the threshold are
 * broken by LOWEST edge id; outputs are compacted in ascending edge id.
 * The steps are exported separately so a multi-GPU caller can all-reduce the digit
 * histograms (2048 x int64) between them; sgs_sample_topq runs the single-GPU sequence.
 * ---------------------------------------------------------------------------------------- */
#define SGS_SAMPLE_TRAIN 0 /* s = (1-coef)*p/S + coef*prob   (sampling.py:93-95) */
#define SGS_SAMPLE_TEST 1  /* s = p/S                        (istest, sampling.py:94) */
#define SGS_SAMPLE_RAW 2   /* s = p  (already a distribution, e.g. softmax(prob), training_hybrid.py:46) */
#define SGS_TOPQ_BINS 2048

/* S_out[0] = sum(p) accumulated in fp64, rounded once to fp32 (deterministic). ws: >= 8*1024 B */
int32_t sgs_sum_f32(const float* p, int64_t n, float* S_out, void* ws, size_t ws_bytes,
                    sgs_stream_t stream);
/* out = softmax(in) over n elements (two-pass, fp32).  ws >= 16 KiB. */
int32_t sgs_softmax_f32(const float* in, int64_t n, float* out, void* ws, size_t ws_bytes,
                        sgs_stream_t stream);
/* noise[i] = Exp(1) sample from a counter-based generator keyed (seed, i). */
int32_t sgs_exponential_f32(float* noise, int64_t n, uint64_t seed, sgs_stream_t stream);
/* noise[i] = element gid[i] of the contiguous draw above (same seed): the noise of a destination-sharded edge
 * list keyed by GLOBAL edge id, so the sampled set does not depend on the number of shards. */
int32_t sgs_exponential_ids_f32(float* noise, const int64_t* gid, int64_t n, uint64_t seed, sgs_stream_t stream);

/* state layout (device, int64[8]): [0]=prefix bits so far, [1]=remaining k, [2]=tau bits,
 * [3]=#keys > tau, [4]=#ties to take, [5]=invalid-input flag, [6]=#keys == tau, [7]=reserved */
/* one_minus_coef and coef are the two fp32 scalars of sampling.py:95 exactly as Python evaluates
 * them: float32(1 - degree_bias_coef) (the subtraction done in double) and float32(degree_bias_coef). */
int32_t sgs_topq_keys(const float* p, const float* prob, const float* noise, int64_t E,
                      float one_minus_coef, float coef, int32_t mode, const float* S, uint32_t* keys,
                      int64_t* hist /*[2048], zeroed here*/, int64_t* state /*[8], zeroed here*/,
                      sgs_stream_t stream);
/* level 0: consume hist of bits [30:20]; level 1: [19:9]; level 2: [8:0] (then tau is final).
 * Zeroes hist for the next level. */
int32_t sgs_topq_find(int64_t* hist, int64_t* state, int64_t k_total, int32_t level, sgs_stream_t stream);
/* histogram of the next digit over keys matching state's prefix (level 1 or 2) */
int32_t sgs_topq_hist(const uint32_t* keys, int64_t E, int64_t* hist, const int64_t* state, int32_t level,
                      sgs_stream_t stream);
size_t sgs_topq_workspace_bytes(int64_t E);
/* Compaction.  tie_skip = number of threshold ties owned by lower-ranked shards (0 on one GPU).
 * Outputs (any may be NULL): sel[q_cap] ascending edge ids, mask[E] (0/1 bytes),
 * n_sel_out (device int64: number written).  ws per sgs_topq_workspace_bytes. */
int32_t sgs_topq_compact(const uint32_t* keys, int64_t E, const int64_t* state, int64_t tie_skip,
                         int32_t* sel, int64_t q_cap, uint8_t* mask, int64_t* n_sel_out, void* ws,
                         size_t ws_bytes, sgs_stream_t stream);
/* Single-GPU convenience: keys -> 3 x (find, hist) -> compact. keys[E] and ws are scratch. */
int32_t sgs_sample_topq(const float* p, const float* prob, const float* noise, int64_t E, int64_t q,
                        float one_minus_coef, float coef, int32_t mode, const float* S, uint32_t* keys,
                        int32_t* sel, uint8_t* mask, int64_t* state /*[8]; [7] = #selected*/, void* ws,
                        size_t ws_bytes /* >= sgs_topq_workspace_bytes(E) + 2048*8 */, sgs_stream_t stream);
/* Per selected edge: p_sel[i] = p[sel[i]]; if w_st != NULL the straight-through weight
 * clamp(p * ((1 - s) + s), 0, 1) of sampling.py:137-155. */
int32_t sgs_gather_selected(const float* p, const float* prob, const int32_t* sel, int64_t q,
                            float one_minus_coef, float coef, int32_t mode, const float* S, float* p_sel,
                            float* w_st, sgs_stream_t stream);
/* dst[sel[i]] (+)= src[i]  (backward of p_full[mask], training_hybrid.py:86) */
int32_t sgs_scatter_selected(const float* src, const int32_t* sel, int64_t q, float* dst,
                             sgs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K5 fused losses (training_hybrid.py:94-132, utils.py:163-169,187-211)
 *   CE over train rows + c1*[sum_label>1]*BCE(p_s, same-class label) over sampled edges with
 *   both endpoints in the train mask + c2*MSE(p_s, cos(logits[src], logits[dst])).
 * acc: device double[8] = {ce_sum, n_train, correct, bce_sum, n_valid, sum_label, mse_sum, q}
 * row_mask (may be NULL = train_mask): the rows that contribute to the CE / accuracy sums; a
 * multi-GPU caller whose rows are sharded passes train_mask AND owned-rows here while the edge
 * terms keep testing both endpoints against the full train_mask.
 * ---------------------------------------------------------------------------------------- */
int32_t sgs_loss_fwd(const float* logits, int64_t N, int64_t C, const int64_t* y,
                     const uint8_t* train_mask, const uint8_t* row_mask, const int32_t* s_src, const int32_t* s_dst,
                     const float* p_s, int64_t q, int32_t with_edges, double* acc, sgs_stream_t stream);
/* loss_out[0] = c0*ce + c1*bce*[sum_label>1] + c2*mse (device float) from acc; a term whose
 * coefficient / flag is 0 is skipped entirely (c0 = 0 gives the stand-alone consistency loss). */
int32_t sgs_loss_finish(const double* acc, float c0, float c1, float c2, int32_t reg1, int32_t reg2,
                        float* loss_out, sgs_stream_t stream);
/* grads scaled by *gscale (device float, upstream dL/dloss): dlogits[N,C] (written: caller
 * zeroes), dp_s[q] (written). */
int32_t sgs_loss_bwd(const float* logits, int64_t N, int64_t C, const int64_t* y,
                     const uint8_t* train_mask, const uint8_t* row_mask, const int32_t* s_src, const int32_t* s_dst,
                     const float* p_s, int64_t q, int32_t with_edges, const double* acc, float c0,
                     float c1, float c2, int32_t reg1, int32_t reg2, const float* gscale,
                     float* dlogits, float* dp_s, sgs_stream_t stream);


/* One-sweep variant of the two calls above for the learned step (training_hybrid.py:103-135): the forward pass
 * over the sampled edges also leaves the UNSCALED gradients of the edge terms behind --
 *   u_reg1[q] = (p - label) / max(p (1 - p), 1e-12) (0 outside the train mask), u_reg2[q] = p - cos,
 *   dlog_e[N,C] = sum over incident sampled edges of (cos - p) * d cos / d logits   (zeroed here) --
 * so that sgs_loss_bwd_fused is two streaming passes (dp_s = g1 u_reg1 + g2 u_reg2, dlogits = CE part + g2 dlog_e)
 * with the scalars g1 = g c1 / n_valid [sum_label > 1], g2 = 2 g c2 / q read from acc on the device.
 * node_code: int32 [N] scratch (train ? y : -1).  Needs C <= 64; the sampled edges must ascend by source (they do:
 * ascending edge id of a (src,dst)-sorted edge list) for the source-side run accumulation to pay off -- any order
 * is still correct. */
int32_t sgs_loss_fwd_fused(const float* logits, int64_t N, int64_t C, const int64_t* y,
                           const uint8_t* train_mask, const uint8_t* row_mask, const int32_t* s_src,
                           const int32_t* s_dst, const float* p_s, int64_t q, double* acc, float* dlog_e,
                           float* u_reg1, float* u_reg2, int32_t* node_code, sgs_stream_t stream);
int32_t sgs_loss_bwd_fused(const float* logits, int64_t N, int64_t C, const int64_t* y,
                           const uint8_t* train_mask, const uint8_t* row_mask, int64_t q, const double* acc,
                           float c0, float c1, float c2, int32_t reg1, int32_t reg2, const float* gscale,
                           const float* dlog_e, const float* u_reg1, const float* u_reg2, float* dlogits,
                           float* dp_s, sgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SGS_B200_H */
