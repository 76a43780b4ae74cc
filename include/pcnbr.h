/*
 * pcnbr.h -- C ABI of libpcnbr.so: the B200 (sm_100a) point-cloud neighbourhood hot path.
 *
 * The reference (piotr-bledowski/3D-Semantic-Segmentation-Benchmark) has no FFI layer: the seam is
 * the set of Python functions in models/utils/common.py and models/dgcnn/dgcnn.py.  Each entry point
 * below replaces the op sequence of the cited reference lines; the Python host layer
 * (3d-semantic-segmentation-benchmark_b200/) binds them with ctypes and keeps the reference's
 * signatures.  INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain device pointers, row-major contiguous unless a stride argument says otherwise;
 *     the caller owns every buffer (including workspaces, sized by the *_ws_bytes helpers);
 *   - fp32 values, int32 indices (the reference's int64 topk indices are widened by the host);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it, never
 *     synchronises the host, keeps no global state and is re-entrant;
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or PCNBR_E_* (<0); never throws.
 *   - selection order everywhere: ascending (key, index) -- "lowest index wins" on equal keys.
 */
#ifndef PCNBR_H_
#define PCNBR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCNBR_ABI_VERSION 1

#define PCNBR_E_BADARG   (-1)   /* null pointer, non-positive size, K > N ... */
#define PCNBR_E_TOOLARGE (-2)   /* outside the compiled limits (K > 128, F > 256, ...) */
#define PCNBR_E_WORKSPACE (-3)  /* workspace missing or too small */

#if defined(__GNUC__)
#define PCNBR_API __attribute__((visibility("default")))
#else
#define PCNBR_API
#endif

typedef void* pcnbr_stream_t;

PCNBR_API int pcnbr_abi_version(void);
/* Human-readable text for a return code of this library. */
PCNBR_API const char* pcnbr_error_string(int code);

/* ---- K1 farthest point sampling ------------------------------------------ common.py:6-34
 * xyz (B,N,3); start (B) first pick per cloud (the reference's randint draw, common.py:22);
 * idx_out (B,C); xyz_out (B,C,3) = xyz[b, idx] (what sample() returns), may be NULL.
 * dist = sqrt_rn(fma(dz,dz,fma(dy,dy,dx*dx))), running min, argmax with lowest index on ties.
 * ws: pcnbr_fps_ws_bytes(B,N) bytes (0 when N <= 8192). */
PCNBR_API size_t pcnbr_fps_ws_bytes(int B, int N);
PCNBR_API int pcnbr_fps_f32(const float* xyz, int B, int N, int C, const int32_t* start,
                  int32_t* idx_out, float* xyz_out, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Length-aware K1 (zero-padded evaluation batches, data_processing/block_datasets.py:19-25,
 * Training/training.py:80-133): cloud b holds n_valid[b] real points (device array, clamped to [1, N]; NULL = N), the rows
 * behind them are padding and take no part -- picks and coordinates are those of the reference's sample() on the cloud
 * passed alone, unpadded (start[b] is clamped into [0, n_valid[b])). */
PCNBR_API int pcnbr_fps_len_f32(const float* xyz, int B, int N, int C, const int32_t* start, const int32_t* n_valid,
                      int32_t* idx_out, float* xyz_out, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* ---- K2 ball query ------------------------------------------------------- common.py:54-61
 * q (B,M,3) centroids, p (B,N,3) points, r2 = (float)((double)r*r).  idx (B,M,K):
 * in-ball points by ascending (d2, index), then out-of-ball points by ascending index.
 * d2 = (dx*dx + dy*dy) + dz*dz without contraction.  1 <= K <= min(N,128). */
PCNBR_API int pcnbr_ball_query_f32(const float* q, const float* p, int B, int M, int N, float r2, int K,
                         int32_t* idx, pcnbr_stream_t stream);

/* ---- K3 kNN on xyz, direct distances ----------------------------------- common.py:110-114
 * Same distance form, no radius.  idx (B,M,K) ascending (d2,index); d2 (B,M,K) may be NULL. */
PCNBR_API int pcnbr_knn_direct_f32(const float* q, const float* p, int B, int M, int N, int K,
                         int32_t* idx, float* d2, pcnbr_stream_t stream);

/* ---- K2 / K3 on a uniform cell grid (csrc/grid.cu): the same tables as pcnbr_ball_query_f32 / pcnbr_knn_direct_f32 (k <= 32),
 * bit for bit, without the M x N scan: every cloud is binned once into cells of edge >= 1.001 r (ball query) or ~2 points
 * per cell (k-NN), a query reads the 27 cells around it (k-NN: rings of cells until the k-th distance is inside the
 * scanned cube), and under-filled balls are padded from the member list (lowest indices not in the ball).
 * ws: pcnbr_grid_ws_bytes(B, N) with N = number of SOURCE points per cloud. */
PCNBR_API size_t pcnbr_grid_ws_bytes(int B, int N);
PCNBR_API int pcnbr_ball_query_grid_f32(const float* q, const float* p, int B, int M, int N, float r2, int K, int32_t* idx,
                    void* ws, size_t ws_bytes, pcnbr_stream_t stream);
PCNBR_API int pcnbr_knn_direct_grid_f32(const float* q, const float* p, int B, int M, int N, int K, int32_t* idx, float* d2,
                    void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Length-aware K2 / K3: n_src[b] real source points, n_qry[b] real queries per cloud (device arrays,
 * NULL = all).  Real rows get exactly the table of the reference's group() / interpolate() on the unpadded cloud; padding
 * query rows get the in-range filler 0..K-1 (and d2 = 0).  K <= min_b n_src[b] is the caller's precondition (the reference's
 * topk raises otherwise); a violated precondition yields in-range indices, never an out-of-bounds one.
 * ws != NULL (pcnbr_grid_ws_bytes(B, N)): cell grid; ws == NULL: M x N scan.  pcnbr_knn_direct_len_f32 with ws needs K <= 32. */
PCNBR_API int pcnbr_ball_query_len_f32(const float* q, const float* p, int B, int M, int N, float r2, int K, const int32_t* n_qry,
                    const int32_t* n_src, int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream);
PCNBR_API int pcnbr_knn_direct_len_f32(const float* q, const float* p, int B, int M, int N, int K, const int32_t* n_qry,
                    const int32_t* n_src, int32_t* idx, float* d2, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Multi-radius ball query ("MSG": several group() calls on ONE centroid set with different (r, K); BASELINE configs[2]).
 * Each idx[i] (B,M,K[i]) is bit-identical to pcnbr_ball_query_f32(q, p, .., r2[i], K[i], idx[i]), but the points are
 * scanned once: one selection with the largest radius and the largest K, the other scales are derived from its sorted
 * list (in-ball prefix + lowest-index out-of-ball padding).  r2, K and idx are HOST arrays of R <= 8 entries (idx holds
 * device pointers); ws: pcnbr_ball_query_multi_ws_bytes(B, M, max K) bytes of device scratch (unused when R == 1). */
PCNBR_API size_t pcnbr_ball_query_multi_ws_bytes(int B, int M, int Kmax);
PCNBR_API int pcnbr_ball_query_multi_f32(const float* q, const float* p, int B, int M, int N, const float* r2, const int* K,
                               int R, int32_t* const* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

PCNBR_API int pcnbr_ball_query_multi_len_f32(const float* q, const float* p, int B, int M, int N, const float* r2, const int* K,
                               int R, const int32_t* n_src, int32_t* const* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* ---- K3/K4 kNN in the reference's expanded form ----------------------- dgcnn.py:7-21
 * x[b, f*stride_f + n*stride_n] (batch stride F*N), any layout of the (F,N) plane.
 * pd_ij = ((-xx_j) - (-2*c_ij)) - xx_i, c = FMA chain over f ascending, xx = ATen cascade sum of
 * squares; idx (B,N,K) by descending pd, lowest index on ties.  F <= 256, K <= min(N,128).
 * F <= 64, 256 <= N <= 65535, K <= 32 run on the tensor cores: tcgen05 kind::f16 tiles of the centred, scaled
 * features fed by TMA, with the |x_j|^2 term, the per-pair fp16 error bound and the row threshold folded into a
 * 16-wide tail K-slice, so the epilogue reads finished scores out of TMEM (group maxima, then sign bits); the
 * survivors are re-ranked in exact fp32: same bits as the CUDA-core path (PCNBR_KNN_GENERIC=1 forces the latter).
 * ws: pcnbr_knn_expand_ws_bytes(B,F,N,K) bytes. */
PCNBR_API size_t pcnbr_knn_expand_ws_bytes(int B, int F, int N, int K);
PCNBR_API int pcnbr_knn_expand_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                         int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Length-aware form: cloud b holds n_valid[b] real points (NULL = N).  Rows i < n_valid[b] get the k
 * nearest among the real points exactly as the reference's knn() on the unpadded cloud (including ATen's length-dependent
 * summation order of |x|^2); padding rows get the filler 0..K-1.  The tensor-core path skips the all-padding work units and
 * column tiles.  K <= min_b n_valid[b] is the caller's precondition. */
PCNBR_API int pcnbr_knn_expand_len_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                             const int32_t* n_valid, int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Test hook of the tensor-core path (F <= 64, 256 <= N <= 65535, K <= 32): same result as
 * pcnbr_knn_expand_f32, plus scores (B,N,N) = the pass-1 tensor-core values a_i.a_j - |a_j|^2/2 - C1|a_i||a_j| of the
 * centred, scaled features a (lower bounds of the exact ranking score; may be NULL) and stats[2] = {sum of survivor-queue lengths, rows that overflowed to the exact full scan}. */
PCNBR_API int pcnbr_knn_tc_debug_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                           int32_t* idx, void* ws, size_t ws_bytes, float* scores, int32_t* stats,
                           pcnbr_stream_t stream);

/* ---- K5 gather + centre-subtract (+ /r) + concat ------------------------ common.py:62-71
 * p (B,N,3), feat (B,N,D) (D may be 0), q (B,M,3), idx (B,M,K) -> out (B,M,K,3+D) with row pitch ldo >= 3+D floats
 * (ldo == 3+D: the reference layout; ldo = 3+D rounded up to 4 gives the 16-byte pitch the tensor-core GEMM reads in
 * place -- the pad columns are written as zeros).
 * rdiv > 0: local coordinates are divided (true fp32 division) by rdiv (common.py:69). */
PCNBR_API int pcnbr_group_f32(const float* p, const float* feat, const float* q, const int32_t* idx,
                    int B, int N, int M, int K, int D, float rdiv, float* out, int ldo, pcnbr_stream_t stream);

/* ---- index_points / square_distance (north-star names; common.py:64-65,117 and :54-56) -----------------------------
 * pcnbr_gather_rows_f32: out (B,E,D) = src[b, idx[b,e], :], src (B,N,D), idx (B,E) (out-of-range indices are clamped).
 * pcnbr_gather_rows_bwd_f32: gsrc (B,N,D) = sum over incoming e of gout[b,e,:], CSR of idx from pcnbr_csr_build; fixed order.
 * pcnbr_square_distance_f32: out (B,N,M) = ((dst[b,m] - src[b,n])^2).sum(-1) as (dx*dx + dy*dy) + dz*dz -- the matrix the
 * reference materialises and no kernel of this library needs; kept for callers that inspect it. */
PCNBR_API int pcnbr_gather_rows_f32(const float* src, const int32_t* idx, int B, int N, long E, int D, float* out, pcnbr_stream_t stream);
PCNBR_API int pcnbr_gather_rows_bwd_f32(const float* gout, const int32_t* offsets, const int32_t* perm, int B, int N, int E, int D,
                              float* gsrc, pcnbr_stream_t stream);
PCNBR_API int pcnbr_square_distance_f32(const float* src, const float* dst, int B, int N, int M, float* out, pcnbr_stream_t stream);

/* ---- K7 inverse index (CSR by source point) for the atomic-free scatter-add backward
 * idx (B,E) values in [0,N).  offsets (B,N+1): segment bounds; perm (B,E): positions e grouped
 * by source, ascending e inside a segment (deterministic).  ws: pcnbr_csr_ws_bytes(B,E,N). */
PCNBR_API size_t pcnbr_csr_ws_bytes(int B, int E, int N);
PCNBR_API int pcnbr_csr_build(const int32_t* idx, int B, int E, int N, int32_t* offsets, int32_t* perm,
                    void* ws, size_t ws_bytes, pcnbr_stream_t stream);
/* The same inverse for a row-structured table idx (B,M,K) (every neighbour table of the path: E = M*K): built by bitmap
 * transposition -- mark (source, row) bits, scan the counts, emit every source's rows in ascending order -- with no sort
 * pass; rows may hold a source more than once.  Falls back to pcnbr_csr_build when the (N x M)-bit matrices would not fit
 * the workspace budget.  ws: pcnbr_csr_rows_ws_bytes(B,M,K,N). */
PCNBR_API size_t pcnbr_csr_rows_ws_bytes(int B, int M, int K, int N);
PCNBR_API int pcnbr_csr_build_rows(const int32_t* idx, int B, int M, int K, int N, int32_t* offsets, int32_t* perm,
                    void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* Backward of K5 w.r.t. feat (autograd IndexBackward of common.py:65):
 * gout (B,M,K,*) with row pitch ldg >= 3+D -> gfeat (B,N,D) = sum over incoming (m,k) of gout[..., 3:3+D], fixed order. */
PCNBR_API int pcnbr_group_bwd_f32(const float* gout, int ldg, const int32_t* offsets, const int32_t* perm,
                        int B, int N, int E, int D, float* gfeat, pcnbr_stream_t stream);

/* ---- K6 grouped max-pool over K (+ argmax) -------------- common.py:85-86, dgcnn.py:76
 * x[r*stride_r + k*stride_k + d*stride_d], r < R rows, reduce k < K; out (R,D) contiguous,
 * arg (R,D) uint8 (first max wins, as torch.max).  One of stride_d / stride_k must be 1. */
PCNBR_API int pcnbr_maxpool_f32(const float* x, long R, int K, int D, long stride_r, long stride_k, long stride_d,
                      float* out, uint8_t* arg, pcnbr_stream_t stream);
/* gx has the same strides as x: gx[r,k,d] = (k == arg[r,d]) ? g[r,d] : 0. */
PCNBR_API int pcnbr_maxpool_bwd_f32(const float* g, const uint8_t* arg, long R, int K, int D, long stride_r,
                          long stride_k, long stride_d, float* gx, pcnbr_stream_t stream);

/* ---- K8 three-point interpolation ------------------------------------- common.py:115-122
 * feat (B,M,D) coarse, idx/d2 (B,N,K) from pcnbr_knn_direct_f32 (K <= 8) -> out (B,N,D);
 * coef (B,N,K) receives w_k/norm (saved for the backward).  w = 1/(d2+1e-9f). */
PCNBR_API int pcnbr_interp_f32(const float* feat, const int32_t* idx, const float* d2, int B, int N, int M, int D,
                     int K, float* out, float* coef, pcnbr_stream_t stream);
/* gfeat (B,M,D) = sum over incoming (n,k) of coef[n,k] * g[n,:]; CSR over idx viewed as (B, N*K). */
PCNBR_API int pcnbr_interp_bwd_f32(const float* g, const float* coef, const int32_t* offsets, const int32_t* perm,
                         int B, int N, int M, int D, int K, float* gfeat, pcnbr_stream_t stream);

/* ---- K9 edge features -------------------------------------------------- dgcnn.py:41-55
 * xt (B,N,F) point-major, idx (B,N,K) -> out (B,N,K,2F) point-major (the host returns it as the
 * (B,2F,N,K) view): out[..., f] = xt[idx] - xt[n]; out[..., F+f] = xt[n]. */
PCNBR_API int pcnbr_edge_feature_f32(const float* xt, const int32_t* idx, int B, int N, int F, int K, float* out,
                           pcnbr_stream_t stream);
/* gxt (B,N,F) = sum_incoming g[e, :F] - sum_k g[n,k,:F] + sum_k g[n,k,F:]. */
PCNBR_API int pcnbr_edge_feature_bwd_f32(const float* g, const int32_t* offsets, const int32_t* perm, int B, int N,
                               int F, int K, float* gxt, pcnbr_stream_t stream);

/* ---- fused EdgeConv: conv1x1 + BatchNorm + LeakyReLU + max over k ---- dgcnn.py:73-76
 * PQ (B,N,2O) = x_t @ [A ; B-A]^T (library GEMM by the host), idx (B,N,K).  selmax[o] = (gamma_o >= 0),
 * shift[o] = any constant near the pre-activations.  Outputs, all (B,N,O): psel = max_j (or min_j) P[idx[n,j]],
 * arg = its j (uint8), s1 = sum_j P[idx[n,j]]; partial (B * pcnbr_edgeconv_fwd_blocks(N), 2O): per-block
 * sum(u - shift) and sum((u - shift)^2) over all N*K pre-activations u = P_j + Q_i (BatchNorm statistics).
 * O in {32,64,128,256}, K <= 255. */
PCNBR_API int pcnbr_edgeconv_fwd_blocks(int N);
PCNBR_API int pcnbr_edgeconv_fwd_f32(const float* PQ, const int32_t* idx, const uint8_t* selmax, const float* shift,
                           int B, int N, int K, int O, float* psel, uint8_t* arg, float* s1, float* partial,
                           pcnbr_stream_t stream);
/* Exact backward of BatchNorm(train)+LeakyReLU+max in (B,N,O) terms.  gs = dL/dy on the selected edge;
 * coef (4,O) = {gamma*rstd, c1, c2r, mean} (c1 = c2r = 0 in eval mode); CSR of idx viewed as (B, N*K).
 * dPQ (B,N,2O) = [dL/dP | dL/dQ]; deterministic, no atomics. */
PCNBR_API int pcnbr_edgeconv_bwd_f32(const float* gs, const uint8_t* arg, const float* PQ, const float* s1,
                           const int32_t* offsets, const int32_t* perm, const float* coef, int B, int N, int K,
                           int O, float* dPQ, pcnbr_stream_t stream);

/* ---- fused BatchNorm + (Leaky)ReLU over point-major rows ---- common.py:125-178, dgcnn.py:66-71,95-126
 * Replaces the library triple BatchNorm -> activation (and its three backward ops) behind MiniPointNet / UnitPointNet /
 * the DGCNN head on a (R, C) row matrix (C % 4 == 0, C/4 a power of two <= 512: pcnbr_bn_supported).
 *   pcnbr_bn_stats_f32      : partial (pcnbr_bn_blocks(R,C), 2, C) = per-block sum(x - x[0,:]) and sum((x - x[0,:])^2)
 *   pcnbr_bn_finalize_f32   : fp64 combine in block order -> stats (4,C) = {mean, rstd, gamma*rstd, beta}; updates
 *                             running_mean / running_var (momentum, unbiased variance) when they are non-NULL.
 *                             nblk == 0 = eval mode: mean / var are the running statistics.  `count` = values per channel.
 *   pcnbr_bn_act_fwd_f32    : y (R,C) = act((x - mean) * gamma*rstd + beta), x[r,c] = a[r*lda+c] (+ b[r*ldb+c] if b != NULL),
 *                             act(t) = t > 0 ? t : slope*t   (slope 0 = ReLU, 0.2 = the DGCNN LeakyReLU)
 *   pcnbr_bn_act_bwd_reduce_f32 : g' = gy * act'(pre);  partial = per-block {sum g', sum g' * xhat};  gs (R,C) <- g' if non-NULL
 *   pcnbr_bn_bwd_finalize_f32   : dgamma, dbeta (C) and coef (4,C) = {gamma*rstd, c1, c2r, mean} (c1 = c2r = 0 unless training)
 *   pcnbr_bn_act_bwd_apply_f32  : dx (R,C) = gamma*rstd * g' - c1 - c2r * (x - mean)
 * drop_seed != NULL (device pointer to one 64-bit word) fuses the nn.Dropout(drop_p) that follows the activation
 * (dgcnn.py:117,122): y is multiplied by the keep mask / (1 - p) derived from a counter-based hash of (seed, element);
 * the backward kernels take the same seed and recompute the mask -- nothing is stored.  NULL = no dropout.
 * Deterministic (fixed-order reductions, no atomics). */
PCNBR_API int pcnbr_bn_supported(long R, int C);
PCNBR_API int pcnbr_bn_blocks(long R, int C);
PCNBR_API int pcnbr_bn_stats_f32(const float* x, long R, int C, float* partial, pcnbr_stream_t stream);
PCNBR_API int pcnbr_bn_finalize_f32(const float* partial, int nblk, const float* shift, double count, int C,
                          const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                          float* running_var, float* stats, pcnbr_stream_t stream);
PCNBR_API int pcnbr_bn_act_fwd_f32(const float* a, long lda, const float* b, long ldb, long R, int C, const float* stats,
                         float slope, float* y, const unsigned long long* drop_seed, float drop_p, float* amax_out,
                         pcnbr_stream_t stream);
PCNBR_API int pcnbr_bn_act_bwd_reduce_f32(const float* gy, const float* a, long lda, const float* b, long ldb, long R, int C,
                                const float* stats, float slope, float* partial, float* gs,
                                const unsigned long long* drop_seed, float drop_p, pcnbr_stream_t stream);
PCNBR_API int pcnbr_bn_bwd_finalize_f32(const float* partial, int nblk, const float* stats, double count, int C, int training,
                              float* dgamma, float* dbeta, float* coef, pcnbr_stream_t stream);
PCNBR_API int pcnbr_bn_act_bwd_apply_f32(const float* gy, const float* x, long R, int C, const float* stats, const float* coef,
                               float slope, float* dx, const unsigned long long* drop_seed, float drop_p, float* amax_out,
                               pcnbr_stream_t stream);

/* ---- BatchNorm + (Leaky)ReLU + max over the K rows of a group, fused ---- common.py:141-147 + 85-86 (SetAbstraction / InvResMLP)
 * h (G*K, C) pre-BatchNorm rows, stats as above.  act(bn(.)) is monotone per channel, so the max over K is taken on h (max
 * for gamma*rstd >= 0, min otherwise) and only the (G,C) result is activated: out (G,C), psel (G,C) the selected h,
 * arg (G,C) its k (uint8, K <= 255).  The activated (G*K, C) tensor is never written.
 * Backward: run pcnbr_bn_act_bwd_reduce_f32 / pcnbr_bn_bwd_finalize_f32 on (gpool, psel) with count = G*K to get
 * gs (G,C) and coef, then pcnbr_pool_bn_bwd_apply_f32: dh[g,k,c] = gr gs[g,c] [k == arg[g,c]] - c1 - c2r (h[g,k,c] - mean). */
PCNBR_API int pcnbr_pool_bn_act_fwd_f32(const float* h, long G, int K, int C, const float* stats, float slope, float* out,
                              float* psel, uint8_t* arg, pcnbr_stream_t stream);
PCNBR_API int pcnbr_pool_bn_bwd_apply_f32(const float* h, const float* gs, const uint8_t* arg, long G, int K, int C,
                                const float* coef, float* dh, pcnbr_stream_t stream);

/* ---- fp32-accurate tensor-core GEMM for the 1x1 convolutions ---- common.py:125-178, dgcnn.py:66-126
 * Not a reference entry point: the reference's Conv1d/Conv2d(kernel 1) are library GEMMs; under the fp32 parity bar
 * they run as SIMT SGEMMs.  Here C (M,N) = A (M,K) . B (N,K)^T (+ bias (N)) runs as 3xTF32 on tcgen05 (hi.hi' + lo.hi' +
 * hi.lo', fp32 accumulation in TMEM; error ~2^-20 |a||b|) directly from the fp32 operands: hi is the operand word
 * itself (the tensor core ignores the low 13 mantissa bits), lo is produced in shared memory inside the kernel.
 *   a_mn == 0: A is row-major (M,K) with row pitch lda;  a_mn != 0: A is stored transposed, row-major (K,M), pitch lda.
 *   b_mn == 0: B is row-major (N,K) with row pitch ldb;  b_mn != 0: B is stored transposed, row-major (K,N), pitch ldb.
 *   (so y = x W^T, dx = gy W and dW = gy^T x all read x, W and gy as they lie in memory).
 *   Bases 16-byte aligned, lda % 4 == 0, ldb % 4 == 0; C is (M,N) contiguous.
 *   splits > 1 (from pcnbr_gemm3x_splits; weight gradients, where M and N are small and K is the number of points)
 *   cuts K over the CTAs; the partial tiles go to ws (pcnbr_gemm3x_ws_bytes) and are summed in a fixed order:
 *   deterministic, no atomics; bias must be NULL then. */
PCNBR_API int pcnbr_gemm3x_splits(int M, int N, int K);
PCNBR_API size_t pcnbr_gemm3x_ws_bytes(int M, int N, int K, int splits);
PCNBR_API int pcnbr_gemm3x_f32(const float* A, long lda, int a_mn, const float* B, long ldb, int b_mn, int M, int N, int K,
                     const float* bias, float* C, int splits, void* ws, size_t ws_bytes, pcnbr_stream_t stream);

/* ---- K10b the same GEMM as a TWO-TERM FP16 SPLIT on kind::f16 (twice the TF32 instruction rate), for the shapes whose
 * tensor-pipe time bound exceeds their HBM time bound (the wide DGCNN layers, models/dgcnn/dgcnn.py:95-126).
 * Each operand is scaled by a power of two that brings its largest element to [2^14, 2^15), split into hi + lo fp16 terms
 * inside the kernel (fp32 tiles arrive by TMA exactly as for pcnbr_gemm3x_ex_f32) and C = (hi.hi' + lo.hi' + hi.lo') / (s s').
 * pcnbr_split_f16 pre-splits a weight matrix once ([hi | lo] fp16 planes, optionally transposed) so that the B operand of a
 * forward / input-gradient GEMM needs no in-kernel conversion (b_split / b_split_ld / b_split_plane; NULL = convert B in
 * the kernel like A).
 * pcnbr_absmax_f32 writes pcnbr_amax_slots() per-block maxima of |x| for a (rows x cols) matrix (row pitch ld; cols, ld
 * multiples of 4); pcnbr_gemm2h_ex_f32 takes those arrays for A, A2 (when given) and B.  Otherwise the contract --
 * operand layouts, split-K (pcnbr_gemm3x_splits / pcnbr_gemm3x_ws_bytes), bias, C pitch -- is that of pcnbr_gemm3x_ex_f32. */
/* amax_out of pcnbr_bn_act_fwd_f32 / pcnbr_bn_act_bwd_apply_f32 (optional, pcnbr_amax_slots() floats): the same per-block
 * maxima for the tensor those kernels WRITE, so the GEMM that consumes it needs no separate pcnbr_absmax_f32 pass. */
PCNBR_API int pcnbr_amax_slots(void);
PCNBR_API int pcnbr_absmax_f32(const float* x, long rows, long cols, long ld, float* partial, pcnbr_stream_t stream);
PCNBR_API int pcnbr_gemm2h_preferred(int M, int N, int K);
PCNBR_API int pcnbr_split_f16(const float* src, int rows, int cols, long ld, int transpose, const float* amax, void* out,
                    long ldo, long plane, pcnbr_stream_t stream);
PCNBR_API int pcnbr_gemm2h_ex_f32(const float* A, long lda, int a_mn, const float* A2, long lda2, int K1, const float* B, long ldb,
                    int b_mn, int M, int N, int K, const float* bias, float* C, long ldc, int splits, void* ws,
                    size_t ws_bytes, const float* amax_a, const float* amax_a2, const float* amax_b,
                    const void* b_split, long b_split_ld, long b_split_plane, pcnbr_stream_t stream);
/* _ex2: the three GEMMs of a layer share their operand splits.  a_planes_out (/ a2_planes_out for A2): with a pre-split B and a
 * K-major A (forward: A = x; input gradient: A = gy) the kernel also WRITES the [hi | lo] fp16 planes of A it computes in
 * shared memory anyway (2 planes of M rows, row pitch *_ld halfs -- a multiple of 8, >= the operand's K --, planes *_plane halfs
 * apart; bit-identical to pcnbr_split_f16(A, M, K, lda, 0, amax_a, ...)).  a_mnsplit + b_mnsplit (a_mn = b_mn = 1, the weight
 * gradient dW = gy^T x): BOTH operands are read as such planes (rows = K), MN-major straight into tcgen05 -- no in-kernel
 * conversion; amax_a / amax_b must be the arrays the planes were written with; A / B may then be NULL (when given, shapes
 * whose tile is narrower than 64 columns fall back to converting them).  All NULL: exactly pcnbr_gemm2h_ex_f32. */
PCNBR_API int pcnbr_gemm2h_ex2_f32(const float* A, long lda, int a_mn, const float* A2, long lda2, int K1, const float* B, long ldb,
                    int b_mn, int M, int N, int K, const float* bias, float* C, long ldc, int splits, void* ws,
                    size_t ws_bytes, const float* amax_a, const float* amax_a2, const float* amax_b,
                    const void* b_split, long b_split_ld, long b_split_plane,
                    void* a_planes_out, long apo_ld, long apo_plane, void* a2_planes_out, long ap2o_ld, long ap2o_plane,
                    const void* a_mnsplit, long ams_ld, long ams_plane, const void* b_mnsplit, long bms_ld, long bms_plane,
                    pcnbr_stream_t stream);
/* Diagnostic (tools/gemm_shapes.py --trace): buf = device array of 148 x 16 uint64, or NULL to switch off.  While set, every
 * gemm2h CTA adds the clock cycles each of its warp roles spent waiting on its barriers (slots: 0 producer on a free ring
 * slot, 1 converters on the TMA, 2 MMA on a free accumulator, 3 MMA on a converted stage, 4 epilogue on a full accumulator,
 * 5 epilogue on a free staging slab, 6 CTA lifetime, 7 converter busy time, 8 CTA lifetime in ns of %globaltimer);
 * gemm3x_kernel fills the same slots (slot 1 there: the MMA warp waiting for the TMA).  Synchronises the device; not under capture. */
PCNBR_API int pcnbr_gemm2h_trace(unsigned long long* buf);
/* Extended form.  A2 != NULL: A is the K-concatenation [A (M,K1) | A2 (M,K-K1)] of two row-major matrices (pitches lda,
 * lda2; K1 % 32 == 0, a_mn must be 0) -- a torch.cat along the channels in front of a convolution (dgcnn.py:147,
 * cat((x1..x4, x5)) -> conv6) that is never materialised.  ldc: row pitch of C (>= N, % 4 == 0), so a GEMM can write a
 * column block of a wider matrix (the two halves of that layer's weight gradient).  A2 == NULL, ldc == N: as above. */
PCNBR_API int pcnbr_gemm3x_ex_f32(const float* A, long lda, int a_mn, const float* A2, long lda2, int K1, const float* B, long ldb,
                        int b_mn, int M, int N, int K, const float* bias, float* C, long ldc, int splits, void* ws,
                        size_t ws_bytes, pcnbr_stream_t stream);

/* ---- evaluation metrics ---------------------------------- Training/metrics.py:3-146
 * pred (B,N,C) scores (softmax or logits: only the argmax matters), onehot (B,N,C) uint8 labels, lengths (B) int64
 * unpadded points per cloud (NULL = N).  matrix (C,C) int64 is ACCUMULATED: matrix[label, predicted] += count over the
 * unpadded points (zero it for a per-batch matrix, keep it across batches for a validation set).  Accuracy and the
 * per-class intersections / unions of metrics.py are functions of it.  C <= 64. */
PCNBR_API int pcnbr_confusion_f32(const float* pred, const uint8_t* onehot, const long long* lengths, int B, int N, int C,
                        long long* matrix, pcnbr_stream_t stream);
/* Extended form: unlabeled (C) int64, ACCUMULATED like matrix: unlabeled[p] += rows whose label row is all zero and whose
 * prediction is p.  The confusion matrix and the accuracy take such a row as class 0 (labels.argmax(-1), metrics.py:20,72),
 * the IoU functions test labels[..., c] == 1 (metrics.py:103,137): it belongs to no class there. */
PCNBR_API int pcnbr_confusion_ex_f32(const float* pred, const uint8_t* onehot, const long long* lengths, int B, int N, int C,
                           long long* matrix, long long* unlabeled, pcnbr_stream_t stream);

/* ---- training loss ---------------------------------------- Training/train_model.py:15-57
 * Masked one-hot cross entropy: loss[0] = mean over the unpadded points (n < lengths[b]) of -sum_c onehot * log_softmax
 * (0 when every point is padding); dlogits (B,L,C), if non-NULL, receives d loss / d logits in the same pass.
 * partial: B * pcnbr_masked_ce_blocks(L) floats of scratch.  No host synchronisation; C <= 64. */
PCNBR_API int pcnbr_masked_ce_blocks(int L);
PCNBR_API int pcnbr_masked_ce_f32(const float* logits, const uint8_t* onehot, const long long* lengths, int B, int L, int C,
                        float* loss, float* dlogits, float* partial, pcnbr_stream_t stream);

/* ---- block dataloader in HBM ------------------------------ data_processing/block_datasets.py:5-29,117-128
 * Replaces, per training step, BlockS3DISDataset.__getitem__ (torch.load of one file per block + row selection) and
 * collate_blocks (zero-padded batch).  points (T,9) f32 and labels (T,L) uint8 hold every block of the split packed back
 * to back; block_start (nblocks+1) int64 row offsets; block_ids (B) int32 = the blocks of this batch.
 *   sel (B,S) int32 != NULL: row sel[b,r] of block block_ids[b] goes to batch row r (the reference's
 *       randperm(n)[:S] / randint(n,(S,)) draw, :119-125); out_len[b] = S.
 *   sel == NULL: rows 0..n_b-1 in order, zero padded up to S (collate_blocks; pass S = the longest block of the batch);
 *       out_len[b] = min(n_b, S).
 * out_points (B,S,9) f32, out_labels (B,S,L) uint8, out_len (B) int64 -- exactly the tuple the reference's DataLoader
 * yields, already on the device.  Out-of-range rows read as zeros. */
PCNBR_API int pcnbr_block_batch(const float* points, const uint8_t* labels, const long long* block_start, const int* block_ids,
                      const int* sel, int B, int S, int L, float* out_points, uint8_t* out_labels, long long* out_len,
                      pcnbr_stream_t stream);

/* ---- sliding-window scene inference, merge step ----------- models/dgcnn/utils.py:67-131
 * predict_single_scene cuts a scene of n_points into W = ceil(n_points / step) windows [w*step, min(n, w*step+window)),
 * runs the model per window, overlap-adds the logits, divides by the cover count and takes argmax / max softmax.
 * logits ((sum of window lengths), C): the windows' logits back to back, window w at row win_off[w] (win_off: W int64).
 * Outputs per scene point: mean_logits (n,C) (may be NULL; bit-identical to all_logits / point_counts: same summation
 * order), pred (n) int64 (first maximum), conf (n) f32.  step <= window, C <= 64. */
PCNBR_API int pcnbr_window_merge_f32(const float* logits, const long long* win_off, int W, long long n_points, int window,
                           int step, int C, float* mean_logits, long long* pred, float* conf, pcnbr_stream_t stream);

/* ---- measurement hook (bench.py roofline, kernel sweep) -- not part of the reference interface
 * pcnbr_prof_enable(1): bracket every kernel this library launches with CUDA events on its launch stream and
 * remember the launch's ALGORITHMIC bytes / flops.  Must be off while a CUDA graph is captured.
 * pcnbr_prof_collect: synchronise, write one line "kernel\tms\tbytes\tflops\n" per launch into out (cap
 * bytes, NUL-terminated), clear the log and return the number of launches recorded. */
PCNBR_API void pcnbr_prof_enable(int on);
PCNBR_API int pcnbr_prof_collect(char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PCNBR_H_ */
