/*
 * oracle/canon.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C restatement of the point-cloud neighbourhood hot path of
 * piotr-bledowski/3D-Semantic-Segmentation-Benchmark, with the CANONICAL tie rule
 * ("lowest index wins") in place of torch.topk's arbitrary tie order.
 *
 * Every function cites the reference lines whose arithmetic it follows.  The
 * arithmetic forms were pinned by probing the reference's torch CPU ops
 * (DESIGN.md "Oracle"):
 *   - sample():       dist = sqrtf(fmaf(dz,dz,fmaf(dy,dy,dx*dx)))      models/utils/common.py:28
 *   - group()/interp: d2   = (dx*dx + dy*dy) + dz*dz  (no contraction)   common.py:56,112
 *   - group() radius: keep iff d2 <= (float)((double)r*(double)r)        common.py:58
 *   - knn():          inner = -2 * FMA-chain over f ascending (MKL sgemm) models/dgcnn/dgcnn.py:16
 *                     xx    = 16-row cascade sum of x*x (ATen sum dim=1)  dgcnn.py:17
 *                     pd    = ((-xx_j) - inner_ij) - xx_i                 dgcnn.py:18
 * Compile with -ffp-contract=off so that the compiler adds no FMAs of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- helpers */

typedef struct { float key; int32_t idx; } orc_cand;

/* (key, idx) lexicographic "a before b" */
static inline int cand_before(float ka, int32_t ia, float kb, int32_t ib) {
    return (ka < kb) || (ka == kb && ia < ib);
}

/* Keep the K smallest (key, idx) pairs, sorted ascending, by insertion. */
static inline void topk_insert(orc_cand* best, int* cnt, int K, float key, int32_t idx) {
    int n = *cnt;
    if (n == K && !cand_before(key, idx, best[K - 1].key, best[K - 1].idx)) return;
    int pos = (n < K) ? n : K - 1;
    while (pos > 0 && cand_before(key, idx, best[pos - 1].key, best[pos - 1].idx)) {
        best[pos] = best[pos - 1];
        --pos;
    }
    best[pos].key = key;
    best[pos].idx = idx;
    if (n < K) *cnt = n + 1;
}

/* ------------------------------------------------------------------- FPS */
/* common.py:17-34.  start[b] replaces the torch.randint draw of common.py:22.
 * idx_out (B,C) int32; xyz_out (B,C,3) may be NULL. */
ORC_API void orc_fps(const float* xyz, int B, int N, int C, const int32_t* start,
                     int32_t* idx_out, float* xyz_out) {
    /* clouds are independent (the batch index is an outer loop of common.py:25-31): one host thread per cloud */
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float* dist = (float*)malloc(sizeof(float) * (size_t)N);
        const float* p = xyz + (size_t)b * N * 3;
        for (int n = 0; n < N; ++n) dist[n] = INFINITY;           /* common.py:21 */
        int32_t far = start[b];
        for (int i = 0; i < C; ++i) {
            idx_out[(size_t)b * C + i] = far;                      /* common.py:26 */
            const float cx = p[far * 3 + 0], cy = p[far * 3 + 1], cz = p[far * 3 + 2];
            if (xyz_out) {
                float* o = xyz_out + ((size_t)b * C + i) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;                   /* common.py:34 */
            }
            float best = -1.0f; int32_t besti = 0;
            for (int n = 0; n < N; ++n) {
                const float dx = p[n * 3 + 0] - cx;
                const float dy = p[n * 3 + 1] - cy;
                const float dz = p[n * 3 + 2] - cz;
                /* linalg.vector_norm over 3 elements: FMA chain then sqrt  (common.py:28) */
                const float d = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                if (d < dist[n]) dist[n] = d;                      /* common.py:29-30 */
                if (dist[n] > best) { best = dist[n]; besti = n; } /* torch.max: lowest index on ties, common.py:31 */
            }
            far = besti;
        }
        free(dist);
    }
}

/* ------------------------------------------------------------ ball query */
/* common.py:54-61 with canonical ties: in-ball points ascending (d2, idx), then the
 * out-of-ball points (d2 := inf) in ascending index.  idx (B,M,K) int32. */
ORC_API void orc_ball_query(const float* q, const float* p, int B, int M, int N,
                            float r2, int K, int32_t* idx) {
    /* rows of the (B,C,N) distance tensor are independent: host threads over (b, m), a private list each */
    #pragma omp parallel
    {
    orc_cand* best = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)K);
    #pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int m = 0; m < M; ++m) {
            const float* c = q + ((size_t)b * M + m) * 3;
            int cnt = 0;
            for (int n = 0; n < N; ++n) {
                const float* s = p + ((size_t)b * N + n) * 3;
                const float dx = s[0] - c[0], dy = s[1] - c[1], dz = s[2] - c[2];
                float d2 = (dx * dx + dy * dy) + dz * dz;          /* common.py:56 */
                if (!(d2 <= r2)) d2 = INFINITY;                    /* common.py:58-59 */
                topk_insert(best, &cnt, K, d2, n);
            }
            int32_t* o = idx + ((size_t)b * M + m) * K;
            for (int k = 0; k < K; ++k) o[k] = best[k].idx;
        }
    free(best);
    }
}

/* ------------------------------------------------- kNN, direct distances */
/* common.py:110-114 (interpolate's top-k): d2 = ((src - query)^2).sum(-1), k smallest,
 * ascending (d2, idx).  idx (B,M,k) int32, d2out (B,M,k) may be NULL. */
ORC_API void orc_knn_direct(const float* q, const float* p, int B, int M, int N, int K,
                            int32_t* idx, float* d2out) {
    #pragma omp parallel
    {
    orc_cand* best = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)K);
    #pragma omp for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int m = 0; m < M; ++m) {
            const float* c = q + ((size_t)b * M + m) * 3;
            int cnt = 0;
            for (int n = 0; n < N; ++n) {
                const float* s = p + ((size_t)b * N + n) * 3;
                const float dx = s[0] - c[0], dy = s[1] - c[1], dz = s[2] - c[2];
                const float d2 = (dx * dx + dy * dy) + dz * dz;    /* common.py:112 */
                topk_insert(best, &cnt, K, d2, n);
            }
            for (int k = 0; k < K; ++k) {
                idx[((size_t)b * M + m) * K + k] = best[k].idx;
                if (d2out) d2out[((size_t)b * M + m) * K + k] = best[k].key;
            }
        }
    free(best);
    }
}

/* --------------------------------------- kNN, expanded form (DGCNN knn) */
/* ATen's sum over a non-innermost dim (dgcnn.py:17, torch CPU build, 8-float vectors):
 * cascade_sumsq(): rows are added into acc0 in runs of 16, each full run is folded into acc1,
 * 16 runs of acc1 into acc2, ...; the tail rows stay in acc0; result = ((acc0+acc1)+acc2)+acc3.
 * Columns n < (N & ~31) are summed that way over all F rows (4x8-float vector path).  The last N % 32
 * columns go through ATen's scalar path instead: four interleaved partial sums (rows i = 4t+k),
 * each cascaded over t, the F % 4 left-over rows added to partial 0, then ((p0+p1)+p2)+p3. */
static float cascade_sumsq(const float* x, int count, size_t stride) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    while (i + 16 <= count) {
        for (int j = 0; j < 16; ++j, ++i) { const float v = x[i * stride]; acc[0] = acc[0] + v * v; }
        for (int lvl = 1; lvl < 4; ++lvl) {
            acc[lvl] = acc[lvl] + acc[lvl - 1];
            acc[lvl - 1] = 0.f;
            if ((i & (15 << (lvl * 4))) != 0) break;
        }
    }
    for (; i < count; ++i) { const float v = x[i * stride]; acc[0] = acc[0] + v * v; }
    float r = acc[0];
    for (int lvl = 1; lvl < 4; ++lvl) r = r + acc[lvl];
    return r;
}

static float column_sumsq(const float* x, int F, int N, int n) {
    const float* col = x + n;
    if (n < (N & ~31)) return cascade_sumsq(col, F, (size_t)N);
    const int q = F / 4;
    float p[4];
    for (int k = 0; k < 4; ++k) p[k] = cascade_sumsq(col + (size_t)k * N, q, (size_t)N * 4);
    for (int i = q * 4; i < F; ++i) { const float v = col[(size_t)i * N]; p[0] = p[0] + v * v; }
    return ((p[0] + p[1]) + p[2]) + p[3];
}

ORC_API void orc_sumsq(const float* x, int B, int F, int N, float* xx) {
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n)
            xx[(size_t)b * N + n] = column_sumsq(x + (size_t)b * F * N, F, N, n);
}

/* dgcnn.py:16-20.  x (B,F,N) channel-first; idx (B,N,k) int32, descending pd, ties by
 * lowest index; pdout (B,N,k) may be NULL. */
ORC_API void orc_knn_expand(const float* x, int B, int F, int N, int K, int32_t* idx, float* pdout) {
    float* xx = (float*)malloc(sizeof(float) * (size_t)N);
    float* xt = (float*)malloc(sizeof(float) * (size_t)N * F);   /* point-major copy */
    for (int b = 0; b < B; ++b) {
        const float* xb = x + (size_t)b * F * N;
        for (int n = 0; n < N; ++n) {
            xx[n] = column_sumsq(xb, F, N, n);
            for (int f = 0; f < F; ++f) xt[(size_t)n * F + f] = xb[(size_t)f * N + n];
        }
        /* every row of the (N,N) matrix is computed independently: host threads over the rows */
        #pragma omp parallel
        {
        orc_cand* best = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)K);
        #pragma omp for schedule(static)
        for (int i = 0; i < N; ++i) {
            int cnt = 0;
            const float* xi = xt + (size_t)i * F;
            for (int j = 0; j < N; ++j) {
                const float* xj = xt + (size_t)j * F;
                float c = xi[0] * xj[0];                           /* sgemm: FMA chain over f */
                for (int f = 1; f < F; ++f) c = fmaf(xi[f], xj[f], c);
                const float inner = -2.0f * c;                     /* dgcnn.py:16 */
                const float pd = ((-xx[j]) - inner) - xx[i];       /* dgcnn.py:18 (xx is (B,1,N): column term first) */
                topk_insert(best, &cnt, K, -pd, j);                /* largest pd first, dgcnn.py:20 */
            }
            for (int k = 0; k < K; ++k) {
                idx[((size_t)b * N + i) * K + k] = best[k].idx;
                if (pdout) pdout[((size_t)b * N + i) * K + k] = -best[k].key;
            }
        }
        free(best);
        }
    }
    free(xx); free(xt);
}

/* dgcnn.py:16-20 for a SUBSET of the query rows of one cloud (large-N spot checks: the full (N,N) scan of a
 * 100 k-point cloud is 1e10 pairs).  x (F,N) channel-first; rows (R) query indices; idx (R,k) int32. */
ORC_API void orc_knn_expand_rows(const float* x, int F, int N, int K, const int32_t* rows, int R, int32_t* idx) {
    float* xx = (float*)malloc(sizeof(float) * (size_t)N);
    float* xt = (float*)malloc(sizeof(float) * (size_t)N * F);
    for (int n = 0; n < N; ++n) {
        xx[n] = column_sumsq(x, F, N, n);
        for (int f = 0; f < F; ++f) xt[(size_t)n * F + f] = x[(size_t)f * N + n];
    }
    #pragma omp parallel
    {
    orc_cand* best = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)K);
    #pragma omp for schedule(static)
    for (int r = 0; r < R; ++r) {
        const int i = rows[r];
        int cnt = 0;
        const float* xi = xt + (size_t)i * F;
        for (int j = 0; j < N; ++j) {
            const float* xj = xt + (size_t)j * F;
            float c = xi[0] * xj[0];
            for (int f = 1; f < F; ++f) c = fmaf(xi[f], xj[f], c);
            const float inner = -2.0f * c;
            const float pd = ((-xx[j]) - inner) - xx[i];
            topk_insert(best, &cnt, K, -pd, j);
        }
        for (int k = 0; k < K; ++k) idx[(size_t)r * K + k] = best[k].idx;
    }
    free(best);
    }
    free(xx); free(xt);
}

/* -------------------------------------------------- group gather+concat */
/* common.py:62-71.  out (B,M,K,3+D).  rdiv<=0: no normalisation; else the local
 * coordinates are DIVIDED by (float)r (common.py:69). */
ORC_API void orc_group(const float* p, const float* feat, const float* q, const int32_t* idx,
                       int B, int N, int M, int K, int D, float rdiv, float* out) {
    const int W = 3 + D;
    for (int b = 0; b < B; ++b)
        for (int m = 0; m < M; ++m) {
            const float* c = q + ((size_t)b * M + m) * 3;
            for (int k = 0; k < K; ++k) {
                const int32_t s = idx[((size_t)b * M + m) * K + k];
                const float* sp = p + ((size_t)b * N + s) * 3;
                float* o = out + (((size_t)b * M + m) * K + k) * W;
                for (int a = 0; a < 3; ++a) {
                    float v = sp[a] - c[a];
                    if (rdiv > 0.f) v = v / rdiv;
                    o[a] = v;
                }
                if (D) memcpy(o + 3, feat + ((size_t)b * N + s) * D, sizeof(float) * (size_t)D);
            }
        }
}

/* ---------------------------------------------------- three interpolate */
/* common.py:115-122 given the 3-NN (idx, d2) of orc_knn_direct.  feat (B,M,D) are the
 * coarse features, out (B,N,D).  w = 1/(d2 + 1e-9f); norm = (w0+w1)+w2;
 * out = ((f0*w0/norm + f1*w1/norm) + f2*w2/norm). */
ORC_API void orc_interp(const float* feat, const int32_t* idx, const float* d2, int B, int N,
                        int M, int D, int K, float* out) {
    float w[64];
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            const int32_t* id = idx + ((size_t)b * N + n) * K;
            const float* dd = d2 + ((size_t)b * N + n) * K;
            float norm = 0.f;
            for (int k = 0; k < K; ++k) { w[k] = 1.0f / (dd[k] + 1e-9f); norm = (k == 0) ? w[0] : norm + w[k]; }
            for (int d = 0; d < D; ++d) {
                float acc = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float t = feat[((size_t)b * M + id[k]) * D + d] * w[k] / norm;
                    acc = (k == 0) ? t : acc + t;
                }
                out[((size_t)b * N + n) * D + d] = acc;
            }
        }
}

/* --------------------------------------------------------- edge feature */
/* dgcnn.py:41-55 (dim9=False).  x (B,F,N), idx (B,N,k) -> out (B,2F,N,k):
 * out[b,f,n,j] = x[b,f,idx[n,j]] - x[b,f,n];  out[b,F+f,n,j] = x[b,f,n]. */
ORC_API void orc_edge_feature(const float* x, const int32_t* idx, int B, int F, int N, int K, float* out) {
    for (int b = 0; b < B; ++b)
        for (int f = 0; f < F; ++f) {
            const float* xr = x + ((size_t)b * F + f) * N;
            float* o1 = out + ((size_t)b * 2 * F + f) * N * K;
            float* o2 = out + ((size_t)b * 2 * F + F + f) * N * K;
            for (int n = 0; n < N; ++n)
                for (int j = 0; j < K; ++j) {
                    const int32_t s = idx[((size_t)b * N + n) * K + j];
                    o1[(size_t)n * K + j] = xr[s] - xr[n];
                    o2[(size_t)n * K + j] = xr[n];
                }
        }
}
