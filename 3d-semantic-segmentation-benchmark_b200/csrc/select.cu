// select.cu -- K2 ball query, K3 xyz kNN (direct distances) and the generic expanded-form kNN.
//
// Reference: models/utils/common.py:54-61 (ball query inside group()), :110-114 (3-NN of
// interpolate()), models/dgcnn/dgcnn.py:7-21 (knn()).  The reference materialises the full (B,M,N)
// distance tensor and runs torch.topk over it; here one warp owns one query, source points stream
// through a shared-memory tile shared by all warps of the CTA, and the warp keeps its K best
// (key,index) pairs as a sorted list spread over its lanes (WarpList).  A candidate is compared
// against the current K-th key first (one compare per point); only the rare survivors are inserted.
// Nothing of size M*N is ever written.
//
// Ordering is the unsigned order of (ord(key) << 32 | index): ascending key, lowest index on ties.
// Ball query needs no special padding code: out-of-ball points simply get key = +inf, so the K-list
// fills with in-ball points by (d2,index) followed by the lowest-index out-of-ball points -- exactly
// what a stable sort of the reference's masked distance row yields.
#include "common.cuh"
#include <cstdlib>

namespace pcnbr {

constexpr int SEL_TILE = 1024;   // source points per shared-memory tile (12 KB SoA)

constexpr int SEL_WARPS = 16;    // queries per CTA: the tile staging is shared by 16 warps

template <int NSLOT, bool RADIUS>
__global__ void __launch_bounds__(SEL_WARPS * 32)
select_xyz_kernel(const float* __restrict__ q, const float* __restrict__ p, int M, int N_alloc, float r2, int K,
                  const int32_t* __restrict__ n_qry, const int32_t* __restrict__ n_src,
                  int32_t* __restrict__ idx, float* __restrict__ d2out) {
    __shared__ float4 spt[SEL_TILE];                      // (x, y, z, -) per source point: one LDS.128 per point
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    // length-aware form (zero-padded batches): only the first n_src[b] sources / n_qry[b] queries of cloud b exist
    const int N = len_valid(n_src, b, N_alloc), Mv = len_valid(n_qry, b, M);
    if (m >= Mv && m < M) len_fill_row(idx, d2out, ((size_t)b * M + m) * K, K, lane);
    const bool active = m < Mv;
    const float* __restrict__ pb = p + (size_t)b * N_alloc * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float* c = q + ((size_t)b * M + m) * 3;
        qx = c[0]; qy = c[1]; qz = c[2];
    }
    WarpList<NSLOT> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;
    bool full = false;                                       // the list holds K real entries (thr is a real key)
    float thr_f = 0.f;                                       // distance of the K-th entry once full

    // exact path for the 32 points [c0, c0+32) of the tile: 64-bit (key, index) order, insert the survivors
    auto exact32 = [&](int t0, int tn, int c0) {
        const int j = c0 + lane;
        u64 key = PCNBR_KEY_MAX;
        if (j < tn) {
            const float4 s = spt[j];
            float d2 = d2_direct(s.x, s.y, s.z, qx, qy, qz);
            if (RADIUS && !(d2 <= r2)) d2 = __int_as_float(0x7f800000);       // common.py:58-59
            key = pack_key(f2ord(d2), (uint32_t)(t0 + j));
        }
        uint32_t pass = __ballot_sync(PCNBR_FULL, key < thr);
        while (pass) {
            const int src = __ffs(pass) - 1;
            pass &= pass - 1;
            const u64 cand = shfl64(key, src);
            if (cand < thr) {
                list.insert(cand, lane);
                thr = list.at(K - 1);
            }
        }
    };

    for (int t0 = 0; t0 < N; t0 += SEL_TILE) {
        const int tn = min(SEL_TILE, N - t0);
        const int tpad = (tn + 127) & ~127;                  // the fast loop walks whole 128-point steps
        __syncthreads();
        for (int i = threadIdx.x; i < tpad; i += blockDim.x) {
            float4 v = make_float4(1e30f, 1e30f, 1e30f, 0.f);   // sentinel: its distance is +inf, never below any threshold
            if (i < tn) {
                const float* s = pb + (size_t)(t0 + i) * 3;
                v = make_float4(s[0], s[1], s[2], 0.f);
            }
            spt[i] = v;
        }
        __syncthreads();
        if (!active) continue;
        int c0 = 0;
        if (t0 == 0) {
            // the first 32 points of the cloud all enter the empty list: ONE warp sort instead of 32 sequential inserts
            // (the inserts were 27 % of the kernel's instructions at K = 32, ncu source page)
            u64 key = PCNBR_KEY_MAX;
            if (lane < tn) {
                const float4 s = spt[lane];
                float d2 = d2_direct(s.x, s.y, s.z, qx, qy, qz);
                if (RADIUS && !(d2 <= r2)) d2 = __int_as_float(0x7f800000);   // common.py:58-59
                key = pack_key(f2ord(d2), (uint32_t)lane);
            }
            list.v[0] = warp_sort64(key, lane);
            thr = list.at(K - 1);
            full = thr != PCNBR_KEY_MAX;
            c0 = 32;
        }
        // start-up: until the list holds K entries every point is a candidate (first tile, first ceil(K/32) groups)
        for (; !full && c0 < tn; c0 += 32) {
            exact32(t0, tn, c0);
            full = thr != PCNBR_KEY_MAX;
        }
        thr_f = ord2f((uint32_t)(thr >> 32));
        for (; c0 < tn && (c0 & 127); c0 += 32) {           // realign to a 128-point boundary
            exact32(t0, tn, c0);
            thr_f = ord2f((uint32_t)(thr >> 32));
        }
        // Hot loop: 128 points per step (4 per lane), 8 flops + ONE float compare per point.  A candidate's index is
        // larger than every index already in the list, so it enters iff its distance is STRICTLY smaller than the K-th
        // entry's: the 64-bit (key, index) order is only materialised for the rare steps with a survivor.
        for (; c0 < tn; c0 += 128) {
            float d[4];
            uint32_t hm = 0;                                 // bit u: this lane's point of sub-group u beats the K-th entry
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 s = spt[c0 + u * 32 + lane];
                float d2 = d2_direct(s.x, s.y, s.z, qx, qy, qz);
                if (RADIUS && !(d2 <= r2)) d2 = __int_as_float(0x7f800000);   // common.py:58-59
                d[u] = d2;
                hm |= (d2 < thr_f) ? (1u << u) : 0u;
            }
            const uint32_t any = __reduce_or_sync(PCNBR_FULL, hm);
            if (!any) continue;
            // survivors only (sentinels have d = +inf and never get here): sub-groups in ascending index order
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (!(any & (1u << u))) continue;            // warp-uniform
                const u64 key = (hm & (1u << u)) ? pack_key(f2ord(d[u]), (uint32_t)(t0 + c0 + u * 32 + lane)) : PCNBR_KEY_MAX;
                uint32_t pass = __ballot_sync(PCNBR_FULL, key < thr);
                while (pass) {
                    const int src = __ffs(pass) - 1;
                    pass &= pass - 1;
                    const u64 cand = shfl64(key, src);
                    if (cand < thr) {
                        list.insert(cand, lane);
                        thr = list.at(K - 1);
                    }
                }
            }
            thr_f = ord2f((uint32_t)(thr >> 32));
        }
    }
    if (!active) return;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int pos = s * 32 + lane;
        if (pos < K) {
            const size_t o = ((size_t)b * M + m) * K + pos;
            idx[o] = (int32_t)min((uint32_t)list.v[s], (uint32_t)(N - 1));      // (an empty slot -- K > valid sources -- stays in range)
            if (d2out) d2out[o] = ord2f((uint32_t)(list.v[s] >> 32));
        }
    }
}

// ---- expanded-form kNN, any F (dgcnn.py:16-20) ------------------------------------------------

// ATen's outer-dimension sum (torch.sum(x**2, dim=1), dgcnn.py:17), restated in oracle/canon.c:
// cascade: rows accumulate in runs of 16 (acc0), runs fold into acc1, 16 runs into acc2, ...; tail rows
// stay in acc0; result ((acc0+acc1)+acc2)+acc3.  Columns n < (N & ~31) cascade over all F rows; the
// last N % 32 columns take ATen's ILP-4 path: four interleaved cascades (rows 4t+k), the F % 4
// left-over rows added to partial 0, then ((p0+p1)+p2)+p3.
__device__ __forceinline__ float cascade_sumsq(const float* __restrict__ xp, int count, size_t stride) {
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    int i = 0;
    while (i + 16 <= count) {
        for (int j = 0; j < 16; ++j, ++i) {
            const float v = xp[(size_t)i * stride];
            acc0 = __fadd_rn(acc0, __fmul_rn(v, v));
        }
        acc1 = __fadd_rn(acc1, acc0); acc0 = 0.f;
        if ((i & 0xf0) == 0) {
            acc2 = __fadd_rn(acc2, acc1); acc1 = 0.f;
            if ((i & 0xf00) == 0) { acc3 = __fadd_rn(acc3, acc2); acc2 = 0.f; }
        }
    }
    for (; i < count; ++i) {
        const float v = xp[(size_t)i * stride];
        acc0 = __fadd_rn(acc0, __fmul_rn(v, v));
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(acc0, acc1), acc2), acc3);
}

__global__ void sumsq_cascade_kernel(const float* __restrict__ x, int F, int N, long sf, long sn,
                                     const int32_t* __restrict__ n_valid, float* __restrict__ xx) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* __restrict__ xp = x + (size_t)b * F * N + (size_t)n * sn;
    // which columns take ATen's vectorised cascade depends on the length of the cloud AS THE REFERENCE SEES IT: a cloud of a
    // zero-padded batch that is evaluated alone (length-aware form) has n_valid[b] columns
    const int NV = len_valid(n_valid, b, N);
    float r;
    if (n < (NV & ~31)) {
        if (sf == 1 && (F & 3) == 0 && F <= 64 && (((uintptr_t)xp) & 15) == 0) {
            // point-major rows: 16-byte loads (4x fewer L1 wavefronts than scalar loads of 256-byte-strided rows);
            // same cascade: runs of 16 into acc0, folded into acc1; no higher level below 256 rows
            float acc0 = 0.f, acc1 = 0.f;
            for (int f4 = 0; f4 < F / 4; ++f4) {
                const float4 v = reinterpret_cast<const float4*>(xp)[f4];
                acc0 = __fadd_rn(acc0, __fmul_rn(v.x, v.x));
                acc0 = __fadd_rn(acc0, __fmul_rn(v.y, v.y));
                acc0 = __fadd_rn(acc0, __fmul_rn(v.z, v.z));
                acc0 = __fadd_rn(acc0, __fmul_rn(v.w, v.w));
                if ((f4 & 3) == 3) { acc1 = __fadd_rn(acc1, acc0); acc0 = 0.f; }
            }
            r = __fadd_rn(__fadd_rn(__fadd_rn(acc0, acc1), 0.f), 0.f);
        } else {
            r = cascade_sumsq(xp, F, (size_t)sf);
        }
    } else {
        const int q = F / 4;
        float p0 = cascade_sumsq(xp, q, (size_t)sf * 4);
        const float p1 = cascade_sumsq(xp + sf, q, (size_t)sf * 4);
        const float p2 = cascade_sumsq(xp + 2 * sf, q, (size_t)sf * 4);
        const float p3 = cascade_sumsq(xp + 3 * sf, q, (size_t)sf * 4);
        for (int i = q * 4; i < F; ++i) {
            const float v = xp[(size_t)i * sf];
            p0 = __fadd_rn(p0, __fmul_rn(v, v));
        }
        r = __fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3);
    }
    xx[(size_t)b * N + n] = r;
}

// dynamic smem: xs[F][TP] (TP = tile + 1 pad when staged transposed) + xq[WARPS][F] + sxx[tile]
template <int NSLOT>
__global__ void __launch_bounds__(1024)
knn_expand_kernel(const float* __restrict__ x, const float* __restrict__ xx, int F, int N_alloc, long sf, long sn,
                  int K, int tile, const int32_t* __restrict__ n_valid, int32_t* __restrict__ idx) {
    extern __shared__ float smem[];
    const int warps = blockDim.x >> 5;
    const int TP = tile + 1;
    float* xs = smem;                        // [F][TP]
    float* xq = xs + (size_t)F * TP;         // [warps][F]
    float* sxx = xq + (size_t)warps * F;     // [tile]
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * warps + warp;
    const int N = len_valid(n_valid, b, N_alloc);
    if (i >= N && i < N_alloc) len_fill_row(idx, nullptr, ((size_t)b * N_alloc + i) * K, K, lane);
    const bool active = i < N;
    const float* __restrict__ xb = x + (size_t)b * F * N_alloc;
    const float* __restrict__ xxb = xx + (size_t)b * N_alloc;

    if (active)
        for (int f = lane; f < F; f += 32) xq[warp * F + f] = xb[(size_t)f * sf + (size_t)i * sn];
    const float xxi = active ? xxb[i] : 0.f;
    const float* __restrict__ myq = xq + warp * F;

    WarpList<NSLOT> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;

    for (int t0 = 0; t0 < N; t0 += tile) {
        const int tn = min(tile, N - t0);
        __syncthreads();
        if (sn == 1) {          // channel-major source: consecutive threads -> consecutive points
            for (int e = threadIdx.x; e < F * tn; e += blockDim.x) {
                const int f = e / tn, j = e - f * tn;
                xs[f * TP + j] = xb[(size_t)f * sf + (size_t)(t0 + j)];
            }
        } else {                // point-major source: consecutive threads -> consecutive channels
            for (int e = threadIdx.x; e < F * tn; e += blockDim.x) {
                const int j = e / F, f = e - j * F;
                xs[f * TP + j] = xb[(size_t)f * sf + (size_t)(t0 + j) * sn];
            }
        }
        for (int j = threadIdx.x; j < tn; j += blockDim.x) sxx[j] = xxb[t0 + j];
        __syncthreads();
        if (!active) continue;
        for (int c0 = 0; c0 < tn; c0 += 32) {
            const int j = c0 + lane;
            u64 key = PCNBR_KEY_MAX;
            if (j < tn) {
                float c = __fmul_rn(myq[0], xs[j]);                     // sgemm: FMA chain over f
                for (int f = 1; f < F; ++f) c = __fmaf_rn(myq[f], xs[f * TP + j], c);
                const float inner = __fmul_rn(-2.0f, c);                // dgcnn.py:16
                const float pd = __fsub_rn(__fsub_rn(-sxx[j], inner), xxi);   // dgcnn.py:18
                key = pack_key(f2ord(-pd), (uint32_t)(t0 + j));        // largest pd first
            }
            uint32_t pass = __ballot_sync(PCNBR_FULL, key < thr);
            while (pass) {
                const int src = __ffs(pass) - 1;
                pass &= pass - 1;
                const u64 cand = shfl64(key, src);
                if (cand < thr) {
                    list.insert(cand, lane);
                    thr = list.at(K - 1);
                }
            }
        }
    }
    if (!active) return;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int pos = s * 32 + lane;
        if (pos < K) idx[((size_t)b * N_alloc + i) * K + pos] = (int32_t)min((uint32_t)list.v[s], (uint32_t)(N - 1));
    }
}

int launch_sumsq(const float* x, int B, int F, int N, long sf, long sn, const int32_t* n_valid, float* xx, cudaStream_t s) {
    PCNBR_TIMED("sumsq_cascade_kernel", s, 4.0 * B * ((double)N * F + N), 2.0 * B * (double)N * F,
                (sumsq_cascade_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(x, F, N, sf, sn, n_valid, xx)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

static inline int expand_tile(int F) {
    int t = (8192 / F) & ~31;
    if (t < 32) t = 32;
    if (t > 1024) t = 1024;
    return t;
}

// ---- multi-radius ball query ("MSG": several group() calls on one centroid set, common.py:37-61 per scale) --------
// One selection pass with the LARGEST radius and the largest K gives, per centroid, the sorted list L of its K_all
// nearest in-ball points (then the canonical padding).  For a smaller radius r_i the in-ball points are a prefix of L
// (L is sorted by distance; if fewer than K_all entries of L lie within r_i, no point within r_i is missing from L), so
// scale i's table is   [first min(c_i, K_i) entries of L, c_i = #{d2 <= r_i^2}]  +  [the lowest-index points with
// !(d2 <= r_i^2) until K_i entries]   -- exactly what a stable sort of the reference's masked distance row yields.  The
// padding scan recomputes the distance with the same rounding sequence as the selection and normally ends after one
// 32-point step (low-index points outside a small ball are plentiful).
constexpr int MSG_MAXR = 8;
struct MsgScales {
    float r2[MSG_MAXR];
    int K[MSG_MAXR];
    int32_t* out[MSG_MAXR];
    int R;
};

__global__ void __launch_bounds__(256)
ball_derive_kernel(const float* __restrict__ q, const float* __restrict__ p, int M, int N_alloc, int Kall, float r2max,
                   const int32_t* __restrict__ n_src, const int32_t* __restrict__ idx_all, const float* __restrict__ d2_all, MsgScales sc) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    const int N = len_valid(n_src, b, N_alloc);
    const float* __restrict__ pb = p + (size_t)b * N_alloc * 3;
    const float* __restrict__ qp = q + ((size_t)b * M + m) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    const size_t row = ((size_t)b * M + m) * Kall;
    for (int i = 0; i < sc.R; ++i) {
        const int K = sc.K[i];
        const float r2 = sc.r2[i];
        int32_t* __restrict__ out = sc.out[i] + ((size_t)b * M + m) * K;
        if (r2 == r2max) {                                   // same ball: a prefix of the list, padding included
            for (int pos = lane; pos < K; pos += 32) out[pos] = idx_all[row + pos];
            continue;
        }
        int c = 0;
        for (int pos0 = 0; pos0 < Kall; pos0 += 32) {
            const int pos = pos0 + lane;
            const bool in = pos < Kall && d2_all[row + pos] <= r2;           // padding entries carry +inf
            c += __popc(__ballot_sync(PCNBR_FULL, in));
        }
        int filled = c < K ? c : K;
        for (int pos = lane; pos < filled; pos += 32) out[pos] = idx_all[row + pos];
        for (int j0 = 0; filled < K && j0 < N; j0 += 32) {
            const int j = j0 + lane;
            bool outside = false;
            if (j < N) outside = !(d2_direct(pb[3 * j], pb[3 * j + 1], pb[3 * j + 2], qx, qy, qz) <= r2);   // common.py:58-59
            const uint32_t mask = __ballot_sync(PCNBR_FULL, outside);
            const int pos = filled + __popc(mask & ((1u << lane) - 1u));
            if (outside && pos < K) out[pos] = j;
            filled += __popc(mask);
        }
    }
}

}  // namespace pcnbr

using namespace pcnbr;

template <bool RADIUS>
static int launch_select(const float* q, const float* p, int B, int M, int N, float r2, int K, const int32_t* n_qry,
                         const int32_t* n_src, int32_t* idx, float* d2, cudaStream_t s) {
    if (!q || !p || !idx || B <= 0 || M <= 0 || N <= 0 || K <= 0 || K > N) return PCNBR_E_BADARG;
    if (K > 128) return PCNBR_E_TOOLARGE;
    dim3 grid((M + SEL_WARPS - 1) / SEL_WARPS, B), block(SEL_WARPS * 32);
    // K2/K3: 8 M N flop + compares; compulsory 12 (N + M) + 4 M K (+ 4 M K distances) bytes per cloud
    const double wb = (double)B * (12.0 * (N + M) + (d2 ? 8.0 : 4.0) * M * K), wf = 8.0 * B * (double)M * N;
    const char* nm = RADIUS ? "select_xyz_kernel<ball>" : "select_xyz_kernel<knn>";
    if (K <= 32)      PCNBR_TIMED(nm, s, wb, wf, (select_xyz_kernel<1, RADIUS><<<grid, block, 0, s>>>(q, p, M, N, r2, K, n_qry, n_src, idx, d2)));
    else if (K <= 64) PCNBR_TIMED(nm, s, wb, wf, (select_xyz_kernel<2, RADIUS><<<grid, block, 0, s>>>(q, p, M, N, r2, K, n_qry, n_src, idx, d2)));
    else              PCNBR_TIMED(nm, s, wb, wf, (select_xyz_kernel<4, RADIUS><<<grid, block, 0, s>>>(q, p, M, N, r2, K, n_qry, n_src, idx, d2)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_ball_query_f32(const float* q, const float* p, int B, int M, int N, float r2, int K,
                                    int32_t* idx, pcnbr_stream_t stream) {
    return launch_select<true>(q, p, B, M, N, r2, K, nullptr, nullptr, idx, nullptr, (cudaStream_t)stream);
}

extern "C" int pcnbr_knn_direct_f32(const float* q, const float* p, int B, int M, int N, int K, int32_t* idx,
                                    float* d2, pcnbr_stream_t stream) {
    return launch_select<false>(q, p, B, M, N, 0.f, K, nullptr, nullptr, idx, d2, (cudaStream_t)stream);
}

namespace pcnbr {
// scan forms with per-cloud lengths (grid.cu dispatches here for small clouds)
int select_ball_len(const float* q, const float* p, int B, int M, int N, float r2, int K, const int32_t* n_qry, const int32_t* n_src,
                    int32_t* idx, cudaStream_t s) {
    return launch_select<true>(q, p, B, M, N, r2, K, n_qry, n_src, idx, nullptr, s);
}
int select_knn_len(const float* q, const float* p, int B, int M, int N, int K, const int32_t* n_qry, const int32_t* n_src,
                   int32_t* idx, float* d2, cudaStream_t s) {
    return launch_select<false>(q, p, B, M, N, 0.f, K, n_qry, n_src, idx, d2, s);
}
}  // namespace pcnbr

extern "C" size_t pcnbr_ball_query_multi_ws_bytes(int B, int M, int Kmax) {
    return (size_t)B * (size_t)M * (size_t)Kmax * (sizeof(int32_t) + sizeof(float));
}

extern "C" int pcnbr_ball_query_multi_f32(const float* q, const float* p, int B, int M, int N, const float* r2, const int* K,
                                          int R, int32_t* const* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    return pcnbr_ball_query_multi_len_f32(q, p, B, M, N, r2, K, R, nullptr, idx, ws, ws_bytes, stream);
}

extern "C" int pcnbr_ball_query_multi_len_f32(const float* q, const float* p, int B, int M, int N, const float* r2, const int* K,
                                              int R, const int32_t* n_src, int32_t* const* idx, void* ws, size_t ws_bytes,
                                              pcnbr_stream_t stream) {
    if (!r2 || !K || !idx || R <= 0) return PCNBR_E_BADARG;
    if (R > MSG_MAXR) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    if (R == 1) return launch_select<true>(q, p, B, M, N, r2[0], K[0], nullptr, n_src, idx[0], nullptr, s);
    MsgScales sc;
    sc.R = R;
    float r2max = r2[0];
    int Kall = 0;
    for (int i = 0; i < R; ++i) {
        if (!idx[i] || K[i] <= 0 || K[i] > N) return PCNBR_E_BADARG;
        sc.r2[i] = r2[i]; sc.K[i] = K[i]; sc.out[i] = idx[i];
        if (r2[i] > r2max) r2max = r2[i];
        if (K[i] > Kall) Kall = K[i];
    }
    if (!ws || ws_bytes < pcnbr_ball_query_multi_ws_bytes(B, M, Kall)) return PCNBR_E_WORKSPACE;
    int32_t* idx_all = (int32_t*)ws;
    float* d2_all = (float*)(idx_all + (size_t)B * M * Kall);
    int rc = launch_select<true>(q, p, B, M, N, r2max, Kall, nullptr, n_src, idx_all, d2_all, s);
    if (rc) return rc;
    double out_bytes = 0.0;
    for (int i = 0; i < R; ++i) out_bytes += 4.0 * K[i];
    PCNBR_TIMED("ball_derive_kernel", s, (double)B * M * (8.0 * Kall + out_bytes + 12.0) + 12.0 * B * N, 0.0,
                (ball_derive_kernel<<<dim3((M + 7) / 8, B), 256, 0, s>>>(q, p, M, N, Kall, r2max, n_src, idx_all, d2_all, sc)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

static bool use_tensor_cores(int F, int N, int K) {
    static const bool forced_generic = getenv("PCNBR_KNN_GENERIC") != nullptr;
    return !forced_generic && knn_tc_supported(F, N, K);
}

extern "C" size_t pcnbr_knn_expand_ws_bytes(int B, int F, int N, int K) {
    const size_t generic = sizeof(float) * (size_t)B * (size_t)N;      // xx
    const size_t tc = knn_tc_supported(F, N, K) ? knn_tc_ws_bytes(B, F, N) : 0;
    return tc > generic ? tc : generic;
}

extern "C" int pcnbr_knn_tc_debug_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                                      int32_t* idx, void* ws, size_t ws_bytes, float* scores, int32_t* stats,
                                      pcnbr_stream_t stream) {
    if (!x || !idx || B <= 0 || F <= 0 || N <= 0 || K <= 0 || K > N) return PCNBR_E_BADARG;
    if (!knn_tc_supported(F, N, K)) return PCNBR_E_TOOLARGE;
    if (!ws || ws_bytes < pcnbr_knn_expand_ws_bytes(B, F, N, K)) return PCNBR_E_WORKSPACE;
    return knn_tc_run(x, B, F, N, stride_f, stride_n, K, nullptr, idx, ws, scores, stats, (cudaStream_t)stream);
}

extern "C" int pcnbr_knn_expand_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                                    int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    return pcnbr_knn_expand_len_f32(x, B, F, N, stride_f, stride_n, K, nullptr, idx, ws, ws_bytes, stream);
}

extern "C" int pcnbr_knn_expand_len_f32(const float* x, int B, int F, int N, long stride_f, long stride_n, int K,
                                        const int32_t* n_valid, int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    if (!x || !idx || B <= 0 || F <= 0 || N <= 0 || K <= 0 || K > N) return PCNBR_E_BADARG;
    if (K > 128 || F > 256) return PCNBR_E_TOOLARGE;
    if (!ws || ws_bytes < pcnbr_knn_expand_ws_bytes(B, F, N, K)) return PCNBR_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    if (use_tensor_cores(F, N, K))
        return knn_tc_run(x, B, F, N, stride_f, stride_n, K, n_valid, idx, ws, nullptr, nullptr, s);
    float* xx = (float*)ws;
    int rc = launch_sumsq(x, B, F, N, stride_f, stride_n, n_valid, xx, s);
    if (rc) return rc;
    const int warps = (F >= 16) ? 32 : 8;
    const int tile = expand_tile(F);
    const size_t smem = sizeof(float) * ((size_t)F * (tile + 1) + (size_t)warps * F + tile);
    dim3 grid((N + warps - 1) / warps, B), block(warps * 32);
#define PCNBR_LAUNCH_EXPAND(NS)                                                                          \
    do {                                                                                                 \
        cudaError_t e = cudaFuncSetAttribute(knn_expand_kernel<NS>,                                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        if (e != cudaSuccess) return (int)e;                                                             \
        PCNBR_TIMED("knn_expand_kernel", s, (double)B * (4.0 * N * F + 4.0 * N * K), 2.0 * B * (double)N * N * F,      \
                    (knn_expand_kernel<NS><<<grid, block, smem, s>>>(x, xx, F, N, stride_f, stride_n, K, tile, n_valid, idx)));  \
    } while (0)
    if (K <= 32)      PCNBR_LAUNCH_EXPAND(1);
    else if (K <= 64) PCNBR_LAUNCH_EXPAND(2);
    else              PCNBR_LAUNCH_EXPAND(4);
#undef PCNBR_LAUNCH_EXPAND
    PCNBR_CHECK_LAUNCH();
    return 0;
}
