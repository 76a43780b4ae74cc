// group.cu -- K5 neighbour gather fused with centre-subtract (+ /r) and concat, and its backward.
//
// Reference: models/utils/common.py:62-71 -- two advanced-index gathers, an in-place subtract, an
// optional in-place divide and a torch.cat, i.e. five passes over the (B,M,K,3+D) tensor.  Here the
// output row (3+D floats) is produced in one pass: consecutive lanes own consecutive output floats,
// so stores are fully coalesced; the source rows (<= 1 KB) come from L2.
// HBM-bound: bytes = 4*M*K*(3+D) written + 4*M*K idx + 4*N*(3+D) + 12*M read per cloud.
#include "common.cuh"
#include "segsum.cuh"

namespace pcnbr {

__global__ void __launch_bounds__(256)
group_fwd_kernel(const float* __restrict__ p, const float* __restrict__ feat, const float* __restrict__ q,
                 const int32_t* __restrict__ idx, int N, int M, int K, int D, float rdiv, int ldo, int split_log2,
                 float* __restrict__ out) {
    // One warp per centroid (b, m): its K output rows are contiguous (K * W floats), the neighbour indices sit in the
    // lanes, and no address needs a division by M*K or K (the first version spent two 64-bit divisions per output element).
    const int W = ldo;                                     // output row pitch >= 3 + D; the pad columns are written as 0
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    // wide rows: 2^split_log2 warps share a centroid, each takes a contiguous slice of every 32-row chunk (the deeper levels
    // have few centroids: 512 at 32 x 16 -- a warp per centroid left most of the chip idle)
    const int wv = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int m = wv >> split_log2, part = wv & ((1 << split_log2) - 1);
    if (m >= M) return;
    const float* __restrict__ pb = p + (size_t)b * N * 3;
    const float* __restrict__ fb = feat + (size_t)b * N * D;
    const float* __restrict__ qc = q + ((size_t)b * M + m) * 3;
    const float qv = lane < 3 ? qc[lane] : 0.f;            // lane c < 3 holds centroid coordinate c
    const int32_t* __restrict__ ib = idx + ((size_t)b * M + m) * K;
    float* __restrict__ ob = out + ((size_t)b * M + m) * (size_t)K * W;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int kn = min(32, K - k0);
        const int mine = lane < kn ? ib[k0 + lane] : 0;
        float* __restrict__ oc = ob + (size_t)k0 * W;
        if (W <= 32) {
            // narrow rows (SA1: 12 floats): the chunk's kn * W floats as one flat, fully coalesced array.  Lane k first
            // computes the local coordinates of ITS neighbour (3 loads, 3 subtractions, 3 divisions per row instead of a
            // divergent load + division inside the element loop); the loop then only shuffles and copies.
            const int total = kn * W;
            float lx = 0.f, ly = 0.f, lz = 0.f;
            if (lane < kn) {
                const float* __restrict__ ps = pb + (size_t)mine * 3;
                lx = __fsub_rn(ps[0], qc[0]); ly = __fsub_rn(ps[1], qc[1]); lz = __fsub_rn(ps[2], qc[2]);
                if (rdiv > 0.f) { lx = __fdiv_rn(lx, rdiv); ly = __fdiv_rn(ly, rdiv); lz = __fdiv_rn(lz, rdiv); }   // common.py:69
            }
            if ((W & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
                // 16-byte pitch (the padded rows the set-abstraction modules ask for): one float4 of output per lane and
                // step -- the index arithmetic is paid once per four floats and the store is a single STG.128
                const int W4 = W >> 2, total4 = kn * W4;
                const uint32_t inv4 = (65536u + (uint32_t)W4 - 1u) / (uint32_t)W4;
                for (int f0 = 0; f0 < total4; f0 += 32) {
                    const int f = f0 + lane;
                    const int k = (int)(((uint32_t)min(f, total4 - 1) * inv4) >> 16), c4 = f - k * W4;
                    const int s = __shfl_sync(PCNBR_FULL, mine, k);
                    const float vx = __shfl_sync(PCNBR_FULL, lx, k), vy = __shfl_sync(PCNBR_FULL, ly, k), vz = __shfl_sync(PCNBR_FULL, lz, k);
                    if (f < total4) {
                        const float* __restrict__ fs = fb + (size_t)s * D;
                        const int cb = 4 * c4 - 3;                     // feature column of the float4's first element
                        float4 v;
                        v.x = c4 == 0 ? vx : (cb < D ? fs[cb] : 0.f);
                        v.y = c4 == 0 ? vy : (cb + 1 < D ? fs[cb + 1] : 0.f);
                        v.z = c4 == 0 ? vz : (cb + 2 < D ? fs[cb + 2] : 0.f);
                        v.w = cb + 3 < D ? fs[cb + 3] : 0.f;
                        *reinterpret_cast<float4*>(oc + 4 * f) = v;
                    }
                }
                continue;
            }
            const uint32_t inv = (65536u + (uint32_t)W - 1u) / (uint32_t)W;   // e / W == (e * inv) >> 16 for e < 1024, W <= 32
#pragma unroll 2
            for (int e0 = 0; e0 < total; e0 += 32) {              // warp-uniform trip count: every lane takes part in the shuffles
                const int e = e0 + lane;
                const int k = (int)(((uint32_t)min(e, total - 1) * inv) >> 16), col = e - k * W;
                const int s = __shfl_sync(PCNBR_FULL, mine, k);
                const float vx = __shfl_sync(PCNBR_FULL, lx, k), vy = __shfl_sync(PCNBR_FULL, ly, k), vz = __shfl_sync(PCNBR_FULL, lz, k);
                float v = col == 0 ? vx : (col == 1 ? vy : vz);
                if (col >= 3) v = (col < 3 + D && e < total) ? fb[(size_t)s * D + (col - 3)] : 0.f;
                if (e < total) oc[e] = v;
            }
        } else {
            // wide rows: lanes over the columns of a row, four rows in flight
            const int slice = 32 >> split_log2;
            const int k_hi = min(kn, (part + 1) * slice);
            for (int k = part * slice; k < k_hi; k += 4) {
                int s[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) s[u] = __shfl_sync(PCNBR_FULL, mine, min(k + u, kn - 1));
                for (int c0 = 0; c0 < W; c0 += 32) {
                    const int c = c0 + lane;
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        v[u] = 0.f;
                        if (c < 3) v[u] = pb[(size_t)s[u] * 3 + c];
                        else if (c < 3 + D) v[u] = fb[(size_t)s[u] * D + (c - 3)];
                    }
                    if (c < 3) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            v[u] = __fsub_rn(v[u], qv);
                            if (rdiv > 0.f) v[u] = __fdiv_rn(v[u], rdiv);
                        }
                    }
                    if (c < W) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (k + u < k_hi) oc[(size_t)(k + u) * W + c] = v[u];
                    }
                }
            }
        }
    }
}

// Backward w.r.t. feat: gfeat[b,s,:] = sum over the CSR segment of s of gout[b,e,3:3+D]  (segsum.cuh).
struct GroupBwdSrc {
    const float* g; long E; int W;
    __device__ __forceinline__ const float* row(int b, int e) const { return g + ((size_t)b * E + e) * W + 3; }
    __device__ __forceinline__ float scale(int, int) const { return 1.0f; }
};
struct RowMajorDst {
    float* out; long N; int D;
    __device__ __forceinline__ void store(int b, int s, int c, float v) const { out[((size_t)b * N + s) * D + c] = v; }
};

// ---- index_points: plain row gather out[b,e,:] = src[b, idx[b,e], :]  (the `points[batch_indices, indices]` of
// common.py:64-65,117 on its own; north-star name index_points).  One warp per output row, float4 when the rows allow it.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int N, long E, int D, float* __restrict__ out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const long nw = (long)gridDim.x * 8;
    const bool vec4 = (D % 4 == 0) && ((((uintptr_t)src | (uintptr_t)out) & 15) == 0);
    for (long e = (long)blockIdx.x * 8 + (threadIdx.x >> 5); e < E; e += nw) {
        const int n = min(max(idx[(size_t)b * E + e], 0), N - 1);                     // an out-of-range index is clamped, never dereferenced
        const float* __restrict__ r = src + ((size_t)b * N + n) * D;
        float* __restrict__ o = out + ((size_t)b * E + e) * D;
        if (vec4) {
            for (int c = lane * 4; c < D; c += 128) *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(r + c);
        } else {
            for (int c = lane; c < D; c += 32) o[c] = r[c];
        }
    }
}
struct GatherBwdSrc {
    const float* g; long E; int D;
    __device__ __forceinline__ const float* row(int b, int e) const { return g + ((size_t)b * E + e) * D; }
    __device__ __forceinline__ float scale(int, int) const { return 1.0f; }
};

// ---- square_distance: the (B,N,M) matrix ((dst - src)^2).sum(-1) of common.py:54-56 / 110-112, materialised.  Inspection /
// compatibility only: no kernel of the path ever builds it.  dst tiles of 256 points in shared memory, one thread per src.
__global__ void __launch_bounds__(256)
square_distance_kernel(const float* __restrict__ src, const float* __restrict__ dst, int N, int M, float* __restrict__ out) {
    __shared__ float sx[256], sy[256], sz[256];
    const int b = blockIdx.z, n = blockIdx.x * 256 + threadIdx.x, m0 = blockIdx.y * 256;
    const int mt = min(256, M - m0);
    if ((int)threadIdx.x < mt) {
        const float* d = dst + ((size_t)b * M + m0 + threadIdx.x) * 3;
        sx[threadIdx.x] = d[0]; sy[threadIdx.x] = d[1]; sz[threadIdx.x] = d[2];
    }
    __syncthreads();
    if (n >= N) return;
    const float* s = src + ((size_t)b * N + n) * 3;
    const float x = s[0], y = s[1], z = s[2];
    float* __restrict__ o = out + ((size_t)b * N + n) * M + m0;
    for (int j = 0; j < mt; ++j) o[j] = d2_direct(sx[j], sy[j], sz[j], x, y, z);          // (dst - src)^2 summed as (dx2 + dy2) + dz2
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_gather_rows_f32(const float* src, const int32_t* idx, int B, int N, long E, int D, float* out, pcnbr_stream_t stream) {
    if (!src || !idx || !out || B <= 0 || N <= 0 || E <= 0 || D <= 0) return PCNBR_E_BADARG;
    if (B > 65535) return PCNBR_E_TOOLARGE;
    const long gx = (E + 7) / 8 < 148L * 16 ? (E + 7) / 8 : 148L * 16;
    PCNBR_TIMED("gather_rows_kernel", (cudaStream_t)stream, (double)B * (8.0 * E * D + 4.0 * E), 0.0,
                (gather_rows_kernel<<<dim3((unsigned)gx, B), 256, 0, (cudaStream_t)stream>>>(src, idx, N, E, D, out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_gather_rows_bwd_f32(const float* gout, const int32_t* offsets, const int32_t* perm, int B, int N, int E, int D,
                                         float* gsrc, pcnbr_stream_t stream) {
    if (!gout || !offsets || !perm || !gsrc || B <= 0 || N <= 0 || E <= 0 || D <= 0) return PCNBR_E_BADARG;
    GatherBwdSrc src{gout, (long)E, D};
    RowMajorDst dst{gsrc, (long)N, D};
    PCNBR_TIMED("segsum_kernel<gather_bwd>", (cudaStream_t)stream, (double)B * (4.0 * E * D + 4.0 * E + 4.0 * N + 4.0 * N * D), (double)B * E * D,
                (segsum_kernel<<<segsum_grid(N, B), 256, 0, (cudaStream_t)stream>>>(src, dst, offsets, perm, N, E, D)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_square_distance_f32(const float* src, const float* dst, int B, int N, int M, float* out, pcnbr_stream_t stream) {
    if (!src || !dst || !out || B <= 0 || N <= 0 || M <= 0) return PCNBR_E_BADARG;
    if (B > 65535 || (M + 255) / 256 > 65535) return PCNBR_E_TOOLARGE;
    PCNBR_TIMED("square_distance_kernel", (cudaStream_t)stream, (double)B * (4.0 * N * M + 12.0 * (N + M)), 8.0 * B * (double)N * M,
                (square_distance_kernel<<<dim3((N + 255) / 256, (M + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(src, dst, N, M, out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_group_f32(const float* p, const float* feat, const float* q, const int32_t* idx, int B,
                               int N, int M, int K, int D, float rdiv, float* out, int ldo, pcnbr_stream_t stream) {
    if (!p || !q || !idx || !out || (D > 0 && !feat) || B <= 0 || N <= 0 || M <= 0 || K <= 0 || D < 0 || ldo < 3 + D || ldo > 3 + D + 32)
        return PCNBR_E_BADARG;
    if (B > 65535) return PCNBR_E_TOOLARGE;
    const int W = ldo;
    // measured (32 clouds): splitting helps the deep levels (16 centroids x 260 floats: 39 -> 16 us; 64 x 132: 30 -> 25 us)
    // and hurts once a warp per centroid already fills the chip (256 centroids x 68 floats: 40 -> 49 us)
    const int split_log2 = (W <= 32 || (long)B * M > 4096) ? 0 : ((long)B * M <= 1024 ? 3 : 2);
    const long warps = (long)M << split_log2;
    // K5: 4 M K (3+D) written + 4 M K idx + 4 N (3+D) + 12 M read per cloud
    PCNBR_TIMED("group_fwd_kernel", (cudaStream_t)stream, (double)B * (4.0 * M * K * W + 4.0 * M * K + 4.0 * N * W + 12.0 * M), 0.0,
                (group_fwd_kernel<<<dim3((unsigned)((warps + 7) / 8), B), 256, 0, (cudaStream_t)stream>>>(p, feat, q, idx, N, M, K, D, rdiv, ldo, split_log2, out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_group_bwd_f32(const float* gout, int ldg, const int32_t* offsets, const int32_t* perm, int B, int N,
                                   int E, int D, float* gfeat, pcnbr_stream_t stream) {
    if (!gout || !offsets || !perm || !gfeat || B <= 0 || N <= 0 || E <= 0 || D <= 0 || ldg < 3 + D) return PCNBR_E_BADARG;
    GroupBwdSrc src{gout, (long)E, ldg};
    RowMajorDst dst{gfeat, (long)N, D};
    // K7: 4 E D read + 4 E perm + 4 N offsets + 4 N D written per cloud
    PCNBR_TIMED("segsum_kernel<group_bwd>", (cudaStream_t)stream, (double)B * (4.0 * E * D + 4.0 * E + 4.0 * N + 4.0 * N * D), (double)B * E * D,
                (segsum_kernel<<<segsum_grid(N, B), 256, 0, (cudaStream_t)stream>>>(src, dst, offsets, perm, N, E, D)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
