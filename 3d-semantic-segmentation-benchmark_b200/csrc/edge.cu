// edge.cu -- K9 DGCNN edge features (get_graph_feature) and their backward.
//
// Reference: models/dgcnn/dgcnn.py:41-55 -- flat gather, repeat of the centre k times, cat, permute,
// contiguous: three full-size temporaries plus a transposing copy of the (B,N,k,2F) result.
// Here the (B,N,k,2F) tensor is written exactly once, point-major (the host hands it to the 1x1 conv
// as the channels-last view (B,2F,N,k), which is what cuDNN wants anyway).  One warp per (n,j) edge
// row: lanes stride over the 2F output floats, so the store is one coalesced 8F-byte burst and the
// two source rows are coalesced 4F-byte reads that hit L2 (x is 4*N*F bytes per cloud).
// HBM-bound: 8*F*N*k written + 4*F*N + 4*N*k read per cloud (41.9 MB at F=64, N=4096, k=20).
#include "common.cuh"
#include "segsum.cuh"

namespace pcnbr {

__global__ void __launch_bounds__(256)
edge_fwd_kernel(const float* __restrict__ xt, const int32_t* __restrict__ idx, int N, int F, int K, long rows,
                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long warp_global = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const long NK = (long)N * K;
    const int W = 2 * F;
    if (W <= 32) {
        const int rpw = 32 / W;
        const int sub = lane / W, col = lane - sub * W;
        for (long r0 = warp_global * rpw; r0 < rows; r0 += nwarps * rpw) {
            const long r = r0 + sub;
            if (sub < rpw && r < rows) {
                const long b = r / NK;
                const long n = (r - b * NK) / K;
                const int f = (col < F) ? col : col - F;
                const float c = xt[((size_t)b * N + n) * F + f];
                float v = c;
                if (col < F) v = __fsub_rn(xt[((size_t)b * N + idx[r]) * F + f], c);   // dgcnn.py:53
                out[(size_t)r * W + col] = v;
            }
        }
    } else {
        for (long r = warp_global; r < rows; r += nwarps) {
            const long b = r / NK;
            const long n = (r - b * NK) / K;
            const float* __restrict__ ctr = xt + ((size_t)b * N + n) * F;
            const float* __restrict__ nbr = xt + ((size_t)b * N + idx[r]) * F;
            float* __restrict__ o = out + (size_t)r * W;
            for (int f = lane; f < F; f += 32) {
                const float c = ctr[f];
                o[f] = __fsub_rn(nbr[f], c);
                o[F + f] = c;
            }
        }
    }
}

// d/dx of the neighbour term: gather of g[b,e,0:F] over the CSR segment of the source point.
struct EdgeBwdSrc {
    const float* g; long E; int W;
    __device__ __forceinline__ const float* row(int b, int e) const { return g + ((size_t)b * E + e) * W; }
    __device__ __forceinline__ float scale(int, int) const { return 1.0f; }
};
// ... plus the dense centre term: - sum_j g[n,j,0:F] + sum_j g[n,j,F:2F], added at store time.
struct EdgeBwdDst {
    const float* g; float* out; long N; int F; int K;
    __device__ __forceinline__ void store(int b, int s, int c, float v) const {
        const float* __restrict__ row = g + (((size_t)b * N + s) * K) * (2 * F);
        float a = 0.f;
        for (int j = 0; j < K; ++j) a += row[(size_t)j * 2 * F + F + c] - row[(size_t)j * 2 * F + c];
        out[((size_t)b * N + s) * F + c] = v + a;
    }
};

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_edge_feature_f32(const float* xt, const int32_t* idx, int B, int N, int F, int K, float* out,
                                      pcnbr_stream_t stream) {
    if (!xt || !idx || !out || B <= 0 || N <= 0 || F <= 0 || K <= 0) return PCNBR_E_BADARG;
    const long rows = (long)B * N * K;
    const long per_warp = (2 * F <= 32) ? 32 / (2 * F) : 1;
    long blocks = (rows + per_warp * 8 - 1) / (per_warp * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    // K9: 8 F N k written + 4 F N + 4 N k read per cloud
    PCNBR_TIMED("edge_fwd_kernel", (cudaStream_t)stream, (double)B * (8.0 * F * N * K + 4.0 * F * N + 4.0 * N * K), (double)B * N * K * F,
                (edge_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(xt, idx, N, F, K, rows, out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_edge_feature_bwd_f32(const float* g, const int32_t* offsets, const int32_t* perm, int B, int N,
                                          int F, int K, float* gxt, pcnbr_stream_t stream) {
    if (!g || !offsets || !perm || !gxt || B <= 0 || N <= 0 || F <= 0 || K <= 0) return PCNBR_E_BADARG;
    EdgeBwdSrc src{g, (long)N * K, 2 * F};
    EdgeBwdDst dst{g, gxt, (long)N, F, K};
    PCNBR_TIMED("segsum_kernel<edge_bwd>", (cudaStream_t)stream, (double)B * (8.0 * F * N * K + 4.0 * F * N + 8.0 * N * K), 3.0 * B * (double)N * K * F,
                (segsum_kernel<<<segsum_grid(N, B), 256, 0, (cudaStream_t)stream>>>(src, dst, offsets, perm, N, N * K, F)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
