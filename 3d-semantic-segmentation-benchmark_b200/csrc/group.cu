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

// rows = B*M*K output rows of width W = 3 + D.  Each warp handles RPW = max(1, 32/W) rows per pass
// when W < 32 (lane -> (row, col)), or one row per pass with lanes striding over the columns.
__global__ void __launch_bounds__(256)
group_fwd_kernel(const float* __restrict__ p, const float* __restrict__ feat, const float* __restrict__ q,
                 const int32_t* __restrict__ idx, int N, int M, int K, int D, float rdiv, long rows, int ldo,
                 float* __restrict__ out) {
    const int W = ldo;                                     // output row pitch >= 3 + D; the pad columns are written as 0
    const int lane = threadIdx.x & 31;
    const long warp_global = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const long MK = (long)M * K;
    if (W <= 32) {
        const int rpw = 32 / W;
        const int sub = lane / W, col = lane - sub * W;
        for (long r0 = warp_global * rpw; r0 < rows; r0 += nwarps * rpw) {
            const long r = r0 + sub;
            if (sub < rpw && r < rows) {
                const long b = r / MK;
                const long m = (r - b * MK) / K;
                const int s = idx[r];
                float v;
                if (col < 3) {
                    v = __fsub_rn(p[((size_t)b * N + s) * 3 + col], q[((size_t)b * M + m) * 3 + col]);
                    if (rdiv > 0.f) v = __fdiv_rn(v, rdiv);                       // common.py:69
                } else if (col < 3 + D) {
                    v = feat[((size_t)b * N + s) * D + (col - 3)];
                } else {
                    v = 0.f;
                }
                out[(size_t)r * W + col] = v;
            }
        }
    } else {
        for (long r = warp_global; r < rows; r += nwarps) {
            const long b = r / MK;
            const long m = (r - b * MK) / K;
            const int s = idx[r];
            float* __restrict__ o = out + (size_t)r * W;
            const float* __restrict__ fs = feat + ((size_t)b * N + s) * D;
            if (lane < 3) {
                float v = __fsub_rn(p[((size_t)b * N + s) * 3 + lane], q[((size_t)b * M + m) * 3 + lane]);
                if (rdiv > 0.f) v = __fdiv_rn(v, rdiv);
                o[lane] = v;
            }
            for (int c = lane; c < D; c += 32) o[3 + c] = fs[c];
            if (lane < W - 3 - D) o[3 + D + lane] = 0.f;
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

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_group_f32(const float* p, const float* feat, const float* q, const int32_t* idx, int B,
                               int N, int M, int K, int D, float rdiv, float* out, int ldo, pcnbr_stream_t stream) {
    if (!p || !q || !idx || !out || (D > 0 && !feat) || B <= 0 || N <= 0 || M <= 0 || K <= 0 || D < 0 || ldo < 3 + D || ldo > 3 + D + 32)
        return PCNBR_E_BADARG;
    const long rows = (long)B * M * K;
    const int W = ldo;
    const long per_warp = (W <= 32) ? 32 / W : 1;
    long blocks = (rows + per_warp * 8 - 1) / (per_warp * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    // K5: 4 M K (3+D) written + 4 M K idx + 4 N (3+D) + 12 M read per cloud
    PCNBR_TIMED("group_fwd_kernel", (cudaStream_t)stream, (double)B * (4.0 * M * K * W + 4.0 * M * K + 4.0 * N * W + 12.0 * M), 0.0,
                (group_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, feat, q, idx, N, M, K, D, rdiv, rows, ldo, out)));
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
