// interp.cu -- K8 inverse-distance interpolation over the k (=3) nearest coarse points, fwd + bwd.
//
// Reference: models/utils/common.py:115-122 -- gather (B,N,k,D), weights 1/(d2 + 1e-9) on the SQUARED
// distance, normalise, multiply, sum over k: four (B,N,k,D)-sized temporaries.  Here one warp owns
// one fine point: it reads the k indices / distances once, forms the weights, and streams the k
// coarse rows (L2-resident, coalesced over D) into one coalesced output row.
// HBM-bound: 4*N*D written + 4*M*D read + 8*N*k (idx, d2) read + 4*N*k coef written per cloud.
// Arithmetic order as the reference: w = 1/(d2 + 1e-9f); norm = (w0+w1)+w2; out = sum_k (f*w_k)/norm.
#include "common.cuh"
#include "segsum.cuh"

namespace pcnbr {

constexpr int INTERP_KMAX = 8;

__global__ void __launch_bounds__(256)
interp_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, const float* __restrict__ d2,
                  int N, int M, int D, int K, float* __restrict__ out, float* __restrict__ coef) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    const float* __restrict__ fb = feat + (size_t)b * M * D;
    const bool vec4 = (D % 4 == 0) && ((((uintptr_t)feat | (uintptr_t)out) & 15) == 0);
    for (int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); n < N; n += nw) {
        const size_t base = ((size_t)b * N + n) * K;
        float w[INTERP_KMAX];
        int id[INTERP_KMAX];
        float norm = 0.f;
#pragma unroll
        for (int k = 0; k < INTERP_KMAX; ++k)
            if (k < K) {
                id[k] = idx[base + k];
                w[k] = __fdiv_rn(1.0f, __fadd_rn(d2[base + k], 1e-9f));       // common.py:119
                norm = (k == 0) ? w[0] : __fadd_rn(norm, w[k]);              // common.py:120
            }
        if (coef && lane < K) {
            float wl = 0.f;
#pragma unroll
            for (int k = 0; k < INTERP_KMAX; ++k) if (k == lane) wl = w[k];
            coef[base + lane] = __fdiv_rn(wl, norm);
        }
        if (vec4) {
            // 4 channels per lane: K coalesced 512-byte row requests in flight, 4K independent divide chains
            for (int c = lane * 4; c < D; c += 128) {
                float4 f[INTERP_KMAX];
#pragma unroll
                for (int k = 0; k < INTERP_KMAX; ++k)
                    if (k < K) f[k] = *reinterpret_cast<const float4*>(fb + (size_t)id[k] * D + c);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < INTERP_KMAX; ++k)
                    if (k < K) {
                        const float tx = __fdiv_rn(__fmul_rn(f[k].x, w[k]), norm), ty = __fdiv_rn(__fmul_rn(f[k].y, w[k]), norm);
                        const float tz = __fdiv_rn(__fmul_rn(f[k].z, w[k]), norm), tw = __fdiv_rn(__fmul_rn(f[k].w, w[k]), norm);
                        if (k == 0) acc = make_float4(tx, ty, tz, tw);
                        else { acc.x = __fadd_rn(acc.x, tx); acc.y = __fadd_rn(acc.y, ty); acc.z = __fadd_rn(acc.z, tz); acc.w = __fadd_rn(acc.w, tw); }
                    }
                *reinterpret_cast<float4*>(out + ((size_t)b * N + n) * D + c) = acc;
            }
            continue;
        }
        for (int c = lane; c < D; c += 32) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < INTERP_KMAX; ++k)
                if (k < K) {
                    const float t = __fdiv_rn(__fmul_rn(fb[(size_t)id[k] * D + c], w[k]), norm);   // common.py:122
                    acc = (k == 0) ? t : __fadd_rn(acc, t);
                }
            out[((size_t)b * N + n) * D + c] = acc;
        }
    }
}

// k = 3 (the only value the models use), D % 4 == 0: one warp takes FOUR consecutive fine points per step -- lanes 0..11
// fetch the 12 (index, distance) pairs in one coalesced request each, the weights and norms are formed once, and the
// 12 coarse rows are in flight as float4 requests (512 B per row and warp) before the first divide.  Same arithmetic
// order as interp_fwd_kernel, element for element.
__global__ void __launch_bounds__(256)
interp3_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, const float* __restrict__ d2,
                   int N, int M, int D, float* __restrict__ out, float* __restrict__ coef) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    const float* __restrict__ fb = feat + (size_t)b * M * D;
    const int lp = lane / 3, lk = lane - 3 * lp;                         // this lane's (point, neighbour) slot, lane < 12
    for (int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; n0 < N; n0 += nw * 4) {
        const size_t base = ((size_t)b * N + n0) * 3;
        const bool slot = lane < 12 && n0 + lp < N;
        const int my_id = slot ? idx[base + lane] : 0;
        const float my_w = slot ? __fdiv_rn(1.0f, __fadd_rn(d2[base + lane], 1e-9f)) : 1.0f;     // common.py:119
        int id[4][3];
        float w[4][3], norm[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                id[p][k] = __shfl_sync(PCNBR_FULL, my_id, 3 * p + k);
                w[p][k] = __shfl_sync(PCNBR_FULL, my_w, 3 * p + k);
            }
            norm[p] = __fadd_rn(__fadd_rn(w[p][0], w[p][1]), w[p][2]);  // common.py:120
        }
        if (coef && slot) {
            float nm = norm[0];
#pragma unroll
            for (int p = 1; p < 4; ++p) if (lp == p) nm = norm[p];
            coef[base + lane] = __fdiv_rn(my_w, nm);
        }
        (void)lk;
        for (int c = lane * 4; c < D; c += 128) {
            float4 f[4][3];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) f[p][k] = *reinterpret_cast<const float4*>(fb + (size_t)id[p][k] * D + c);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (n0 + p >= N) break;
                float4 acc;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float tx = __fdiv_rn(__fmul_rn(f[p][k].x, w[p][k]), norm[p]), ty = __fdiv_rn(__fmul_rn(f[p][k].y, w[p][k]), norm[p]);
                    const float tz = __fdiv_rn(__fmul_rn(f[p][k].z, w[p][k]), norm[p]), tw = __fdiv_rn(__fmul_rn(f[p][k].w, w[p][k]), norm[p]);
                    if (k == 0) acc = make_float4(tx, ty, tz, tw);
                    else { acc.x = __fadd_rn(acc.x, tx); acc.y = __fadd_rn(acc.y, ty); acc.z = __fadd_rn(acc.z, tz); acc.w = __fadd_rn(acc.w, tw); }
                }
                *reinterpret_cast<float4*>(out + ((size_t)b * N + n0 + p) * D + c) = acc;
            }
        }
    }
}

// gfeat[b,m,:] = sum over incoming positions e = n*K + k of coef[e] * g[b,n,:]
struct InterpBwdSrc {
    const float* g; const float* cf; long N; int D; int K;
    __device__ __forceinline__ const float* row(int b, int e) const { return g + ((size_t)b * N + e / K) * D; }
    __device__ __forceinline__ float scale(int b, int e) const { return cf[(size_t)b * N * K + e]; }
};
struct InterpDst {
    float* out; long M; int D;
    __device__ __forceinline__ void store(int b, int s, int c, float v) const { out[((size_t)b * M + s) * D + c] = v; }
};

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_interp_f32(const float* feat, const int32_t* idx, const float* d2, int B, int N, int M, int D,
                                int K, float* out, float* coef, pcnbr_stream_t stream) {
    if (!feat || !idx || !d2 || !out || B <= 0 || N <= 0 || M <= 0 || D <= 0 || K <= 0) return PCNBR_E_BADARG;
    if (K > INTERP_KMAX) return PCNBR_E_TOOLARGE;
    int gx = (N + 7) / 8;
    if (gx > 148 * 8) gx = 148 * 8;
    // K8: 4 N D written + 4 M D read + 8 N k (idx, d2) read + 4 N k coef written per cloud
    const double wb = (double)B * (4.0 * N * D + 4.0 * M * D + 12.0 * N * K), wf = 3.0 * B * (double)N * D * K;
    if (K == 3 && D % 4 == 0 && ((((uintptr_t)feat | (uintptr_t)out) & 15) == 0)) {
        int g3 = (N + 31) / 32;                                           // 8 warps x 4 points per CTA step
        if (g3 > 148 * 8) g3 = 148 * 8;
        PCNBR_TIMED("interp_fwd_kernel", (cudaStream_t)stream, wb, wf,
                    (interp3_fwd_kernel<<<dim3(g3, B), 256, 0, (cudaStream_t)stream>>>(feat, idx, d2, N, M, D, out, coef)));
    } else {
        PCNBR_TIMED("interp_fwd_kernel", (cudaStream_t)stream, wb, wf,
                    (interp_fwd_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(feat, idx, d2, N, M, D, K, out, coef)));
    }
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_interp_bwd_f32(const float* g, const float* coef, const int32_t* offsets, const int32_t* perm,
                                    int B, int N, int M, int D, int K, float* gfeat, pcnbr_stream_t stream) {
    if (!g || !coef || !offsets || !perm || !gfeat || B <= 0 || N <= 0 || M <= 0 || D <= 0 || K <= 0)
        return PCNBR_E_BADARG;
    InterpBwdSrc src{g, coef, (long)N, D, K};
    InterpDst dst{gfeat, (long)M, D};
    // bwd: 4 N D read + 8 N k (coef, perm) + 4 M offsets + 4 M D written per cloud (the k re-reads of g hit L2)
    PCNBR_TIMED("segsum_kernel<interp_bwd>", (cudaStream_t)stream, (double)B * (4.0 * N * D + 8.0 * N * K + 4.0 * M + 4.0 * M * D), 2.0 * B * (double)N * K * D,
                (segsum_kernel<<<segsum_grid(M, B), 256, 0, (cudaStream_t)stream>>>(src, dst, offsets, perm, M, N * K, D)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
