// pool.cu -- K6 grouped max-pool over the neighbour axis (+ argmax) and its backward.
//
// Reference: reduce(x,'max') = torch.max(x, dim=2)[0] on the permuted view of the MLP output
// (models/utils/common.py:85-86,211) and x.max(dim=-1)[0] in EdgeConv (models/dgcnn/dgcnn.py:76).
// Element (r,k,d) lives at x[r*stride_r + k*stride_k + d*stride_d]; r enumerates (cloud, centroid).
// Two layouts are handled natively, nothing is copied:
//   stride_d == 1  channels-last conv output: float4 over d, loop over k      (HBM-coalesced)
//   stride_k == 1  NCHW conv output: one thread per (r,d) row of K contiguous floats
// HBM-bound: 4*R*K*D bytes read, 5*R*D written (value + uint8 argmax).
// torch.max semantics: first maximum wins; NaN propagates (a NaN beats everything).
#include "common.cuh"

namespace pcnbr {

__device__ __forceinline__ bool pool_better(float v, float best) { return v > best || (v != v && best == best); }

// stride_d == 1, D % 4 == 0, 16-byte aligned rows
__global__ void __launch_bounds__(256)
maxpool_cl4_kernel(const float* __restrict__ x, long R, int K, int D4, long sr, long sk, float* __restrict__ out,
                   uint8_t* __restrict__ arg) {
    const long total = R * D4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / D4;
        const int d4 = (int)(i - r * D4);
        const float4* __restrict__ src = reinterpret_cast<const float4*>(x + r * sr) + d4;
        float4 best = src[0];
        uchar4 a = make_uchar4(0, 0, 0, 0);
        for (int k = 1; k < K; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (long)k * sk);
            if (pool_better(v.x, best.x)) { best.x = v.x; a.x = (unsigned char)k; }
            if (pool_better(v.y, best.y)) { best.y = v.y; a.y = (unsigned char)k; }
            if (pool_better(v.z, best.z)) { best.z = v.z; a.z = (unsigned char)k; }
            if (pool_better(v.w, best.w)) { best.w = v.w; a.w = (unsigned char)k; }
        }
        reinterpret_cast<float4*>(out)[i] = best;
        reinterpret_cast<uchar4*>(arg)[i] = a;
    }
}

// generic strides, one thread per (r,d)
__global__ void __launch_bounds__(256)
maxpool_generic_kernel(const float* __restrict__ x, long R, int K, int D, long sr, long sk, long sd,
                       float* __restrict__ out, uint8_t* __restrict__ arg) {
    const long total = R * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / D;
        const int d = (int)(i - r * D);
        const float* __restrict__ src = x + r * sr + (long)d * sd;
        float best = src[0];
        int a = 0;
        for (int k = 1; k < K; ++k) {
            const float v = src[(long)k * sk];
            if (pool_better(v, best)) { best = v; a = k; }
        }
        out[i] = best;
        arg[i] = (uint8_t)a;
    }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_cl4_kernel(const float* __restrict__ g, const uint8_t* __restrict__ arg, long R, int K, int D4, long sr,
                       long sk, float* __restrict__ gx) {
    const long total = R * D4;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / D4;
        const int d4 = (int)(i - r * D4);
        const float4 gv = reinterpret_cast<const float4*>(g)[i];
        const uchar4 a = reinterpret_cast<const uchar4*>(arg)[i];
        float* __restrict__ dst = gx + r * sr + 4 * d4;
        for (int k = 0; k < K; ++k) {
            float4 o;
            o.x = (a.x == k) ? gv.x : 0.f;
            o.y = (a.y == k) ? gv.y : 0.f;
            o.z = (a.z == k) ? gv.z : 0.f;
            o.w = (a.w == k) ? gv.w : 0.f;
            *reinterpret_cast<float4*>(dst + (long)k * sk) = o;
        }
    }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_generic_kernel(const float* __restrict__ g, const uint8_t* __restrict__ arg, long R, int K, int D,
                           long sr, long sk, long sd, float* __restrict__ gx) {
    const long total = R * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / D;
        const int d = (int)(i - r * D);
        const float gv = g[i];
        const int a = arg[i];
        float* __restrict__ dst = gx + r * sr + (long)d * sd;
        for (int k = 0; k < K; ++k) dst[(long)k * sk] = (k == a) ? gv : 0.f;
    }
}

static inline unsigned pool_blocks(long total) {
    long b = (total + 255) / 256;
    if (b > 148L * 32) b = 148L * 32;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace pcnbr

using namespace pcnbr;

static bool cl4_ok(const void* x, int D, long sr, long sk, long sd) {
    return sd == 1 && (D % 4) == 0 && (sr % 4) == 0 && (sk % 4) == 0 && ((uintptr_t)x % 16) == 0;
}

extern "C" int pcnbr_maxpool_f32(const float* x, long R, int K, int D, long stride_r, long stride_k, long stride_d,
                                 float* out, uint8_t* arg, pcnbr_stream_t stream) {
    if (!x || !out || !arg || R <= 0 || K <= 0 || D <= 0) return PCNBR_E_BADARG;
    if (K > 255) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const double wb = (double)R * D * (4.0 * K + 5.0);       // K6: 4 R K D read + 5 R D written
    if (cl4_ok(x, D, stride_r, stride_k, stride_d) && ((uintptr_t)out % 16) == 0 && ((uintptr_t)arg % 4) == 0)
        PCNBR_TIMED("maxpool_cl4_kernel", s, wb, 0.0, (maxpool_cl4_kernel<<<pool_blocks(R * (D / 4)), 256, 0, s>>>(x, R, K, D / 4, stride_r, stride_k, out, arg)));
    else
        PCNBR_TIMED("maxpool_generic_kernel", s, wb, 0.0, (maxpool_generic_kernel<<<pool_blocks(R * D), 256, 0, s>>>(x, R, K, D, stride_r, stride_k, stride_d, out, arg)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_maxpool_bwd_f32(const float* g, const uint8_t* arg, long R, int K, int D, long stride_r,
                                     long stride_k, long stride_d, float* gx, pcnbr_stream_t stream) {
    if (!g || !gx || !arg || R <= 0 || K <= 0 || D <= 0) return PCNBR_E_BADARG;
    if (K > 255) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const double wb = (double)R * D * (4.0 * K + 5.0);       // 5 R D read + 4 R K D written
    if (cl4_ok(gx, D, stride_r, stride_k, stride_d) && ((uintptr_t)g % 16) == 0 && ((uintptr_t)arg % 4) == 0)
        PCNBR_TIMED("maxpool_bwd_cl4_kernel", s, wb, 0.0, (maxpool_bwd_cl4_kernel<<<pool_blocks(R * (D / 4)), 256, 0, s>>>(g, arg, R, K, D / 4, stride_r, stride_k, gx)));
    else
        PCNBR_TIMED("maxpool_bwd_generic_kernel", s, wb, 0.0, (maxpool_bwd_generic_kernel<<<pool_blocks(R * D), 256, 0, s>>>(g, arg, R, K, D, stride_r, stride_k, stride_d, gx)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
