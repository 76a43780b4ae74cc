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

// ---- exact division by a shared divisor ---------------------------------------------------------------------------
// The reference divides every weighted feature by the point's norm (common.py:122: points * weights / norm): N*k*D true
// divisions.  div.rn.f32 costs ~15 SASS instructions (MUFU.RCP, FCHK, 5 FFMA, a guarded call to the slow path): they
// were most of interp3_fwd_kernel's instructions.  All D*k quotients of a point share ONE divisor, so its correctly
// rounded reciprocal y = RN(1/n) is formed once and each quotient takes the Markstein sequence
//     q0 = RN(a y);  r0 = a - n q0 (exact, FMA);  q1 = RN(q0 + r0 y);  r1 = a - n q1 (exact);  q = RN(q1 + r1 y)
// which returns the correctly rounded a / n (q1 is faithful, and a faithful quotient corrected once with the correctly
// rounded reciprocal rounds correctly unless the significand of n is all ones -- Markstein 1990; Muller et al., Handbook
// of Floating-Point Arithmetic, 2nd ed., sec. 4.7) as long as nothing under/overflows: n in [2^-60, 2^60] and a = 0 or
// |a| in [2^-100, 2^100] keep q, r0 and r1 normal and the residuals exact.  Anything else takes div.rn.f32.
struct SharedDiv {
    float n, y;
    bool ok;
};
__device__ __forceinline__ SharedDiv shared_div(float n) {
    SharedDiv d;
    d.n = n;
    d.y = __frcp_rn(n);
    const uint32_t bits = __float_as_uint(n);
    d.ok = (n >= 8.6736174e-19f) && (n <= 1.1529215e18f) && ((bits & 0x7fffffu) != 0x7fffffu);
    return d;
}
// a within the exact range of the fast sequence (0, or 2^-100 <= |a| <= 2^100): x = |a|'s bit pattern - 1 wraps to the
// top for a = 0, so one unsigned range test covers "zero or at least 2^-100", a second one "at most 2^100 (and finite)"
__device__ __forceinline__ bool shared_div_safe4(const float4& t) {
    constexpr uint32_t LO = 0x0d800000u, HI = 0x71800000u;               // 2^-100, 2^100
    const uint32_t x = (__float_as_uint(t.x) & 0x7fffffffu) - 1u, y = (__float_as_uint(t.y) & 0x7fffffffu) - 1u;
    const uint32_t z = (__float_as_uint(t.z) & 0x7fffffffu) - 1u, w = (__float_as_uint(t.w) & 0x7fffffffu) - 1u;
    const uint32_t lo = min(min(x, y), min(z, w));                       // smallest: must not fall in [0, LO - 1)
    // largest finite magnitude: zeros (0xffffffff after the decrement) are taken out of the maximum first
    const uint32_t hx = x + 1u, hy = y + 1u, hz = z + 1u, hw = w + 1u;
    const uint32_t hi = max(max(hx, hy), max(hz, hw));
    return lo >= LO - 1u && hi <= HI;
}
__device__ __forceinline__ float shared_div_fast(float a, const SharedDiv& d) {
    float q = __fmul_rn(a, d.y);
    float r = __fmaf_rn(-d.n, q, a);
    q = __fmaf_rn(r, d.y, q);
    r = __fmaf_rn(-d.n, q, a);
    return __fmaf_rn(r, d.y, q);
}
__device__ __noinline__ float4 shared_div4_slow(float4 t, float n) {
    return make_float4(__fdiv_rn(t.x, n), __fdiv_rn(t.y, n), __fdiv_rn(t.z, n), __fdiv_rn(t.w, n));
}
// (t.x, t.y, t.z, t.w) / d.n, each correctly rounded
__device__ __forceinline__ float4 shared_div4(const float4& t, const SharedDiv& d) {
    if (d.ok && shared_div_safe4(t))
        return make_float4(shared_div_fast(t.x, d), shared_div_fast(t.y, d), shared_div_fast(t.z, d), shared_div_fast(t.w, d));
    return shared_div4_slow(t, d.n);
}

// k = 3 (the only value the models use), D % 4 == 0: one warp takes PTS consecutive fine points per step -- the first
// 3*PTS lanes fetch the (index, distance) pairs in one coalesced request each, the weights and norms are formed once, and
// the 3*PTS coarse rows are in flight as float4 requests (512 B per row and warp) before the first quotient.  Same
// arithmetic order as interp_fwd_kernel, element for element.  PTS = 2 keeps the kernel at <= 80 registers (3 CTAs per
// SM): the kernel is a chain of dependent latencies (table -> rows -> quotients -> store), hidden by resident warps.
template <int PTS>
__global__ void __launch_bounds__(256, 3)
interp3_fwd_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, const float* __restrict__ d2,
                   int N, int M, int D, float* __restrict__ out, float* __restrict__ coef) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    const float* __restrict__ fb = feat + (size_t)b * M * D;
    const int lp = lane / 3;                                             // this lane's point slot, lane < 3 * PTS
    for (int n0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PTS; n0 < N; n0 += nw * PTS) {
        const size_t base = ((size_t)b * N + n0) * 3;
        const bool slot = lane < 3 * PTS && n0 + lp < N;
        const int my_id = slot ? idx[base + lane] : 0;
        const float my_w = slot ? __fdiv_rn(1.0f, __fadd_rn(d2[base + lane], 1e-9f)) : 1.0f;     // common.py:119
        int id[PTS][3];
        float w[PTS][3];
        SharedDiv dv[PTS];
#pragma unroll
        for (int p = 0; p < PTS; ++p) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                id[p][k] = __shfl_sync(PCNBR_FULL, my_id, 3 * p + k);
                w[p][k] = __shfl_sync(PCNBR_FULL, my_w, 3 * p + k);
            }
            dv[p] = shared_div(__fadd_rn(__fadd_rn(w[p][0], w[p][1]), w[p][2]));  // common.py:120
        }
        if (coef && slot) {
            float nm = dv[0].n;
#pragma unroll
            for (int p = 1; p < PTS; ++p) if (lp == p) nm = dv[p].n;
            coef[base + lane] = __fdiv_rn(my_w, nm);
        }
        for (int c = lane * 4; c < D; c += 128) {
            float4 f[PTS][3];
#pragma unroll
            for (int p = 0; p < PTS; ++p)
#pragma unroll
                for (int k = 0; k < 3; ++k) f[p][k] = *reinterpret_cast<const float4*>(fb + (size_t)id[p][k] * D + c);
#pragma unroll
            for (int p = 0; p < PTS; ++p) {
                if (n0 + p >= N) break;
                float4 acc;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 t = shared_div4(make_float4(__fmul_rn(f[p][k].x, w[p][k]), __fmul_rn(f[p][k].y, w[p][k]),
                                                             __fmul_rn(f[p][k].z, w[p][k]), __fmul_rn(f[p][k].w, w[p][k])), dv[p]);
                    if (k == 0) acc = t;
                    else { acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y); acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w); }
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
        int g3 = (N + 15) / 16;                                           // 8 warps x 2 points per CTA step
        if (g3 > 148 * 8) g3 = 148 * 8;
        PCNBR_TIMED("interp_fwd_kernel", (cudaStream_t)stream, wb, wf,
                    (interp3_fwd_kernel<2><<<dim3(g3, B), 256, 0, (cudaStream_t)stream>>>(feat, idx, d2, N, M, D, out, coef)));
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
