// edgeconv.cu -- fused EdgeConv: conv1x1 + BatchNorm + LeakyReLU + max over k WITHOUT the k-inflated tensors.
//
// Reference: models/dgcnn/dgcnn.py:73-76 builds the (B,2F,N,k) edge tensor (671 MB at B=16), runs a 1x1
// convolution to (B,O,N,k) (335-671 MB), BatchNorm, LeakyReLU and a max over k: ~4 GB of HBM traffic per
// layer forward and ~3x that backward.  Algebra:
//     W . [x_j - x_i ; x_i] = A x_j + (B - A) x_i =: P[j] + Q[i]            (W = [A | B], two per-point GEMMs)
//     max_j act(bn(u_j)) = act(bn(max_j u_j))   when gamma >= 0  (min_j when gamma < 0):  BN+LeakyReLU is
//     monotone per channel, and fl(P_j + Q_i) is monotone in P_j, so max_j (P_j + Q_i) = (max_j P_j) + Q_i exactly.
// So one gather pass over the kNN table gives, per (point, channel): the selected P (max or min over the k
// neighbours) with its argmax, sum_j P_j (for the backward), and the BatchNorm batch statistics of ALL N*k
// pre-activations u = P_j + Q_i (shifted sums, per-block partials, combined in fp64 in a fixed order).
// HBM traffic per layer: the (B,N,k) table + a few (B,N,O) tensors (~100 MB instead of ~4 GB).
// The convolution itself stays a library GEMM on (B*N, F) x (F, 2O).  Rounding differs from the reference's
// order of operations (within 1e-4, tests/test_gpu_fused.py); get_graph_feature/EdgeConv keep an exact path.
#include "common.cuh"
#include "segsum.cuh"

namespace pcnbr {

// A lane owns VEC = O/32 CONSECUTIVE channels, so one neighbour row is one 128/256/512/1024-byte warp request.
template <int VEC>
__device__ __forceinline__ void ec_ld(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        v[0] = *p;
    } else if constexpr (VEC == 2) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}
template <int VEC>
__device__ __forceinline__ void ec_st(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 1) {
        *p = v[0];
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i)
            *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
}
template <int VEC>
__device__ __forceinline__ void ec_ld_u8(const uint8_t* __restrict__ p, int (&v)[VEC]) {
    if constexpr (VEC == 1) {
        v[0] = *p;
    } else if constexpr (VEC == 2) {
        const uchar2 t = *reinterpret_cast<const uchar2*>(p);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i) {
            const uchar4 t = *reinterpret_cast<const uchar4*>(p + 4 * i);
            v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
        }
    }
}
template <int VEC>
__device__ __forceinline__ void ec_st_u8(uint8_t* __restrict__ p, const int (&v)[VEC]) {
    if constexpr (VEC == 1) {
        *p = (uint8_t)v[0];
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<uchar2*>(p) = make_uchar2((uint8_t)v[0], (uint8_t)v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < VEC / 4; ++i)
            *reinterpret_cast<uchar4*>(p + 4 * i) = make_uchar4((uint8_t)v[4 * i], (uint8_t)v[4 * i + 1], (uint8_t)v[4 * i + 2], (uint8_t)v[4 * i + 3]);
    }
}

// PQ (B,N,2O): P = [..., :O], Q = [..., O:].  idx (B,N,K).  selmax[o] != 0: keep max_j P_j, else min_j.
// shift[o]: any constant near the typical u (variance is accumulated about it).  VEC = O / 32.
template <int VEC>
__global__ void __launch_bounds__(256)
edgeconv_fwd_kernel(const float* __restrict__ PQ, const int32_t* __restrict__ idx, const uint8_t* __restrict__ selmax,
                    const float* __restrict__ shift, int N, int K, float* __restrict__ psel, uint8_t* __restrict__ arg,
                    float* __restrict__ s1, float* __restrict__ partial) {
    constexpr int O = 32 * VEC;
    __shared__ float red[8][2 * O];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = lane * VEC;
    const float* __restrict__ pq = PQ + (size_t)b * N * 2 * O;
    const int32_t* __restrict__ ib = idx + (size_t)b * N * K;
    float c[VEC], a1[VEC], a2[VEC];
    int smax_i[VEC];
    ec_ld<VEC>(shift + c0, c);
    ec_ld_u8<VEC>(selmax + c0, smax_i);
    float sg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { a1[v] = 0.f; a2[v] = 0.f; sg[v] = smax_i[v] ? 1.f : -1.f; }
    for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
        float q[VEC], best[VEC], sum[VEC];
        int ba[VEC];
        ec_ld<VEC>(pq + (size_t)n * 2 * O + O + c0, q);
        // max and min share one branch-free path: t = +-p (exact), keep the first strict maximum of t.  (The per-element
        // `if (take)` of the first version compiled to a BSSY / BRA / BSYNC group per channel and row: 29 % of the kernel's
        // instructions were control flow, ncu source page.)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { best[v] = __int_as_float(0xff800000); sum[v] = 0.f; ba[v] = 0; }
        auto take_row = [&](const float (&p)[VEC], int j) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float d = (p[v] + q[v]) - c[v];
                a1[v] += d;
                a2[v] = fmaf(d, d, a2[v]);
                sum[v] += p[v];
                const float t = p[v] * sg[v];
                const bool take = t > best[v];
                best[v] = take ? t : best[v];
                ba[v] = take ? j : ba[v];
            }
        };
        for (int j0 = 0; j0 < K; j0 += 32) {
            const int mine = (j0 + lane < K) ? ib[(size_t)n * K + j0 + lane] : 0;
            const int cnt = min(32, K - j0);
            int l = 0;
            for (; l + 4 <= cnt; l += 4) {                      // four neighbour rows in flight
                float p0[VEC], p1[VEC], p2[VEC], p3[VEC];
                ec_ld<VEC>(pq + (size_t)__shfl_sync(PCNBR_FULL, mine, l) * 2 * O + c0, p0);
                ec_ld<VEC>(pq + (size_t)__shfl_sync(PCNBR_FULL, mine, l + 1) * 2 * O + c0, p1);
                ec_ld<VEC>(pq + (size_t)__shfl_sync(PCNBR_FULL, mine, l + 2) * 2 * O + c0, p2);
                ec_ld<VEC>(pq + (size_t)__shfl_sync(PCNBR_FULL, mine, l + 3) * 2 * O + c0, p3);
                take_row(p0, j0 + l); take_row(p1, j0 + l + 1); take_row(p2, j0 + l + 2); take_row(p3, j0 + l + 3);
            }
            for (; l < cnt; ++l) {
                float p0[VEC];
                ec_ld<VEC>(pq + (size_t)__shfl_sync(PCNBR_FULL, mine, l) * 2 * O + c0, p0);
                take_row(p0, j0 + l);
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) best[v] *= sg[v];
        const size_t o = ((size_t)b * N + n) * O + c0;
        ec_st<VEC>(psel + o, best);
        ec_st_u8<VEC>(arg + o, ba);
        ec_st<VEC>(s1 + o, sum);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) { red[warp][c0 + v] = a1[v]; red[warp][O + c0 + v] = a2[v]; }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * O; t += 256) {
        float s = red[0][t];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += red[w][t];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 2 * O + t] = s;
    }
}

// Backward: exact BatchNorm(train) + LeakyReLU + max backward in terms of (B,N,O) tensors.
//   u[n,j] = P[idx[n,j]] + Q[n];  y = gamma (u - mu) r + beta;  out[n] = act(y[n, arg[n]])
//   gs[n,o]  = dL/dy on the selected edge (g_out * act'(out))
//   dL/du[n,j] = gr gs [j == arg] - c1 - c2r (u[n,j] - mu)            gr = gamma r, c1 = gr sum(gs)/M, c2r = gr r sum(gs yhat)/M
//   dP[m] = sum over incoming edges (n,j) of m of dL/du[n,j]
//         = gr T1[m] - deg(m) c1 - c2r (deg(m) (P[m] - mu) + T2[m]),  T1 = sum_{incoming, arg[n]==j} gs[n],  T2 = sum_incoming Q[n]
//   dQ[n] = sum_j dL/du[n,j] = gr gs[n] - K c1 - c2r (s1[n] + K (Q[n] - mu))
// One warp per point m gathers its incoming edges through the CSR inverse (ascending position: deterministic,
// no atomics); segments longer than 64 edges (hub points) are split over the 8 warps of the CTA.
template <int VEC>
__device__ __forceinline__ void edgeconv_bwd_accumulate(const float* __restrict__ gs, const uint8_t* __restrict__ arg,
                                                        const float* __restrict__ PQ, const int32_t* __restrict__ pm,
                                                        size_t rowbase, int beg, int end, int K, int lane,
                                                        float (&t1)[VEC], float (&t2)[VEC]) {
    constexpr int O = 32 * VEC;
    const int c0 = lane * VEC;
    auto load = [&](int n, float (&g)[VEC], int (&av)[VEC], float (&qv)[VEC]) {
        const size_t r = rowbase + n;
        ec_ld<VEC>(gs + r * O + c0, g);
        ec_ld_u8<VEC>(arg + r * O + c0, av);
        ec_ld<VEC>(PQ + r * 2 * O + O + c0, qv);
    };
    auto add = [&](int j, const float (&g)[VEC], const int (&av)[VEC], const float (&qv)[VEC]) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            t1[v] += (av[v] == j) ? g[v] : 0.f;
            t2[v] += qv[v];
        }
    };
    for (int t0 = beg; t0 < end; t0 += 32) {
        const int e = (t0 + lane < end) ? pm[t0 + lane] : 0;
        const int nl = e / K, jl = e - nl * K;
        const int cnt = min(32, end - t0);
        int l = 0;
        for (; l + 4 <= cnt; l += 4) {                          // four incoming edges (12 row loads) in flight
            float g0[VEC], g1[VEC], g2[VEC], g3[VEC], q0[VEC], q1[VEC], q2[VEC], q3[VEC];
            int a0[VEC], a1[VEC], a2[VEC], a3[VEC];
            load(__shfl_sync(PCNBR_FULL, nl, l), g0, a0, q0);
            load(__shfl_sync(PCNBR_FULL, nl, l + 1), g1, a1, q1);
            load(__shfl_sync(PCNBR_FULL, nl, l + 2), g2, a2, q2);
            load(__shfl_sync(PCNBR_FULL, nl, l + 3), g3, a3, q3);
            add(__shfl_sync(PCNBR_FULL, jl, l), g0, a0, q0);
            add(__shfl_sync(PCNBR_FULL, jl, l + 1), g1, a1, q1);
            add(__shfl_sync(PCNBR_FULL, jl, l + 2), g2, a2, q2);
            add(__shfl_sync(PCNBR_FULL, jl, l + 3), g3, a3, q3);
        }
        for (; l < cnt; ++l) {
            float g0[VEC], q0[VEC];
            int a0[VEC];
            load(__shfl_sync(PCNBR_FULL, nl, l), g0, a0, q0);
            add(__shfl_sync(PCNBR_FULL, jl, l), g0, a0, q0);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(256)
edgeconv_bwd_kernel(const float* __restrict__ gs, const uint8_t* __restrict__ arg, const float* __restrict__ PQ,
                    const float* __restrict__ s1, const int32_t* __restrict__ offsets, const int32_t* __restrict__ perm,
                    const float* __restrict__ coef, int N, int K, float* __restrict__ dPQ) {
    constexpr int O = 32 * VEC;
    __shared__ int s_beg[8], s_len[8];
    __shared__ float s_part[8][2 * O];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = lane * VEC;
    const int32_t* o = offsets + (size_t)b * (N + 1);
    const int32_t* pm = perm + (size_t)b * N * K;
    const size_t rowbase = (size_t)b * N;
    float gr[VEC], c1[VEC], c2r[VEC], mu[VEC];
    ec_ld<VEC>(coef + c0, gr);
    ec_ld<VEC>(coef + O + c0, c1);
    ec_ld<VEC>(coef + 2 * O + c0, c2r);
    ec_ld<VEC>(coef + 3 * O + c0, mu);
    auto finish = [&](int m, int deg, const float (&t1)[VEC], const float (&t2)[VEC]) {
        const size_t r = rowbase + m;
        float P[VEC], Q[VEC], g[VEC], sv[VEC], dP[VEC], dQ[VEC];
        ec_ld<VEC>(PQ + r * 2 * O + c0, P);
        ec_ld<VEC>(PQ + r * 2 * O + O + c0, Q);
        ec_ld<VEC>(gs + r * O + c0, g);
        ec_ld<VEC>(s1 + r * O + c0, sv);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            dP[v] = gr[v] * t1[v] - (float)deg * c1[v] - c2r[v] * ((float)deg * (P[v] - mu[v]) + t2[v]);
            dQ[v] = gr[v] * g[v] - (float)K * c1[v] - c2r[v] * (sv[v] + (float)K * (Q[v] - mu[v]));
        }
        ec_st<VEC>(dPQ + r * 2 * O + c0, dP);
        ec_st<VEC>(dPQ + r * 2 * O + O + c0, dQ);
    };
    for (int m0 = blockIdx.x * 8; m0 < N; m0 += gridDim.x * 8) {
        const int m = m0 + warp;
        const int beg = (m < N) ? o[m] : 0;
        const int len = (m < N) ? o[m + 1] - beg : 0;
        if (lane == 0) { s_beg[warp] = beg; s_len[warp] = len; }
        if (m < N && len <= SEG_HEAVY) {
            float t1[VEC], t2[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { t1[v] = 0.f; t2[v] = 0.f; }
            edgeconv_bwd_accumulate<VEC>(gs, arg, PQ, pm, rowbase, beg, beg + len, K, lane, t1, t2);
            finish(m, len, t1, t2);
        }
        __syncthreads();
        for (int w = 0; w < 8; ++w) {
            const int hl = s_len[w];
            if (hl <= SEG_HEAVY) continue;
            const int hb = s_beg[w];
            const int piece = (hl + 7) / 8;
            const int pb = min(hb + warp * piece, hb + hl), pe = min(pb + piece, hb + hl);
            float t1[VEC], t2[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { t1[v] = 0.f; t2[v] = 0.f; }
            edgeconv_bwd_accumulate<VEC>(gs, arg, PQ, pm, rowbase, pb, pe, K, lane, t1, t2);
#pragma unroll
            for (int v = 0; v < VEC; ++v) { s_part[warp][c0 + v] = t1[v]; s_part[warp][O + c0 + v] = t2[v]; }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    t1[v] = s_part[0][c0 + v]; t2[v] = s_part[0][O + c0 + v];
#pragma unroll
                    for (int k = 1; k < 8; ++k) { t1[v] += s_part[k][c0 + v]; t2[v] += s_part[k][O + c0 + v]; }
                }
                finish(m0 + w, hl, t1, t2);
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_edgeconv_fwd_blocks(int N) {
    int gx = (N + 7) / 8;
    return gx > 74 ? 74 : gx;
}

extern "C" int pcnbr_edgeconv_fwd_f32(const float* PQ, const int32_t* idx, const uint8_t* selmax, const float* shift,
                                      int B, int N, int K, int O, float* psel, uint8_t* arg, float* s1, float* partial,
                                      pcnbr_stream_t stream) {
    if (!PQ || !idx || !selmax || !shift || !psel || !arg || !s1 || !partial || B <= 0 || N <= 0 || K <= 0) return PCNBR_E_BADARG;
    if (K > 255 || (O != 32 && O != 64 && O != 128 && O != 256)) return PCNBR_E_TOOLARGE;
    dim3 grid(pcnbr_edgeconv_fwd_blocks(N), B);
    cudaStream_t s = (cudaStream_t)stream;
    // algorithmic bytes per cloud: table 4 N K + PQ rows 8 N O (gathers hit L2) + psel, s1 (8 N O) + arg (N O)
    const double wb = (double)B * (4.0 * N * K + 17.0 * N * O), wf = 5.0 * B * (double)N * K * O;
    switch (O) {
        case 32:  PCNBR_TIMED("edgeconv_fwd_kernel", s, wb, wf, (edgeconv_fwd_kernel<1><<<grid, 256, 0, s>>>(PQ, idx, selmax, shift, N, K, psel, arg, s1, partial))); break;
        case 64:  PCNBR_TIMED("edgeconv_fwd_kernel", s, wb, wf, (edgeconv_fwd_kernel<2><<<grid, 256, 0, s>>>(PQ, idx, selmax, shift, N, K, psel, arg, s1, partial))); break;
        case 128: PCNBR_TIMED("edgeconv_fwd_kernel", s, wb, wf, (edgeconv_fwd_kernel<4><<<grid, 256, 0, s>>>(PQ, idx, selmax, shift, N, K, psel, arg, s1, partial))); break;
        default:  PCNBR_TIMED("edgeconv_fwd_kernel", s, wb, wf, (edgeconv_fwd_kernel<8><<<grid, 256, 0, s>>>(PQ, idx, selmax, shift, N, K, psel, arg, s1, partial))); break;
    }
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_edgeconv_bwd_f32(const float* gs, const uint8_t* arg, const float* PQ, const float* s1,
                                      const int32_t* offsets, const int32_t* perm, const float* coef, int B, int N,
                                      int K, int O, float* dPQ, pcnbr_stream_t stream) {
    if (!gs || !arg || !PQ || !s1 || !offsets || !perm || !coef || !dPQ || B <= 0 || N <= 0 || K <= 0) return PCNBR_E_BADARG;
    if (K > 255 || (O != 32 && O != 64 && O != 128 && O != 256)) return PCNBR_E_TOOLARGE;
    dim3 grid = segsum_grid(N, B);
    cudaStream_t s = (cudaStream_t)stream;
    // algorithmic bytes per cloud: perm 4 N K + offsets 4 N + gs, s1 (8 N O) + arg (N O) + PQ (8 N O) + dPQ (8 N O)
    const double wb = (double)B * (4.0 * N * K + 4.0 * N + 25.0 * N * O), wf = 3.0 * B * (double)N * K * O;
    switch (O) {
        case 32:  PCNBR_TIMED("edgeconv_bwd_kernel", s, wb, wf, (edgeconv_bwd_kernel<1><<<grid, 256, 0, s>>>(gs, arg, PQ, s1, offsets, perm, coef, N, K, dPQ))); break;
        case 64:  PCNBR_TIMED("edgeconv_bwd_kernel", s, wb, wf, (edgeconv_bwd_kernel<2><<<grid, 256, 0, s>>>(gs, arg, PQ, s1, offsets, perm, coef, N, K, dPQ))); break;
        case 128: PCNBR_TIMED("edgeconv_bwd_kernel", s, wb, wf, (edgeconv_bwd_kernel<4><<<grid, 256, 0, s>>>(gs, arg, PQ, s1, offsets, perm, coef, N, K, dPQ))); break;
        default:  PCNBR_TIMED("edgeconv_bwd_kernel", s, wb, wf, (edgeconv_bwd_kernel<8><<<grid, 256, 0, s>>>(gs, arg, PQ, s1, offsets, perm, coef, N, K, dPQ))); break;
    }
    PCNBR_CHECK_LAUNCH();
    return 0;
}
