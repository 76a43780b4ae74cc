// segsum.cuh -- K7: atomic-free segmented scatter-add, shared by the backward of K5/K8/K9.
//
// out[b, s, c] = sum_{t in segment(s)} src.term(b, e_t, c),   e_t = perm[t]  (ascending e)
//
// A CTA of 8 warps takes 8 consecutive source points.  Lanes stride over the columns, so every
// gathered row is read with coalesced 128-byte requests.  Short segments are summed by one warp;
// long ones (padded ball-query heavy hitters, duplicated points: up to M entries) are cut into 8
// contiguous pieces, one per warp, whose partial sums are combined in warp order -- the summation
// tree depends only on the segment length, never on scheduling: bitwise deterministic, no atomics.
#pragma once
#include "common.cuh"

namespace pcnbr {

constexpr int SEG_HEAVY = 64;      // segments longer than this are summed by the whole CTA
constexpr int SEG_COLS = 128;      // columns per pass (4 per lane)

// Accumulate rows perm[beg..end) into acc[4] for columns c0 + lane + 32*i < ncols.
// Src: row(b, e) -> pointer to the gathered row of position e, scale(b, e) -> its weight (1 for plain sums; the
// product is a single-rounding fma, so scale 1 adds exactly).  Row pointer and weight are formed once per position,
// four positions (16 loads per lane) are in flight before the first add.
template <class Src>
__device__ __forceinline__ void seg_accumulate(const Src& src, int b, const int32_t* __restrict__ pm, int beg,
                                               int end, int c0, int ncols, int lane, float acc[4]) {
    const int rem = ncols - c0;
    for (int t0 = beg; t0 < end; t0 += 32) {
        const int my_e = (t0 + lane < end) ? pm[t0 + lane] : 0;
        const int n = min(32, end - t0);
        int l = 0;
        for (; l + 4 <= n; l += 4) {
            const float* r[4];
            float w[4], v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = __shfl_sync(PCNBR_FULL, my_e, l + u);
                r[u] = src.row(b, e) + c0 + lane;
                w[u] = src.scale(b, e);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][i] = (lane + 32 * i < rem) ? r[u][32 * i] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = __fmaf_rn(w[u], v[u][i], acc[i]);
        }
        for (; l < n; ++l) {
            const int e = __shfl_sync(PCNBR_FULL, my_e, l);
            const float* r = src.row(b, e) + c0 + lane;
            const float w = src.scale(b, e);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (lane + 32 * i < rem) acc[i] = __fmaf_rn(w, r[32 * i], acc[i]);
        }
    }
}

// Src: row(b, e) / scale(b, e), see seg_accumulate.
// Dst: store(b, s, col, value).
template <class Src, class Dst>
__global__ void __launch_bounds__(256)
segsum_kernel(Src src, Dst dst, const int32_t* __restrict__ offsets, const int32_t* __restrict__ perm, int N,
              int E, int ncols) {
    __shared__ int s_beg[8], s_len[8];
    __shared__ float s_part[8][SEG_COLS];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* o = offsets + (size_t)b * (N + 1);
    const int32_t* pm = perm + (size_t)b * E;
    for (int s0 = blockIdx.x * 8; s0 < N; s0 += gridDim.x * 8) {
        const int s = s0 + warp;
        const int beg = (s < N) ? o[s] : 0;
        const int len = (s < N) ? o[s + 1] - beg : 0;
        if (lane == 0) { s_beg[warp] = beg; s_len[warp] = len; }
        if (s < N && len <= SEG_HEAVY) {
            for (int c0 = 0; c0 < ncols; c0 += SEG_COLS) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                seg_accumulate(src, b, pm, beg, beg + len, c0, ncols, lane, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (lane + 32 * i < ncols - c0) dst.store(b, s, c0 + lane + 32 * i, acc[i]);
            }
        }
        __syncthreads();
        for (int w = 0; w < 8; ++w) {
            const int hl = s_len[w];
            if (hl <= SEG_HEAVY) continue;                 // uniform over the CTA
            const int hb = s_beg[w];
            const int piece = (hl + 7) / 8;
            const int pb = min(hb + warp * piece, hb + hl), pe = min(pb + piece, hb + hl);
            for (int c0 = 0; c0 < ncols; c0 += SEG_COLS) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                seg_accumulate(src, b, pm, pb, pe, c0, ncols, lane, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i) s_part[warp][lane + 32 * i] = acc[i];
                __syncthreads();
                if (warp == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = lane + 32 * i;
                        if (c < ncols - c0) {
                            float t = s_part[0][c];
#pragma unroll
                            for (int k = 1; k < 8; ++k) t = __fadd_rn(t, s_part[k][c]);
                            dst.store(b, s0 + w, c0 + c, t);
                        }
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
}

static inline dim3 segsum_grid(int N, int B) {
    int gx = (N + 7) / 8;
    if (gx > 148 * 8) gx = 148 * 8;
    return dim3(gx, B);
}

}  // namespace pcnbr
