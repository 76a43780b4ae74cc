// bnact.cu -- BatchNorm (training or eval) + LeakyReLU/ReLU over the rows of a (R, C) point-major matrix,
// forward and backward, in the minimum number of passes over HBM.
//
// Reference: every "Conv -> BatchNorm -> ReLU" of models/utils/common.py:125-178 (MiniPointNet / UnitPointNet) and
// "Conv -> BatchNorm -> LeakyReLU(0.2)" of models/dgcnn/dgcnn.py:66-71,95-126.  The reference runs them as three
// library ops per direction (statistics, normalise, activation; activation', reduce, input gradient): 6 passes
// over the activation tensor forward and 8 backward.  Here:
//   forward : bn_stats_kernel (1 read)  -> bn_finalize_kernel (tiny) -> bn_act_fwd_kernel (1 read, 1 write)
//   backward: bn_act_bwd_reduce_kernel (2 reads) -> bn_bwd_finalize_kernel (tiny) -> bn_act_bwd_apply_kernel (2 reads, 1 write)
// The activation output is never needed by the backward: its sign is recomputed from the saved pre-BatchNorm rows.
// The row source may be the sum of two strided matrices (x = a[r*lda + c] + b[r*ldb + c]): the fused EdgeConv path
// (edgeconv.cu) normalises psel + Q without materialising it, and takes its statistics partials from its own
// gather kernel.  Statistics: per-block sums of (x - shift) and (x - shift)^2 about a per-channel shift (row 0),
// combined over the blocks in fp64 in a fixed order -- deterministic, no atomics.
//
// Thread mapping: C % 4 == 0 and CV = C/4 a power of two <= 512.  A thread owns max(1, CV/256) float4 columns and
// walks rows; consecutive threads read consecutive 16-byte words, so every warp request is a full 512-byte span.
#include "common.cuh"

namespace pcnbr {

constexpr int BA_T = 256;

struct BaMap {
    int cvb;    // float4 columns covered by one sweep of the block (min(CV, 256))
    int ncol;   // float4 columns per thread (1 or 2)
    int ty;     // row lane of this thread
    int TY;     // rows per sweep
    int col;    // first float (not float4) column of this thread
};
__device__ __forceinline__ BaMap ba_map(int C) {
    BaMap m;
    const int CV = C >> 2;
    m.cvb = CV < BA_T ? CV : BA_T;
    m.ncol = CV > BA_T ? CV / BA_T : 1;
    m.ty = threadIdx.x / m.cvb;
    m.TY = BA_T / m.cvb;
    m.col = (threadIdx.x % m.cvb) * 4;
    return m;
}
__device__ __forceinline__ float4 ba_ld(const float* __restrict__ a, long lda, const float* __restrict__ b, long ldb, long r, int c) {
    float4 v = *reinterpret_cast<const float4*>(a + r * lda + c);
    if (b) {
        const float4 w = *reinterpret_cast<const float4*>(b + r * ldb + c);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    return v;
}

// block tree reduction over the row lanes of (s1, s2) per column, result written by row lane 0 to partial[blk][2][C].
// pivot != nullptr: the sums were taken about pivot[j] (per column) and are moved to the common pivot x0[column] on
// the way out: s1 + n d, s2 + 2 d s1 + n d^2, d = pivot - x0.
__device__ __forceinline__ void ba_rebase(float& a, float& q, float piv, float x0, float n) {
    // fp32 FMAs: each step rounds at 2^-24 of n d^2, the same order as storing the result as a float (fp64 here cost 30 us
    // per PointNet++ step on the B200's fp64 pipe for nothing)
    const float d = piv - x0, nd = n * d, s1 = a;
    a = s1 + nd;
    q = fmaf(nd, d, fmaf(2.0f * d, s1, q));
}
template <int NCOL>
__device__ __forceinline__ void ba_block_reduce(const BaMap& m, float4 (&s1)[NCOL], float4 (&s2)[NCOL], int C, float* __restrict__ partial,
                                                const float4* pivot = nullptr, const float* __restrict__ x0 = nullptr, float nrows = 0.f) {
    __shared__ float4 red[2 * BA_T];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        red[threadIdx.x] = s1[j];
        red[BA_T + threadIdx.x] = s2[j];
        __syncthreads();
        for (int s = m.TY >> 1; s > 0; s >>= 1) {
            if (m.ty < s) {
                const float4 u = red[threadIdx.x + s * m.cvb], w = red[BA_T + threadIdx.x + s * m.cvb];
                float4& p = red[threadIdx.x];
                float4& q = red[BA_T + threadIdx.x];
                p.x += u.x; p.y += u.y; p.z += u.z; p.w += u.w;
                q.x += w.x; q.y += w.y; q.z += w.z; q.w += w.w;
            }
            __syncthreads();
        }
        if (m.ty == 0) {
            float4 a = red[threadIdx.x], q = red[BA_T + threadIdx.x];
            if (pivot) {
                const float4 z = *reinterpret_cast<const float4*>(x0 + m.col + j * 4 * BA_T);
                ba_rebase(a.x, q.x, pivot[j].x, z.x, nrows); ba_rebase(a.y, q.y, pivot[j].y, z.y, nrows);
                ba_rebase(a.z, q.z, pivot[j].z, z.z, nrows); ba_rebase(a.w, q.w, pivot[j].w, z.w, nrows);
            }
            float* dst = partial + (size_t)blockIdx.x * 2 * C + m.col + j * 4 * BA_T;
            *reinterpret_cast<float4*>(dst) = a;
            *reinterpret_cast<float4*>(dst + C) = q;
        }
        __syncthreads();
    }
}

// partial[blk] = { sum_r (x - shift), sum_r (x - shift)^2 } over the block's rows; shift = row 0 of x.
// Every block ACCUMULATES about its own first row (a pivot inside its own data) and moves the two sums to the common pivot
// once when it writes them: s1' = s1 + n d, s2' = s2 + 2 d s1 + n d^2 with d = pivot_block - x[0].  If row 0 of the
// tensor is an outlier (|x0 - mean| = D sigma), accumulating about it loses ~rows-per-block * 2^-24 * D^2 of the variance
// in the fp32 running sums; moved afterwards, only the final rounding of the partial (2^-24 * D^2) remains.
template <int NCOL>
__global__ void __launch_bounds__(BA_T)
bn_stats_kernel(const float* __restrict__ x, long R, int C, int rows_per_block, float* __restrict__ partial) {
    const BaMap m = ba_map(C);
    float4 sh[NCOL], s1[NCOL], s2[NCOL];
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        sh[j] = *reinterpret_cast<const float4*>(x + (r0 < R ? r0 : 0) * C + m.col + j * 4 * BA_T);      // the block's own pivot
        s1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        s2[j] = s1[j];
    }
#pragma unroll 4
    for (long r = r0 + m.ty; r < r1; r += m.TY) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(x + r * C + m.col + j * 4 * BA_T);
            const float dx = v.x - sh[j].x, dy = v.y - sh[j].y, dz = v.z - sh[j].z, dw = v.w - sh[j].w;
            s1[j].x += dx; s1[j].y += dy; s1[j].z += dz; s1[j].w += dw;
            s2[j].x = fmaf(dx, dx, s2[j].x); s2[j].y = fmaf(dy, dy, s2[j].y);
            s2[j].z = fmaf(dz, dz, s2[j].z); s2[j].w = fmaf(dw, dw, s2[j].w);
        }
    }
    const float nrows = r1 > r0 ? (float)(r1 - r0) : 0.f;
    ba_block_reduce<NCOL>(m, s1, s2, C, partial, sh, x, nrows);
}

// Combine the block partials (fp64, ascending block order) into the BatchNorm constants
//   stats (4,C) = { mean, rstd, gamma*rstd, beta }      and update the running statistics (momentum, unbiased variance).
// nblk == 0: eval mode, mean/var come from running_mean / running_var.
// Finalize kernels: a CTA of 1024 threads owns BA_FC = 8 channels and cuts the partial rows into BA_FS = 128 slices
// (few channels per CTA = many CTAs and a short dependent-load chain per thread: the kernel is pure latency).
constexpr int BA_FC = 8, BA_FS = 128;

// sum over the partial rows of partial[:, 0, c] and partial[:, 1, c] in fp64; fixed order (slice-major tree); valid in
// the threads with slice 0
__device__ __forceinline__ void ba_finalize_sums(const float* __restrict__ partial, int nblk, int C, int c, int sy, int cx,
                                                 double& s1, double& s2) {
    __shared__ double a1[BA_FS][BA_FC], a2[BA_FS][BA_FC];
    s1 = 0.0; s2 = 0.0;
    if (c < C) {
        const int per = (nblk + BA_FS - 1) / BA_FS;
        const int b0 = sy * per, b1 = min(nblk, b0 + per);
#pragma unroll 4
        for (int b = b0; b < b1; ++b) {
            s1 += (double)partial[(size_t)b * 2 * C + c];
            s2 += (double)partial[(size_t)b * 2 * C + C + c];
        }
    }
    a1[sy][cx] = s1; a2[sy][cx] = s2;
    __syncthreads();
    if (sy < 8) {                                             // 8 x 16 slices
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { t1 += a1[sy * 16 + k][cx]; t2 += a2[sy * 16 + k][cx]; }
        __syncthreads();
        a1[sy][cx] = t1; a2[sy][cx] = t2;
    } else {
        __syncthreads();
    }
    __syncthreads();
    if (sy == 0) {
        s1 = 0.0; s2 = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s1 += a1[k][cx]; s2 += a2[k][cx]; }
    }
}

__global__ void __launch_bounds__(BA_FC * BA_FS)
bn_finalize_kernel(const float* __restrict__ partial, int nblk, const float* __restrict__ shift, double count, int C,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ stats) {
    const int cx = threadIdx.x % BA_FC, sy = threadIdx.x / BA_FC;
    const int c = blockIdx.x * BA_FC + cx;
    double s1, s2;
    ba_finalize_sums(partial, nblk, C, c, sy, cx, s1, s2);
    if (sy != 0 || c >= C) return;
    double mean, var;
    if (nblk > 0) {
        const double m1 = s1 / count;
        mean = (double)shift[c] + m1;
        var = s2 / count - m1 * m1;
        if (var < 0.0) var = 0.0;
        if (running_mean) {
            const double unb = count > 1.0 ? var * (count / (count - 1.0)) : var;
            running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
            running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unb);
        }
    } else {
        mean = (double)running_mean[c];
        var = (double)running_var[c];
    }
    const float meanf = (float)mean;
    const float rstd = (float)(1.0 / sqrt((double)(float)var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f;
    stats[c] = meanf;
    stats[C + c] = rstd;
    stats[2 * C + c] = g * rstd;
    stats[3 * C + c] = beta ? beta[c] : 0.f;
}

// Per-block maximum of |value written| for the fp16-split GEMM that consumes this kernel's output (gemm_h2.cu): slot
// blockIdx.x of `amax` (PCNBR_AMAX_SLOTS floats) takes the block's maximum, the slots no block owns are zeroed -- the
// consumer reduces all slots, so no separate absmax pass over the tensor is needed.
constexpr int BA_AMAX_SLOTS = 1280;
__device__ __forceinline__ void ba_store_amax(float mx, float* __restrict__ amax) {
    __shared__ float s_mx[BA_T / 32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PCNBR_FULL, mx, d));
    if ((threadIdx.x & 31) == 0) s_mx[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < BA_T / 32; ++w) mx = fmaxf(mx, s_mx[w]);
        amax[blockIdx.x] = mx;
        for (int i = blockIdx.x + gridDim.x; i < BA_AMAX_SLOTS; i += gridDim.x) amax[i] = 0.f;
    }
}
__device__ __forceinline__ float ba_amax4(float m, const float4& o) {
    return fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
}

__device__ __forceinline__ float ba_act(float x, float mean, float scale, float beta, float slope) {
    const float y = fmaf(x - mean, scale, beta);
    return y > 0.f ? y : y * slope;
}

// Dropout fused behind the activation (models/dgcnn/dgcnn.py:117,122: BatchNorm -> LeakyReLU -> Dropout): element i is
// kept (and scaled by 1/(1-p)) iff a counter-based hash of (seed, i) falls above p.  The seed is a device word drawn by
// the host layer from torch's generator for every forward pass; the backward kernels recompute the same mask from it,
// so no mask tensor is stored or re-read.  keep4 returns the four multipliers of one float4.
__device__ __forceinline__ float ba_keep(unsigned long long seed, unsigned long long i, float p, float scale) {
    unsigned long long z = seed + (i + 1ull) * 0x9E3779B97F4A7C15ull;       // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f) >= p ? scale : 0.f;
}
__device__ __forceinline__ float4 ba_keep4(const unsigned long long* __restrict__ seedp, long r, int C, int c, float p) {
    if (!seedp) return make_float4(1.f, 1.f, 1.f, 1.f);
    const unsigned long long seed = *seedp, i = (unsigned long long)r * (unsigned long long)C + (unsigned long long)c;
    const float scale = 1.0f / (1.0f - p);
    return make_float4(ba_keep(seed, i, p, scale), ba_keep(seed, i + 1, p, scale), ba_keep(seed, i + 2, p, scale),
                       ba_keep(seed, i + 3, p, scale));
}

// y[r, c] = act((x[r, c] - mean) * gamma * rstd + beta)  (* dropout multiplier)
template <int NCOL>
__global__ void __launch_bounds__(BA_T)
bn_act_fwd_kernel(const float* __restrict__ a, long lda, const float* __restrict__ b, long ldb, long R, int C,
                  const float* __restrict__ stats, float slope, float* __restrict__ y,
                  const unsigned long long* __restrict__ drop_seed, float drop_p, float* __restrict__ amax) {
    const BaMap m = ba_map(C);
    float mx = 0.f;
    float4 mu[NCOL], sc[NCOL], be[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int c = m.col + j * 4 * BA_T;
        mu[j] = *reinterpret_cast<const float4*>(stats + c);
        sc[j] = *reinterpret_cast<const float4*>(stats + 2 * C + c);
        be[j] = *reinterpret_cast<const float4*>(stats + 3 * C + c);
    }
#pragma unroll 4
    for (long r = (long)blockIdx.x * m.TY + m.ty; r < R; r += (long)gridDim.x * m.TY) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int c = m.col + j * 4 * BA_T;
            const float4 v = ba_ld(a, lda, b, ldb, r, c);
            float4 o;
            o.x = ba_act(v.x, mu[j].x, sc[j].x, be[j].x, slope);
            o.y = ba_act(v.y, mu[j].y, sc[j].y, be[j].y, slope);
            o.z = ba_act(v.z, mu[j].z, sc[j].z, be[j].z, slope);
            o.w = ba_act(v.w, mu[j].w, sc[j].w, be[j].w, slope);
            if (drop_seed) {
                const float4 k = ba_keep4(drop_seed, r, C, c, drop_p);
                o.x *= k.x; o.y *= k.y; o.z *= k.z; o.w *= k.w;
            }
            mx = ba_amax4(mx, o);
            *reinterpret_cast<float4*>(y + r * C + c) = o;
        }
    }
    if (amax) ba_store_amax(mx, amax);
}

// g' = gy * act'(pre);  partial[blk] = { sum_r g', sum_r g' * xhat };  optionally gs = g' is written out.
template <int NCOL>
__global__ void __launch_bounds__(BA_T)
bn_act_bwd_reduce_kernel(const float* __restrict__ gy, const float* __restrict__ a, long lda, const float* __restrict__ b,
                         long ldb, long R, int C, int rows_per_block, const float* __restrict__ stats, float slope,
                         float* __restrict__ partial, float* __restrict__ gs,
                         const unsigned long long* __restrict__ drop_seed, float drop_p) {
    const BaMap m = ba_map(C);
    float4 mu[NCOL], rs[NCOL], sc[NCOL], be[NCOL], s1[NCOL], s2[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int c = m.col + j * 4 * BA_T;
        mu[j] = *reinterpret_cast<const float4*>(stats + c);
        rs[j] = *reinterpret_cast<const float4*>(stats + C + c);
        sc[j] = *reinterpret_cast<const float4*>(stats + 2 * C + c);
        be[j] = *reinterpret_cast<const float4*>(stats + 3 * C + c);
        s1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        s2[j] = s1[j];
    }
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
#pragma unroll 4
    for (long r = r0 + m.ty; r < r1; r += m.TY) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int c = m.col + j * 4 * BA_T;
            const float4 v = ba_ld(a, lda, b, ldb, r, c);
            float4 g = *reinterpret_cast<const float4*>(gy + r * C + c);
            if (drop_seed) {
                const float4 k = ba_keep4(drop_seed, r, C, c, drop_p);
                g.x *= k.x; g.y *= k.y; g.z *= k.z; g.w *= k.w;
            }
            const float dx = v.x - mu[j].x, dy = v.y - mu[j].y, dz = v.z - mu[j].z, dw = v.w - mu[j].w;
            g.x = fmaf(dx, sc[j].x, be[j].x) > 0.f ? g.x : g.x * slope;
            g.y = fmaf(dy, sc[j].y, be[j].y) > 0.f ? g.y : g.y * slope;
            g.z = fmaf(dz, sc[j].z, be[j].z) > 0.f ? g.z : g.z * slope;
            g.w = fmaf(dw, sc[j].w, be[j].w) > 0.f ? g.w : g.w * slope;
            if (gs) *reinterpret_cast<float4*>(gs + r * C + c) = g;
            s1[j].x += g.x; s1[j].y += g.y; s1[j].z += g.z; s1[j].w += g.w;
            s2[j].x = fmaf(g.x, dx * rs[j].x, s2[j].x); s2[j].y = fmaf(g.y, dy * rs[j].y, s2[j].y);
            s2[j].z = fmaf(g.z, dz * rs[j].z, s2[j].z); s2[j].w = fmaf(g.w, dw * rs[j].w, s2[j].w);
        }
    }
    ba_block_reduce<NCOL>(m, s1, s2, C, partial);
}

// dbeta = sum g', dgamma = sum g' xhat (fp64 over the blocks, fixed order);
// coef (4,C) = { gr = gamma*rstd, gr*dbeta/count, gr*rstd*dgamma/count, mean }   (the two middle rows are 0 in eval mode)
__global__ void __launch_bounds__(BA_FC * BA_FS)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, const float* __restrict__ stats, double count, int C,
                       int training, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
    const int cx = threadIdx.x % BA_FC, sy = threadIdx.x / BA_FC;
    const int c = blockIdx.x * BA_FC + cx;
    double s1, s2;
    ba_finalize_sums(partial, nblk, C, c, sy, cx, s1, s2);
    if (sy != 0 || c >= C) return;
    dbeta[c] = (float)s1;
    dgamma[c] = (float)s2;
    const double gr = (double)stats[2 * C + c], rstd = (double)stats[C + c];
    coef[c] = (float)gr;
    coef[C + c] = training ? (float)(gr * s1 / count) : 0.f;
    coef[2 * C + c] = training ? (float)(gr * rstd * s2 / count) : 0.f;
    coef[3 * C + c] = stats[c];
}

// dx = gr g' - c1 - c2r (x - mean)
template <int NCOL>
__global__ void __launch_bounds__(BA_T)
bn_act_bwd_apply_kernel(const float* __restrict__ gy, const float* __restrict__ x, long R, int C,
                        const float* __restrict__ stats, const float* __restrict__ coef, float slope, float* __restrict__ dx,
                        const unsigned long long* __restrict__ drop_seed, float drop_p, float* __restrict__ amax) {
    const BaMap m = ba_map(C);
    float mx = 0.f;
    float4 mu[NCOL], sc[NCOL], be[NCOL], gr[NCOL], c1[NCOL], c2[NCOL];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
        const int c = m.col + j * 4 * BA_T;
        mu[j] = *reinterpret_cast<const float4*>(stats + c);
        sc[j] = *reinterpret_cast<const float4*>(stats + 2 * C + c);
        be[j] = *reinterpret_cast<const float4*>(stats + 3 * C + c);
        gr[j] = *reinterpret_cast<const float4*>(coef + c);
        c1[j] = *reinterpret_cast<const float4*>(coef + C + c);
        c2[j] = *reinterpret_cast<const float4*>(coef + 2 * C + c);
    }
#pragma unroll 4
    for (long r = (long)blockIdx.x * m.TY + m.ty; r < R; r += (long)gridDim.x * m.TY) {
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            const int c = m.col + j * 4 * BA_T;
            const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
            float4 g = *reinterpret_cast<const float4*>(gy + r * C + c);
            if (drop_seed) {
                const float4 k = ba_keep4(drop_seed, r, C, c, drop_p);
                g.x *= k.x; g.y *= k.y; g.z *= k.z; g.w *= k.w;
            }
            const float ex = v.x - mu[j].x, ey = v.y - mu[j].y, ez = v.z - mu[j].z, ew = v.w - mu[j].w;
            g.x = fmaf(ex, sc[j].x, be[j].x) > 0.f ? g.x : g.x * slope;
            g.y = fmaf(ey, sc[j].y, be[j].y) > 0.f ? g.y : g.y * slope;
            g.z = fmaf(ez, sc[j].z, be[j].z) > 0.f ? g.z : g.z * slope;
            g.w = fmaf(ew, sc[j].w, be[j].w) > 0.f ? g.w : g.w * slope;
            float4 o;
            o.x = fmaf(gr[j].x, g.x, -c1[j].x) - c2[j].x * ex;
            o.y = fmaf(gr[j].y, g.y, -c1[j].y) - c2[j].y * ey;
            o.z = fmaf(gr[j].z, g.z, -c1[j].z) - c2[j].z * ez;
            o.w = fmaf(gr[j].w, g.w, -c1[j].w) - c2[j].w * ew;
            mx = ba_amax4(mx, o);
            *reinterpret_cast<float4*>(dx + r * C + c) = o;
        }
    }
    if (amax) ba_store_amax(mx, amax);
}

// ---- BatchNorm + activation + max over the K rows of a group, without writing the activated (G*K, C) tensor.
// act(bn(.)) is monotone per channel (increasing for gamma*rstd >= 0, decreasing otherwise), so
//     max_k act(bn(h[g,k,c])) = act(bn(sel_k h[g,k,c])),  sel = max for scale >= 0, min otherwise   (exactly, in fp32),
// the same identity the fused EdgeConv uses.  Forward: one pass over h -> out (G,C), the selected pre-BN value psel and
// its k (uint8).  Backward: g' and the BatchNorm sums live on (G,C) (bn_act_bwd_reduce on psel), and
//     dh[g,k,c] = gr g'[g,c] [k == arg] - c1 - c2r (h[g,k,c] - mean)
// is one read of h and one write of dh.  Thread = (group, float4 column), K rows walked with 8 loads in flight.
__global__ void __launch_bounds__(BA_T)
pool_bn_act_fwd_kernel(const float* __restrict__ h, long G, int K, int C, const float* __restrict__ stats, float slope,
                       float* __restrict__ out, float* __restrict__ psel, uint8_t* __restrict__ arg) {
    const int CV = C >> 2;
    const long total = G * CV;
    for (long t = (long)blockIdx.x * BA_T + threadIdx.x; t < total; t += (long)gridDim.x * BA_T) {
        const long g = t / CV;
        const int c = (int)(t - g * CV) * 4;
        const float4 mu = *reinterpret_cast<const float4*>(stats + c);
        const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + c);
        const float4 be = *reinterpret_cast<const float4*>(stats + 3 * C + c);
        const float* __restrict__ hp = h + (g * K) * C + c;
        float4 best = *reinterpret_cast<const float4*>(hp);
        int ax = 0, ay = 0, az = 0, aw = 0;
#pragma unroll 8
        for (int k = 1; k < K; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(hp + (long)k * C);
            if (sc.x >= 0.f ? v.x > best.x : v.x < best.x) { best.x = v.x; ax = k; }
            if (sc.y >= 0.f ? v.y > best.y : v.y < best.y) { best.y = v.y; ay = k; }
            if (sc.z >= 0.f ? v.z > best.z : v.z < best.z) { best.z = v.z; az = k; }
            if (sc.w >= 0.f ? v.w > best.w : v.w < best.w) { best.w = v.w; aw = k; }
        }
        float4 o;
        o.x = ba_act(best.x, mu.x, sc.x, be.x, slope);
        o.y = ba_act(best.y, mu.y, sc.y, be.y, slope);
        o.z = ba_act(best.z, mu.z, sc.z, be.z, slope);
        o.w = ba_act(best.w, mu.w, sc.w, be.w, slope);
        *reinterpret_cast<float4*>(out + g * C + c) = o;
        *reinterpret_cast<float4*>(psel + g * C + c) = best;
        *reinterpret_cast<uchar4*>(arg + g * C + c) = make_uchar4((uint8_t)ax, (uint8_t)ay, (uint8_t)az, (uint8_t)aw);
    }
}

__global__ void __launch_bounds__(BA_T)
pool_bn_bwd_apply_kernel(const float* __restrict__ h, const float* __restrict__ gs, const uint8_t* __restrict__ arg, long G,
                         int K, int C, const float* __restrict__ coef, float* __restrict__ dh) {
    const int CV = C >> 2;
    const long total = G * CV;
    for (long t = (long)blockIdx.x * BA_T + threadIdx.x; t < total; t += (long)gridDim.x * BA_T) {
        const long g = t / CV;
        const int c = (int)(t - g * CV) * 4;
        const float4 gr = *reinterpret_cast<const float4*>(coef + c);
        const float4 c1 = *reinterpret_cast<const float4*>(coef + C + c);
        const float4 c2 = *reinterpret_cast<const float4*>(coef + 2 * C + c);
        const float4 mu = *reinterpret_cast<const float4*>(coef + 3 * C + c);
        const float4 gv = *reinterpret_cast<const float4*>(gs + g * C + c);
        const uchar4 a = *reinterpret_cast<const uchar4*>(arg + g * C + c);
        const float sx = gr.x * gv.x, sy = gr.y * gv.y, sz = gr.z * gv.z, sw = gr.w * gv.w;
        const float* __restrict__ hp = h + (g * K) * C + c;
        float* __restrict__ dp = dh + (g * K) * C + c;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(hp + (long)k * C);
            float4 o;
            o.x = ((k == a.x) ? sx : 0.f) - c1.x - c2.x * (v.x - mu.x);
            o.y = ((k == a.y) ? sy : 0.f) - c1.y - c2.y * (v.y - mu.y);
            o.z = ((k == a.z) ? sz : 0.f) - c1.z - c2.z * (v.z - mu.z);
            o.w = ((k == a.w) ? sw : 0.f) - c1.w - c2.w * (v.w - mu.w);
            *reinterpret_cast<float4*>(dp + (long)k * C) = o;
        }
    }
}

static bool ba_supported(long R, int C) {
    if (R <= 0 || C < 4 || C % 4) return false;
    const int cv = C / 4;
    return (cv & (cv - 1)) == 0 && cv <= 512;
}
static int ba_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
static int ba_sweep_rows(int C) { const int cv = C / 4; return cv >= BA_T ? 1 : BA_T / cv; }
// elementwise kernels: enough CTAs for 8 per SM, never more than one sweep each
static int ba_grid(long R, int C) {
    const long sweeps = (R + ba_sweep_rows(C) - 1) / ba_sweep_rows(C);
    const long cap = 8L * ba_sms();
    return (int)(sweeps < cap ? sweeps : cap);
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_bn_supported(long R, int C) { return ba_supported(R, C) ? 1 : 0; }

// number of partial blocks the reducing kernels use for an (R, C) matrix: >= 32 KB of rows per block, <= 4 per SM
// (128 KB left the mid-sized layers -- 2..16 MB, every deeper PointNet++ level and the 64-channel DGCNN layers -- on
// 16..128 CTAs, fewer than the chip has SMs: 0.7 TB/s for an 8 MB tensor on 64 CTAs)
extern "C" int pcnbr_bn_blocks(long R, int C) {
    if (!ba_supported(R, C)) return 0;
    const long by_bytes = (R * (long)C * 4 + 32767) / 32768;
    const long sweeps = (R + ba_sweep_rows(C) - 1) / ba_sweep_rows(C);
    long n = by_bytes < 4L * 148 ? by_bytes : 4L * 148;
    if (n > sweeps) n = sweeps;
    if (n < 1) n = 1;
    const long rpb = (R + n - 1) / n;
    return (int)((R + rpb - 1) / rpb);
}

extern "C" int pcnbr_bn_stats_f32(const float* x, long R, int C, float* partial, pcnbr_stream_t stream) {
    if (!x || !partial) return PCNBR_E_BADARG;
    if (!ba_supported(R, C) || ((uintptr_t)x & 15)) return PCNBR_E_TOOLARGE;
    const int nblk = pcnbr_bn_blocks(R, C);
    const int rpb = (int)((R + nblk - 1) / nblk);
    cudaStream_t s = (cudaStream_t)stream;
    const double wb = 4.0 * R * C, wf = 3.0 * R * C;
    if (C / 4 > BA_T) PCNBR_TIMED("bn_stats_kernel", s, wb, wf, (bn_stats_kernel<2><<<nblk, BA_T, 0, s>>>(x, R, C, rpb, partial)));
    else              PCNBR_TIMED("bn_stats_kernel", s, wb, wf, (bn_stats_kernel<1><<<nblk, BA_T, 0, s>>>(x, R, C, rpb, partial)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_bn_finalize_f32(const float* partial, int nblk, const float* shift, double count, int C,
                                     const float* gamma, const float* beta, float eps, float momentum,
                                     float* running_mean, float* running_var, float* stats, pcnbr_stream_t stream) {
    if (!stats || C <= 0 || nblk < 0) return PCNBR_E_BADARG;
    if (nblk > 0 && (!partial || !shift || count <= 0.0)) return PCNBR_E_BADARG;
    if (nblk == 0 && (!running_mean || !running_var)) return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    PCNBR_TIMED("bn_finalize_kernel", s, 8.0 * nblk * C + 32.0 * C, 4.0 * nblk * C,
                (bn_finalize_kernel<<<(C + BA_FC - 1) / BA_FC, BA_FC * BA_FS, 0, s>>>(partial, nblk, shift, count, C, gamma, beta, eps, momentum,
                                                                  running_mean, running_var, stats)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_bn_act_fwd_f32(const float* a, long lda, const float* b, long ldb, long R, int C, const float* stats,
                                    float slope, float* y, const unsigned long long* drop_seed, float drop_p,
                                    float* amax_out, pcnbr_stream_t stream) {
    if (drop_seed && !(drop_p >= 0.f && drop_p < 1.f)) return PCNBR_E_BADARG;
    if (!a || !stats || !y) return PCNBR_E_BADARG;
    if (!ba_supported(R, C) || (lda % 4) || (b && (ldb % 4)) || (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y | (uintptr_t)stats) & 15))
        return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = ba_grid(R, C);
    const double wb = 4.0 * R * C * (b ? 3.0 : 2.0), wf = 4.0 * R * C;
    if (C / 4 > BA_T) PCNBR_TIMED("bn_act_fwd_kernel", s, wb, wf, (bn_act_fwd_kernel<2><<<grid, BA_T, 0, s>>>(a, lda, b, ldb, R, C, stats, slope, y, drop_seed, drop_p, amax_out)));
    else              PCNBR_TIMED("bn_act_fwd_kernel", s, wb, wf, (bn_act_fwd_kernel<1><<<grid, BA_T, 0, s>>>(a, lda, b, ldb, R, C, stats, slope, y, drop_seed, drop_p, amax_out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_bn_act_bwd_reduce_f32(const float* gy, const float* a, long lda, const float* b, long ldb, long R, int C,
                                           const float* stats, float slope, float* partial, float* gs,
                                           const unsigned long long* drop_seed, float drop_p, pcnbr_stream_t stream) {
    if (drop_seed && !(drop_p >= 0.f && drop_p < 1.f)) return PCNBR_E_BADARG;
    if (!gy || !a || !stats || !partial) return PCNBR_E_BADARG;
    if (!ba_supported(R, C) || (lda % 4) || (b && (ldb % 4)) ||
        (((uintptr_t)a | (uintptr_t)b | (uintptr_t)gy | (uintptr_t)gs | (uintptr_t)stats) & 15))
        return PCNBR_E_TOOLARGE;
    const int nblk = pcnbr_bn_blocks(R, C);
    const int rpb = (int)((R + nblk - 1) / nblk);
    cudaStream_t s = (cudaStream_t)stream;
    const double wb = 4.0 * R * C * (2.0 + (b ? 1.0 : 0.0) + (gs ? 1.0 : 0.0)), wf = 8.0 * R * C;
    if (C / 4 > BA_T) PCNBR_TIMED("bn_act_bwd_reduce_kernel", s, wb, wf, (bn_act_bwd_reduce_kernel<2><<<nblk, BA_T, 0, s>>>(gy, a, lda, b, ldb, R, C, rpb, stats, slope, partial, gs, drop_seed, drop_p)));
    else              PCNBR_TIMED("bn_act_bwd_reduce_kernel", s, wb, wf, (bn_act_bwd_reduce_kernel<1><<<nblk, BA_T, 0, s>>>(gy, a, lda, b, ldb, R, C, rpb, stats, slope, partial, gs, drop_seed, drop_p)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_bn_bwd_finalize_f32(const float* partial, int nblk, const float* stats, double count, int C, int training,
                                         float* dgamma, float* dbeta, float* coef, pcnbr_stream_t stream) {
    if (!partial || !stats || !dgamma || !dbeta || !coef || C <= 0 || nblk <= 0 || count <= 0.0) return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    PCNBR_TIMED("bn_bwd_finalize_kernel", s, 8.0 * nblk * C + 40.0 * C, 4.0 * nblk * C,
                (bn_bwd_finalize_kernel<<<(C + BA_FC - 1) / BA_FC, BA_FC * BA_FS, 0, s>>>(partial, nblk, stats, count, C, training, dgamma, dbeta, coef)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_bn_act_bwd_apply_f32(const float* gy, const float* x, long R, int C, const float* stats, const float* coef,
                                          float slope, float* dx, const unsigned long long* drop_seed, float drop_p,
                                          float* amax_out, pcnbr_stream_t stream) {
    if (drop_seed && !(drop_p >= 0.f && drop_p < 1.f)) return PCNBR_E_BADARG;
    if (!gy || !x || !stats || !coef || !dx) return PCNBR_E_BADARG;
    if (!ba_supported(R, C) || (((uintptr_t)x | (uintptr_t)gy | (uintptr_t)dx | (uintptr_t)stats | (uintptr_t)coef) & 15))
        return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = ba_grid(R, C);
    const double wb = 12.0 * R * C, wf = 8.0 * R * C;
    if (C / 4 > BA_T) PCNBR_TIMED("bn_act_bwd_apply_kernel", s, wb, wf, (bn_act_bwd_apply_kernel<2><<<grid, BA_T, 0, s>>>(gy, x, R, C, stats, coef, slope, dx, drop_seed, drop_p, amax_out)));
    else              PCNBR_TIMED("bn_act_bwd_apply_kernel", s, wb, wf, (bn_act_bwd_apply_kernel<1><<<grid, BA_T, 0, s>>>(gy, x, R, C, stats, coef, slope, dx, drop_seed, drop_p, amax_out)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_pool_bn_act_fwd_f32(const float* h, long G, int K, int C, const float* stats, float slope, float* out,
                                         float* psel, uint8_t* arg, pcnbr_stream_t stream) {
    if (!h || !stats || !out || !psel || !arg || G <= 0 || K <= 0) return PCNBR_E_BADARG;
    if (K > 255 || C < 4 || C % 4 || (((uintptr_t)h | (uintptr_t)out | (uintptr_t)psel | (uintptr_t)stats) & 15) || ((uintptr_t)arg & 3))
        return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const long total = G * (C / 4);
    const long cap = 16L * ba_sms();
    const int grid = (int)((total + BA_T - 1) / BA_T < cap ? (total + BA_T - 1) / BA_T : cap);
    PCNBR_TIMED("pool_bn_act_fwd_kernel", s, 4.0 * G * K * C + 9.0 * G * C, 3.0 * G * K * C,
                (pool_bn_act_fwd_kernel<<<grid, BA_T, 0, s>>>(h, G, K, C, stats, slope, out, psel, arg)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_pool_bn_bwd_apply_f32(const float* h, const float* gs, const uint8_t* arg, long G, int K, int C,
                                           const float* coef, float* dh, pcnbr_stream_t stream) {
    if (!h || !gs || !arg || !coef || !dh || G <= 0 || K <= 0) return PCNBR_E_BADARG;
    if (K > 255 || C < 4 || C % 4 || (((uintptr_t)h | (uintptr_t)gs | (uintptr_t)dh | (uintptr_t)coef) & 15) || ((uintptr_t)arg & 3))
        return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const long total = G * (C / 4);
    const long cap = 16L * ba_sms();
    const int grid = (int)((total + BA_T - 1) / BA_T < cap ? (total + BA_T - 1) / BA_T : cap);
    PCNBR_TIMED("pool_bn_bwd_apply_kernel", s, 8.0 * G * K * C + 5.0 * G * C, 4.0 * G * K * C,
                (pool_bn_bwd_apply_kernel<<<grid, BA_T, 0, s>>>(h, gs, arg, G, K, C, coef, dh)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
