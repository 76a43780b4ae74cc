// fps.cu -- K1 farthest point sampling (reference: models/utils/common.py:6-34).
//
// One CTA per cloud.  Each thread keeps PPT points (xyz + running min distance) in registers for the
// whole kernel; per pick the only traffic is one 64-bit partial per warp through shared memory.
// The argmax with "lowest index wins" is a max over the key (dist_bits << 32 | ~index): distances
// are >= +0 so their bit patterns are monotone.  Warp stage = two REDUX ops, block stage = one
// __syncthreads per pick (double-buffered partials).  Compulsory HBM traffic is 12N + 16C bytes per
// cloud, so the kernel is ALU/latency bound on the SMs it occupies.
#include "common.cuh"
#include <cstdlib>
#include <cooperative_groups.h>

namespace pcnbr {

__device__ __forceinline__ float fps_dist(float x, float y, float z, float cx, float cy, float cz) {
    // linalg.vector_norm over 3 components on the reference's CPU path: FMA chain, then sqrt
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// the radicand of fps_dist (sqrt_rn is applied once per thread and pick, see fps_reg_kernel)
__device__ __forceinline__ float fps_dist2(float x, float y, float z, float cx, float cy, float cz) {
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy), dz = __fsub_rn(z, cz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// points of cloud b that exist: N, or n_valid[b] clamped to [1, N] (length-aware form: zero-padded evaluation batches,
// data_processing/block_datasets.py:19-25 -- the padding rows take no part, as if the cloud had been passed alone)
__device__ __forceinline__ int fps_valid(const int32_t* __restrict__ n_valid, int b, int N) {
    return n_valid ? min(max(n_valid[b], 1), N) : N;
}

// warp-wide max of (hi, lo) pairs in lexicographic order, result broadcast to all lanes
__device__ __forceinline__ void warp_max_pair(uint32_t& hi, uint32_t& lo) {
    const uint32_t mh = __reduce_max_sync(PCNBR_FULL, hi);
    const uint32_t ml = __reduce_max_sync(PCNBR_FULL, hi == mh ? lo : 0u);
    hi = mh;
    lo = ml;
}

template <int PPT, int T>
__global__ void __launch_bounds__(T, 1)
fps_reg_kernel(const float* __restrict__ xyz, int N, int C, const int32_t* __restrict__ start,
               const int32_t* __restrict__ n_valid, int32_t* __restrict__ idx_out, float* __restrict__ xyz_out) {
    constexpr int W = T / 32;
    __shared__ uint32_t s_hi[2][W], s_lo[2][W];

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ p = xyz + (size_t)b * N * 3;
    const int NV = fps_valid(n_valid, b, N);               // rows >= NV are padding of a zero-padded batch: they do not exist

    float x[PPT], y[PPT], z[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int n = j * T + tid;
        if (n < NV) {
            x[j] = p[n * 3 + 0]; y[j] = p[n * 3 + 1]; z[j] = p[n * 3 + 2];
            md[j] = __int_as_float(0x7f800000);          // +inf (common.py:21)
        } else {
            x[j] = y[j] = z[j] = 0.f;
            md[j] = 0.f;                                 // padding lanes can never beat a real point
        }
    }

    int cur = min(max(start[b], 0), NV - 1);       // a caller-supplied first pick outside [0, NV) is clamped, never dereferenced
    float cx = p[cur * 3 + 0], cy = p[cur * 3 + 1], cz = p[cur * 3 + 2];

    for (int i = 0; i < C; ++i) {
        if (tid == 0) {
            idx_out[(size_t)b * C + i] = cur;
            if (xyz_out) {
                float* o = xyz_out + ((size_t)b * C + i) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
        }
        if (i + 1 == C) break;
        const int buf = i & 1;

        // The running minimum is kept SQUARED: sqrt_rn is monotone, so min and max commute with it and only the
        // thread's winner needs a square root (1 instead of PPT per pick).  What sqrt does change is ties: two different
        // squared distances can round to the same norm, and the reference (argmax over the norms, lowest index on ties)
        // then takes the lower index.  A rounding class spans at most 3 adjacent floats, so any point within 4 ulps of
        // the thread's maximum is re-checked exactly (rare).
        uint32_t mb = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            md[j] = fminf(md[j], fps_dist2(x[j], y[j], z[j], cx, cy, cz));      // common.py:28-30 (squared)
            mb = max(mb, __float_as_uint(md[j]));
        }
        const float smax = __fsqrt_rn(__uint_as_float(mb));
        int bj = PPT;
        bool near = false;                               // some other point lies 1..4 ulps below the maximum (rare)
#pragma unroll
        for (int j = PPT - 1; j >= 0; --j) {
            const uint32_t h = __float_as_uint(md[j]);
            if (h == mb) bj = j;
            near |= (mb - h - 1u) < 4u;
        }
        if (near) {                                      // ONE branch per thread and pick; exact re-check of the rounding class
#pragma unroll
            for (int j = PPT - 1; j >= 0; --j) {
                const uint32_t h = __float_as_uint(md[j]);
                if ((mb - h - 1u) < 4u && j < bj && __fsqrt_rn(md[j]) == smax) bj = j;
            }
        }
        const uint32_t bh = __float_as_uint(smax);
        const uint32_t bl = 0xffffffffu - (uint32_t)(bj * T + tid);
        uint32_t wh = bh, wl = bl;
        warp_max_pair(wh, wl);
        if (lane == 0) { s_hi[buf][warp] = wh; s_lo[buf][warp] = wl; }
        __syncthreads();
        uint32_t gh = (lane < W) ? s_hi[buf][lane] : 0u;
        uint32_t gl = (lane < W) ? s_lo[buf][lane] : 0u;
        warp_max_pair(gh, gl);
        cur = (int)(0xffffffffu - gl);                   // torch.max: lowest index on ties (common.py:31)
        // the winner's coordinates come back from the cloud's L1-resident copy (one uniform 12-byte read) instead of being
        // looked up in the registers of every warp's candidate (PPT x 4 select instructions per warp and pick: 20 % of the
        // kernel's instructions at 16 points per thread, ncu source page)
        cx = p[cur * 3 + 0]; cy = p[cur * 3 + 1]; cz = p[cur * 3 + 2];
    }
}

// 8192 < N <= 131072 (the 24 k-point chunks of PointNeXt, the N sweep up to 100 k): one cloud on a CLUSTER of 4, 8 or 16 CTAs, a
// contiguous share of the points in the registers of each.  Per pick: CTA-local argmax exactly as fps_reg_kernel (same
// keys, so the same point wins); warp 0 then sends the CTA's winner (key + xyz, 20 bytes) to EVERY CTA of the cluster with
// st.async -- a distributed-shared-memory store that signals the receiver's mbarrier (complete_tx) -- and every thread
// waits on its OWN CTA's mbarrier until the CL messages of this pick have landed, then takes the maximum of the
// candidates from local shared memory.  No cluster-wide hardware barrier: the first version used cluster.sync() per pick
// and spent ~2 us per pick in it whatever N (profiles/r4_kernel_sweep.md: 2.07 ms for 1024 picks at N = 16 k .. 32 k).
// Slots and mbarriers are double-buffered by pick parity: a CTA sends pick i+2 only after it has received every CTA's
// pick i+1, i.e. after every CTA has finished reading the slots of pick i.
constexpr int FPS_CL_MAX = 16;        // 16 CTAs per cluster is the non-portable maximum (one cluster per GPC)

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t fps_mapa(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
__device__ __forceinline__ void fps_send4(uint32_t raddr, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar) : "memory");
}
__device__ __forceinline__ void fps_send1(uint32_t raddr, uint32_t rbar, uint32_t a) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(raddr), "r"(a), "r"(rbar) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int PPT, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_kernel(const float* __restrict__ xyz, int N, int C, const int32_t* __restrict__ start,
                   const int32_t* __restrict__ n_valid, int32_t* __restrict__ idx_out, float* __restrict__ xyz_out) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int W = T / 32;
    __shared__ uint32_t s_hi[2][W], s_lo[2][W];
    __shared__ float s_xyz[2][W][3];
    __shared__ __align__(16) uint32_t c_msg[2][FPS_CL_MAX][4];   // {key hi, key lo, x, y} of every CTA's winner
    __shared__ uint32_t c_z[2][FPS_CL_MAX];
    __shared__ __align__(8) uint64_t c_bar[2];

    const int CL = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ p = xyz + (size_t)b * N * 3;
    const int NV = fps_valid(n_valid, b, N);
    const int Q = (N + CL - 1) / CL;                          // points per CTA, contiguous share
    const int n0 = rank * Q;

    float x[PPT], y[PPT], z[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int q = j * T + tid, n = n0 + q;
        if (q < Q && n < NV) {
            x[j] = p[n * 3 + 0]; y[j] = p[n * 3 + 1]; z[j] = p[n * 3 + 2];
            md[j] = __int_as_float(0x7f800000);
        } else {
            x[j] = y[j] = z[j] = 0.f;
            md[j] = 0.f;
        }
    }
    int cur = min(max(start[b], 0), NV - 1);       // a caller-supplied first pick outside [0, NV) is clamped, never dereferenced
    float cx = p[cur * 3 + 0], cy = p[cur * 3 + 1], cz = p[cur * 3 + 2];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&c_bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(&c_bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster.sync();                                            // every CTA is resident and its mbarriers exist before remote stores

    for (int i = 0; i < C; ++i) {
        if (rank == 0 && tid == 0) {
            idx_out[(size_t)b * C + i] = cur;
            if (xyz_out) {
                float* o = xyz_out + ((size_t)b * C + i) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
        }
        if (i + 1 == C) break;
        const int buf = i & 1;
        const uint32_t bar = fps_smem_u32(&c_bar[buf]);
        if (tid == 0)                                          // this pick's CL messages: 20 bytes each
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(CL * 20)) : "memory");
        uint32_t mb = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            md[j] = fminf(md[j], fps_dist2(x[j], y[j], z[j], cx, cy, cz));
            mb = max(mb, __float_as_uint(md[j]));
        }
        const float smax = __fsqrt_rn(__uint_as_float(mb));
        int bj = PPT;
        bool near = false;                               // some other point lies 1..4 ulps below the maximum (rare)
#pragma unroll
        for (int j = PPT - 1; j >= 0; --j) {
            const uint32_t h = __float_as_uint(md[j]);
            if (h == mb) bj = j;
            near |= (mb - h - 1u) < 4u;
        }
        if (near) {                                      // ONE branch per thread and pick; exact re-check of the rounding class
#pragma unroll
            for (int j = PPT - 1; j >= 0; --j) {
                const uint32_t h = __float_as_uint(md[j]);
                if ((mb - h - 1u) < 4u && j < bj && __fsqrt_rn(md[j]) == smax) bj = j;
            }
        }
        const uint32_t bh = __float_as_uint(smax);
        // padding slots (beyond the share or the cloud) carry md = 0 and the lowest-priority index: their slot numbers would
        // otherwise alias real points of the next CTA's share
        const int bq = bj * T + tid;
        const uint32_t bl = (bq < Q && n0 + bq < NV) ? 0xffffffffu - (uint32_t)(n0 + bq) : 0u;
        uint32_t wh = bh, wl = bl;
        warp_max_pair(wh, wl);
        if (bh == wh && bl == wl) {
            float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
            for (int j = 0; j < PPT; ++j)
                if (j == bj) { wx = x[j]; wy = y[j]; wz = z[j]; }
            if (__ffs(__ballot_sync(__activemask(), true)) - 1 == lane) {   // all-padding warps can tie on (0, 0): one writer
                s_hi[buf][warp] = wh; s_lo[buf][warp] = wl;
                s_xyz[buf][warp][0] = wx; s_xyz[buf][warp][1] = wy; s_xyz[buf][warp][2] = wz;
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t gh = (lane < W) ? s_hi[buf][lane] : 0u;
            uint32_t gl = (lane < W) ? s_lo[buf][lane] : 0u;
            const uint32_t mh = gh, ml = gl;
            warp_max_pair(gh, gl);
            const uint32_t who = __ballot_sync(PCNBR_FULL, lane < W && mh == gh && ml == gl);
            const int w = __ffs(who) - 1;
            if (lane < CL) {                                   // lane r sends this CTA's winner to CTA r (slot [buf][rank] there)
                const uint32_t rbar = fps_mapa(bar, (uint32_t)lane);
                fps_send4(fps_mapa(fps_smem_u32(&c_msg[buf][rank][0]), (uint32_t)lane), rbar, gh, gl,
                          __float_as_uint(s_xyz[buf][w][0]), __float_as_uint(s_xyz[buf][w][1]));
                fps_send1(fps_mapa(fps_smem_u32(&c_z[buf][rank]), (uint32_t)lane), rbar, __float_as_uint(s_xyz[buf][w][2]));
            }
        }
        fps_mbar_wait_cluster(bar, (uint32_t)((i >> 1) & 1));     // all CL winners of this pick have landed here
        uint32_t gh = c_msg[buf][0][0], gl = c_msg[buf][0][1];
        int w = 0;
        for (int r = 1; r < CL; ++r) {
            const uint32_t h = c_msg[buf][r][0], l = c_msg[buf][r][1];
            if (h > gh || (h == gh && l > gl)) { gh = h; gl = l; w = r; }
        }
        cur = (int)(0xffffffffu - gl);                       // torch.max: lowest index on ties (common.py:31)
        cx = __uint_as_float(c_msg[buf][w][2]); cy = __uint_as_float(c_msg[buf][w][3]); cz = __uint_as_float(c_z[buf][w]);
    }
    cluster.sync();                                            // no CTA exits while a peer may still store into it
}

template <int PPT, int T>
static int fps_launch_cluster(int CL, int B, const float* xyz, int N, int C, const int32_t* start, const int32_t* n_valid,
                              int32_t* idx_out, float* xyz_out, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CL > 8) {                                              // beyond the portable cluster size: opt in once per function
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<PPT, T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PPT, T>, xyz, N, C, start, n_valid, idx_out, xyz_out);
}

// Any N: running distances live in a global workspace, coordinates are re-read through L1/L2.
template <int T>
__global__ void __launch_bounds__(T, 1)
fps_big_kernel(const float* __restrict__ xyz, int N, int C, const int32_t* __restrict__ start,
               const int32_t* __restrict__ n_valid, int32_t* __restrict__ idx_out, float* __restrict__ xyz_out, float* __restrict__ ws) {
    constexpr int W = T / 32;
    __shared__ uint32_t s_hi[2][W], s_lo[2][W];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* __restrict__ p = xyz + (size_t)b * N * 3;
    float* __restrict__ md = ws + (size_t)b * N;
    const int NV = fps_valid(n_valid, b, N);
    for (int n = tid; n < NV; n += T) md[n] = __int_as_float(0x7f800000);
    int cur = min(max(start[b], 0), NV - 1);       // a caller-supplied first pick outside [0, NV) is clamped, never dereferenced
    for (int i = 0; i < C; ++i) {
        const float cx = p[cur * 3 + 0], cy = p[cur * 3 + 1], cz = p[cur * 3 + 2];
        if (tid == 0) {
            idx_out[(size_t)b * C + i] = cur;
            if (xyz_out) {
                float* o = xyz_out + ((size_t)b * C + i) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
        }
        if (i + 1 == C) break;
        const int buf = i & 1;
        uint32_t bh = 0, bl = 0;
        for (int n = tid; n < NV; n += T) {
            const float d = fps_dist(p[n * 3 + 0], p[n * 3 + 1], p[n * 3 + 2], cx, cy, cz);
            const float m = fminf(md[n], d);
            md[n] = m;
            const uint32_t h = __float_as_uint(m), l = 0xffffffffu - (uint32_t)n;
            if (h > bh || (h == bh && l > bl)) { bh = h; bl = l; }
        }
        warp_max_pair(bh, bl);
        if (lane == 0) { s_hi[buf][warp] = bh; s_lo[buf][warp] = bl; }
        __syncthreads();
        uint32_t gh = (lane < W) ? s_hi[buf][lane] : 0u;
        uint32_t gl = (lane < W) ? s_lo[buf][lane] : 0u;
        warp_max_pair(gh, gl);
        cur = (int)(0xffffffffu - gl);
    }
}

}  // namespace pcnbr

extern "C" size_t pcnbr_fps_ws_bytes(int B, int N) {
    return (N > 8192) ? sizeof(float) * (size_t)B * (size_t)N : 0;
}

extern "C" int pcnbr_fps_f32(const float* xyz, int B, int N, int C, const int32_t* start, int32_t* idx_out,
                             float* xyz_out, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    return pcnbr_fps_len_f32(xyz, B, N, C, start, nullptr, idx_out, xyz_out, ws, ws_bytes, stream);
}

extern "C" int pcnbr_fps_len_f32(const float* xyz, int B, int N, int C, const int32_t* start, const int32_t* n_valid,
                                 int32_t* idx_out, float* xyz_out, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    using namespace pcnbr;
    if (!xyz || !start || !idx_out || B <= 0 || N <= 0 || C <= 0) return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    // algorithmic work: 12N + 16C compulsory bytes, 10 N C lane-ops per cloud.  One CTA = one SM per
    // cloud, so the ceiling is the lane-op rate of the B SMs that are occupied; the profiler's ALU roofline
    // is the whole chip's FMA rate (2 flops per lane-op slot), hence the 2 * 148 / min(B,148) scaling of the stated work.
    const double wb = (double)B * (12.0 * N + 16.0 * C);
    const double wf = 10.0 * B * (double)N * C * 2.0 * 148.0 / (double)(B < 148 ? B : 148);
    // Threads per cloud (N <= 8192, points in registers): fewer, fatter threads halve the per-warp reduction overhead of a
    // pick but leave fewer warps to hide the pick-to-pick latency chain.  Measured at 32 x (4096 -> 1024): 1024 threads x 4
    // points 752 us, 512 x 8 625 us, 256 x 16 588 us (with the single-branch tie re-check; the per-point branches of the
    // first version made the fat threads slower).  PCNBR_FPS_T=64|128|256|512|1024 forces a width (A/B in the sweep).
    static const int fps_t_env = getenv("PCNBR_FPS_T") ? atoi(getenv("PCNBR_FPS_T")) : 0;
    int T = 256;
    if (fps_t_env == 64 || fps_t_env == 128 || fps_t_env == 256 || fps_t_env == 512 || fps_t_env == 1024) T = fps_t_env;
    while (T < 1024 && (N + T - 1) / T > (T == 128 ? 32 : 16)) T *= 2;
    int ppt = 1;
    while (ppt * T < N) ppt *= 2;
#define PCNBR_FPS_CASE(P, TT)                                                                                        \
    if (ppt == P && T == TT) {                                                                                       \
        PCNBR_TIMED("fps_reg_kernel", s, wb, wf,                                                                     \
                    (fps_reg_kernel<P, TT><<<B, TT, 0, s>>>(xyz, N, C, start, n_valid, idx_out, xyz_out)));                   \
        PCNBR_CHECK_LAUNCH();                                                                                        \
        return 0;                                                                                                    \
    }
    if (N <= 8192) {
        PCNBR_FPS_CASE(1, 64) PCNBR_FPS_CASE(2, 64) PCNBR_FPS_CASE(4, 64) PCNBR_FPS_CASE(8, 64) PCNBR_FPS_CASE(16, 64)
        PCNBR_FPS_CASE(1, 128) PCNBR_FPS_CASE(2, 128) PCNBR_FPS_CASE(4, 128) PCNBR_FPS_CASE(8, 128) PCNBR_FPS_CASE(16, 128)
        PCNBR_FPS_CASE(32, 128)
        PCNBR_FPS_CASE(1, 256) PCNBR_FPS_CASE(2, 256) PCNBR_FPS_CASE(4, 256) PCNBR_FPS_CASE(8, 256) PCNBR_FPS_CASE(16, 256)
        PCNBR_FPS_CASE(1, 512) PCNBR_FPS_CASE(2, 512) PCNBR_FPS_CASE(4, 512) PCNBR_FPS_CASE(8, 512) PCNBR_FPS_CASE(16, 512)
        PCNBR_FPS_CASE(1, 1024) PCNBR_FPS_CASE(2, 1024) PCNBR_FPS_CASE(4, 1024) PCNBR_FPS_CASE(8, 1024)
        return PCNBR_E_TOOLARGE;                            // unreachable: every (ppt, T) above is instantiated
    }
#undef PCNBR_FPS_CASE
    if (N <= 16 * 8192 && B * 8 <= 148 * 2) {
        // One cloud on a cluster, 16 points per thread as in the single-CTA kernel (fat threads: the per-pick cost is the
        // block reduction plus the exchange, both cheaper with fewer warps -- 1024 threads x 8 points cost 1.8 us per pick):
        //   N <= 32768 : ceil(N / 4096) CTAs (2..8) of 256 threads;   N <= 65536 : 8 CTAs of 512 threads;
        //   N <= 131072: 16 CTAs of 512 threads (non-portable cluster size: one cluster per GPC).
        // ceiling: the CL SMs per cloud
        const bool wide = N > 8 * 4096;
        const int CL = !wide ? (N + 4095) / 4096 : (N <= 8 * 8192 ? 8 : 16);
        const double wfc = 10.0 * B * (double)N * C * 2.0 * 148.0 / (double)(B * CL < 148 ? B * CL : 148);
        int rc = 0;
        if (wide) PCNBR_TIMED("fps_cluster_kernel", s, wb, wfc, (rc = fps_launch_cluster<16, 512>(CL, B, xyz, N, C, start, n_valid, idx_out, xyz_out, s)));
        else      PCNBR_TIMED("fps_cluster_kernel", s, wb, wfc, (rc = fps_launch_cluster<16, 256>(CL, B, xyz, N, C, start, n_valid, idx_out, xyz_out, s)));
        if (rc) return rc;
    } else {
        if (!ws || ws_bytes < pcnbr_fps_ws_bytes(B, N)) return PCNBR_E_WORKSPACE;
        PCNBR_TIMED("fps_big_kernel", s, wb, wf, (fps_big_kernel<1024><<<B, 1024, 0, s>>>(xyz, N, C, start, n_valid, idx_out, xyz_out, (float*)ws)));
    }
    PCNBR_CHECK_LAUNCH();
    return 0;
}
