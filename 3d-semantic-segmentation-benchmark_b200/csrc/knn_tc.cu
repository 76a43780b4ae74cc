// knn_tc.cu -- K4: feature-space kNN (DGCNN layers 1-4) on the 5th-generation tensor cores.
//
// Reference: models/dgcnn/dgcnn.py:16-20 -- a dense (N,N) = X^T X contraction (cuBLAS/MKL SGEMM), two
// broadcast adds and torch.topk over 4096-wide rows: 3 x 64 MB of temporaries per cloud.
//
// Here the (N,N) matrix only ever exists as 128x128 fp32 tiles in TMEM, and the tensor core delivers the
// finished RANKING SCORE, not just the dot product: the operands are the centred, power-of-two scaled
// features in fp16 (a = (x - mean)/S, |a_f| <= 1) with a 16-wide "tail" K-slice appended,
//
//      A-role row i : [ a_i (F, zero-padded to 16s) | 1  1  1  u0 u1 u2  m_i  0 ... ]
//      B-role row j : [ a_j                         | t0 t1 t2 1  1  1   nb_j 0 ... ]
//
//   t0+t1+t2 = -|a_j|^2/2 (3-term fp16 split, exact to 2^-25), nb_j >= |a_j|, m_i = +-C1 |a_i|, u = split(-thr_i),
// so one accumulator element is   acc_ij = a_i.a_j - |a_j|^2/2  -+  C1 |a_i||a_j|  - thr_i :
// half the negative squared distance up to a row constant, with the worst-case fp16 rounding error of THIS pair
// (Cauchy-Schwarz: 2^-10 |a_i||a_j|) already added or subtracted.  The epilogue therefore never touches a score
// with an FMA -- it only needs 3-input maxima and sign bits:
//
//   producer warp : TMA (cp.async.bulk.tensor, 128-byte swizzle) of the 128-column feature slab + a 4 KB bulk copy
//                   of its tail slice into a 6-slot shared-memory ring; 256 query rows (two 128-row blocks, so every
//                   slab is used twice) stay resident per work unit;
//   MMA warp      : one thread issues tcgen05.mma kind::f16, M=128 x N=128 x K=16, F/16 + 1 instructions per
//                   (row block, tile), into one of four 128-column TMEM accumulators (all 512 TMEM columns);
//   8 epilogue warps (one thread per query row), tcgen05.ld straight out of TMEM:
//        pass 1  (m_i = -C1|a_i|: LOWER bounds): 64 running group maxima per row, one FMNMX3 per two columns;
//                the k-th largest of them is a lower bound tau_i on the row's k-th best exact score;
//        pass 2  (m_i = +C1|a_i|, u = -thr_i, thr_i = tau_i - slack: UPPER bounds minus the threshold; the tiles are
//                recomputed, cheaper than storing N^2 floats): a column survives iff its accumulator is >= 0 -- one
//                funnel shift per column collects the sign bits, the ~k+8 survivors of 4096 go to a per-row queue.
//
// A second kernel re-ranks each row's survivors with the reference's EXACT arithmetic (sequential fp32 FMA chain
// over the feature index, ATen's cascade |x|^2, ((-xx_j) - inner) - xx_i) and sorts them by (value, index).  Because
// pass 1 uses lower and pass 2 upper bounds of the exact score, the survivor set provably contains the exact top-k:
// the indices are bit-identical to oracle/canon.c although the bulk of the flops ran in fp16.  Rows whose queue
// overflows (degenerate clouds) fall back to an exact full scan inside the re-rank kernel.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace pcnbr {

constexpr int TC_ROWS = 256;       // query rows per work unit = two TMEM row blocks of 128 lanes
constexpr int TC_N = 128;          // candidate columns per tile (TMEM columns per accumulator)
constexpr int TC_RING = 6;         // ring slots: 128 columns x (128 B features + 32 B tail) = 20 KB
constexpr int TC_QCAP = 96;        // survivors kept per row (uint16 indices)
constexpr int TC_QSTRIDE = 98;     // uint16 per queue row in shared memory (49 words: conflict-free)
constexpr int TC_THREADS = 320;    // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-9 epilogue
constexpr int TC_KMAIN = 64;       // fp16 per feature row (one 128-byte swizzle atom), F <= 64
constexpr uint32_t TC_MAIN_BYTES = 128 * 128;      // 128 rows x 128 B
constexpr uint32_t TC_TAIL_BYTES = 128 * 32;       // 128 rows x 16 fp16, interleaved core-matrix layout
constexpr uint32_t TC_SLOT_BYTES = TC_MAIN_BYTES + TC_TAIL_BYTES;
constexpr float TC_PAD_T0 = -60000.f;              // tail t0 of the padding rows n >= N: never selected
// Error model, in accumulator units (scores of the scaled features, |a_f| <= 1, |a| <= 8):
//   |acc_ij - exact_ij| <= C1 |a_i||a_j| + C0,   C1 covers the fp16 rounding of both operands (2^-10 + 2^-20, Cauchy-
//   Schwarz) with 2 % headroom for the rounding of m_i / nb_j themselves;  C0 = TC_C_ACC Bmax^2 + TC_C_LIN Bmax +
//   TC_C_ABS covers the fp32 accumulation inside the tensor core, fp16 subnormals and the t / u split residues.
// On top, in score units of the UNSCALED features (as in the CUDA-core path's analysis):
//   c_ref |x_i| max|x_j|     rounding of the REFERENCE's own fp32 value (F-term FMA chain, cascade |x|^2): its ranking
//                            deviates from the true distances by that much, and we must follow it; = 2 (F+8) 2^-24
//   C_CTR max|x_j| max|x'_j| rounding of the centring subtraction and of |x'_j|^2
constexpr float TC_C1 = 1.0e-3f, TC_C_ACC = 2.0e-5f, TC_C_LIN = 1.0e-6f, TC_C_ABS = 1.0e-7f, TC_C_CTR = 1e-6f;

// ------------------------------------------------------------------------------------ PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) in [32,46),
// version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)(1024 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}
// K-major, no swizzle ("interleave"): core matrices of 8 rows x 16 B; the two K chunks of a 16-wide fp16 slice are
// LBO = 128 B apart, consecutive 8-row groups SBO = 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_tail(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(128 >> 4) << 16);
    const uint64_t hi = (uint64_t)(256 >> 4) | ((uint64_t)1 << 14);
    return lo | (hi << 32);
}
// kind::f16 (A, B = fp16), D = fp32, A/B K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// sign bits of 16 accumulator values, value i -> bit 15 - i; four independent funnel-shift chains
__device__ __forceinline__ uint32_t sign_bits16(const uint32_t (&r)[16]) {
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c0 = __funnelshift_l(r[i], c0, 1);
        c1 = __funnelshift_l(r[4 + i], c1, 1);
        c2 = __funnelshift_l(r[8 + i], c2, 1);
        c3 = __funnelshift_l(r[12 + i], c3, 1);
    }
    return (c0 << 12) | (c1 << 8) | (c2 << 4) | c3;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// v -> three fp16 terms with h0 + h1 + h2 = v up to 2^-25 (|v| < 60000)
__device__ __forceinline__ void split3_f16(float v, __half& h0, __half& h1, __half& h2) {
    h0 = __float2half_rn(v);
    const float r1 = __fsub_rn(v, __half2float(h0));
    h1 = __float2half_rn(r1);
    h2 = __float2half_rn(__fsub_rn(r1, __half2float(h1)));
}

// byte offset of row r's 16-byte chunk c (c = 0: tail elements 0..7, c = 1: 8..15) inside a 128-row tail block
__device__ __host__ __forceinline__ uint32_t tail_offset(int r, int c) {
    return (uint32_t)((r >> 3) * 256 + c * 128 + (r & 7) * 16);
}

// ------------------------------------------------------------------------------------ operand preparation

// Per-cloud channel sums and max |x| over a slice of the points: part[b][chunk][0..F) sums, [64] max (fixed order).
constexpr int TC_MEAN_CHUNKS = 16;
constexpr int TC_PART = 65;
__global__ void __launch_bounds__(256)
knn_tc_mean_kernel(const float* __restrict__ x, int F, int N, long sf, long sn, const int32_t* __restrict__ n_valid,
                   float* __restrict__ part) {
    __shared__ float red[256], redm[256];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int groups = 256 / F;                                   // F <= 64
    const int f = threadIdx.x % F, g = threadIdx.x / F;
    const int NV = len_valid(n_valid, b, N);                      // length-aware form: the padding rows do not exist
    const int per = (NV + TC_MEAN_CHUNKS - 1) / TC_MEAN_CHUNKS;
    const int n0 = min(NV, chunk * per), n1 = min(NV, n0 + per);
    const float* __restrict__ xb = x + (size_t)b * F * N;
    float acc = 0.f, mx = 0.f;
    if (g < groups)
        for (int n = n0 + g; n < n1; n += groups) {
            const float v = xb[(size_t)f * sf + (size_t)n * sn];
            acc += v;
            mx = fmaxf(mx, fabsf(v));
        }
    red[threadIdx.x] = acc;
    redm[threadIdx.x] = mx;
    __syncthreads();
    float* __restrict__ out = part + ((size_t)b * TC_MEAN_CHUNKS + chunk) * TC_PART;
    if (g == 0) {
        for (int k = 1; k < groups; ++k) acc += red[k * F + f];
        out[f] = acc;
    }
    if (threadIdx.x == 0) {
        for (int t = 1; t < groups * F; ++t) mx = fmaxf(mx, redm[t]);
        out[64] = mx;
    }
}

// One warp per point n < Npad: a = (x - mean)/S as fp16 (point-major, K-major for the MMA, zero-padded to 64);
// B-role tail [t0 t1 t2 1 1 1 nb 0 | 0 x 8] in the interleaved core-matrix layout the MMA descriptor reads;
// xt = exact copy of x (only when the input is not already point-major); xxc = |x - mean|^2; per cloud: S and the
// maxima of |x|^2 and |x'|^2 (bit patterns of non-negative floats order like the values).
// scal[b] = {max |x|^2, max |x'|^2, S, -}
__global__ void __launch_bounds__(256)
knn_tc_prep_kernel(const float* __restrict__ x, const float* __restrict__ part, const float* __restrict__ xx,
                   int F, int N, int Npad, long sf, long sn, const int32_t* __restrict__ n_valid, float* __restrict__ xt,
                   __half* __restrict__ ymain, uint8_t* __restrict__ tailb, float* __restrict__ xxc, uint32_t* __restrict__ scal) {
    __shared__ float mu[64];
    __shared__ float s_scale;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NV = len_valid(n_valid, b, N);                      // rows >= NV are treated like the rows >= N: padding columns
    if (threadIdx.x < 64) {
        float m = 0.f;
        if (threadIdx.x < F)
            for (int c = 0; c < TC_MEAN_CHUNKS; ++c) m += part[((size_t)b * TC_MEAN_CHUNKS + c) * TC_PART + threadIdx.x];
        mu[threadIdx.x] = m / (float)NV;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = 0.f, mm = 0.f;
        for (int c = 0; c < TC_MEAN_CHUNKS; ++c) mx = fmaxf(mx, part[((size_t)b * TC_MEAN_CHUNKS + c) * TC_PART + 64]);
        for (int f = 0; f < F; ++f) mm = fmaxf(mm, fabsf(mu[f]));
        const float v = mx + mm;                                   // >= max |x_f - mean_f|
        // power of two above v (next binade): |a_f| < 1
        const float S = (v > 1e-30f && v < 1e30f) ? __uint_as_float(((__float_as_uint(v) >> 23) + 1u) << 23) : 1.0f;
        s_scale = S;
        if (blockIdx.x == 0) scal[4 * b + 2] = __float_as_uint(S);
    }
    __syncthreads();
    const float S = s_scale, inv = 1.0f / S;                       // exact: S is a power of two
    const float* __restrict__ xb = x + (size_t)b * F * N;
    uint8_t* __restrict__ tb = tailb + (size_t)b * Npad * 32;
    float mx = 0.f, mxc = 0.f;
    for (int n = blockIdx.x * 8 + warp; n < Npad; n += gridDim.x * 8) {
        const int f0 = 2 * lane, f1 = 2 * lane + 1;
        float v0 = 0.f, v1 = 0.f;
        if (n < NV) {
            if (f0 < F) {
                const float v = xb[(size_t)f0 * sf + (size_t)n * sn];
                v0 = __fsub_rn(v, mu[f0]);
                if (xt) xt[((size_t)b * N + n) * F + f0] = v;
            }
            if (f1 < F) {
                const float v = xb[(size_t)f1 * sf + (size_t)n * sn];
                v1 = __fsub_rn(v, mu[f1]);
                if (xt) xt[((size_t)b * N + n) * F + f1] = v;
            }
        }
        float ss = fmaf(v1, v1, v0 * v0);
        reinterpret_cast<__half2*>(ymain + ((size_t)b * Npad + n) * TC_KMAIN)[lane] = __floats2half2_rn(v0 * inv, v1 * inv);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(PCNBR_FULL, ss, d);
        const int blk = n >> 7, r = n & 127;
        uint8_t* trow = tb + (size_t)blk * TC_TAIL_BYTES;
        if (lane == 0) {
            __half t[8];
            for (int i = 0; i < 8; ++i) t[i] = __float2half_rn(0.f);
            t[3] = t[4] = t[5] = __float2half_rn(1.0f);
            if (n < NV) {
                xxc[(size_t)b * N + n] = ss;
                split3_f16(-0.5f * ss * inv * inv, t[0], t[1], t[2]);
                t[6] = __float2half_ru(sqrtf(ss) * inv * 1.0001f);            // nb_j >= |a_j|
            } else {
                t[0] = __float2half_rn(TC_PAD_T0);
            }
            uint4 pk;
            memcpy(&pk, t, 16);
            *reinterpret_cast<uint4*>(trow + tail_offset(r, 0)) = pk;
        } else if (lane == 1) {
            *reinterpret_cast<uint4*>(trow + tail_offset(r, 1)) = make_uint4(0, 0, 0, 0);
        }
        if (n < NV) {
            mxc = fmaxf(mxc, ss);
            mx = fmaxf(mx, xx[(size_t)b * N + n]);
        }
    }
    // one pair of atomics per CTA, not per warp: 1184 warps per cloud hammering two addresses serialised in L2 (~25 of the
    // kernel's 35 us)
    __shared__ float s_mx[8], s_mxc[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(PCNBR_FULL, mx, d));
        mxc = fmaxf(mxc, __shfl_xor_sync(PCNBR_FULL, mxc, d));
    }
    if (lane == 0) { s_mx[warp] = mx; s_mxc[warp] = mxc; }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, s_mx[w]); mxc = fmaxf(mxc, s_mxc[w]); }
        atomicMax(&scal[4 * b], __float_as_uint(mx));
        atomicMax(&scal[4 * b + 1], __float_as_uint(mxc));
    }
}

// ------------------------------------------------------------------------------------ main kernel

template <int KSTEPS, bool DUMP>      // KSTEPS = ceil(F / 16); DUMP: also write the pass-1 lower-bound scores (tests only)
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_main, const uint8_t* __restrict__ tailb,
              const float* __restrict__ xx, const float* __restrict__ xxc, const uint32_t* __restrict__ scal,
              int B, int N, int Npad, int K, float c_ref, const int32_t* __restrict__ n_valid,
              int32_t* __restrict__ qcnt, uint16_t* __restrict__ qidx, float* __restrict__ dump) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B-swizzle atoms) by OFFSET, so the compiler keeps the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sAmain = smem;                                       // 2 x 16 KB
    uint8_t* sAtail = sAmain + 2 * TC_MAIN_BYTES;                 // 2 x 4 KB
    uint8_t* sRing = sAtail + 2 * TC_TAIL_BYTES;                  // TC_RING x 20 KB
    uint16_t* sQ = (uint16_t*)(sRing + TC_RING * TC_SLOT_BYTES);  // 256 x 98 x 2 B
    uint64_t* bars = (uint64_t*)(sQ + TC_ROWS * TC_QSTRIDE);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* atail_full = bars + 2;
    uint64_t* ring_full = bars + 3;                               // [TC_RING]
    uint64_t* ring_empty = bars + 3 + TC_RING;                    // [TC_RING]
    uint64_t* tmem_full = bars + 3 + 2 * TC_RING;                 // [4]
    uint64_t* tmem_empty = bars + 7 + 2 * TC_RING;                // [4]
    uint32_t* tmem_slot = (uint32_t*)(bars + 11 + 2 * TC_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RT = (Npad + TC_ROWS - 1) / TC_ROWS;
    const int units = B * RT;
    // Length-aware form: cloud b has NV = n_valid[b] real points.  Work units whose rows are all padding and column tiles
    // beyond the last real point are skipped -- by the three roles alike, so ring slots, TMEM buffers and phases stay in step.
#define TC_UNIT_GEOMETRY()                                                     \
    const int b = unit / RT, row0 = (unit - b * RT) * TC_ROWS;                 \
    const int NV = len_valid(n_valid, b, N);                                   \
    if (row0 >= NV) continue;                                                  \
    const int CT = (NV + TC_N - 1) / TC_N

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_main) : "memory");
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(atail_full, 8);
        for (int i = 0; i < TC_RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            uint32_t slot = 0, ring_phase = 0, a_phase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                TC_UNIT_GEOMETRY();
                const uint8_t* tb = tailb + (size_t)b * Npad * 32;
                mbar_wait(a_empty, a_phase ^ 1);
                mbar_expect_tx(a_full, 2 * TC_MAIN_BYTES);
                tma_load_3d(sAmain, &tm_main, a_full, 0, row0, b);                    // rows >= Npad read as zeros
                tma_load_3d(sAmain + TC_MAIN_BYTES, &tm_main, a_full, 0, row0 + 128, b);
                a_phase ^= 1;
                for (int pass = 0; pass < 2; ++pass)
                    for (int ct = 0; ct < CT; ++ct) {
                        mbar_wait(&ring_empty[slot], ring_phase ^ 1);
                        mbar_expect_tx(&ring_full[slot], TC_SLOT_BYTES);
                        uint8_t* dst = sRing + slot * TC_SLOT_BYTES;
                        tma_load_3d(dst, &tm_main, &ring_full[slot], 0, ct * TC_N, b);
                        bulk_load(dst + TC_MAIN_BYTES, tb + (size_t)ct * TC_TAIL_BYTES, TC_TAIL_BYTES, &ring_full[slot]);
                        if (++slot == TC_RING) { slot = 0; ring_phase ^= 1; }
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // The whole warp runs the loops in uniform control flow (barrier waits by all lanes, descriptors in uniform
        // registers); one elected lane issues the tcgen05 instructions.  A scalar `if (lane == 0)` around the loops
        // makes the compiler serialise every uniform-datapath operand (R2UR + BRA.U.ANY loops, ~200 cycles per MMA).
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint32_t a_lo[2] = {(smem_u32(sAmain) & 0x3ffffu) >> 4, (smem_u32(sAmain + TC_MAIN_BYTES) & 0x3ffffu) >> 4};
        const uint64_t atail[2] = {umma_desc_tail(smem_u32(sAtail)), umma_desc_tail(smem_u32(sAtail + TC_TAIL_BYTES))};
        const uint64_t sw_hi = ((uint64_t)(1024 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29)) << 32 | ((uint64_t)1 << 16);
        uint32_t slot = 0, ring_phase = 0, a_phase = 0, t_phase = 0, tile = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            TC_UNIT_GEOMETRY();
            mbar_wait(a_full, a_phase);
            a_phase ^= 1;
            for (int pass = 0; pass < 2; ++pass) {
                mbar_wait(atail_full, t_phase);                        // this pass's A tails are in place
                t_phase ^= 1;
                tc_fence_after();
                for (int ct = 0; ct < CT; ++ct, ++tile) {
                    const uint32_t par = tile & 1, tphase = (tile >> 1) & 1;
                    mbar_wait(&ring_full[slot], ring_phase);
                    tc_fence_after();
                    const uint32_t bmain = smem_u32(sRing + slot * TC_SLOT_BYTES);
                    const uint32_t b_lo = (bmain & 0x3ffffu) >> 4;
                    const uint64_t btail = umma_desc_tail(bmain + TC_MAIN_BYTES);
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        const uint32_t buf = par * 2 + rb;
                        mbar_wait(&tmem_empty[buf], tphase ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + buf * TC_N;
                        if (leader) {
#pragma unroll
                            for (int s = 0; s < KSTEPS; ++s)           // +32 B per 16-wide K step inside the swizzle atom
                                umma_f16(d, sw_hi | (uint64_t)(a_lo[rb] + 2 * s), sw_hi | (uint64_t)(b_lo + 2 * s), s > 0);
                            umma_f16(d, atail[rb], btail, 1);
                            umma_commit(&tmem_full[buf]);
                        }
                        __syncwarp();
                    }
                    if (leader) umma_commit(&ring_empty[slot]);        // slot reusable once these MMAs retire
                    __syncwarp();
                    if (++slot == TC_RING) { slot = 0; ring_phase ^= 1; }
                }
            }
            if (leader) umma_commit(a_empty);
            __syncwarp();
        }
    } else {
        // ===================================================== epilogue: warps 2-5 row block 0, 6-9 row block 1;
        // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const int rb = (warp - 2) >> 2, quarter = warp & 3;
        const int rblk = quarter * 32 + lane;                      // row inside the 128-row block
        const int rloc = rb * 128 + rblk;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        uint8_t* mytail = sAtail + rb * TC_TAIL_BYTES + tail_offset(rblk, 0);
        const uint32_t qaddr = smem_u32(sQ + rloc * TC_QSTRIDE);
        uint32_t tile = 0;
        const float NEG_INF = __int_as_float(0xff800000);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            TC_UNIT_GEOMETRY();
            const int row = row0 + rloc;
            const bool vrow = row < NV;
            const float xxi = vrow ? xx[(size_t)b * N + row] : 0.f;
            const float xxci = vrow ? xxc[(size_t)b * N + row] : 0.f;
            const float xm = __uint_as_float(scal[4 * b]), xcm = __uint_as_float(scal[4 * b + 1]);
            const float S = __uint_as_float(scal[4 * b + 2]), inv = 1.0f / S;
            const float an = sqrtf(xxci) * inv, bmax = sqrtf(xcm) * inv;
            const __half mi = __float2half_ru(TC_C1 * an);
            {   // pass-1 tail: [1 1 1 0 0 0 -m 0 | 0 x 8]  (the previous unit's MMAs have all retired: its last tile was read)
                __half t[8];
                t[0] = t[1] = t[2] = __float2half_rn(1.0f);
                t[3] = t[4] = t[5] = t[7] = __float2half_rn(0.f);
                t[6] = __hneg(mi);
                uint4 pk;
                memcpy(&pk, t, 16);
                *reinterpret_cast<uint4*>(mytail) = pk;
                *reinterpret_cast<uint4*>(mytail + 128) = make_uint4(0, 0, 0, 0);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(atail_full);
            }
            float gmax[64];
#pragma unroll
            for (int g = 0; g < 64; ++g) gmax[g] = NEG_INF;
            int cnt = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int ct = 0; ct < CT; ++ct, ++tile) {
                    const uint32_t buf = (tile & 1) * 2 + rb, tphase = (tile >> 1) & 1;
                    const uint32_t tcol = tmem_base + tlane + buf * TC_N;
                    mbar_wait(&tmem_full[buf], tphase);
                    tc_fence_after();
                    // four 32-column quarters, TMEM loads double-buffered: quarter q+1 is in flight while q is processed
                    uint32_t ra[2][16], rb2[2][16];
                    tmem_ld16(tcol, ra[0]);
                    tmem_ld16(tcol + 16, rb2[0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        tmem_wait_ld();
                        if (q < 3) {
                            tmem_ld16(tcol + (q + 1) * 32, ra[(q + 1) & 1]);
                            tmem_ld16(tcol + (q + 1) * 32 + 16, rb2[(q + 1) & 1]);
                        } else {                                           // accumulator fully read: hand it back early
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                        }
                        const uint32_t (&a)[16] = ra[q & 1];
                        const uint32_t (&c)[16] = rb2[q & 1];
                        if (pass == 0) {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                gmax[q * 16 + i] = fmax3(gmax[q * 16 + i], __uint_as_float(a[i]), __uint_as_float(c[i]));
                            if (DUMP) {
                                if (dump && vrow)
                                    for (int i = 0; i < 32; ++i) {
                                        const int j = ct * TC_N + q * 32 + i;
                                        const uint32_t v = (i < 16) ? a[i & 15] : c[i & 15];
                                        if (j < N) dump[((size_t)b * N + row) * N + j] = __uint_as_float(v);
                                    }
                            }
                        } else {
                            // acc >= +0  <=>  survivor; column i of the quarter -> bit 31 - i
                            uint32_t hit = ~((sign_bits16(a) << 16) | sign_bits16(c));
                            const int j0 = ct * TC_N + q * 32;
                            while (hit) {                                      // rare: ~k+5 survivors per 4096 columns
                                const int i = __clz(hit);
                                hit &= ~(0x80000000u >> i);
                                if (cnt < TC_QCAP)
                                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(qaddr + 2u * cnt), "h"((uint16_t)(j0 + i)) : "memory");
                                ++cnt;
                            }
                        }
                    }
                }
                if (pass == 0) {
                    // k-th largest DISTINCT group maximum: at least k columns have a lower bound >= tau
                    float tau = __int_as_float(0x7f800000);
                    for (int t = 0; t < K; ++t) {
                        float m = NEG_INF;
#pragma unroll
                        for (int g = 0; g < 64; ++g) m = (gmax[g] < tau) ? fmaxf(m, gmax[g]) : m;
                        tau = m;
                    }
                    const float c0 = TC_C_ACC * bmax * bmax + TC_C_LIN * bmax + TC_C_ABS;
                    const float refdev = 0.5f * inv * inv * (c_ref * sqrtf(xxi * xm) + TC_C_CTR * sqrtf(xm * xcm));
                    float thr = tau - 2.0f * c0 - refdev;
                    thr = fminf(fmaxf(thr, -30000.f), 30000.f);            // finite in fp16; padding columns sit at -60000
                    // pass-2 tail: [1 1 1 u0 u1 u2 +m 0], u0+u1+u2 = -thr
                    __half t[8];
                    t[0] = t[1] = t[2] = __float2half_rn(1.0f);
                    split3_f16(-thr, t[3], t[4], t[5]);
                    t[6] = mi;
                    t[7] = __float2half_rn(0.f);
                    uint4 pk;
                    memcpy(&pk, t, 16);
                    *reinterpret_cast<uint4*>(mytail) = pk;
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(atail_full);
                }
            }
            // queues out: one coalesced 192-byte row per iteration
            if (vrow) qcnt[(size_t)b * N + row] = cnt;
            __syncwarp();
            const int wrow0 = (unit - b * RT) * TC_ROWS + rb * 128 + quarter * 32;
            const uint32_t* sQ32 = reinterpret_cast<const uint32_t*>(sQ + (rb * 128 + quarter * 32) * TC_QSTRIDE);
            for (int r = 0; r < 32; ++r) {
                if (wrow0 + r >= NV) break;
                uint32_t* dst = reinterpret_cast<uint32_t*>(qidx + ((size_t)b * N + wrow0 + r) * TC_QCAP);
                dst[lane] = sQ32[r * (TC_QSTRIDE / 2) + lane];
                if (lane < TC_QCAP / 2 - 32) dst[32 + lane] = sQ32[r * (TC_QSTRIDE / 2) + 32 + lane];
            }
            __syncwarp();
        }
    }
#undef TC_UNIT_GEOMETRY
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ exact re-rank

constexpr int RR_STRIDE = 68;          // floats per staged survivor row: 16-byte aligned, LDS.128 conflict-free

// exact reference score (dgcnn.py:16-18) of column j, whose features are at xj (stride 1), for the query row q,
// as an ascending sort key
__device__ __forceinline__ u64 rerank_key(const float* xj, const float* q, float xxj, float xxi, int j, int F) {
    float acc;
    if ((F & 3) == 0) {
        acc = 0.f;
        for (int f4 = 0; f4 < F / 4; ++f4) {
            const float4 v = reinterpret_cast<const float4*>(xj)[f4];
            const float4 w = reinterpret_cast<const float4*>(q)[f4];
            acc = (f4 == 0) ? __fmul_rn(w.x, v.x) : __fmaf_rn(w.x, v.x, acc);           // sgemm: FMA chain over f
            acc = __fmaf_rn(w.y, v.y, acc);
            acc = __fmaf_rn(w.z, v.z, acc);
            acc = __fmaf_rn(w.w, v.w, acc);
        }
    } else {
        acc = __fmul_rn(q[0], xj[0]);
        for (int f = 1; f < F; ++f) acc = __fmaf_rn(q[f], xj[f], acc);
    }
    const float inner = __fmul_rn(-2.0f, acc);                                         // dgcnn.py:16
    const float pd = __fsub_rn(__fsub_rn(-xxj, inner), xxi);                           // dgcnn.py:18
    return pack_key(f2ord(-pd), (uint32_t)j);
}

// One warp per query row.  The survivors' feature rows are gathered with COALESCED loads (one row per warp
// instruction; a lane-per-row gather costs 25+ L1 wavefronts per instruction) into a padded shared-memory tile,
// then each lane runs the reference's sequential FMA chain for one survivor out of shared memory, and the
// (value, index) keys are sorted by rank counting across the warp: the lane with rank r < K writes idx[r].
// Rows whose queue overflowed are re-ranked by an exact full scan with the same WarpList as the CUDA-core kernels.
// dynamic smem: 8 warps x (32 x RR_STRIDE + 64) floats
__global__ void __launch_bounds__(256, 3)       // 71 KB of shared memory allow three CTAs per SM anyway: up to 80 registers, no spills
knn_tc_rerank_kernel(const float* __restrict__ xt, const float* __restrict__ xx, const int32_t* __restrict__ qcnt,
                     const uint16_t* __restrict__ qidx, int N, int F, int K, const int32_t* __restrict__ n_valid,
                     int32_t* __restrict__ idx, int32_t* __restrict__ stats) {
    extern __shared__ __align__(16) float rr_smem[];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const int NV = len_valid(n_valid, b, N);
    if (i >= NV) { len_fill_row(idx, nullptr, ((size_t)b * N + i) * K, K, lane); return; }
    float* xs = rr_smem + warp * (32 * RR_STRIDE + 64);      // [32][RR_STRIDE]
    float* myq = xs + 32 * RR_STRIDE;                         // [64]
    const float* __restrict__ xb = xt + (size_t)b * N * F;
    const float* __restrict__ xxb = xx + (size_t)b * N;
    for (int f = lane; f < F; f += 32) myq[f] = xb[(size_t)i * F + f];
    const float xxi = xxb[i];
    const int cnt = qcnt[(size_t)b * N + i];
    const bool overflow = cnt > TC_QCAP;
    if (stats && lane == 0) {
        atomicAdd(&stats[0], cnt);
        if (overflow) atomicAdd(&stats[1], 1);
    }
    int32_t* __restrict__ out = idx + ((size_t)b * N + i) * K;
    if (!overflow) {
        constexpr int R = TC_QCAP / 32;
        const uint16_t* __restrict__ q = qidx + ((size_t)b * N + i) * TC_QCAP;
        u64 key[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            key[r] = PCNBR_KEY_MAX;
            if (r * 32 < cnt) {                                       // warp-uniform
                const int c = r * 32 + lane;
                const int j = (c < cnt) ? (int)q[c] : 0;
                const float xxj = xxb[j];
                const int n = min(32, cnt - r * 32);
                __syncwarp();
                if (F == 64) {
                    // half a warp per row: one LDG.128 + one STS.128 move two 256-byte rows; 8 loads in flight
                    const int hl = lane & 15, hi = lane >> 4;
                    for (int l0 = 0; l0 < n; l0 += 16) {
                        float4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int jj = __shfl_sync(PCNBR_FULL, j, min(l0 + 2 * u + hi, n - 1));
                            v[u] = reinterpret_cast<const float4*>(xb + (size_t)jj * 64)[hl];
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (l0 + 2 * u + hi < n)
                                reinterpret_cast<float4*>(xs + (l0 + 2 * u + hi) * RR_STRIDE)[hl] = v[u];
                    }
                } else {
                    for (int l0 = 0; l0 < n; l0 += 8) {               // 8 coalesced rows (16 loads) in flight per batch
                        float v0[8], v1[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float* __restrict__ src = xb + (size_t)__shfl_sync(PCNBR_FULL, j, min(l0 + u, n - 1)) * F;
                            v0[u] = (lane < F) ? src[lane] : 0.f;
                            v1[u] = (lane + 32 < F) ? src[lane + 32] : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (l0 + u < n) {
                                xs[(l0 + u) * RR_STRIDE + lane] = v0[u];
                                xs[(l0 + u) * RR_STRIDE + lane + 32] = v1[u];
                            }
                    }
                }
                __syncwarp();
                if (c < cnt) key[r] = rerank_key(xs + lane * RR_STRIDE, myq, xxj, xxi, j, F);
            }
        }
        if (cnt <= 32) {
            // the common case (k + 9..12 survivors): one key per lane, a 15-step bitonic sort instead of cnt rounds of rank
            // counting (2 shuffles + a 64-bit compare per survivor and register: ~360 of the kernel's ~960 instructions per row)
            const u64 sorted = warp_sort64(key[0], lane);
            if (lane < K) out[lane] = (int32_t)min((uint32_t)sorted, (uint32_t)(NV - 1));    // K <= 32 on this path; cnt >= K: the survivors contain the top K
            return;
        }
        int rank[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rank[r] = 0;
#pragma unroll
        for (int sr = 0; sr < R; ++sr) {
            if (sr * 32 < cnt) {                                      // warp-uniform
                const int n = min(32, cnt - sr * 32);
                for (int l = 0; l < n; ++l) {
                    const u64 other = shfl64(key[sr], l);
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r * 32 < cnt) rank[r] += (other < key[r]) ? 1 : 0;       // warp-uniform guard
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r * 32 + lane < cnt && rank[r] < K) out[rank[r]] = (int32_t)(uint32_t)key[r];
        return;
    }
    __syncwarp();
    WarpList<1> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;
    for (int c0 = 0; c0 < NV; c0 += 32) {
        const int c = c0 + lane;
        u64 key = PCNBR_KEY_MAX;
        if (c < NV) {
            const float* __restrict__ xj = xb + (size_t)c * F;        // direct (uncoalesced) reads: rare path
            float acc = __fmul_rn(myq[0], xj[0]);
            for (int f = 1; f < F; ++f) acc = __fmaf_rn(myq[f], xj[f], acc);
            const float inner = __fmul_rn(-2.0f, acc);
            const float pd = __fsub_rn(__fsub_rn(-xxb[c], inner), xxi);
            key = pack_key(f2ord(-pd), (uint32_t)c);
        }
        uint32_t pass = __ballot_sync(PCNBR_FULL, key < thr);
        while (pass) {
            const int src = __ffs(pass) - 1;
            pass &= pass - 1;
            const u64 cand = shfl64(key, src);
            if (cand < thr) {
                list.insert(cand, lane);
                thr = list.at(K - 1);
            }
        }
    }
    if (lane < K) out[lane] = (int32_t)min((uint32_t)list.v[0], (uint32_t)(NV - 1));
}

// ------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// (64, Npad, B) fp16 tensor, box = 64 halves (128 B) x 128 rows x 1 cloud, 128-byte swizzle; rows >= Npad read as zeros.
static int make_map(CUtensorMap* map, const __half* base, int B, int Npad) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[3] = {(cuuint64_t)TC_KMAIN, (cuuint64_t)Npad, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)TC_KMAIN * 2, (cuuint64_t)Npad * TC_KMAIN * 2};
    cuuint32_t box[3] = {TC_KMAIN, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TcWorkspace {
    float *xx, *xxc, *part, *xt;
    __half* ymain;
    uint8_t* tailb;
    uint32_t* scal;
    int32_t *qcnt, *stats;
    uint16_t* qidx;
    size_t bytes;
};

static inline int tc_npad(int N) { return (N + 127) & ~127; }

static TcWorkspace tc_carve(void* ws, int B, int F, int N, bool need_xt) {
    const int Npad = tc_npad(N);
    TcWorkspace w;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align256(n); return (uint8_t*)ws + o; };
    const size_t bn = (size_t)B * N, bp = (size_t)B * Npad;
    w.xx = (float*)take(bn * 4);
    w.xxc = (float*)take(bn * 4);
    w.part = (float*)take((size_t)B * TC_MEAN_CHUNKS * TC_PART * 4);
    w.scal = (uint32_t*)take((size_t)B * 16);
    w.stats = (int32_t*)take(16);
    w.ymain = (__half*)take(bp * TC_KMAIN * 2);
    w.tailb = (uint8_t*)take(bp * 32);
    w.xt = need_xt ? (float*)take(bn * F * 4) : nullptr;
    w.qcnt = (int32_t*)take(bn * 4);
    w.qidx = (uint16_t*)take(bn * TC_QCAP * 2);
    w.bytes = off;
    return w;
}

size_t knn_tc_ws_bytes(int B, int F, int N) { return tc_carve(nullptr, B, F, N, true).bytes; }

bool knn_tc_supported(int F, int N, int K) {
    return F >= 1 && F <= 64 && K <= 32 && N >= 256 && N <= 65535;     // F is zero-padded to a multiple of 16 for the MMA
}

template <int KSTEPS, bool DUMP>
static int launch_tc(const CUtensorMap& mm, const TcWorkspace& w, int B, int F, int N, int K, float c_ref, const int32_t* n_valid,
                     float* dump, cudaStream_t s) {
    const size_t smem = 2 * TC_MAIN_BYTES + 2 * TC_TAIL_BYTES + TC_RING * TC_SLOT_BYTES + TC_ROWS * TC_QSTRIDE * 2 + 40 * 8 + 1024;
    cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<KSTEPS, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int Npad = tc_npad(N);
    const int units = B * ((Npad + TC_ROWS - 1) / TC_ROWS);
    const int grid = units < sms ? units : sms;
    // K4: 2 N^2 F flop per cloud counted ONCE (whatever the pass count issues); compulsory bytes: the
    // fp16 operands (160 B per point) and the survivor queues
    PCNBR_TIMED("knn_tc_kernel", s, (double)B * N * (160.0 + 4.0 + 2.0 * TC_QCAP), 2.0 * B * (double)N * N * F,
                (knn_tc_kernel<KSTEPS, DUMP><<<grid, TC_THREADS, smem, s>>>(mm, w.tailb, w.xx, w.xxc, w.scal, B, N, Npad, K, c_ref, n_valid,
                                                                              w.qcnt, w.qidx, dump)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// x: (B,F,N) with strides (sf, sn).  idx (B,N,K).  dump: optional (B,N,N) pass-1 tensor-core scores (tests).
int knn_tc_run(const float* x, int B, int F, int N, long sf, long sn, int K, const int32_t* n_valid, int32_t* idx, void* ws,
               float* dump, int32_t* stats_out, cudaStream_t s) {
    const bool point_major = (sf == 1 && sn == F);
    TcWorkspace w = tc_carve(ws, B, F, N, !point_major);
    const float* xt = point_major ? x : w.xt;
    const int Npad = tc_npad(N);
    cudaError_t e = cudaMemsetAsync(w.scal, 0, (size_t)((uint8_t*)w.stats + 16 - (uint8_t*)w.scal), s);   // scal + stats
    if (e != cudaSuccess) return (int)e;
    // exact |x|^2 in the reference's summation order (select.cu); channel means; centred, scaled fp16 operands
    int rc0 = launch_sumsq(x, B, F, N, sf, sn, n_valid, w.xx, s);
    if (rc0) return rc0;
    PCNBR_TIMED("knn_tc_mean_kernel", s, 4.0 * B * (double)N * F, (double)B * N * F,
                (knn_tc_mean_kernel<<<dim3(TC_MEAN_CHUNKS, B), 256, 0, s>>>(x, F, N, sf, sn, n_valid, w.part)));
    PCNBR_CHECK_LAUNCH();
    int pb = (Npad + 7) / 8;
    if (pb > 148) pb = 148;
    PCNBR_TIMED("knn_tc_prep_kernel", s, (double)B * N * (4.0 * F + 160.0 + 8.0), 6.0 * B * (double)N * F,
                (knn_tc_prep_kernel<<<dim3(pb, B), 256, 0, s>>>(x, w.part, w.xx, F, N, Npad, sf, sn, n_valid, point_major ? nullptr : w.xt,
                                                                w.ymain, w.tailb, w.xxc, w.scal)));
    PCNBR_CHECK_LAUNCH();
    CUtensorMap mm;
    int rc = make_map(&mm, w.ymain, B, Npad);
    if (rc) return rc;
    const float c_ref = 2.0f * (float)(F + 8) * 5.9604645e-8f;             // 2 (F+8) 2^-24, see TC_C_* above
    switch ((F + 15) / 16) {
        case 1:  rc = dump ? launch_tc<1, true>(mm, w, B, F, N, K, c_ref, n_valid, dump, s) : launch_tc<1, false>(mm, w, B, F, N, K, c_ref, n_valid, dump, s); break;
        case 2:  rc = dump ? launch_tc<2, true>(mm, w, B, F, N, K, c_ref, n_valid, dump, s) : launch_tc<2, false>(mm, w, B, F, N, K, c_ref, n_valid, dump, s); break;
        case 3:  rc = dump ? launch_tc<3, true>(mm, w, B, F, N, K, c_ref, n_valid, dump, s) : launch_tc<3, false>(mm, w, B, F, N, K, c_ref, n_valid, dump, s); break;
        default: rc = dump ? launch_tc<4, true>(mm, w, B, F, N, K, c_ref, n_valid, dump, s) : launch_tc<4, false>(mm, w, B, F, N, K, c_ref, n_valid, dump, s); break;
    }
    if (rc) return rc;
    const size_t rr_smem = 8 * (32 * RR_STRIDE + 64) * sizeof(float);
    e = cudaFuncSetAttribute(knn_tc_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr_smem);
    if (e != cudaSuccess) return (int)e;
    PCNBR_TIMED("knn_tc_rerank_kernel", s, (double)B * N * (4.0 * F + 8.0 + 2.0 * TC_QCAP + 4.0 * K), 2.0 * B * (double)N * F * K,
                (knn_tc_rerank_kernel<<<dim3((N + 7) / 8, B), 256, rr_smem, s>>>(xt, w.xx, w.qcnt, w.qidx, N, F, K, n_valid, idx, w.stats)));
    PCNBR_CHECK_LAUNCH();
    if (stats_out) {
        e = cudaMemcpyAsync(stats_out, w.stats, 8, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace pcnbr
