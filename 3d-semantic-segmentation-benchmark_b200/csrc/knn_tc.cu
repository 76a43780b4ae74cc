// knn_tc.cu -- K4: feature-space kNN (DGCNN layers 2-4) on the 5th-generation tensor cores.
//
// Reference: models/dgcnn/dgcnn.py:16-20 -- a dense (N,N) = X^T X contraction (cuBLAS/MKL SGEMM), two
// broadcast adds and torch.topk over 4096-wide rows: 3 x 64 MB of temporaries per cloud.
//
// Here the contraction runs on tcgen05 and the (N,N) matrix only ever exists as 128x256 tiles in TMEM:
//
//   producer warp : TMA (cp.async.bulk.tensor, 128-byte swizzle) of K-major operand slabs into a
//                   4-slot shared-memory ring, mbarrier full/empty pipeline;
//   MMA warp      : one thread issues tcgen05.mma kind::tf32, M=128 x N=256 x K=8, accumulating the
//                   3xTF32 split  hi*hi' + lo*hi' + hi*lo'  (~fp32 accuracy) into one of two TMEM
//                   accumulators (2 x 256 columns = all 512 TMEM columns);
//   4 epilogue warps: tcgen05.ld the finished tile (thread = query row) while the next tile is being
//                   multiplied, form the ranking score s = 2*x'_i.x'_j - |x'_j|^2 (x' = x - per-cloud channel
//                   mean: distances are translation invariant, and centring keeps the TF32 error relative to
//                   the SPREAD of the features instead of their common offset), and
//        pass 1: keep 64 interleaved group maxima per row in registers (branch-free); the k-th largest
//                of them is a lower bound tau on the row's k-th best score;
//        pass 2: (tiles are recomputed -- cheaper than storing N^2 floats) append every column with
//                s >= tau - margin to a small per-row queue (~k+4 survivors of 4096).
//
// A second kernel re-ranks each row's survivors with the reference's EXACT arithmetic (sequential fp32
// FMA chain over the feature index, ATen's cascade |x|^2, ((-xx_j) - inner) - xx_i) and sorts them by
// (value, index).  `margin` bounds twice the worst-case deviation of the tensor-core score from the
// exact one, so the survivor set provably contains the exact top-k: the indices are bit-identical to
// oracle/canon.c even though the bulk of the flops ran in TF32.  Rows whose queue overflows (degenerate
// clouds) fall back to an exact full scan inside the re-rank kernel.
#include "common.cuh"
#include <cuda.h>

namespace pcnbr {

constexpr int TC_M = 128;          // query rows per work unit (TMEM lanes)
constexpr int TC_N = 256;          // candidate columns per tile (TMEM columns per accumulator)
constexpr int TC_RING = 4;         // ring slots of 256 rows x 32 floats (32 KB)
constexpr int TC_QCAP = 64;        // survivors kept per row (uint16 indices)
constexpr int TC_THREADS = 192;    // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr uint32_t TC_SLAB_A = 128 * 128;      // bytes: 128 rows x 128 B (one swizzle atom of K)
constexpr uint32_t TC_SLAB_B = 256 * 128;
// Survivor margin (see header): 2 x |tensor-core score - exact reference score| is bounded by
//   C_FILT * |x'_i| max|x'_j|   3xTF32 split (3*2^-22) + fp32 accumulation of 24 MMAs, centred features x' = x - mean
// + C_REF  * |x_i|  max|x_j|    rounding of the REFERENCE's own fp32 value (64-term FMA chain, cascade |x|^2): its
//                               ranking deviates from the true distances by that much, and we must follow it
// + C_CTR  * max|x_j| max|x'_j| rounding of the centring subtraction itself
//                               = 2 (F + 8) 2^-24 (F chain roundings + cascade |x|^2 + the two subtractions), 8.6e-6 at F=64
constexpr float TC_C_FILT = 5e-5f, TC_C_CTR = 1e-6f;

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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) in [32,46),
// version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)(1024 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}
// kind::tf32, D = fp32, A/B K-major, M = 128, N = 256 (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

// ------------------------------------------------------------------------------------ operand preparation

// Per-cloud channel sums over a slice of the points: part[b][chunk][f] (fixed order -> deterministic).
constexpr int TC_MEAN_CHUNKS = 16;
__global__ void __launch_bounds__(256)
knn_tc_mean_kernel(const float* __restrict__ x, int F, int N, long sf, long sn, float* __restrict__ part) {
    __shared__ float red[256];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int groups = 256 / F;                                   // F <= 64
    const int f = threadIdx.x % F, g = threadIdx.x / F;
    const int per = (N + TC_MEAN_CHUNKS - 1) / TC_MEAN_CHUNKS;
    const int n0 = chunk * per, n1 = min(N, n0 + per);
    const float* __restrict__ xb = x + (size_t)b * F * N;
    float acc = 0.f;
    if (g < groups)
        for (int n = n0 + g; n < n1; n += groups) acc += xb[(size_t)f * sf + (size_t)n * sn];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (g == 0) {
        for (int k = 1; k < groups; ++k) acc += red[k * F + f];
        part[((size_t)b * TC_MEAN_CHUNKS + chunk) * F + f] = acc;
    }
}

// One warp per point: x' = x - mean; xhi = tf32(x'), xlo = tf32(x' - xhi) (point-major, K-major for the
// MMA); xt = exact copy of x (only when the input is not already point-major); xxc = |x'|^2;
// maxima of |x|^2 and |x'|^2 per cloud (bit patterns of non-negative floats order like the values).
__global__ void __launch_bounds__(256)
knn_tc_prep_kernel(const float* __restrict__ x, const float* __restrict__ part, const float* __restrict__ xx,
                   int F, int Fp, int N, long sf, long sn, float* __restrict__ xt, float* __restrict__ xhi,
                   float* __restrict__ xlo, float* __restrict__ xxc, uint32_t* __restrict__ maxes) {
    __shared__ float mu[64];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < F) {
        float m = 0.f;
        for (int c = 0; c < TC_MEAN_CHUNKS; ++c) m += part[((size_t)b * TC_MEAN_CHUNKS + c) * F + threadIdx.x];
        mu[threadIdx.x] = m / (float)N;
    }
    __syncthreads();
    const float* __restrict__ xb = x + (size_t)b * F * N;
    float mx = 0.f, mxc = 0.f;
    for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
        float ss = 0.f;
        for (int f = lane; f < Fp; f += 32) {                     // channels F..Fp-1 are zero padding of the K dimension
            float v = 0.f, vc = 0.f;
            if (f < F) {
                v = xb[(size_t)f * sf + (size_t)n * sn];
                vc = __fsub_rn(v, mu[f]);
                if (xt) xt[((size_t)b * N + n) * F + f] = v;
            }
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(vc));
            const float rem = __fsub_rn(vc, __uint_as_float(h));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rem));
            const size_t o = ((size_t)b * N + n) * Fp + f;
            xhi[o] = __uint_as_float(h);
            xlo[o] = __uint_as_float(l);
            ss = fmaf(vc, vc, ss);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) ss += __shfl_xor_sync(PCNBR_FULL, ss, d);
        if (lane == 0) xxc[(size_t)b * N + n] = ss;
        mxc = fmaxf(mxc, ss);
        mx = fmaxf(mx, xx[(size_t)b * N + n]);
    }
    if (lane == 0) {
        atomicMax(&maxes[2 * b], __float_as_uint(mx));
        atomicMax(&maxes[2 * b + 1], __float_as_uint(mxc));
    }
}

// ------------------------------------------------------------------------------------ main kernel

template <int KATOMS, bool DUMP>      // F = 32 * KATOMS; DUMP: also write the raw scores (tests only)
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
              const float* __restrict__ xx, const float* __restrict__ xxc, const uint32_t* __restrict__ maxes,
              int B, int N, int K, float c_ref,
              int32_t* __restrict__ qcnt, uint16_t* __restrict__ qidx, float* __restrict__ dump) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (128B-swizzle atoms) by OFFSET, so the compiler keeps the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sAhi = smem;                                         // KATOMS x 16 KB
    uint8_t* sAlo = sAhi + KATOMS * TC_SLAB_A;                    // KATOMS x 16 KB
    uint8_t* sRing = sAlo + KATOMS * TC_SLAB_A;                   // TC_RING x 32 KB
    uint16_t* sQ = (uint16_t*)(sRing + TC_RING * TC_SLAB_B);      // 128 x 64 x 2 B
    float* sXX = (float*)(sQ + TC_M * TC_QCAP);                   // 2 x 256 floats
    uint64_t* bars = (uint64_t*)(sXX + 2 * TC_N);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* ring_full = bars + 2;                               // [TC_RING]
    uint64_t* ring_empty = bars + 2 + TC_RING;                    // [TC_RING]
    uint64_t* tmem_full = bars + 2 + 2 * TC_RING;                 // [2]
    uint64_t* tmem_empty = bars + 4 + 2 * TC_RING;                // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 6 + 2 * TC_RING);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RT = (N + TC_M - 1) / TC_M, CT = (N + TC_N - 1) / TC_N;
    const int units = B * RT;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_lo) : "memory");
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int i = 0; i < TC_RING; ++i) { mbar_init(&ring_full[i], 1); mbar_init(&ring_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
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
                const int b = unit / RT, row0 = (unit - b * RT) * TC_M;
                mbar_wait(a_empty, a_phase ^ 1);
                mbar_expect_tx(a_full, 2 * KATOMS * TC_SLAB_A);
                for (int a = 0; a < KATOMS; ++a) {
                    tma_load_3d(sAhi + a * TC_SLAB_A, &tm_hi, a_full, a * 32, row0, b);
                    tma_load_3d(sAlo + a * TC_SLAB_A, &tm_lo, a_full, a * 32, row0, b);
                }
                a_phase ^= 1;
                for (int pass = 0; pass < 2; ++pass)
                    for (int ct = 0; ct < CT; ++ct)
                        for (int arr = 0; arr < 2; ++arr)
                            for (int a = 0; a < KATOMS; ++a) {
                                mbar_wait(&ring_empty[slot], ring_phase ^ 1);
                                mbar_expect_tx(&ring_full[slot], TC_SLAB_B);
                                uint8_t* dst = sRing + slot * TC_SLAB_B;
                                const CUtensorMap* tm = arr ? &tm_lo : &tm_hi;
                                tma_load_3d(dst, tm, &ring_full[slot], a * 32, ct * TC_N, b);
                                tma_load_3d(dst + TC_SLAB_A, tm, &ring_full[slot], a * 32, ct * TC_N + 128, b);
                                if (++slot == TC_RING) { slot = 0; ring_phase ^= 1; }
                            }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            uint32_t slot = 0, ring_phase = 0, a_phase = 0, tile = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                mbar_wait(a_full, a_phase);
                a_phase ^= 1;
                tc_fence_after();
                for (int pass = 0; pass < 2; ++pass)
                    for (int ct = 0; ct < CT; ++ct, ++tile) {
                        const uint32_t buf = tile & 1, tphase = (tile >> 1) & 1;
                        mbar_wait(&tmem_empty[buf], tphase ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + buf * TC_N;
                        uint32_t acc = 0;
                        for (int arr = 0; arr < 2; ++arr)
                            for (int a = 0; a < KATOMS; ++a) {
                                mbar_wait(&ring_full[slot], ring_phase);
                                tc_fence_after();
                                const uint32_t bb = smem_u32(sRing + slot * TC_SLAB_B);
                                const uint32_t ah = smem_u32(sAhi + a * TC_SLAB_A), al = smem_u32(sAlo + a * TC_SLAB_A);
#pragma unroll
                                for (int s = 0; s < 4; ++s) {              // hi(A) x {hi,lo}(B)
                                    umma_tf32(d, umma_desc(ah + s * 32), umma_desc(bb + s * 32), acc);
                                    acc = 1;
                                }
                                if (arr == 0) {
#pragma unroll
                                    for (int s = 0; s < 4; ++s)            // lo(A) x hi(B)
                                        umma_tf32(d, umma_desc(al + s * 32), umma_desc(bb + s * 32), 1);
                                }
                                umma_commit(&ring_empty[slot]);            // slot reusable once these MMAs retire
                                if (++slot == TC_RING) { slot = 0; ring_phase ^= 1; }
                            }
                        umma_commit(&tmem_full[buf]);
                    }
                umma_commit(a_empty);
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const int et = threadIdx.x - 64;
        const int rloc = quarter * 32 + lane;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        uint32_t tile = 0;
        const float NEG_INF = __int_as_float(0xff800000);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int b = unit / RT, row = (unit - b * RT) * TC_M + rloc;
            const bool vrow = row < N;
            const float xxi = vrow ? xx[(size_t)b * N + row] : 0.f;
            const float xxci = vrow ? xxc[(size_t)b * N + row] : 0.f;
            float gmax[64];
#pragma unroll
            for (int g = 0; g < 64; ++g) gmax[g] = NEG_INF;
            float thr = 0.f;
            int cnt = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int ct = 0; ct < CT; ++ct, ++tile) {
                    const uint32_t buf = tile & 1, tphase = (tile >> 1) & 1;
                    const int j0 = ct * TC_N;
                    float* sx = sXX + buf * TC_N;                   // staged -|x'_j|^2 (masked columns: -inf)
                    for (int t = et; t < TC_N; t += 128)
                        sx[t] = (j0 + t < N) ? -xxc[(size_t)b * N + j0 + t] : NEG_INF;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    mbar_wait(&tmem_full[buf], tphase);
                    tc_fence_after();
                    const float4* sx4 = reinterpret_cast<const float4*>(sx);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + tlane + buf * TC_N + q * 32, r);
                        tmem_wait_ld();
                        if (pass == 0) {
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                const float4 nx = sx4[q * 8 + i4];          // one broadcast LDS.128 per 4 columns
                                const float s0 = fmaf(2.0f, __uint_as_float(r[4 * i4 + 0]), nx.x);
                                const float s1 = fmaf(2.0f, __uint_as_float(r[4 * i4 + 1]), nx.y);
                                const float s2 = fmaf(2.0f, __uint_as_float(r[4 * i4 + 2]), nx.z);
                                const float s3 = fmaf(2.0f, __uint_as_float(r[4 * i4 + 3]), nx.w);
                                const int g = (q & 1) * 32 + 4 * i4;
                                gmax[g + 0] = fmaxf(gmax[g + 0], s0);
                                gmax[g + 1] = fmaxf(gmax[g + 1], s1);
                                gmax[g + 2] = fmaxf(gmax[g + 2], s2);
                                gmax[g + 3] = fmaxf(gmax[g + 3], s3);
                                if (DUMP) {
                                    const float sv[4] = {s0, s1, s2, s3};
                                    for (int u = 0; u < 4; ++u) {
                                        const int j = j0 + q * 32 + 4 * i4 + u;
                                        if (dump && vrow && j < N) dump[((size_t)b * N + row) * N + j] = sv[u];
                                    }
                                }
                            }
                        } else {
                            uint32_t hit = 0;                              // branch-free filter: one bit per column
#pragma unroll
                            for (int i4 = 0; i4 < 8; ++i4) {
                                const float4 nx = sx4[q * 8 + i4];
                                hit |= (fmaf(2.0f, __uint_as_float(r[4 * i4 + 0]), nx.x) >= thr ? 1u : 0u) << (4 * i4 + 0);
                                hit |= (fmaf(2.0f, __uint_as_float(r[4 * i4 + 1]), nx.y) >= thr ? 1u : 0u) << (4 * i4 + 1);
                                hit |= (fmaf(2.0f, __uint_as_float(r[4 * i4 + 2]), nx.z) >= thr ? 1u : 0u) << (4 * i4 + 2);
                                hit |= (fmaf(2.0f, __uint_as_float(r[4 * i4 + 3]), nx.w) >= thr ? 1u : 0u) << (4 * i4 + 3);
                            }
                            while (hit) {                                  // rare: ~k+4 survivors per 4096 columns
                                const int i = __ffs(hit) - 1;
                                hit &= hit - 1;
                                if (cnt < TC_QCAP) sQ[rloc * TC_QCAP + cnt] = (uint16_t)(j0 + q * 32 + i);
                                ++cnt;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                if (pass == 0) {
                    // k-th largest DISTINCT group maximum: at least k columns score >= tau
                    float tau = __int_as_float(0x7f800000);
                    for (int t = 0; t < K; ++t) {
                        float m = NEG_INF;
#pragma unroll
                        for (int g = 0; g < 64; ++g) m = (gmax[g] < tau) ? fmaxf(m, gmax[g]) : m;
                        tau = m;
                    }
                    const float xm = __uint_as_float(maxes[2 * b]), xcm = __uint_as_float(maxes[2 * b + 1]);
                    const float margin = TC_C_FILT * sqrtf(xxci * xcm) + c_ref * sqrtf(xxi * xm) + TC_C_CTR * sqrtf(xm * xcm);
                    thr = fmaxf(tau - margin, -3.0e38f);           // finite: masked columns (s = -inf) never pass
                }
            }
            if (vrow) {
                qcnt[(size_t)b * N + row] = cnt;
                const uint4* src = reinterpret_cast<const uint4*>(sQ + rloc * TC_QCAP);
                uint4* dst = reinterpret_cast<uint4*>(qidx + ((size_t)b * N + row) * TC_QCAP);
#pragma unroll
                for (int i = 0; i < TC_QCAP * 2 / 16; ++i) dst[i] = src[i];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------ exact re-rank

// One warp per query row: exact reference arithmetic on the survivors (or on all N columns when the
// queue overflowed), sorted by (value, index) with the same WarpList as the CUDA-core kernels.
__global__ void __launch_bounds__(256)
knn_tc_rerank_kernel(const float* __restrict__ xt, const float* __restrict__ xx, const int32_t* __restrict__ qcnt,
                     const uint16_t* __restrict__ qidx, int N, int F, int K, int32_t* __restrict__ idx,
                     int32_t* __restrict__ stats) {
    extern __shared__ float sq[];                 // [8 warps][F]
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const float* __restrict__ xb = xt + (size_t)b * N * F;
    float* myq = sq + warp * F;
    for (int f = lane; f < F; f += 32) myq[f] = xb[(size_t)i * F + f];
    __syncwarp();
    const float xxi = xx[(size_t)b * N + i];
    const int cnt = qcnt[(size_t)b * N + i];
    const bool overflow = cnt > TC_QCAP;
    const int total = overflow ? N : cnt;
    if (stats && lane == 0) {
        atomicAdd(&stats[0], cnt);
        if (overflow) atomicAdd(&stats[1], 1);
    }
    WarpList<1> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;
    for (int c0 = 0; c0 < total; c0 += 32) {
        const int c = c0 + lane;
        u64 key = PCNBR_KEY_MAX;
        if (c < total) {
            const int j = overflow ? c : (int)qidx[((size_t)b * N + i) * TC_QCAP + c];
            float acc = 0.f;
            if ((F & 3) == 0) {
                const float4* __restrict__ xj = reinterpret_cast<const float4*>(xb + (size_t)j * F);
                for (int f4 = 0; f4 < F / 4; ++f4) {
                    const float4 v = xj[f4];
                    acc = (f4 == 0) ? __fmul_rn(myq[0], v.x) : __fmaf_rn(myq[4 * f4], v.x, acc);   // sgemm: FMA chain over f
                    acc = __fmaf_rn(myq[4 * f4 + 1], v.y, acc);
                    acc = __fmaf_rn(myq[4 * f4 + 2], v.z, acc);
                    acc = __fmaf_rn(myq[4 * f4 + 3], v.w, acc);
                }
            } else {
                const float* __restrict__ xj = xb + (size_t)j * F;
                acc = __fmul_rn(myq[0], xj[0]);
                for (int f = 1; f < F; ++f) acc = __fmaf_rn(myq[f], xj[f], acc);
            }
            const float inner = __fmul_rn(-2.0f, acc);                                         // dgcnn.py:16
            const float pd = __fsub_rn(__fsub_rn(-xx[(size_t)b * N + j], inner), xxi);         // dgcnn.py:18
            key = pack_key(f2ord(-pd), (uint32_t)j);
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
    if (lane < K) idx[((size_t)b * N + i) * K + lane] = (int32_t)(uint32_t)list.v[0];
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

// (F, N, B) fp32 tensor, box = 32 floats x 128 rows x 1 cloud, 128-byte swizzle; rows >= N read as zeros.
static int make_map(CUtensorMap* map, const float* base, int B, int N, int F) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[3] = {(cuuint64_t)F, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)F * 4, (cuuint64_t)N * F * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct TcWorkspace {
    float *xx, *xxc, *part, *xt, *xhi, *xlo;
    uint32_t* maxes;
    int32_t *qcnt, *stats;
    uint16_t* qidx;
    size_t bytes;
};

static inline int tc_padded(int F) { return F <= 32 ? 32 : 64; }

static TcWorkspace tc_carve(void* ws, int B, int F, int N, bool need_xt) {
    const int Fp = tc_padded(F);
    TcWorkspace w;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += align256(n); return (uint8_t*)ws + o; };
    const size_t bn = (size_t)B * N;
    w.xx = (float*)take(bn * 4);
    w.xxc = (float*)take(bn * 4);
    w.part = (float*)take((size_t)B * TC_MEAN_CHUNKS * 64 * 4);
    w.maxes = (uint32_t*)take((size_t)B * 8);
    w.stats = (int32_t*)take(16);
    w.xhi = (float*)take(bn * Fp * 4);
    w.xlo = (float*)take(bn * Fp * 4);
    w.xt = need_xt ? (float*)take(bn * F * 4) : nullptr;
    w.qcnt = (int32_t*)take(bn * 4);
    w.qidx = (uint16_t*)take(bn * TC_QCAP * 2);
    w.bytes = off;
    return w;
}

size_t knn_tc_ws_bytes(int B, int F, int N) { return tc_carve(nullptr, B, F, N, true).bytes; }

bool knn_tc_supported(int F, int N, int K) {
    return F >= 1 && F <= 64 && K <= 32 && N >= 256 && N <= 65535;     // F is zero-padded to 32 or 64 for the MMA
}

template <int KATOMS, bool DUMP>
static int launch_tc(const CUtensorMap& mh, const CUtensorMap& ml, const TcWorkspace& w, int B, int N, int K,
                     float c_ref, float* dump, cudaStream_t s) {
    const size_t smem = 2 * KATOMS * TC_SLAB_A + TC_RING * TC_SLAB_B + TC_M * TC_QCAP * 2 + 2 * TC_N * 4 + 32 * 8 + 1024;
    cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<KATOMS, DUMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = B * ((N + TC_M - 1) / TC_M);
    const int grid = units < sms ? units : sms;
    // K4: 2 N^2 F flop per cloud counted ONCE (whatever the split / pass count issues); compulsory
    // bytes: the operands (4 N F) and the survivor queues
    PCNBR_TIMED("knn_tc_kernel", s, (double)B * N * (4.0 * 32 * KATOMS + 4.0 + 2.0 * TC_QCAP), 2.0 * B * (double)N * N * (32.0 * KATOMS),
                (knn_tc_kernel<KATOMS, DUMP><<<grid, TC_THREADS, smem, s>>>(mh, ml, w.xx, w.xxc, w.maxes, B, N, K, c_ref, w.qcnt, w.qidx, dump)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// x: (B,F,N) with strides (sf, sn).  idx (B,N,K).  dump: optional (B,N,N) tensor-core scores (tests).
int knn_tc_run(const float* x, int B, int F, int N, long sf, long sn, int K, int32_t* idx, void* ws,
               float* dump, int32_t* stats_out, cudaStream_t s) {
    const bool point_major = (sf == 1 && sn == F);
    TcWorkspace w = tc_carve(ws, B, F, N, !point_major);
    const float* xt = point_major ? x : w.xt;
    cudaError_t e = cudaMemsetAsync(w.maxes, 0, (size_t)((uint8_t*)w.stats + 16 - (uint8_t*)w.maxes), s);   // maxes + stats
    if (e != cudaSuccess) return (int)e;
    // exact |x|^2 in the reference's summation order (select.cu); channel means; centred TF32 split
    int rc0 = launch_sumsq(x, B, F, N, sf, sn, w.xx, s);
    if (rc0) return rc0;
    PCNBR_TIMED("knn_tc_mean_kernel", s, 4.0 * B * (double)N * F, (double)B * N * F,
                (knn_tc_mean_kernel<<<dim3(TC_MEAN_CHUNKS, B), 256, 0, s>>>(x, F, N, sf, sn, w.part)));
    PCNBR_CHECK_LAUNCH();
    int pb = (N + 7) / 8;
    if (pb > 148) pb = 148;
    const int Fp = tc_padded(F);
    PCNBR_TIMED("knn_tc_prep_kernel", s, (double)B * N * (4.0 * F + 8.0 * Fp + 8.0), 6.0 * B * (double)N * F,
                (knn_tc_prep_kernel<<<dim3(pb, B), 256, 0, s>>>(x, w.part, w.xx, F, Fp, N, sf, sn, point_major ? nullptr : w.xt, w.xhi,
                                                                w.xlo, w.xxc, w.maxes)));
    PCNBR_CHECK_LAUNCH();
    CUtensorMap mh, ml;
    int rc = make_map(&mh, w.xhi, B, N, Fp);
    if (rc) return rc;
    rc = make_map(&ml, w.xlo, B, N, Fp);
    if (rc) return rc;
    const float c_ref = 2.0f * (float)(F + 8) * 5.9604645e-8f;             // 2 (F+8) 2^-24, see TC_C_* above
    if (dump) rc = (Fp == 64) ? launch_tc<2, true>(mh, ml, w, B, N, K, c_ref, dump, s) : launch_tc<1, true>(mh, ml, w, B, N, K, c_ref, dump, s);
    else      rc = (Fp == 64) ? launch_tc<2, false>(mh, ml, w, B, N, K, c_ref, dump, s) : launch_tc<1, false>(mh, ml, w, B, N, K, c_ref, dump, s);
    if (rc) return rc;
    PCNBR_TIMED("knn_tc_rerank_kernel", s, (double)B * N * (4.0 * F + 8.0 + 2.0 * TC_QCAP + 4.0 * K), 2.0 * B * (double)N * F * K,
                (knn_tc_rerank_kernel<<<dim3((N + 7) / 8, B), 256, 8 * F * sizeof(float), s>>>(xt, w.xx, w.qcnt, w.qidx, N, F, K, idx, w.stats)));
    PCNBR_CHECK_LAUNCH();
    if (stats_out) {
        e = cudaMemcpyAsync(stats_out, w.stats, 8, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace pcnbr
