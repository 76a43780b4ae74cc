// gemm_h2.cu -- fp32-accurate GEMM on the tensor cores with a TWO-TERM FP16 SPLIT (kind::f16: twice the TF32 rate).
//
// Reference: the wide 1x1 convolutions of models/dgcnn/dgcnn.py:95-126 (conv5 384->1024, conv6 1408->512, conv7 512->256
// over 65536 rows) are tensor-bound GEMMs.  gemm_tc.cu runs them as three TF32 products and sits at that design's ceiling
// (89 % of the TF32 issue rate).  kind::f16 issues at twice the TF32 rate, and an fp32 value splits into two fp16 terms
// as exactly as into two tf32 terms once the tensor is brought into fp16's range:
//     y  = x * s            s = power of two with  max|x| * s  in [2^14, 2^15)      (exact)
//     hi = fp16_rn(y)       lo = fp16_rn(y - hi)                                   y = hi + lo + e,  |e| <= 2^-24 |y|
//     C  = (hi.hi' + lo.hi' + hi.lo') / (s s')                                     dropped: lo.lo' <= 2^-24 |y||y'|
// (for |y| < 2^-9 the residual lo falls into fp16 subnormals: its ABSOLUTE error is 2^-25, i.e. 2^-40 of the tensor's
// largest element -- irrelevant against the 2^-24 of the large elements it is added to).  Per product the error is
// ~3 * 2^-24, better than the 3xTF32 kernel's 2^-21 (whose `hi` is a truncation).
//
// max|x| per operand comes from pcnbr_absmax_f32 (one extra read of the tensor: 256 per-block maxima, reduced here).
// Structure as gemm_tc.cu: warp 0 TMA producer (fp32 tiles exactly as they lie in HBM, K-major or MN-major), warp 1 MMA
// issuer, warps 2-5 epilogue (tcgen05.ld -> * 1/(s s') (+ bias) -> swizzled staging -> TMA store), warps 6-13 (6-21 when both
// operands are converted in the kernel: two groups on alternate stages) converters.
// The converters rewrite each landed fp32 tile IN PLACE (4 B per element before and after) as [hi | lo] fp16 tiles in the
// K-major 64-byte-swizzle UMMA layout -- MN-major operands are transposed on the way, so the MMA only ever sees K-major
// fp16 -- with one named barrier between "everything read into registers" and "first store".
#include "gemm_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace pcnbr {

constexpr int H2_AMAX_SLOTS = 1280;        // per-block maxima: absmax_kernel fills 256, the fused producers (bnact.cu) up to 1184
constexpr int H2_CONV_THREADS = 32 * GM_CONV_WARPS;          // threads of ONE converter group
// Converter groups: a group of 8 warps converts a whole ring stage (its in-place rewrite needs one barrier over everybody
// who reads the stage).  When both operands are converted in the kernel (the weight gradients: 12288 values per stage at
// BN = 256, ~80 % of the stage's MMA time in issue slots alone) ONE group leaves two warps per scheduler to hide the
// LDS -> ALU -> STS latencies and the barrier, and the MMA warp waits for it; TWO groups take alternate stages, so the
// conversion of stage i+1 overlaps the tail of stage i.  With a pre-split B (forward / input gradient) one group is plenty.
#ifndef H2_PRE_GROUPS
#define H2_PRE_GROUPS 1
#endif
template <bool B_PRE> constexpr int h2_groups() { return B_PRE ? H2_PRE_GROUPS : 2; }
// + one "plane writer" warp behind the converters of the pre-split-B variants (forward / input gradient): it stores the
// converted A tiles to HBM for the weight-gradient GEMM (see PRE2 below) when the caller asks for them.
template <bool B_PRE, bool PRE2 = false> constexpr int h2_threads() {
    return PRE2 ? 192 : 192 + h2_groups<B_PRE>() * H2_CONV_THREADS + (B_PRE ? 32 : 0);
}

// kind::f16 (A, B = fp16, both K-major), D = fp32, M = 128, N = BN (cute::UMMA::InstrDescriptor)
template <int BN, bool MN = false>
__device__ __forceinline__ void h2_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, bool accumulate) {
    constexpr uint32_t idesc = (1u << 4) | ((MN ? 3u : 0u) << 15) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, 64-byte swizzle (rows of 32 fp16 = 64 B, 8-row groups 512 B apart; layout type 4 = SWIZZLE_64B);
// a K step of 16 fp16 = +32 B inside the swizzle atom
__device__ __forceinline__ uint64_t h2_desc(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)(512 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)4 << 29);
    return lo | (hi << 32);
}
// MN-major fp16 tile as TMA SWIZZLE_128B boxes of 64 MN-elements (128 B) x 32 K rows: an 8-row group is one swizzle atom
// (1024 B, SBO), the next 64 MN-elements are one box (4096 B, LBO) further; layout type 2 = SWIZZLE_128B; a K step of 16
// = two atoms = +2048 B
__device__ __forceinline__ uint64_t h2_desc_mn(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(4096 >> 4) << 16);
    const uint64_t hi = (uint64_t)(1024 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}
// byte offset of the 16-byte chunk g (8 fp16 of K) of row r in a [rows x 32 fp16] K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t h2_dst_off(int r, int g) { return (uint32_t)(r * 64 + ((g ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void h2_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void h2_tma_load_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(gm_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// scale s = 2^e with amax * s in [2^14, 2^15), and 1/s; degenerate tensors (all zero / denormal / non-finite) get s = 1
__device__ __forceinline__ void h2_scale_of(float amax, float& s, float& inv) {
    const uint32_t E = (__float_as_uint(amax) >> 23) & 0xffu;
    if (E < 16u || E == 255u) { s = 1.f; inv = 1.f; return; }
    s = __uint_as_float((268u - E) << 23);
    inv = __uint_as_float((E - 14u) << 23);
}

// 8 consecutive K values of one row -> hi / lo fp16 chunks (16 bytes each)
__device__ __forceinline__ void h2_split8(const float (&v)[8], float s, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float y0 = __fmul_rn(v[2 * j], s), y1 = __fmul_rn(v[2 * j + 1], s);
        const __half2 hh = __floats2half2_rn(y0, y1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(__fsub_rn(y0, hf.x), __fsub_rn(y1, hf.y));
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Read the 8 K values [8g, 8g+8) of row `r` of a landed fp32 tile (TMA SWIZZLE_128B in both layouts).
//   K-major : rows of 32 floats (128 B); tiles of more than 128 rows are stacked 128-row boxes.
//   MN-major: chunks of 32 rows; inside a chunk K row k holds the 32 row-values as 128 B (so a lane per row reads
//             consecutive words: conflict-free).
template <bool MN>
__device__ __forceinline__ void h2_load8(const uint8_t* tile, int r, int g, float (&v)[8]) {
    if (MN) {
        const uint8_t* ch = tile + (r >> 5) * GM_CHUNK;
        const int rl = r & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * g + j;
            v[j] = *reinterpret_cast<const float*>(ch + k * 128 + (((rl >> 2) ^ (k & 7)) << 4) + (rl & 3) * 4);
        }
    } else {
        const uint8_t* row = tile + (r >> 7) * GM_SLAB + (r & 127) * 128;
        const float4 a = *reinterpret_cast<const float4*>(row + (((2 * g) ^ (r & 7)) << 4));
        const float4 b = *reinterpret_cast<const float4*>(row + (((2 * g + 1) ^ (r & 7)) << 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}

// one operand tile of `ROWS` rows, a task = (row, 8-wide K group): load phase / store phase of the in-place conversion.
// (Tried: MN-major tiles as (4 rows, K group) tasks with one LDS.128 per K row -- a quarter of the load instructions, but the
// stores of four rows of the same parity land on half of the banks (2-4-way conflicts): 369 us instead of 238 us for the
// conv6 weight gradient.  The converters move 4 B in and 4 B out per value through the LSU, 768 wavefronts per stage at
// BN = 256 -- as many cycles as the stage's six MMAs: the weight-gradient GEMMs are bound by that, not by issue.)
template <int ROWS> constexpr int h2_per_thread() { return (ROWS * 4 + H2_CONV_THREADS - 1) / H2_CONV_THREADS; }

template <bool MN, int ROWS, int NT>
__device__ __forceinline__ void h2_tile_load(const uint8_t* tile, int t, float (&v)[NT][8]) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int q = t + i * H2_CONV_THREADS;
        if (q < ROWS * 4) h2_load8<MN>(tile, q % ROWS, q / ROWS, v[i]);
    }
}
template <int ROWS, int NT>
__device__ __forceinline__ void h2_tile_store(uint8_t* tile, uint32_t plane, int t, float s, const float (&v)[NT][8]) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int q = t + i * H2_CONV_THREADS;
        if (q < ROWS * 4) {
            uint4 hi, lo;
            h2_split8(v[i], s, hi, lo);
            const uint32_t off = h2_dst_off(q % ROWS, q / ROWS);
            *reinterpret_cast<uint4*>(tile + off) = hi;
            *reinterpret_cast<uint4*>(tile + plane + off) = lo;
        }
    }
}

// Optional wait-time trace (tools/gemm_shapes.py --trace): when set, every role of every CTA adds the clock cycles it spent
// waiting on each of its barriers into trace[blockIdx.x][slot] (16 slots of 8 bytes per CTA):
//   0 producer: ring slot free   1 converters (group 0, warp 0): TMA landed   2 MMA: accumulator free   3 MMA: stage converted
//   4 epilogue: accumulator full   5 epilogue: staging slab free (bulk store read)   6 CTA lifetime   7 converter busy time
//   8 CTA lifetime in ns of %globaltimer (6 / 8 = the SM clock the kernel ran at)
__device__ unsigned long long* g_h2_trace = nullptr;

// B_PRE: the B operand arrives already split ([hi | lo] fp16 planes written once by split_f16_kernel -- the weights of a
// forward / input-gradient GEMM, which every CTA would otherwise convert again): TMA drops the two planes straight into
// the UMMA layout (SWIZZLE_64B) and the converters only handle A (a third of the work at BN = 256).
// PRE2 (the weight gradients dW = gy^T x): BOTH operands arrive as [hi | lo] fp16 planes in their natural row-major
// (rows = K = points, columns = M or N = channels) layout, i.e. MN-major -- written by the plane-writer warps of the
// input-gradient GEMM (gy) and of the forward GEMM (x), whose converters had the split tiles in shared memory anyway.
// tcgen05 reads MN-major fp16 straight from the TMA boxes, so this variant has no converter warps at all (they were its
// bound: 4 B in + 4 B out per value through the LSU for BOTH operands, 0.45-0.5 of the pipe).
// planes_out (B_PRE variants, K-major A; bit 0: A, bit 1: A2): tm_ap / tm_ap2 = the planes to write (units of the first column tile).
template <int BN, int STAGES, bool A_MN, bool B_MN, bool B_PRE, bool PRE2 = false>
__global__ void __launch_bounds__(h2_threads<B_PRE, PRE2>(), 1)
gemm2h_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
              const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_c,
              const __grid_constant__ CUtensorMap tm_ap, const __grid_constant__ CUtensorMap tm_ap2, int planes_out,
              int M, int N, int K,
              int kb_split, int splits, const float* __restrict__ bias, const float* __restrict__ amax_a,
              const float* __restrict__ amax_a2, const float* __restrict__ amax_b) {
    static_assert(!PRE2 || (A_MN && B_MN && !B_PRE && BN >= 64 && BN % 64 == 0), "PRE2: both operands MN-major planes");
    constexpr int WRITER_WARP = 6 + h2_groups<B_PRE>() * GM_CONV_WARPS;        // B_PRE variants only
    const bool write_planes = B_PRE && !A_MN && !PRE2 && planes_out != 0;
    extern __shared__ uint8_t gm_smem_raw[];
    uint8_t* smem = gm_smem_raw + ((1024u - (gm_smem_u32(gm_smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t A_BYTES = GM_SLAB, B_BYTES = (uint32_t)BN * 128u;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;               // fp32 landing zone == [hi | lo] fp16 afterwards
    constexpr int NBUF = (512 / BN) < 8 ? (512 / BN) : 8;             // TMEM accumulators of BN columns
    uint8_t* cstage = smem + STAGES * STAGE_BYTES;                    // 2 x (128 rows x 128 B), swizzled: TMA-store staging
    uint64_t* bars = (uint64_t*)(cstage + 2 * GM_SLAB);
    uint64_t* full = bars;                                            // [STAGES]  TMA landed          (count 1 + tx)
    uint64_t* conv = bars + STAGES;                                   // [STAGES]  fp16 tiles ready     (count GM_CONV_WARPS)
    uint64_t* empty = bars + 2 * STAGES;                              // [STAGES]  MMAs retired        (count 1)
    uint64_t* tmem_full = bars + 3 * STAGES;                          // [NBUF]
    uint64_t* tmem_empty = bars + 3 * STAGES + NBUF;                  // [NBUF]
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * STAGES + 2 * NBUF);
    float* s_scale = (float*)(tmem_slot + 2);                         // {s_a, s_b, 1/s_a, 1/s_b}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = (M + GM_BM - 1) / GM_BM, NT = (N + BN - 1) / BN;
    const int KB = (K + GM_BK - 1) / GM_BK;
    const int kb_per = (KB + splits - 1) / splits;
    const int units = MT * NT * splits;
    unsigned long long* const trace = g_h2_trace;
    const long long t_cta = trace ? clock64() : 0;
    unsigned long long ns_cta = 0;
    if (trace) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_cta));

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
        if (write_planes) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ap2) : "memory");
        }
        // empty: the stage's MMAs have retired (+ the plane writer's bulk stores have read it)
        for (int i = 0; i < STAGES; ++i) { gm_mbar_init(&full[i], 1); gm_mbar_init(&conv[i], GM_CONV_WARPS); gm_mbar_init(&empty[i], write_planes ? 2 : 1); }
        for (int i = 0; i < NBUF; ++i) { gm_mbar_init(&tmem_full[i], 1); gm_mbar_init(&tmem_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 2) {
        // per-tensor scales from the per-block maxima of pcnbr_absmax_f32 (A and its concatenated second half share one)
        float ma = 0.f, mb = 0.f;
        for (int i = lane; i < H2_AMAX_SLOTS; i += 32) {
            ma = fmaxf(ma, amax_a[i]);
            if (amax_a2) ma = fmaxf(ma, amax_a2[i]);
            mb = fmaxf(mb, amax_b[i]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            ma = fmaxf(ma, __shfl_xor_sync(PCNBR_FULL, ma, d));
            mb = fmaxf(mb, __shfl_xor_sync(PCNBR_FULL, mb, d));
        }
        if (lane == 0) {
            h2_scale_of(ma, s_scale[0], s_scale[2]);
            h2_scale_of(mb, s_scale[1], s_scale[3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer (fp32 tiles, as in gemm3x_kernel)
        if (lane == 0) {
            H2Wait w_ring(trace, 0);
            uint32_t stage = 0, phase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int nt = unit % NT, mt = (unit / NT) % MT, sp = unit / (MT * NT);
                const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const long long t0 = w_ring.begin();
                    gm_mbar_wait(&empty[stage], phase ^ 1);
                    w_ring.end(t0);
                    gm_mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
                    const uint32_t st = gm_smem_u32(smem + stage * STAGE_BYTES);
                    if (PRE2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int j = 0; j < GM_BM / 64; ++j)
                                h2_tma_load_3d(st + h * (A_BYTES / 2) + j * 4096, &tm_a, &full[stage], mt * GM_BM + j * 64, kb * GM_BK, h);
                    } else if (A_MN) {
#pragma unroll
                        for (int c = 0; c < GM_BM / 32; ++c)
                            gm_tma_load_2d(st + c * GM_CHUNK, &tm_a, &full[stage], mt * GM_BM + c * 32, kb * GM_BK);
                    } else {
                        if (kb < kb_split) gm_tma_load_2d(st, &tm_a, &full[stage], kb * GM_BK, mt * GM_BM);
                        else               gm_tma_load_2d(st, &tm_a2, &full[stage], (kb - kb_split) * GM_BK, mt * GM_BM);
                    }
                    const uint32_t sb = st + A_BYTES;
                    if (PRE2) {
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int j = 0; j < BN / 64; ++j)
                                h2_tma_load_3d(sb + h * (B_BYTES / 2) + j * 4096, &tm_b, &full[stage], nt * BN + j * 64, kb * GM_BK, h);
                    } else if (B_PRE) {
                        h2_tma_load_3d(sb, &tm_b, &full[stage], kb * GM_BK, nt * BN, 0);                      // hi plane
                        h2_tma_load_3d(sb + B_BYTES / 2, &tm_b, &full[stage], kb * GM_BK, nt * BN, 1);        // lo plane
                    } else if (B_MN) {
#pragma unroll
                        for (int c = 0; c < BN / 32; ++c)
                            gm_tma_load_2d(sb + c * GM_CHUNK, &tm_b, &full[stage], nt * BN + c * 32, kb * GM_BK);
                    } else {
#pragma unroll
                        for (int h = 0; h < (BN + 127) / 128; ++h)
                            gm_tma_load_2d(sb + h * GM_SLAB, &tm_b, &full[stage], kb * GM_BK, nt * BN + h * 128);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            w_ring.flush();
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer: 3 products x 2 K steps of 16 per stage
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        uint32_t stage = 0, phase = 0, tile = 0;
        H2Wait w_acc(lane == 0 ? trace : nullptr, 2), w_conv(lane == 0 ? trace : nullptr, 3);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++tile) {
            const int sp = unit / (MT * NT);
            const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
            const uint32_t buf = tile % NBUF, tphase = (tile / NBUF) & 1;
            long long t0 = w_acc.begin();
            gm_mbar_wait(&tmem_empty[buf], tphase ^ 1);
            w_acc.end(t0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + buf * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                const uint32_t st = gm_smem_u32(smem + stage * STAGE_BYTES);
                const uint64_t ah = PRE2 ? h2_desc_mn(st) : h2_desc(st), al = PRE2 ? h2_desc_mn(st + A_BYTES / 2) : h2_desc(st + A_BYTES / 2);
                const uint64_t bh = PRE2 ? h2_desc_mn(st + A_BYTES) : h2_desc(st + A_BYTES);
                const uint64_t bl = PRE2 ? h2_desc_mn(st + A_BYTES + B_BYTES / 2) : h2_desc(st + A_BYTES + B_BYTES / 2);
                constexpr uint64_t KS = PRE2 ? (2048 >> 4) : 2;       // descriptor step of 16 K values
                t0 = w_conv.begin();
                if (PRE2) gm_mbar_wait(&full[stage], phase);          // planes land ready for the MMA
                else      gm_mbar_wait(&conv[stage], phase);          // [hi | lo] tiles written and published
                w_conv.end(t0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (leader) {
#pragma unroll
                    for (int s = 0; s < 2; ++s) h2_umma<BN, PRE2>(d, ah + KS * s, bh + KS * s, kb > kb0 || s > 0);    // hi . hi'
#pragma unroll
                    for (int s = 0; s < 2; ++s) h2_umma<BN, PRE2>(d, al + KS * s, bh + KS * s, true);                 // lo . hi'
#pragma unroll
                    for (int s = 0; s < 2; ++s) h2_umma<BN, PRE2>(d, ah + KS * s, bl + KS * s, true);                 // hi . lo'
                    gm_umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (leader) gm_umma_commit(&tmem_full[buf]);
            __syncwarp();
        }
        w_acc.flush(); w_conv.flush();
    } else if (warp < 6) {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        const int rloc = quarter * 32 + lane;
        const bool issuer = (warp == 2 && lane == 0);
        const float inv_a = s_scale[2], inv_b = s_scale[3];
        uint32_t tile = 0, slab = 0;
        H2Wait w_full(issuer ? trace : nullptr, 4), w_slab(issuer ? trace : nullptr, 5);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++tile) {
            const int nt = unit % NT, mt = (unit / NT) % MT, sp = unit / (MT * NT);
            const uint32_t buf = tile % NBUF, tphase = (tile / NBUF) & 1;
            const int ncols = min(BN, N - nt * BN);
            const int nq = (ncols + 31) / 32;
            const long long t0 = w_full.begin();
            gm_mbar_wait(&tmem_full[buf], tphase);
            w_full.end(t0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int q = 0; q < nq; ++q, ++slab) {
                uint32_t r[32];
                gm_tmem_ld32(tmem_base + tlane + buf * BN + q * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (q == nq - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) gm_mbar_arrive(&tmem_empty[buf]);
                }
                const int c0 = nt * BN + q * 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = __fmul_rn(__fmul_rn(__uint_as_float(r[i]), inv_a), inv_b);      // exact: powers of two
                    if (bias && c0 + i < N) v += __ldg(bias + c0 + i);
                    r[i] = __float_as_uint(v);
                }
                uint8_t* sb = cstage + (slab & 1) * GM_SLAB;
                if (issuer) {
                    const long long t1 = w_slab.begin();
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    w_slab.end(t1);
                }
                gm_epi_barrier();
                uint8_t* rowp = sb + rloc * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(rowp + ((c ^ (rloc & 7)) << 4)) = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                gm_epi_barrier();
                if (issuer) {
                    gm_tma_store_3d(&tm_c, gm_smem_u32(sb), nt * BN + q * 32, mt * GM_BM, sp);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        w_full.flush(); w_slab.flush();
    } else if (B_PRE && warp == WRITER_WARP) {
        // ===================================================== plane writer: converted A tiles -> HBM (first column tile only)
        // Every A tile (mt, kb) is converted by the NT units of its row of tiles; the one with nt == hash(mt) % NT stores it (a
        // fixed nt -- or mt % NT -- would put all the stores on a subset of the CTAs: the grid stride is a multiple of NT).  The ring slot
        // is handed back one stage LATE -- when the next stage's stores have been issued and this stage's have read their
        // shared memory -- so the writer never sits between a conversion and the slot's release.
        if (write_planes && lane == 0) {
            uint32_t stage = 0, phase = 0;
            int prev = -1;                                            // stage whose release is still owed
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int nt = unit % NT, mt = (unit / NT) % MT, sp = unit / (MT * NT);
                const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
                const bool mine = nt == (int)((((uint32_t)mt * 2654435761u) >> 16) % (uint32_t)NT);
                for (int kb = kb0; kb < kb1; ++kb) {
                    gm_mbar_wait(&conv[stage], phase);                // the converters fenced their writes for the async proxy
                    if (mine && (planes_out & (kb < kb_split ? 1 : 2))) {
                        const uint32_t st = gm_smem_u32(smem + stage * STAGE_BYTES);
                        const CUtensorMap* map = kb < kb_split ? &tm_ap : &tm_ap2;
                        const int kc = (kb < kb_split ? kb : kb - kb_split) * GM_BK;
                        h2_tma_store_3d(map, st, kc, mt * GM_BM, 0);
                        h2_tma_store_3d(map, st + A_BYTES / 2, kc, mt * GM_BM, 1);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");       // (possibly empty) group of this stage
                    if (prev >= 0) {
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // all but this stage's group have read
                        gm_mbar_arrive(&empty[prev]);
                    }
                    prev = (int)stage;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (prev >= 0) {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                gm_mbar_arrive(&empty[prev]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        __syncwarp();
    } else if (!PRE2) {
        // ===================================================== converters: fp32 tile -> [hi | lo] fp16 tiles, in place
        constexpr int NG = h2_groups<B_PRE>();
        const int grp = (threadIdx.x - 192) / H2_CONV_THREADS;        // this warp's group: it converts the stages seq % NG == grp
        const int t = (threadIdx.x - 192) % H2_CONV_THREADS;          // 0 .. 255 inside the group
        constexpr int TA = h2_per_thread<GM_BM>();                    // conversion tasks per thread (see h2_tile_load)
        constexpr int TB = B_PRE ? 0 : h2_per_thread<BN>();
        const float sa = s_scale[0], sbs = s_scale[1];
        uint32_t stage = 0, phase = 0, seq = 0;
        H2Wait w_land(threadIdx.x == 192 ? trace : nullptr, 1), w_busy(threadIdx.x == 192 ? trace : nullptr, 7);
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
            const int sp = unit / (MT * NT);
            const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
            for (int kb = kb0; kb < kb1; ++kb, ++seq) {
                if (NG > 1 && (int)(seq % NG) != grp) {               // the other group's stage
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    continue;
                }
                uint8_t* st = smem + stage * STAGE_BYTES;
                uint8_t* sb = st + A_BYTES;
                long long t0 = w_land.begin();
                gm_mbar_wait(&full[stage], phase);
                w_land.end(t0);
                t0 = w_busy.begin();
                // One operand tile at a time (each is rewritten in place inside its own region): every fp32 word of the tile is
                // in registers before the first store -- one named barrier per group and tile.  Keeping the two tiles apart
                // keeps the live registers low (the CTA's 22 warps leave 80 each).
                {
                    float va[TA][8];
                    h2_tile_load<A_MN, GM_BM, TA>(st, t, va);
                    if (grp == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
                    else          asm volatile("bar.sync 3, 256;" ::: "memory");
                    h2_tile_store<GM_BM, TA>(st, A_BYTES / 2, t, sa, va);
                }
                if (TB > 0) {
                    float vb[TB > 0 ? TB : 1][8];
                    h2_tile_load<B_MN, BN, (TB > 0 ? TB : 1)>(sb, t, vb);
                    if (grp == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
                    else          asm volatile("bar.sync 3, 256;" ::: "memory");
                    h2_tile_store<BN, (TB > 0 ? TB : 1)>(sb, B_BYTES / 2, t, sbs, vb);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) gm_mbar_arrive(&conv[stage]);
                w_busy.end(t0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        w_land.flush(); w_busy.flush();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (trace && threadIdx.x == 0) {
        unsigned long long ns1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns1));
        trace[(size_t)blockIdx.x * 16 + 6] = (unsigned long long)(clock64() - t_cta);
        trace[(size_t)blockIdx.x * 16 + 8] = ns1 - ns_cta;
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// partial[blockIdx.x] = max |x| over this block's share of a (rows x cols) fp32 matrix with row pitch ld (cols, ld
// multiples of 4, 16-byte aligned base); every one of the H2_AMAX_SLOTS slots is written
__global__ void __launch_bounds__(512)
absmax_kernel(const float* __restrict__ x, long rows, int cols, long ld, float* __restrict__ partial) {
    __shared__ float red[16];
    const int c4 = cols >> 2;
    const long total = rows * c4;
    float m = 0.f;
    if (ld == cols) {
        const float4* __restrict__ v = reinterpret_cast<const float4*>(x);
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
            const float4 a = v[i];
            m = fmaxf(fmaxf(m, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        }
    } else {
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
            const long r = i / c4;
            const int c = (int)(i - r * c4);
            const float4 a = *reinterpret_cast<const float4*>(x + r * ld + 4 * c);
            m = fmaxf(fmaxf(m, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(PCNBR_FULL, m, d));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 16 ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int d = 8; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(PCNBR_FULL, m, d));
        if (threadIdx.x == 0) {
            partial[blockIdx.x] = m;
            for (int i = blockIdx.x + gridDim.x; i < H2_AMAX_SLOTS; i += gridDim.x) partial[i] = 0.f;     // slots no block owns
        }
    }
}

// out[plane][r][k] (plane 0 = hi, 1 = lo; row pitch ldo halfs, plane pitch `plane` halfs) = the split of
// src[r][k] (transpose = 0) or src[k][r] (transpose = 1), scaled by the power of two derived from the per-block maxima --
// the same derivation as in gemm2h_kernel, so the epilogue's 1/s matches.  Weights only: a few MB.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ src, int rows, int cols, long ld, int transpose, const float* __restrict__ amax,
                 __half* __restrict__ out, long ldo, long plane) {
    __shared__ float s_s;
    if (threadIdx.x < 32) {
        float m = 0.f;
        for (int i = threadIdx.x; i < H2_AMAX_SLOTS; i += 32) m = fmaxf(m, amax[i]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(PCNBR_FULL, m, d));
        float s, inv;
        h2_scale_of(m, s, inv);
        if (threadIdx.x == 0) s_s = s;
    }
    __syncthreads();
    const float s = s_s;
    const long total = (long)rows * cols;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        // consecutive threads walk the SOURCE's inner dimension (coalesced reads; the scattered 2-byte writes of the
        // transposed case stay in L2: the matrices are a few MB)
        int r, k;
        float v;
        if (transpose) { k = (int)(i / rows); r = (int)(i - (long)k * rows); v = src[(long)k * ld + r]; }
        else           { r = (int)(i / cols); k = (int)(i - (long)r * cols); v = src[(long)r * ld + k]; }
        const float y = __fmul_rn(v, s);
        const __half h = __float2half_rn(y);
        const __half l = __float2half_rn(__fsub_rn(y, __half2float(h)));
        out[(long)r * ldo + k] = h;
        out[plane + (long)r * ldo + k] = l;
    }
}

// ------------------------------------------------------------------------------------ host side

// pre-split B: fp16 tensor (K, rows, 2 planes), box = 32 halfs x box_rows x 1 plane, 64-byte swizzle = the UMMA layout
static int h2_make_map_split(CUtensorMap* map, const void* base, long K, long rows, long ldo, long plane, int box_rows) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, 2};
    cuuint64_t gstr[2] = {(cuuint64_t)ldo * 2, (cuuint64_t)plane * 2};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// fp32 matrix of `outer` rows x `inner` contiguous floats, box = 32 floats x box_rows, plain 128-byte swizzle (the
// converters, not the MMA, read the landed tile); out-of-range elements read as 0
static int h2_make_map(CUtensorMap* map, const float* base, long inner, long outer, long ld, int box_rows) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// MN-major planes (the PRE2 operands): fp16 tensor (cols = M or N, rows = K, 2 planes), box = 64 halfs x 32 rows x 1 plane,
// 128-byte swizzle; out-of-range elements read as 0
static int h2_make_map_mnsplit(CUtensorMap* map, const void* base, long cols, long rows, long ld, long plane) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

struct H2Planes {                       // planes of A (bit 0 of on) and A2 (bit 1) the forward / input-gradient kernel writes
    CUtensorMap ap, ap2;
    int on;
};

template <int BN, int STAGES, bool A_MN, bool B_MN, bool B_PRE, bool PRE2 = false>
static int h2_launch(const CUtensorMap& ta, const CUtensorMap& ta2, int kb_split, const CUtensorMap& tb, const CUtensorMap& tc,
                     int M, int N, int K, int splits, const float* bias, const float* amax_a, const float* amax_a2,
                     const float* amax_b, const H2Planes& pl, cudaStream_t s) {
    const size_t smem = (size_t)STAGES * (GM_SLAB + (size_t)BN * 128) + 2 * GM_SLAB + 64 * 8 + 64 + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm2h_kernel<BN, STAGES, A_MN, B_MN, B_PRE, PRE2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = ((M + GM_BM - 1) / GM_BM) * ((N + BN - 1) / BN) * splits;
    const int grid = units < sms ? units : sms;
    // algorithmic work as gemm3x_kernel: 2 M N K flop counted once (3 fp16 products are issued); operands once + result
    const double bytes = 4.0 * ((double)M * K + (double)N * K + (double)M * N * splits), flops = 2.0 * M * (double)N * K;
    // shape class of the profiler label: tensor-bound when the three fp16 products the kernel issues per product need longer on
    // the f16 pipe (1356.7 TFLOP/s sustained, MEASURED_PEAKS.json) than the operands need on HBM
    const char* name = 3.0 * flops / 1356.7e12 > bytes / 6551e9 ? "gemm2h_kernel[tensor]" : "gemm2h_kernel[hbm]";
    PCNBR_TIMED(name, s, bytes, flops,
                (gemm2h_kernel<BN, STAGES, A_MN, B_MN, B_PRE, PRE2><<<grid, h2_threads<B_PRE, PRE2>(), smem, s>>>(
                    ta, ta2, tb, tc, pl.ap, pl.ap2, pl.on, M, N, K, kb_split, splits, bias, amax_a, amax_a2, amax_b)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// Tile width.  As gm_tile_n, plus 192 columns where that wastes fewer padding columns than 256 (N = 384: two full tiles
// instead of one and a half -- the 384-channel skip concatenation of DGCNNWithColor is the N of conv5's input gradient
// and weight gradient and of conv6's first K block: a quarter of their MMA time was spent on zero columns).  Not for a
// raw K-major B, whose TMA boxes are 128-row slabs.
static int h2_tile_n(int M, int N, int K, bool b_kmajor_raw) {
    static const bool off = getenv("PCNBR_H2_NO192") != nullptr;               // A/B switch (tools/gemm_shapes.py)
    const int bn = gm_tile_n(M, N, K);
    if (bn == 256 && !off && !b_kmajor_raw && ((N + 191) / 192) * 192 < ((N + 255) / 256) * 256) return 192;
    return bn;
}

template <bool A_MN, bool B_MN, bool B_PRE>
static int h2_dispatch(const CUtensorMap& ta, const CUtensorMap& ta2, int kb_split, const CUtensorMap& tb, const CUtensorMap& tc,
                       int M, int N, int K, int splits, const float* bias, const float* amax_a, const float* amax_a2,
                       const float* amax_b, const H2Planes& pl, cudaStream_t s) {
    switch (h2_tile_n(M, N, K, !B_MN && !B_PRE)) {                           // stages: 48 / 40 / 32 / 24 / 20 KB each beside 32 KB of staging
        case 256: return h2_launch<256, 4, A_MN, B_MN, B_PRE>(ta, ta2, kb_split, tb, tc, M, N, K, splits, bias, amax_a, amax_a2, amax_b, pl, s);
        case 192:
            if constexpr (B_MN || B_PRE)
                return h2_launch<192, 4, A_MN, B_MN, B_PRE>(ta, ta2, kb_split, tb, tc, M, N, K, splits, bias, amax_a, amax_a2, amax_b, pl, s);
            else
                return PCNBR_E_BADARG;
        case 128: return h2_launch<128, 5, A_MN, B_MN, B_PRE>(ta, ta2, kb_split, tb, tc, M, N, K, splits, bias, amax_a, amax_a2, amax_b, pl, s);
        case 64:  return h2_launch<64, 6, A_MN, B_MN, B_PRE>(ta, ta2, kb_split, tb, tc, M, N, K, splits, bias, amax_a, amax_a2, amax_b, pl, s);
        default:  return h2_launch<32, 6, A_MN, B_MN, B_PRE>(ta, ta2, kb_split, tb, tc, M, N, K, splits, bias, amax_a, amax_a2, amax_b, pl, s);
    }
}

// both operands as MN-major planes (weight gradients): tile widths of whole 64-column boxes
static bool h2_pre2_tile_ok(int M, int N, int K) { return h2_tile_n(M, N, K, false) >= 64; }
static int h2_dispatch_pre2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, int M, int N, int K, int splits,
                            const float* amax_a, const float* amax_b, cudaStream_t s) {
    H2Planes none;
    none.ap = ta; none.ap2 = ta; none.on = 0;
    const int kb_split = (K + GM_BK - 1) / GM_BK;
    switch (h2_tile_n(M, N, K, false)) {
        case 256: return h2_launch<256, 4, true, true, false, true>(ta, ta, kb_split, tb, tc, M, N, K, splits, nullptr, amax_a, nullptr, amax_b, none, s);
        case 192: return h2_launch<192, 4, true, true, false, true>(ta, ta, kb_split, tb, tc, M, N, K, splits, nullptr, amax_a, nullptr, amax_b, none, s);
        case 128: return h2_launch<128, 5, true, true, false, true>(ta, ta, kb_split, tb, tc, M, N, K, splits, nullptr, amax_a, nullptr, amax_b, none, s);
        case 64:  return h2_launch<64, 6, true, true, false, true>(ta, ta, kb_split, tb, tc, M, N, K, splits, nullptr, amax_a, nullptr, amax_b, none, s);
        default:  return PCNBR_E_BADARG;
    }
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_amax_slots(void) { return H2_AMAX_SLOTS; }

// Diagnostic: buf = device array of 148 x 16 uint64 (or NULL to switch off) that the next gemm2h launches fill with the clock
// cycles every role spent waiting (see H2Wait).  Synchronises the device; not for use under graph capture.
extern "C" int pcnbr_gemm2h_trace(unsigned long long* buf) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpyToSymbol(g_h2_trace, &buf, sizeof(buf));
    if (e != cudaSuccess) return (int)e;
    return gm3_set_trace(buf);                                                 // gemm3x_kernel fills the same slots
}

extern "C" int pcnbr_absmax_f32(const float* x, long rows, long cols, long ld, float* partial, pcnbr_stream_t stream) {
    if (!x || !partial || rows <= 0 || cols <= 0 || (cols % 4) || (ld % 4) || ld < cols || ((uintptr_t)x & 15)) return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    PCNBR_TIMED("absmax_kernel", s, 4.0 * (double)rows * cols, 0.0,
                (absmax_kernel<<<256, 512, 0, s>>>(x, rows, (int)cols, ld, partial)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// 1 when the two-term fp16 kernel should take this GEMM: its tensor-pipe time at the TF32 rate (what the 3xTF32 kernel is
// tied to, three instructions per product) reaches 60 % of its HBM time bound -- from there on the 3xTF32 kernel is
// limited by instruction issue, not by bytes (measured: 65536 x 512 -> 256 runs at 0.30 of either bound on it), and the
// extra absmax scan of the activations (one read) costs less than the halved issue time saves.  GEMMs below 10 GFLOP stay
// on the 3xTF32 kernel whatever their ratio: they are launch / pipeline-fill bound and the three extra launches (two
// absmax scans, one weight split) cost more than they save (measured on the PointNet++ decoder layers: +0.37 ms per step).
extern "C" int pcnbr_gemm2h_preferred(int M, int N, int K) {
    const double bytes = 4.0 * ((double)M * K + (double)N * K + (double)M * N), flops = 2.0 * M * (double)N * K;
    return (flops / 678.35e12 > 0.6 * (bytes / 6551e9) && flops >= 1.0e10) ? 1 : 0;
}

// Weights pre-split once for the forward (transpose = 0: B = W as stored, (rows, cols) = (N, K)) or the input-gradient GEMM
// (transpose = 1: B[n][k] = W[k][n], src is (cols, rows) row-major): out = 2 planes (hi, lo) of `rows` rows with pitch ldo
// halfs (a multiple of 8), `plane` halfs apart.  amax = pcnbr_absmax_f32 of the source.
extern "C" int pcnbr_split_f16(const float* src, int rows, int cols, long ld, int transpose, const float* amax, void* out,
                               long ldo, long plane, pcnbr_stream_t stream) {
    if (!src || !amax || !out || rows <= 0 || cols <= 0 || ldo < cols || (ldo % 8) || plane < (long)rows * ldo || ((uintptr_t)out & 15))
        return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const long total = (long)rows * cols;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    PCNBR_TIMED("split_f16_kernel", s, 8.0 * (double)total, 0.0,
                (split_f16_kernel<<<grid, 256, 0, s>>>(src, rows, cols, ld, transpose, amax, (__half*)out, ldo, plane)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// Same contract as pcnbr_gemm3x_ex_f32 (splits from pcnbr_gemm3x_splits, ws from pcnbr_gemm3x_ws_bytes) plus the
// per-block maxima (pcnbr_amax_slots() floats each, from pcnbr_absmax_f32) of A, of A2 when given, and of B.
// b_split != NULL: B is taken from the planes written by pcnbr_split_f16 (rows = N, cols = K; B / ldb / b_mn are ignored),
// amax_b must be the array that call was given.
// _ex2 adds the operand planes that link the three GEMMs of a layer:
//   a_planes_out (/ a2_planes_out): with a pre-split B and a K-major A (forward: A = x; input gradient: A = gy) the kernel
//       also WRITES the [hi | lo] fp16 split of A (of A2) it computes anyway: 2 planes of M rows, row pitch *_ld halfs (a multiple
//       of 8, >= the operand's K), planes *_plane halfs apart -- bit-identical to pcnbr_split_f16(A, M, K, lda, 0, amax_a, ...);
//   a_mnsplit + b_mnsplit: a_mn = b_mn = 1 (weight gradient dW = gy^T x) with BOTH operands given as such planes (rows = K):
//       no in-kernel conversion; amax_a / amax_b must be the arrays the planes were written with.  A / B (fp32) may then be NULL.
extern "C" int pcnbr_gemm2h_ex2_f32(const float* A, long lda, int a_mn, const float* A2, long lda2, int K1, const float* B, long ldb,
                                    int b_mn, int M, int N, int K, const float* bias, float* C, long ldc, int splits, void* ws,
                                    size_t ws_bytes, const float* amax_a, const float* amax_a2, const float* amax_b,
                                    const void* b_split, long b_split_ld, long b_split_plane,
                                    void* a_planes_out, long apo_ld, long apo_plane, void* a2_planes_out, long ap2o_ld, long ap2o_plane,
                                    const void* a_mnsplit, long ams_ld, long ams_plane, const void* b_mnsplit, long bms_ld, long bms_plane,
                                    pcnbr_stream_t stream) {
    const bool pre2 = a_mnsplit && b_mnsplit;
    if ((!A && !pre2) || (!B && !b_split && !pre2) || !C || !amax_a || !amax_b || M <= 0 || N <= 0 || K <= 0 || splits < 1) return PCNBR_E_BADARG;
    if ((a_mnsplit != nullptr) != (b_mnsplit != nullptr)) return PCNBR_E_BADARG;
    if (ldc < N || (ldc % 4) || ((uintptr_t)C & 15)) return PCNBR_E_BADARG;
    if (splits > 1 && (!ws || ws_bytes < sizeof(float) * (size_t)splits * (size_t)M * (size_t)N)) return PCNBR_E_WORKSPACE;
    if (splits > 1 && bias) return PCNBR_E_BADARG;
    {
        const int kb = (K + GM_BK - 1) / GM_BK, per = (kb + splits - 1) / splits;
        if ((kb + per - 1) / per != splits) return PCNBR_E_BADARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* out = splits > 1 ? (float*)ws : C;
    const float* b = splits > 1 ? nullptr : bias;
    if ((uintptr_t)out & 15) return PCNBR_E_BADARG;
    CUtensorMap tc;
    int rc = gm_make_map_c(&tc, out, M, N, splits > 1 ? N : ldc, splits);
    if (rc) return rc;
    if (pre2 && !(A && B && !h2_pre2_tile_ok(M, N, K))) {
        // ---- both operands as MN-major planes
        if (!a_mn || !b_mn || A2 || bias) return PCNBR_E_BADARG;
        if ((ams_ld % 8) || ams_ld < M || ((uintptr_t)a_mnsplit & 15) || ams_plane < (long)K * ams_ld) return PCNBR_E_BADARG;
        if ((bms_ld % 8) || bms_ld < N || ((uintptr_t)b_mnsplit & 15) || bms_plane < (long)K * bms_ld) return PCNBR_E_BADARG;
        if (!h2_pre2_tile_ok(M, N, K)) return PCNBR_E_BADARG;
        CUtensorMap ta, tb;
        rc = h2_make_map_mnsplit(&ta, a_mnsplit, M, K, ams_ld, ams_plane);
        if (!rc) rc = h2_make_map_mnsplit(&tb, b_mnsplit, N, K, bms_ld, bms_plane);
        if (rc) return rc;
        rc = h2_dispatch_pre2(ta, tb, tc, M, N, K, splits, amax_a, amax_b, s);
        if (rc) return rc;
        if (splits > 1) rc = gm_launch_reduce((const float*)ws, M, N, ldc, splits, C, s);
        return rc;
    }
    if ((lda % 4) || ((uintptr_t)A & 15)) return PCNBR_E_BADARG;
    if (!b_split && ((ldb % 4) || ((uintptr_t)B & 15) || ldb < (b_mn ? N : K))) return PCNBR_E_BADARG;
    if (b_split && ((b_split_ld % 8) || b_split_ld < K || ((uintptr_t)b_split & 15) || b_split_plane < (long)N * b_split_ld)) return PCNBR_E_BADARG;
    const int Ka = A2 ? K1 : K;
    if (lda < (a_mn ? M : Ka)) return PCNBR_E_BADARG;
    if (A2 && (!amax_a2 || a_mn || K1 <= 0 || K1 >= K || (K1 % GM_BK) || (lda2 % 4) || ((uintptr_t)A2 & 15) || lda2 < K - K1)) return PCNBR_E_BADARG;
    const bool planes = a_planes_out != nullptr || a2_planes_out != nullptr;
    if (planes) {
        // written by the units of the first column tile while they convert A: needs the pre-split-B kernel and a K-major A
        if (!b_split || a_mn || (a2_planes_out && !A2)) return PCNBR_E_BADARG;
        if (a_planes_out && ((apo_ld % 8) || apo_ld < Ka || ((uintptr_t)a_planes_out & 15) || apo_plane < (long)M * apo_ld)) return PCNBR_E_BADARG;
        if (a2_planes_out && ((ap2o_ld % 8) || ap2o_ld < K - K1 || ((uintptr_t)a2_planes_out & 15) || ap2o_plane < (long)M * ap2o_ld)) return PCNBR_E_BADARG;
    }
    const int bn = h2_tile_n(M, N, K, !b_split && !b_mn);
    CUtensorMap ta, ta2, tb;
    rc = a_mn ? h2_make_map(&ta, A, M, K, lda, 32) : h2_make_map(&ta, A, Ka, M, lda, 128);
    if (!rc && A2) rc = h2_make_map(&ta2, A2, K - K1, M, lda2, 128);
    if (!A2) ta2 = ta;
    if (!rc) {
        if (b_split) rc = h2_make_map_split(&tb, b_split, K, N, b_split_ld, b_split_plane, bn);
        else         rc = b_mn ? h2_make_map(&tb, B, N, K, ldb, 32) : h2_make_map(&tb, B, K, N, ldb, bn < 128 ? bn : 128);
    }
    if (rc) return rc;
    H2Planes pl;
    pl.ap = ta; pl.ap2 = ta; pl.on = 0;
    if (planes) {
        if (a_planes_out) rc = h2_make_map_split(&pl.ap, a_planes_out, Ka, M, apo_ld, apo_plane, 128);
        if (!rc && a2_planes_out) rc = h2_make_map_split(&pl.ap2, a2_planes_out, K - K1, M, ap2o_ld, ap2o_plane, 128);
        if (rc) return rc;
        pl.on = (a_planes_out ? 1 : 0) | (a2_planes_out ? 2 : 0);           // which of A / A2 the plane writer stores
    }
    const int kb_split = A2 ? K1 / GM_BK : (K + GM_BK - 1) / GM_BK;
    const float* am2 = A2 ? amax_a2 : nullptr;
    if (b_split) {
        rc = a_mn ? h2_dispatch<true, false, true>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s)
                  : h2_dispatch<false, false, true>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s);
    } else if (a_mn) {
        rc = b_mn ? h2_dispatch<true, true, false>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s)
                  : h2_dispatch<true, false, false>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s);
    } else {
        rc = b_mn ? h2_dispatch<false, true, false>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s)
                  : h2_dispatch<false, false, false>(ta, ta2, kb_split, tb, tc, M, N, K, splits, b, amax_a, am2, amax_b, pl, s);
    }
    if (rc) return rc;
    if (splits > 1) rc = gm_launch_reduce((const float*)ws, M, N, ldc, splits, C, s);
    return rc;
}

extern "C" int pcnbr_gemm2h_ex_f32(const float* A, long lda, int a_mn, const float* A2, long lda2, int K1, const float* B, long ldb,
                                   int b_mn, int M, int N, int K, const float* bias, float* C, long ldc, int splits, void* ws,
                                   size_t ws_bytes, const float* amax_a, const float* amax_a2, const float* amax_b,
                                   const void* b_split, long b_split_ld, long b_split_plane, pcnbr_stream_t stream) {
    return pcnbr_gemm2h_ex2_f32(A, lda, a_mn, A2, lda2, K1, B, ldb, b_mn, M, N, K, bias, C, ldc, splits, ws, ws_bytes, amax_a, amax_a2,
                                amax_b, b_split, b_split_ld, b_split_plane, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0, 0,
                                stream);
}
