// gemm_tc.cu -- fp32-accurate GEMM on the 5th-generation tensor cores (3xTF32), for the 1x1 convolutions that
// sit between the neighbourhood kernels.
//
// Reference: every Conv1d/Conv2d with kernel size 1 in models/utils/common.py:125-178 and
// models/dgcnn/dgcnn.py:66-71,95-126 is a GEMM over the (points, channels) matrix.  Under the parity bar (fp32,
// 1e-4 relative; TF32 is off) the library runs them as SIMT SGEMMs at ~55 TFLOP/s, which is 60 % of a DGCNN train
// step.  Here   C[M,N] = A[M,K] . B[N,K]^T (+ bias[N])   runs as three TF32 tensor-core products of pre-split
// operands,  hi.hi' + lo.hi' + hi.lo'  with hi = tf32(x), lo = tf32(x - hi)  (error ~2^-21 |a||b|: fp32-grade), all
// accumulated in one fp32 TMEM accumulator:
//
//   split kernel  : x -> hi, lo (and, for the weight-gradient GEMM, their transposes, so every GEMM is K-major);
//   producer warp : TMA (128-byte swizzle) of 128-row x 32-float slabs of A_hi, A_lo, B_hi, B_lo into a ring;
//   MMA warp      : one elected lane issues tcgen05.mma kind::tf32, M=128 x N=BN x K=8, 12 per 32-wide K block,
//                   into one of two TMEM accumulators (2 x 256 columns);
//   4 epilogue warps: tcgen05.ld the finished tile while the next one is being multiplied, add the bias, store
//                   full 128-byte lines.
//   Weight gradients (M, N small, K = number of points) are split along K over the CTAs; the partial tiles are
//   summed in a fixed order by a second kernel (deterministic, no atomics).
#include "common.cuh"
#include <cuda.h>

namespace pcnbr {

constexpr int GM_BM = 128;                 // rows per tile (TMEM lanes)
constexpr int GM_BK = 32;                  // floats per K block = one 128-byte swizzle atom
constexpr int GM_THREADS = 192;            // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue
constexpr uint32_t GM_SLAB = 128 * 128;    // bytes: 128 rows x 128 B

__device__ __forceinline__ uint32_t gm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gm_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gm_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gm_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(gm_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void gm_tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(gm_smem_u32(dst)), "l"(map), "r"(gm_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// kind::tf32, D = fp32, A/B K-major, M = 128, N = BN (cute::UMMA::InstrDescriptor)
template <int BN>
__device__ __forceinline__ void gm_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, bool accumulate) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
    const uint32_t acc = accumulate ? 1u : 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void gm_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gm_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

// ------------------------------------------------------------------------------------ operand split

// x (R,C) row-major -> hi = tf32_rna(x), lo = tf32_rna(x - hi), both (R,C); optionally the transposes (C,R).
// 32x32 tiles through shared memory: all four outputs are written with coalesced 128-byte rows.
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, long R, long C, float* __restrict__ hi, float* __restrict__ lo,
                  float* __restrict__ hiT, float* __restrict__ loT) {
    __shared__ float th[32][33], tl[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
    const long tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
    for (long t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
        const long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long r = r0 + ty + 8 * i, c = c0 + tx;
            float h = 0.f, l = 0.f;
            if (r < R && c < C) {
                const float v = x[r * C + c];
                uint32_t hb, lb;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
                h = __uint_as_float(hb);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(__fsub_rn(v, h)));
                l = __uint_as_float(lb);
                if (hi) { hi[r * C + c] = h; lo[r * C + c] = l; }
            }
            th[ty + 8 * i][tx] = h;
            tl[ty + 8 * i][tx] = l;
        }
        if (hiT) {
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long c = c0 + ty + 8 * i, r = r0 + tx;
                if (r < R && c < C) { hiT[c * R + r] = th[tx][ty + 8 * i]; loT[c * R + r] = tl[tx][ty + 8 * i]; }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------ main kernel

// K-major, 128-byte swizzle shared-memory matrix descriptor (see knn_tc.cu)
__device__ __forceinline__ uint64_t gm_desc(uint32_t saddr) {
    const uint64_t lo = (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16);
    const uint64_t hi = (uint64_t)(1024 >> 4) | ((uint64_t)1 << 14) | ((uint64_t)2 << 29);
    return lo | (hi << 32);
}

// Work unit = (split, m tile, n tile), n fastest: the CTAs that run side by side share the same A slab, so the
// streamed operand (the activations) is read from HBM once and hits L2 for the other n tiles.  C tile goes to out + split * M * ldc (partials) -- splits == 1
// writes the result (plus bias) directly.
template <int BN, int STAGES>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tm_ah, const __grid_constant__ CUtensorMap tm_al,
              const __grid_constant__ CUtensorMap tm_bh, const __grid_constant__ CUtensorMap tm_bl,
              int M, int N, int K, int splits, const float* __restrict__ bias, float* __restrict__ out, long ldc) {
    extern __shared__ uint8_t gm_smem_raw[];
    uint8_t* smem = gm_smem_raw + ((1024u - (gm_smem_u32(gm_smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t A_BYTES = GM_SLAB, B_BYTES = (BN / 128) * GM_SLAB;
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;       // A_hi | A_lo | B_hi | B_lo
    uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                                            // [STAGES]
    uint64_t* empty = bars + STAGES;                                  // [STAGES]
    uint64_t* tmem_full = bars + 2 * STAGES;                          // [2]
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;                     // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MT = (M + GM_BM - 1) / GM_BM, NT = (N + BN - 1) / BN;
    const int KB = (K + GM_BK - 1) / GM_BK;                           // K blocks in total
    const int kb_per = (KB + splits - 1) / splits;
    const int units = MT * NT * splits;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_ah) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_al) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_bl) : "memory");
        for (int i = 0; i < STAGES; ++i) { gm_mbar_init(&full[i], 1); gm_mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { gm_mbar_init(&tmem_full[i], 1); gm_mbar_init(&tmem_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
                const int nt = unit % NT, mt = (unit / NT) % MT, sp = unit / (MT * NT);
                const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    gm_mbar_wait(&empty[stage], phase ^ 1);
                    gm_mbar_expect_tx(&full[stage], STAGE_BYTES);
                    uint8_t* st = smem + stage * STAGE_BYTES;
                    gm_tma_load_2d(st, &tm_ah, &full[stage], kb * GM_BK, mt * GM_BM);
                    gm_tma_load_2d(st + A_BYTES, &tm_al, &full[stage], kb * GM_BK, mt * GM_BM);
#pragma unroll
                    for (int h = 0; h < BN / 128; ++h) {
                        gm_tma_load_2d(st + 2 * A_BYTES + h * GM_SLAB, &tm_bh, &full[stage], kb * GM_BK, nt * BN + h * 128);
                        gm_tma_load_2d(st + 2 * A_BYTES + B_BYTES + h * GM_SLAB, &tm_bl, &full[stage], kb * GM_BK, nt * BN + h * 128);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================== MMA issuer (warp-uniform loops, one elected lane issues)
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        uint32_t stage = 0, phase = 0, tile = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++tile) {
            const int sp = unit / (MT * NT);
            const int kb0 = sp * kb_per, kb1 = min(KB, kb0 + kb_per);
            const uint32_t buf = tile & 1, tphase = (tile >> 1) & 1;
            gm_mbar_wait(&tmem_empty[buf], tphase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem_base + buf * 256;
            for (int kb = kb0; kb < kb1; ++kb) {
                gm_mbar_wait(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = gm_smem_u32(smem + stage * STAGE_BYTES);
                const uint64_t ah = gm_desc(st), al = gm_desc(st + A_BYTES);
                const uint64_t bh = gm_desc(st + 2 * A_BYTES), bl = gm_desc(st + 2 * A_BYTES + B_BYTES);
                if (leader) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) gm_umma_tf32<BN>(d, ah + 2 * s, bh + 2 * s, kb > kb0 || s > 0);   // hi . hi'
#pragma unroll
                    for (int s = 0; s < 4; ++s) gm_umma_tf32<BN>(d, al + 2 * s, bh + 2 * s, true);                // lo . hi'
#pragma unroll
                    for (int s = 0; s < 4; ++s) gm_umma_tf32<BN>(d, ah + 2 * s, bl + 2 * s, true);                // hi . lo'
                    gm_umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (leader) gm_umma_commit(&tmem_full[buf]);
            __syncwarp();
        }
    } else {
        // ===================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int quarter = warp & 3;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        uint32_t tile = 0;
        for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++tile) {
            const int nt = unit % NT, mt = (unit / NT) % MT, sp = unit / (MT * NT);
            const uint32_t buf = tile & 1, tphase = (tile >> 1) & 1;
            const long row = (long)mt * GM_BM + quarter * 32 + lane;
            float* __restrict__ dst = out + ((long)sp * M + row) * ldc + (long)nt * BN;
            const int ncols = min(BN, N - nt * BN);
            const bool vec = (ldc % 4 == 0) && ((((uintptr_t)out) & 15) == 0);
            gm_mbar_wait(&tmem_full[buf], tphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int q = 0; q < BN / 32; ++q) {
                if (q * 32 >= ncols) break;                            // warp-uniform
                uint32_t r[32];
                gm_tmem_ld32(tmem_base + tlane + buf * 256 + q * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < M) {
                    if (vec && q * 32 + 32 <= ncols) {
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            float4 v = make_float4(__uint_as_float(r[4 * i4]), __uint_as_float(r[4 * i4 + 1]),
                                                   __uint_as_float(r[4 * i4 + 2]), __uint_as_float(r[4 * i4 + 3]));
                            if (bias) {
                                const float4 bv = *reinterpret_cast<const float4*>(bias + nt * BN + q * 32 + 4 * i4);
                                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            }
                            *reinterpret_cast<float4*>(dst + q * 32 + 4 * i4) = v;
                        }
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (q * 32 + i < ncols)
                                dst[q * 32 + i] = __uint_as_float(r[i]) + (bias ? bias[nt * BN + q * 32 + i] : 0.f);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) gm_mbar_arrive(&tmem_empty[buf]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// out[i] = sum_s part[s][i] in ascending s (fixed order)
__global__ void __launch_bounds__(256)
gemm_reduce_kernel(const float* __restrict__ part, long n, int splits, float* __restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float a = part[i];
        for (int s = 1; s < splits; ++s) a = __fadd_rn(a, part[(long)s * n + i]);
        out[i] = a;
    }
}

// ------------------------------------------------------------------------------------ host side

typedef CUresult (*GmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static GmEncodeFn gm_encode_fn() {
    static GmEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (GmEncodeFn)p;
    }
    return fn;
}

// (K, rows) fp32 matrix with leading dimension K, box = 32 floats x 128 rows, 128-byte swizzle; out of range reads as 0
static int gm_make_map(CUtensorMap* map, const float* base, long rows, long K) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int BN, int STAGES>
static int gm_launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, int M, int N,
                     int K, int splits, const float* bias, float* out, long ldc, cudaStream_t s) {
    const size_t smem = (size_t)STAGES * (2 * GM_SLAB + 2 * (BN / 128) * GM_SLAB) + 32 * 8 + 1024;
    cudaError_t e = cudaFuncSetAttribute(gemm3x_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = ((M + GM_BM - 1) / GM_BM) * ((N + BN - 1) / BN) * splits;
    const int grid = units < sms ? units : sms;
    // algorithmic work: 2 M N K flop counted once (the kernel issues 3x); operands + result bytes
    PCNBR_TIMED("gemm3x_kernel", s, 4.0 * ((double)M * K * 2 + (double)N * K * 2 + (double)M * N * splits), 2.0 * M * (double)N * K,
                (gemm3x_kernel<BN, STAGES><<<grid, GM_THREADS, smem, s>>>(ah, al, bh, bl, M, N, K, splits, bias, out, ldc)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_split_tf32(const float* x, long R, long C, float* hi, float* lo, float* hiT, float* loT,
                                pcnbr_stream_t stream) {
    if (!x || R <= 0 || C <= 0 || (!hi && !hiT) || (hi && !lo) || (hiT && !loT)) return PCNBR_E_BADARG;
    const long tiles = ((R + 31) / 32) * ((C + 31) / 32);
    const int grid = (int)(tiles < 148L * 16 ? tiles : 148L * 16);
    cudaStream_t s = (cudaStream_t)stream;
    PCNBR_TIMED("split_tf32_kernel", s, 4.0 * R * C * (1.0 + (hi ? 2.0 : 0.0) + (hiT ? 2.0 : 0.0)), 3.0 * R * C,
                (split_tf32_kernel<<<grid, 256, 0, s>>>(x, R, C, hi, lo, hiT, loT)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_gemm3x_splits(int M, int N, int K) {
    // Split K when the output has too few tiles to fill the chip (weight gradients: K = number of points).
    // Cost model: waves of work units over the SMs x K blocks per unit; the smallest split count that minimises it.
    const int bn = (N > 128) ? 256 : 128;
    const long tiles = (long)((M + GM_BM - 1) / GM_BM) * ((N + bn - 1) / bn);
    const int kb = (K + GM_BK - 1) / GM_BK;
    if (tiles >= 148 || kb < 64) return 1;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long best_cost = -1;
    int best = 1;
    for (int s = 1; s <= 32 && s <= kb / 16; ++s) {
        const long per = (kb + s - 1) / s;
        const long eff = (kb + per - 1) / per;                            // every split owns at least one K block
        const long waves = (tiles * eff + sms - 1) / sms;
        const long cost = waves * (per + 4);                              // + tile prologue / epilogue
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = (int)eff; }
    }
    return best;
}

extern "C" size_t pcnbr_gemm3x_ws_bytes(int M, int N, int K, int splits) {
    (void)K;
    return splits > 1 ? sizeof(float) * (size_t)splits * (size_t)M * (size_t)N : 0;
}

extern "C" int pcnbr_gemm3x_f32(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, int M, int N,
                                int K, const float* bias, float* C, int splits, void* ws, size_t ws_bytes,
                                pcnbr_stream_t stream) {
    if (!a_hi || !a_lo || !b_hi || !b_lo || !C || M <= 0 || N <= 0 || K <= 0 || splits < 1) return PCNBR_E_BADARG;
    if (K % 4 != 0) return PCNBR_E_TOOLARGE;                              // TMA needs 16-byte row pitch
    if (((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) & 15) return PCNBR_E_BADARG;
    if (splits > 1 && (!ws || ws_bytes < pcnbr_gemm3x_ws_bytes(M, N, K, splits))) return PCNBR_E_WORKSPACE;
    if (splits > 1 && bias) return PCNBR_E_BADARG;                        // bias with split-K is not supported
    {
        const int kb = (K + GM_BK - 1) / GM_BK, per = (kb + splits - 1) / splits;
        if ((kb + per - 1) / per != splits) return PCNBR_E_BADARG;        // use pcnbr_gemm3x_splits(): no empty split
    }
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap ah, al, bh, bl;
    int rc = gm_make_map(&ah, a_hi, M, K);
    if (!rc) rc = gm_make_map(&al, a_lo, M, K);
    if (!rc) rc = gm_make_map(&bh, b_hi, N, K);
    if (!rc) rc = gm_make_map(&bl, b_lo, N, K);
    if (rc) return rc;
    float* out = splits > 1 ? (float*)ws : C;
    const float* b = splits > 1 ? nullptr : bias;
    if (N > 128) rc = gm_launch<256, 2>(ah, al, bh, bl, M, N, K, splits, b, out, N, s);
    else         rc = gm_launch<128, 3>(ah, al, bh, bl, M, N, K, splits, b, out, N, s);
    if (rc) return rc;
    if (splits > 1) {
        const long n = (long)M * N;
        const int grid = (int)((n + 255) / 256 < 148L * 8 ? (n + 255) / 256 : 148L * 8);
        PCNBR_TIMED("gemm_reduce_kernel", s, 4.0 * n * (splits + 1), (double)n * splits,
                    (gemm_reduce_kernel<<<grid, 256, 0, s>>>((const float*)ws, n, splits, C)));
        PCNBR_CHECK_LAUNCH();
    }
    return 0;
}
