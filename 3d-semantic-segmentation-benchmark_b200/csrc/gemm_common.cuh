// gemm_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05) and tensor-map builders shared by the two tensor-core GEMMs
// (gemm_tc.cu: 3xTF32; gemm_h2.cu: two-term fp16 split).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace pcnbr {

constexpr int GM_BM = 128;                 // rows per tile (TMEM lanes)
constexpr int GM_BK = 32;                  // floats per K block = one 128-byte swizzle atom
constexpr int GM_CONV_WARPS = 8;
constexpr int GM_THREADS = 192 + 32 * GM_CONV_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue, warps 6.. converters
constexpr uint32_t GM_SLAB = 128 * 128;    // bytes of a 128 x 32-float operand tile
constexpr uint32_t GM_CHUNK = 32 * 128;    // bytes of a 32 x 32-float MN-major chunk

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
__device__ __forceinline__ void gm_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(gm_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void gm_tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void gm_epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps
// kind::tf32, D = fp32, M = 128, N = BN; A / B major-ness in bits 15 / 16 (cute::UMMA::InstrDescriptor)
template <int BN, bool A_MN, bool B_MN>
__device__ __forceinline__ void gm_umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, bool accumulate) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GM_BM >> 4) << 24);
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


// ------------------------------------------------------------------------------------ host side: tensor maps

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

// fp32 matrix of `outer` rows x `inner` contiguous floats with row pitch ld (floats), box = 32 floats x box_rows,
// swizzled for a K-major (inner = K) or MN-major (inner = M or N) UMMA operand; out-of-range elements read as 0.
static int gm_make_map(CUtensorMap* map, const float* base, long inner, long outer, long ld, int box_rows, bool mn_major) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// C as a (N, M, splits) fp32 tensor (row pitch ldc, split pitch M * ldc), box = 32 floats x 128 rows, 128-byte swizzle
static int gm_make_map_c(CUtensorMap* map, float* base, long M, long N, long ldc, int splits) {
    GmEncodeFn enc = gm_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
    cuuint64_t gstr[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)M * (cuuint64_t)ldc * 4};
    cuuint32_t box[3] = {32, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}


// Tile width: the widest BN <= N (rounded up) that still gives every SM a work unit (measured against two units per SM:
// PointNet++ 5.63 -> 5.59 ms per step, the others unchanged; always-widest: 5.67) -- narrower tiles mean
// more units and a deeper smem ring (more bytes in flight per SM), which is what the many small, latency-bound layer
// GEMMs need; the big tensor-bound ones (>= 296 units at BN = 256) keep the widest tile and its operand reuse.
// (Long-K GEMMs -- the weight gradients -- get their units from split-K instead and keep the widest tile.)
static int gm_tile_n(int M, int N, int K) {
    const int widest = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
    if (K >= 64 * GM_BK) return widest;
    const long mt = (M + GM_BM - 1) / GM_BM;
    int bn = widest;
    while (bn > 64 && mt * ((N + bn - 1) / bn) < 148) bn >>= 1;
    return bn;
}

int gm_launch_reduce(const float* ws, int M, int N, long ldc, int splits, float* C, cudaStream_t s);   // gemm_tc.cu
int gm3_set_trace(unsigned long long* buf);                                                             // gemm_tc.cu

// wait-time trace of the GEMM kernels' warp roles (pcnbr_gemm2h_trace): cycles spent inside [begin, end) pairs, per CTA and slot
struct H2Wait {
    unsigned long long* dst;
    long long acc;
    __device__ __forceinline__ H2Wait(unsigned long long* base, int slot) : dst(base ? base + (size_t)blockIdx.x * 16 + slot : nullptr), acc(0) {}
    __device__ __forceinline__ long long begin() const { return dst ? clock64() : 0; }
    __device__ __forceinline__ void end(long long t0) { if (dst) acc += clock64() - t0; }
    __device__ __forceinline__ void flush() { if (dst) *dst = (unsigned long long)acc; }
};


}  // namespace pcnbr
