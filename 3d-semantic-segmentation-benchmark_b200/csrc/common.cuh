// common.cuh -- shared device helpers for libpcnbr (sm_100a).
//
// Exact-arithmetic rule: every distance that decides an index is computed with explicit
// round-to-nearest intrinsics (__fsub_rn/__fmul_rn/__fadd_rn/__fmaf_rn/__fsqrt_rn) so nvcc can never
// contract or reassociate; the forms are the ones pinned against the reference in oracle/canon.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pcnbr.h"

#define PCNBR_FULL 0xffffffffu
#define PCNBR_KEY_MAX 0xffffffffffffffffull

#define PCNBR_CHECK_LAUNCH()                         \
    do {                                             \
        cudaError_t e__ = cudaGetLastError();        \
        if (e__ != cudaSuccess) return (int)e__;     \
    } while (0)

namespace pcnbr {

typedef unsigned long long u64;

// Monotone map float -> uint32 (total order of the reals incl. negatives; -0 < +0).
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
// (key, index) -> one u64 whose unsigned order is ascending (key, index).
__device__ __forceinline__ u64 pack_key(uint32_t ord, uint32_t idx) { return ((u64)ord << 32) | idx; }

__device__ __forceinline__ u64 shfl64(u64 v, int src) {
    uint32_t lo = __shfl_sync(PCNBR_FULL, (uint32_t)v, src);
    uint32_t hi = __shfl_sync(PCNBR_FULL, (uint32_t)(v >> 32), src);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl64_up1(u64 v) {
    uint32_t lo = __shfl_up_sync(PCNBR_FULL, (uint32_t)v, 1);
    uint32_t hi = __shfl_up_sync(PCNBR_FULL, (uint32_t)(v >> 32), 1);
    return ((u64)hi << 32) | lo;
}

// reference group()/interpolate() distance: ((dx*dx + dy*dy) + dz*dz), dx = src - query.
__device__ __forceinline__ float d2_direct(float px, float py, float pz, float qx, float qy, float qz) {
    const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy), dz = __fsub_rn(pz, qz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Length-aware forms (zero-padded evaluation batches, data_processing/block_datasets.py:19-25): cloud b holds len[b] real
// rows (clamped to [1, n_alloc]; len == NULL: all n_alloc), the rest is padding that takes no part -- the result for the
// real rows is what the reference computes when the cloud is passed alone, unpadded.
__device__ __forceinline__ int len_valid(const int32_t* __restrict__ len, int b, int n_alloc) {
    return len ? min(max(len[b], 1), n_alloc) : n_alloc;
}
// output row of a padding query: in-range indices 0..K-1 and zero distances, so downstream gathers stay in bounds
__device__ __forceinline__ void len_fill_row(int32_t* __restrict__ idx, float* __restrict__ d2, size_t row, int K, int lane) {
    for (int pos = lane; pos < K; pos += 32) {
        idx[row + pos] = pos;
        if (d2) d2[row + pos] = 0.f;
    }
}

// Sorted K-list spread over a warp: position p = slot*32 + lane holds the p-th smallest key.
// NSLOT*32 >= K.  All lanes call insert() with the same candidate.
template <int NSLOT>
struct WarpList {
    u64 v[NSLOT];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) v[s] = PCNBR_KEY_MAX;
    }
    // key at list position p (warp-uniform p)
    __device__ __forceinline__ u64 at(int p) const {
        u64 r = 0;
#pragma unroll
        for (int s = 0; s < NSLOT; ++s)
            if ((p >> 5) == s) r = shfl64(v[s], p & 31);
        return r;
    }
    __device__ __forceinline__ void insert(u64 cand, int lane) {
        u64 carry = 0;  // key at position p-1 of lane 0 of the current slot (0 = -inf)
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            u64 up = shfl64_up1(v[s]);
            if (lane == 0) up = carry;
            const u64 last = shfl64(v[s], 31);
            if (v[s] > cand) v[s] = (up > cand) ? up : cand;
            carry = last;
        }
    }
};

// ascending bitonic sort of one 64-bit key per lane across the warp (15 compare-exchange steps)
__device__ __forceinline__ u64 warp_sort64(u64 key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t lo = __shfl_xor_sync(PCNBR_FULL, (uint32_t)key, j);
            const uint32_t hi = __shfl_xor_sync(PCNBR_FULL, (uint32_t)(key >> 32), j);
            const u64 other = ((u64)hi << 32) | lo;
            const bool take_min = ((lane & j) == 0) == ((lane & k) == 0);
            key = (take_min == (other < key)) ? other : key;
        }
    }
    return key;
}

// ---- optional per-kernel timing (prof.cu): events on the launch stream + the launch's algorithmic work
bool prof_enabled();
struct ProfScope {
    ProfScope(const char* name, cudaStream_t s, double bytes, double flops);
    ~ProfScope();
    int slot_;
    cudaStream_t stream_;
};
// PCNBR_TIMED("kernel", stream, algorithmic_bytes, algorithmic_flops, kernel<<<...>>>(...));
#define PCNBR_TIMED(name, s, bytes, flops, ...)                       \
    do {                                                              \
        pcnbr::ProfScope prof_scope__((name), (s), (bytes), (flops)); \
        __VA_ARGS__;                                                  \
    } while (0)

// cross-file host helpers
int launch_sumsq(const float* x, int B, int F, int N, long sf, long sn, const int32_t* n_valid, float* xx, cudaStream_t s);     // select.cu
int select_ball_len(const float* q, const float* p, int B, int M, int N, float r2, int K, const int32_t* n_qry, const int32_t* n_src,
                    int32_t* idx, cudaStream_t s);
int select_knn_len(const float* q, const float* p, int B, int M, int N, int K, const int32_t* n_qry, const int32_t* n_src,
                   int32_t* idx, float* d2, cudaStream_t s);
bool knn_tc_supported(int F, int N, int K);                                                              // knn_tc.cu
size_t knn_tc_ws_bytes(int B, int F, int N);
int knn_tc_run(const float* x, int B, int F, int N, long sf, long sn, int K, const int32_t* n_valid, int32_t* idx, void* ws,
               float* dump, int32_t* stats_out, cudaStream_t s);

}  // namespace pcnbr
