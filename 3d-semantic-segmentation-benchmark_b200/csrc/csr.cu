// csr.cu -- K7 support: inverse of a neighbour index (CSR by source point).
//
// The reference's backward of every gather (autograd IndexBackward -> index_put_(accumulate=True),
// models/utils/common.py:65,117 and models/dgcnn/dgcnn.py:48) is a scatter-add.  We turn it into a
// gather: for each source point s, the list of positions e (flattened (m,k)) with idx[e] == s, in
// ascending e.  Each backward kernel then sums a segment in that fixed order: no float atomics,
// bitwise deterministic.  Built once per forward and reused by the backward.
//
// count (int atomics: order-independent) -> per-cloud exclusive scan -> fill (int atomic cursor,
// arbitrary order) -> per-segment rank-by-counting sort (out of place) which restores ascending e.
#include "common.cuh"

namespace pcnbr {

__global__ void csr_count_kernel(const int32_t* __restrict__ idx, int E, int N, int32_t* __restrict__ cnt) {
    const int b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int s = idx[(size_t)b * E + e];
        atomicAdd(&cnt[(size_t)b * (N + 1) + s], 1);
    }
}

// one CTA per cloud: offsets[b, 0..N] = exclusive scan of cnt[b, 0..N-1]; cursor = copy of offsets
__global__ void __launch_bounds__(1024)
csr_scan_kernel(const int32_t* cnt, int N, int32_t* __restrict__ offsets, int32_t* cursor) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* c = cnt + (size_t)b * (N + 1);
    int32_t* o = offsets + (size_t)b * (N + 1);
    int32_t* cur = cursor + (size_t)b * (N + 1);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < N + 1; base += 1024) {
        const int i = base + tid;
        const int v = (i < N) ? c[i] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(PCNBR_FULL, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(PCNBR_FULL, w, d);
                if (lane >= d) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp ? s_warp[warp - 1] : 0) + x - v;
        if (i < N + 1) { o[i] = excl; cur[i] = excl; }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
}

__global__ void csr_fill_kernel(const int32_t* __restrict__ idx, int E, int N, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ tmp) {
    const int b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int s = idx[(size_t)b * E + e];
        const int pos = atomicAdd(&cursor[(size_t)b * (N + 1) + s], 1);
        tmp[(size_t)b * E + pos] = e;
    }
}

// rank(x) = #{y in segment : y < x} (positions are unique), written out of place.
// One warp ranks the 32 entries [c0, c0+32) of a segment against the whole segment.
__device__ __forceinline__ void csr_rank_chunk(const int32_t* __restrict__ t, int32_t* __restrict__ pm, int beg,
                                               int len, int c0, int lane) {
    const bool mine = c0 + lane < len;
    const int x = mine ? t[beg + c0 + lane] : 0x7fffffff;
    int rank = 0;
    for (int y0 = 0; y0 < len; y0 += 32) {
        const int y = (y0 + lane < len) ? t[beg + y0 + lane] : 0x7fffffff;
        const int n = min(32, len - y0);
        for (int l = 0; l < n; ++l) rank += (__shfl_sync(PCNBR_FULL, y, l) < x);
    }
    if (mine) pm[beg + rank] = x;
}

// A CTA of 8 warps takes 8 consecutive segments: short ones are rank-sorted by one warp each.  Long ones -- the
// padded / duplicated heavy hitters of an under-filled ball query (up to M*K entries on a few low-index points) and the
// hubs of a feature-space kNN graph -- would cost O(len^2) that way; positions are distinct integers below E, so the
// whole CTA instead marks them in a shared-memory bitmap of E bits and reads the set bits back in order
// (popcount prefix over the words): O(E/32 + len) per long segment.
constexpr int CSR_HEAVY = 128;
constexpr int CSR_BITMAP_MAX_E = 1 << 20;          // 128 KB of bitmap at most (dynamic shared memory)

__device__ __forceinline__ void csr_bitmap_sort(const int32_t* __restrict__ t, int32_t* __restrict__ pm, int beg, int len,
                                                int E, uint32_t* bm, int* s_scan) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = (E + 31) >> 5;
    for (int i = tid; i < W; i += 256) bm[i] = 0u;
    __syncthreads();
    for (int i = tid; i < len; i += 256) {
        const int e = t[beg + i];
        atomicOr(&bm[e >> 5], 1u << (e & 31));
    }
    __syncthreads();
    const int wpt = (W + 255) / 256;                // words per thread, contiguous slice
    const int w0 = min(tid * wpt, W), w1 = min(w0 + wpt, W);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(bm[w]);
    int x = cnt;                                     // block exclusive scan of cnt
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(PCNBR_FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_scan[warp] = x;
    __syncthreads();
    int base = x - cnt;
    for (int k = 0; k < warp; ++k) base += s_scan[k];
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = bm[w];
        while (bits) {
            const int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            pm[beg + base++] = (w << 5) + bpos;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
csr_sort_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ tmp, int E, int N,
                int32_t* __restrict__ perm, int use_bitmap) {
    extern __shared__ uint32_t csr_bm[];
    __shared__ int s_beg[8], s_len[8], s_scan[8];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* o = offsets + (size_t)b * (N + 1);
    const int32_t* t = tmp + (size_t)b * E;
    int32_t* pm = perm + (size_t)b * E;
    for (int s0 = blockIdx.x * 8; s0 < N; s0 += gridDim.x * 8) {
        const int s = s0 + warp;
        const int beg = (s < N) ? o[s] : 0;
        const int len = (s < N) ? o[s + 1] - beg : 0;
        if (lane == 0) { s_beg[warp] = beg; s_len[warp] = len; }
        if (len == 1) { if (lane == 0) pm[beg] = t[beg]; }
        else if (len <= CSR_HEAVY)
            for (int c0 = 0; c0 < len; c0 += 32) csr_rank_chunk(t, pm, beg, len, c0, lane);
        __syncthreads();
        for (int w = 0; w < 8; ++w) {
            const int hl = s_len[w];
            if (hl <= CSR_HEAVY) continue;                             // uniform over the CTA
            if (use_bitmap) csr_bitmap_sort(t, pm, s_beg[w], hl, E, csr_bm, s_scan);
            else
                for (int c0 = warp * 32; c0 < hl; c0 += 8 * 32) csr_rank_chunk(t, pm, s_beg[w], hl, c0, lane);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ stable counting sort (no sort pass)
// For N <= CSR_STABLE_MAX_N the inverse is built as a STABLE counting sort, so the positions of a segment come out in
// ascending order by construction and the O(len^2) rank sort above (the whole cost on hub-heavy kNN graphs and padded
// ball tables) disappears.  Each warp owns a contiguous range of CSR_CH positions and a private shared-memory counter
// row of N words:
//   csr_hist_kernel  : per (cloud, warp range) key histogram -> cm (B, P, N)           (__match_any_sync groups equal keys
//                      inside a 32-position step; the lowest lane of a group adds the group size: no atomics)
//   csr_colscan_kernel: exclusive prefix over the P ranges for every key, in place; the per-key totals go to cnt
//   csr_scan_kernel  : exclusive scan of the totals over the keys -> offsets              (as before)
//   csr_place_kernel : each warp reloads its prefix row (+ offsets) as cursors and walks its range in order:
//                      position = cursor[key] + (rank of the lane among equal keys of the step).
// Deterministic: no atomics, no scheduling dependence.
constexpr int CSR_CH = 2048;               // positions per warp range (64 steps of 32)
constexpr int CSR_SW = 4;                  // warps per CTA (each with N words of shared memory)
constexpr int CSR_STABLE_MAX_N = 12288;    // 4 warps x 48 KB = 192 KB of dynamic shared memory at most

__global__ void __launch_bounds__(32 * CSR_SW)
csr_hist_kernel(const int32_t* __restrict__ idx, int E, int N, int P, uint32_t* __restrict__ cm) {
    extern __shared__ uint32_t csr_sm[];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * CSR_SW + warp;
    if (p >= P) return;
    uint32_t* h = csr_sm + (size_t)warp * N;
    for (int i = lane; i < N; i += 32) h[i] = 0u;
    __syncwarp();
    const int e0 = p * CSR_CH, e1 = min(E, e0 + CSR_CH);
    const int32_t* __restrict__ ib = idx + (size_t)b * E;
    // the whole range is fetched up front (64 independent coalesced loads per lane): the walk below is a serial chain and
    // used to pay one exposed global-load latency per 32-position step
    int keys[CSR_CH / 32];
#pragma unroll
    for (int i = 0; i < CSR_CH / 32; ++i) {
        const int e = e0 + i * 32 + lane;
        keys[i] = (e < e1) ? ib[e] : -1 - lane;                           // invalid lanes get unique dummy keys
    }
#pragma unroll
    for (int i = 0; i < CSR_CH / 32; ++i) {
        if (e0 + i * 32 >= e1) break;                                     // warp-uniform
        const int key = keys[i];
        const bool v = key >= 0;
        const uint32_t peers = __match_any_sync(PCNBR_FULL, key);
        if (v && (peers & ((1u << lane) - 1u)) == 0u) h[key] += __popc(peers);
        __syncwarp();
    }
    uint32_t* __restrict__ row = cm + ((size_t)b * P + p) * N;
    for (int i = lane; i < N; i += 32) row[i] = h[i];
}

__global__ void __launch_bounds__(256)
csr_colscan_kernel(uint32_t* __restrict__ cm, int N, int P, int32_t* __restrict__ cnt) {
    const int b = blockIdx.y;
    const int key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= N) return;
    uint32_t* __restrict__ col = cm + (size_t)b * P * N + key;
    uint32_t run = 0;
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
        const uint32_t c = col[(size_t)p * N];
        col[(size_t)p * N] = run;
        run += c;
    }
    cnt[(size_t)b * (N + 1) + key] = (int32_t)run;
}

__global__ void __launch_bounds__(32 * CSR_SW)
csr_place_kernel(const int32_t* __restrict__ idx, int E, int N, int P, const uint32_t* __restrict__ cm,
                 const int32_t* __restrict__ offsets, int32_t* __restrict__ perm) {
    extern __shared__ uint32_t csr_sm[];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * CSR_SW + warp;
    if (p >= P) return;
    uint32_t* cur = csr_sm + (size_t)warp * N;
    const uint32_t* __restrict__ row = cm + ((size_t)b * P + p) * N;
    const int32_t* __restrict__ off = offsets + (size_t)b * (N + 1);
    for (int i = lane; i < N; i += 32) cur[i] = row[i] + (uint32_t)off[i];
    __syncwarp();
    const int e0 = p * CSR_CH, e1 = min(E, e0 + CSR_CH);
    const int32_t* __restrict__ ib = idx + (size_t)b * E;
    int32_t* __restrict__ pm = perm + (size_t)b * E;
    int keys[CSR_CH / 32];                                               // as in csr_hist_kernel: no load inside the serial walk
#pragma unroll
    for (int i = 0; i < CSR_CH / 32; ++i) {
        const int e = e0 + i * 32 + lane;
        keys[i] = (e < e1) ? ib[e] : -1 - lane;
    }
#pragma unroll
    for (int i = 0; i < CSR_CH / 32; ++i) {
        if (e0 + i * 32 >= e1) break;                                     // warp-uniform
        const int key = keys[i];
        const bool v = key >= 0;
        const uint32_t peers = __match_any_sync(PCNBR_FULL, key);
        const uint32_t below = peers & ((1u << lane) - 1u);
        if (v) {
            pm[cur[key] + __popc(below)] = e0 + i * 32 + lane;            // every peer reads the cursor before it moves
        }
        __syncwarp();
        if (v && below == 0u) cur[key] += __popc(peers);
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------ bitmap transposition (row-structured tables)
// A neighbour table is a sparse (M queries x N sources) matrix with K entries per row; its CSR-by-source inverse with
// ascending positions is the TRANSPOSE with sorted columns.  Sorted order falls out for free from a bitmap: bit (s, m)
// of a (N x M)-bit matrix says "row m references source s".
//   csr_mark_kernel : one thread per table entry: cnt[s] += 1, bm[s][m] |= 1 (integer atomics: order-independent results)
//   csr_scan1_kernel: exclusive scan of the counts -> offsets                     (one CTA per cloud, one pass)
//   csr_emit_kernel : one warp per source s walks its M bits in ascending m; for every set bit it re-reads row m of the
//                     table (K entries, one or a few 16-byte loads) and emits the matching position(s) m*K + r.
// No sort, no per-range count matrix, every SM busy (the old stable counting sort walked 2048-position ranges with one
// warp each: a serial chain of exposed global-load latencies, 0.68 ms per DGCNN step on 10-16 CTAs).  Rows that hold the
// same source twice (possible only in caller-made tables; kNN / ball-query rows are distinct) are flagged by the mark
// kernel; the emit kernel then counts matches per bit before placing them.
constexpr size_t CSR_BITMAP_BUDGET = (size_t)512 << 20;       // bytes of bitmap the workspace may take

__global__ void __launch_bounds__(256)
csr_mark_kernel(const int32_t* __restrict__ idx, int M, int K, int N, int W, int32_t* __restrict__ cnt,
                uint32_t* __restrict__ bm, int32_t* __restrict__ dupflag) {
    const int b = blockIdx.y;
    const long E = (long)M * K;
    const int32_t* __restrict__ ib = idx + (size_t)b * E;
    bool any_dup = false;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long)gridDim.x * blockDim.x) {
        const int m = (int)(e / K), r = (int)(e - (long)m * K);
        const int s = ib[e];
        if ((unsigned)s >= (unsigned)N) continue;                       // out-of-range entries are ignored, never dereferenced
        bool dup = false;
        for (int r2 = 0; r2 < r; ++r2) dup |= (ib[(long)m * K + r2] == s);           // the row is L1-resident
        atomicAdd(&cnt[(size_t)b * (N + 1) + s], 1);
        if (!dup) atomicOr(&bm[((size_t)b * N + s) * W + (m >> 5)], 1u << (m & 31));
        any_dup |= dup;
    }
    if (__any_sync(PCNBR_FULL, any_dup) && (threadIdx.x & 31) == 0) atomicOr(&dupflag[b], 1);
}

// offsets[b, 0..N] = exclusive scan of cnt[b, 0..N-1]: 1024 threads x ITEMS contiguous counts each, one pass
__global__ void __launch_bounds__(1024)
csr_scan1_kernel(const int32_t* __restrict__ cnt, int N, int32_t* __restrict__ offsets) {
    __shared__ int s_warp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t* __restrict__ c = cnt + (size_t)b * (N + 1);
    int32_t* __restrict__ o = offsets + (size_t)b * (N + 1);
    const int items = (N + 1 + 1023) / 1024;
    const int i0 = tid * items, i1 = min(N + 1, i0 + items);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += (i < N) ? c[i] : 0;
    int x = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(PCNBR_FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(PCNBR_FULL, w, d);
            if (lane >= d) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    int run = (warp ? s_warp[warp - 1] : 0) + x - sum;
    for (int i = i0; i < i1; ++i) {
        o[i] = run;
        run += (i < N) ? c[i] : 0;
    }
}

// number of entries of table row `row` (K ints) equal to s; the first match position goes to r_first
__device__ __forceinline__ int csr_row_matches(const int32_t* __restrict__ row, int K, int s, int& r_first) {
    int n = 0;
    r_first = -1;
    if ((K & 3) == 0) {
        const int4* __restrict__ v = reinterpret_cast<const int4*>(row);
#pragma unroll 4
        for (int q = 0; q < K / 4; ++q) {
            const int4 t = v[q];
            if (t.x == s) { if (!n) r_first = 4 * q; ++n; }
            if (t.y == s) { if (!n) r_first = 4 * q + 1; ++n; }
            if (t.z == s) { if (!n) r_first = 4 * q + 2; ++n; }
            if (t.w == s) { if (!n) r_first = 4 * q + 3; ++n; }
        }
    } else {
        for (int r = 0; r < K; ++r)
            if (row[r] == s) { if (!n) r_first = r; ++n; }
    }
    return n;
}

__global__ void __launch_bounds__(256)
csr_emit_kernel(const int32_t* __restrict__ idx, int M, int K, int N, int W, const int32_t* __restrict__ offsets,
                const uint32_t* __restrict__ bm, const int32_t* __restrict__ dupflag, int32_t* __restrict__ perm, int only_flagged) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long E = (long)M * K;
    const int32_t* __restrict__ ib = idx + (size_t)b * E;
    int32_t* __restrict__ pm = perm + (size_t)b * E;
    const int32_t* __restrict__ off = offsets + (size_t)b * (N + 1);
    const bool multi = dupflag[b] != 0;
    if (only_flagged && !multi) return;                  // this cloud was handled by csr_emit_entries_kernel
    for (int s = blockIdx.x * 8 + warp; s < N; s += gridDim.x * 8) {
        int out = off[s];
        if (off[s + 1] == out) continue;
        const uint32_t* __restrict__ bits_row = bm + ((size_t)b * N + s) * W;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const uint32_t word = (w0 + lane < W) ? bits_row[w0 + lane] : 0u;
            if (!__any_sync(PCNBR_FULL, word != 0u)) continue;
            int c = __popc(word);
            if (multi) {                                               // rows may hold s more than once: count first
                c = 0;
                uint32_t bits = word;
                while (bits) {
                    const int m = ((w0 + lane) << 5) + __ffs(bits) - 1;
                    bits &= bits - 1;
                    int rf;
                    c += csr_row_matches(ib + (long)m * K, K, s, rf);
                }
            }
            int x = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(PCNBR_FULL, x, d);
                if (lane >= d) x += y;
            }
            int base = out + x - c;
            out += __shfl_sync(PCNBR_FULL, x, 31);
            uint32_t bits = word;
            while (bits) {
                const int m = ((w0 + lane) << 5) + __ffs(bits) - 1;
                bits &= bits - 1;
                const int32_t* __restrict__ row = ib + (long)m * K;
                if (!multi) {
                    int rf;
                    csr_row_matches(row, K, s, rf);
                    pm[base++] = m * K + rf;
                } else {
                    for (int r = 0; r < K; ++r)
                        if (row[r] == s) pm[base++] = m * K + r;
                }
            }
        }
    }
}

}  // namespace pcnbr

using namespace pcnbr;

// The stable counting sort pays a fixed 3 N words of shared/global traffic per warp range, so it wins where segments are
// long on average (kNN graphs, E/N = k >= 16: 0.58 -> 0.47 ms per DGCNN step) and loses on the short-segment tables of
// PointNet++ (E/N = 8-12), which keep count / scan / fill / sort with the bitmap path for their few long segments.
static bool csr_use_stable(int E, int N) { return N <= CSR_STABLE_MAX_N && (long)E >= 16L * N; }

extern "C" size_t pcnbr_csr_ws_bytes(int B, int E, int N) {
    // cnt/cursor (B,N+1) + max(tmp (B,E), per-range count matrix (B,P,N))
    const size_t P = ((size_t)E + CSR_CH - 1) / CSR_CH;
    const size_t tmp = (size_t)B * E, cm = csr_use_stable(E, N) ? (size_t)B * P * N : 0;
    return sizeof(int32_t) * ((size_t)B * (N + 1) + (tmp > cm ? tmp : cm));
}

extern "C" int pcnbr_csr_build(const int32_t* idx, int B, int E, int N, int32_t* offsets, int32_t* perm,
                               void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    if (!idx || !offsets || !perm || B <= 0 || E <= 0 || N <= 0) return PCNBR_E_BADARG;
    if (!ws || ws_bytes < pcnbr_csr_ws_bytes(B, E, N)) return PCNBR_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* cnt = (int32_t*)ws;
    int32_t* tmp = cnt + (size_t)B * (N + 1);
    if (csr_use_stable(E, N)) {
        const int P = (E + CSR_CH - 1) / CSR_CH;
        uint32_t* cm = (uint32_t*)tmp;
        const size_t smem = (size_t)CSR_SW * N * sizeof(uint32_t);
        if (smem > 48 * 1024) {
            cudaError_t ea = cudaFuncSetAttribute(csr_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ea == cudaSuccess) ea = cudaFuncSetAttribute(csr_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ea != cudaSuccess) return (int)ea;
        }
        const dim3 gw((P + CSR_SW - 1) / CSR_SW, B);
        PCNBR_TIMED("csr_hist_kernel", s, (double)B * (4.0 * E + 4.0 * P * N), 0.0,
                    (csr_hist_kernel<<<gw, 32 * CSR_SW, smem, s>>>(idx, E, N, P, cm)));
        PCNBR_CHECK_LAUNCH();
        PCNBR_TIMED("csr_colscan_kernel", s, (double)B * (8.0 * P * N + 4.0 * N), 0.0,
                    (csr_colscan_kernel<<<dim3((N + 255) / 256, B), 256, 0, s>>>(cm, N, P, cnt)));
        PCNBR_CHECK_LAUNCH();
        PCNBR_TIMED("csr_scan_kernel", s, (double)B * 12.0 * N, 0.0, (csr_scan_kernel<<<B, 1024, 0, s>>>(cnt, N, offsets, cnt)));
        PCNBR_CHECK_LAUNCH();
        PCNBR_TIMED("csr_place_kernel", s, (double)B * (8.0 * E + 4.0 * P * N + 4.0 * N), 0.0,
                    (csr_place_kernel<<<gw, 32 * CSR_SW, smem, s>>>(idx, E, N, P, cm, offsets, perm)));
        PCNBR_CHECK_LAUNCH();
        return 0;
    }
    cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)B * (N + 1), s);
    if (e != cudaSuccess) return (int)e;
    const int gx = min((E + 255) / 256, 1184);
    // algorithmic bytes: the table is read once per pass (4E), plus the pass's own output
    PCNBR_TIMED("csr_count_kernel", s, (double)B * (4.0 * E + 4.0 * N), 0.0, (csr_count_kernel<<<dim3(gx, B), 256, 0, s>>>(idx, E, N, cnt)));
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("csr_scan_kernel", s, (double)B * 12.0 * N, 0.0, (csr_scan_kernel<<<B, 1024, 0, s>>>(cnt, N, offsets, cnt)));       // cursor overwrites cnt in place
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("csr_fill_kernel", s, (double)B * 8.0 * E, 0.0, (csr_fill_kernel<<<dim3(gx, B), 256, 0, s>>>(idx, E, N, cnt, tmp)));
    PCNBR_CHECK_LAUNCH();
    const int gs = min((N + 7) / 8, 1184);
    const int use_bitmap = E <= CSR_BITMAP_MAX_E ? 1 : 0;
    const size_t bm_bytes = use_bitmap ? (size_t)((E + 31) / 32) * 4 : 0;
    if (bm_bytes > 48 * 1024) {
        cudaError_t ea = cudaFuncSetAttribute(csr_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_bytes);
        if (ea != cudaSuccess) return (int)ea;
    }
    PCNBR_TIMED("csr_sort_kernel", s, (double)B * (8.0 * E + 4.0 * N), 0.0,
                (csr_sort_kernel<<<dim3(gs, B), 256, bm_bytes, s>>>(offsets, tmp, E, N, perm, use_bitmap)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}


namespace pcnbr {
// Entry-side emit for tables whose rows hold distinct sources (every kNN / ball-query / 3-NN table): the position of entry
// (m, r) inside its source's segment is the number of rows below m that reference the same source = a prefix popcount of
// that source's bitmap row.  One thread per entry, at most W/4 16-byte loads each, the same cost for a hub referenced by
// every row as for a leaf -- the source-side walk above spends 32 x K loads per lane on a hub (139 us for the SA1 ball
// table, whose 32 padding indices appear in almost every row).  Clouds flagged by csr_mark_kernel (a row holds a source
// twice) are left to csr_emit_kernel.
__global__ void __launch_bounds__(256)
csr_emit_entries_kernel(const int32_t* __restrict__ idx, int M, int K, int N, int W, const int32_t* __restrict__ offsets,
                        const uint32_t* __restrict__ bm, const int32_t* __restrict__ dupflag, int32_t* __restrict__ perm) {
    const int b = blockIdx.y;
    if (dupflag[b] != 0) return;
    const long E = (long)M * K;
    const int32_t* __restrict__ ib = idx + (size_t)b * E;
    int32_t* __restrict__ pm = perm + (size_t)b * E;
    const int32_t* __restrict__ off = offsets + (size_t)b * (N + 1);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long)gridDim.x * blockDim.x) {
        const int m = (int)(e / K);
        const int s = ib[e];
        if ((unsigned)s >= (unsigned)N) continue;
        const uint32_t* __restrict__ row = bm + ((size_t)b * N + s) * W;
        const int wfull = m >> 5;
        int rank = 0, w = 0;
        if ((W & 3) == 0) {
            const uint4* __restrict__ row4 = reinterpret_cast<const uint4*>(row);
            for (; w + 4 <= wfull; w += 4) {
                const uint4 v = row4[w >> 2];
                rank += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
            }
        }
        for (; w < wfull; ++w) rank += __popc(row[w]);
        rank += __popc(row[wfull] & ((1u << (m & 31)) - 1u));
        pm[off[s] + rank] = (int32_t)e;
    }
}
}  // namespace pcnbr

static bool csr_rows_use_bitmap(int B, int M, int N) {
    const size_t W = ((size_t)M + 31) / 32;
    return (size_t)B * N * W * 4 <= CSR_BITMAP_BUDGET;
}

extern "C" size_t pcnbr_csr_rows_ws_bytes(int B, int M, int K, int N) {
    const long E = (long)M * K;
    if (E > 0x7fffffffL) return 0;
    if (csr_use_stable((int)E, N) || !csr_rows_use_bitmap(B, M, N)) return pcnbr_csr_ws_bytes(B, (int)E, N);
    const size_t W = ((size_t)M + 31) / 32;
    // cnt (B,N+1) + duplicate flags (B, padded to 64 words) + bitmap (B,N,W)
    return sizeof(int32_t) * ((((size_t)B * (N + 1) + 3) & ~(size_t)3) + (size_t)((B + 63) / 64 * 64) + (size_t)B * N * W);
}

// Inverse of a row-structured table idx (B,M,K) into N sources: same result as pcnbr_csr_build(idx, B, M*K, N, ...).
extern "C" int pcnbr_csr_build_rows(const int32_t* idx, int B, int M, int K, int N, int32_t* offsets, int32_t* perm,
                                    void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    if (!idx || !offsets || !perm || B <= 0 || M <= 0 || K <= 0 || N <= 0) return PCNBR_E_BADARG;
    const long E = (long)M * K;
    if (E > 0x7fffffffL) return PCNBR_E_TOOLARGE;
    if (!ws || ws_bytes < pcnbr_csr_rows_ws_bytes(B, M, K, N)) return PCNBR_E_WORKSPACE;
    // dense tables (kNN graphs, E/N >= 16) keep the stable counting sort: it moves ~4x fewer bytes than the bitmap path and,
    // running on the side stream under the forward pass, total work matters more than its own latency (measured: DGCNN
    // 6.12 ms/step with it, 6.40 with the bitmap path, which takes SMs from the main stream)
    if (csr_use_stable((int)E, N) || !csr_rows_use_bitmap(B, M, N))
        return pcnbr_csr_build(idx, B, (int)E, N, offsets, perm, ws, ws_bytes, stream);
    cudaStream_t s = (cudaStream_t)stream;
    const int W = (M + 31) / 32;
    int32_t* cnt = (int32_t*)ws;
    int32_t* dup = cnt + (((size_t)B * (N + 1) + 3) & ~(size_t)3);          // keeps the bitmap 16-byte aligned (uint4 reads)
    uint32_t* bm = (uint32_t*)(dup + (size_t)((B + 63) / 64 * 64));
    cudaError_t e = cudaMemsetAsync(ws, 0, pcnbr_csr_rows_ws_bytes(B, M, K, N), s);
    if (e != cudaSuccess) return (int)e;
    const int gx = (int)((E + 255) / 256 < 2368 ? (E + 255) / 256 : 2368);
    // algorithmic bytes: the table (4E) is read by mark and emit, perm (4E) and offsets are written; the bitmap is scratch
    PCNBR_TIMED("csr_mark_kernel", s, (double)B * (4.0 * E + 4.0 * N), 0.0,
                (csr_mark_kernel<<<dim3(gx, B), 256, 0, s>>>(idx, M, K, N, W, cnt, bm, dup)));
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("csr_scan1_kernel", s, (double)B * 8.0 * N, 0.0, (csr_scan1_kernel<<<B, 1024, 0, s>>>(cnt, N, offsets)));
    PCNBR_CHECK_LAUNCH();
    const int gs = (N + 7) / 8 < 2368 ? (N + 7) / 8 : 2368;
    if (W <= 256) {
        // rows with distinct sources (the normal case): one thread per entry; flagged clouds fall through to the walk below
        PCNBR_TIMED("csr_emit_entries_kernel", s, (double)B * (8.0 * E + 4.0 * N), 0.0,
                    (csr_emit_entries_kernel<<<dim3(gx, B), 256, 0, s>>>(idx, M, K, N, W, offsets, bm, dup, perm)));
        PCNBR_CHECK_LAUNCH();
        PCNBR_TIMED("csr_emit_kernel", s, 0.0, 0.0,
                    (csr_emit_kernel<<<dim3(gs, B), 256, 0, s>>>(idx, M, K, N, W, offsets, bm, dup, perm, 1)));
    } else {
        PCNBR_TIMED("csr_emit_kernel", s, (double)B * (8.0 * E + 4.0 * N), 0.0,
                    (csr_emit_kernel<<<dim3(gs, B), 256, 0, s>>>(idx, M, K, N, W, offsets, bm, dup, perm, 0)));
    }
    PCNBR_CHECK_LAUNCH();
    return 0;
}
