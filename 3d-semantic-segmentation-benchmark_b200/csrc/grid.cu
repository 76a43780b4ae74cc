// grid.cu -- K2 / K3 on a uniform cell grid: ball query and xyz k-NN without the M x N scan.
//
// Reference: models/utils/common.py:54-61 (the (B,C,N) distance tensor + mask + topk inside group()) and :110-114 (the
// 3-NN of interpolate()).  select.cu evaluates all M*N pairs (121 us / 192 us for 32 x 4096 x 1024, issue-bound at 5-12 %
// of the FP32 roofline) although fewer than 0.2 % of the pairs can be inside a ball.  Here every cloud is binned once
// into cells of edge h >= 1.001 r (counting sort: count -> scan -> fill), a query only reads the 27 cells around its own,
// and the selection is the SAME total order as the brute-force kernel -- ascending (d2, index), d2 in the reference's
// exact arithmetic -- so the result is bit-identical:
//   ball query : the in-ball points sorted by (d2, index); an under-filled ball is padded with the lowest indices that are
//                not in the ball (what a stable sort of the masked distance row yields: all +inf keys, ascending index),
//                generated from the member list without a single distance;
//   k-NN       : rings of cells around the query are scanned until the k-th best distance is strictly inside the
//                scanned cube (any unseen point is farther), so the k best over ALL points have been seen.
// Cell membership only prunes; every decision uses d2_direct() on the original coordinates.
#include "common.cuh"

namespace pcnbr {

struct GridParams {
    float ox, oy, oz, h, inv_h;
    int nx, ny, nz, ncell;
};

constexpr int GRID_CELL_CAP = 32768;       // cells per cloud (the cell edge grows until the grid fits)
constexpr int GRID_DIM_CAP = 512;          // cells per axis (keeps fp32 cell coordinates exact to ~1e-4 of a cell)

// One CTA per cloud: bounding box -> grid origin, cell edge, dimensions.  h_req > 0: requested edge (ball query);
// h_req <= 0: edge from the density, ~ppc points per cell of the bounding box (k-NN).
__global__ void __launch_bounds__(256)
grid_bbox_kernel(const float* __restrict__ p, int N_alloc, const int32_t* __restrict__ n_src, float h_req, float ppc,
                 GridParams* __restrict__ gp) {
    __shared__ float red[6][8];
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = len_valid(n_src, b, N_alloc);              // length-aware form: the padding rows are not binned
    const float* __restrict__ pb = p + (size_t)b * N_alloc * 3;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int n = threadIdx.x; n < N; n += 256)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pb[3 * n + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(PCNBR_FULL, lo[a], d));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(PCNBR_FULL, hi[a], d));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ext[3];
        GridParams g;
        for (int a = 0; a < 3; ++a) {
            float l = red[a][0], u = red[3 + a][0];
            for (int w = 1; w < 8; ++w) { l = fminf(l, red[a][w]); u = fmaxf(u, red[3 + a][w]); }
            (a == 0 ? g.ox : a == 1 ? g.oy : g.oz) = l;
            ext[a] = fmaxf(u - l, 0.f);
        }
        const float emax = fmaxf(fmaxf(ext[0], ext[1]), fmaxf(ext[2], 1e-12f));
        float h = h_req;
        if (!(h > 0.f)) {
            const float ex = fmaxf(ext[0], 1e-3f * emax), ey = fmaxf(ext[1], 1e-3f * emax), ez = fmaxf(ext[2], 1e-3f * emax);
            h = cbrtf(ex * ey * ez * ppc / (float)N);
        }
        h = fmaxf(h, emax / (float)(GRID_DIM_CAP - 1));
        if (!(h > 0.f) || !(h < 3.0e38f)) h = 1.0f;             // non-finite coordinates: one cell, every query scans everything
        int nx = 1, ny = 1, nz = 1;
        for (int it = 0; it < 200; ++it) {
            nx = (int)fminf(ext[0] / h, (float)GRID_DIM_CAP) + 1; ny = (int)fminf(ext[1] / h, (float)GRID_DIM_CAP) + 1;
            nz = (int)fminf(ext[2] / h, (float)GRID_DIM_CAP) + 1;
            if ((long)nx * ny * nz <= GRID_CELL_CAP) break;
            h *= 1.25f;
        }
        if ((long)nx * ny * nz > GRID_CELL_CAP) { nx = ny = nz = 1; h = 3.0e37f; }
        g.h = h; g.inv_h = 1.0f / h; g.nx = nx; g.ny = ny; g.nz = nz; g.ncell = nx * ny * nz;
        gp[b] = g;
    }
}

// cell coordinate along one axis, clamped to [-1, n] BEFORE the conversion to int (a query far outside the bounding box, or
// a degenerate cloud with a tiny cell edge, would overflow the conversion): -1 / n mean "outside, below / above the grid",
// which is all the consumers need (no point lives there)
__device__ __forceinline__ int grid_coord(float x, float o, float inv_h, int n) {
    const float t = floorf((x - o) * inv_h);
    return (int)fminf(fmaxf(t, -1.0f), (float)n);
}
__device__ __forceinline__ int grid_clamp(int c, int n) { return c < 0 ? 0 : (c >= n ? n - 1 : c); }

// cell of every point (clamped into the grid) + per-cell counts (integer atomics: the counts do not depend on the order)
__global__ void __launch_bounds__(256)
grid_count_kernel(const float* __restrict__ p, int N, const int32_t* __restrict__ n_src, const GridParams* __restrict__ gp,
                  int32_t* __restrict__ cellid, int32_t* __restrict__ count) {
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    if (n >= len_valid(n_src, b, N)) return;
    const GridParams g = gp[b];
    const float* __restrict__ s = p + ((size_t)b * N + n) * 3;
    const int cx = grid_clamp(grid_coord(s[0], g.ox, g.inv_h, g.nx), g.nx), cy = grid_clamp(grid_coord(s[1], g.oy, g.inv_h, g.ny), g.ny),
              cz = grid_clamp(grid_coord(s[2], g.oz, g.inv_h, g.nz), g.nz);
    const int c = (cz * g.ny + cy) * g.nx + cx;
    cellid[(size_t)b * N + n] = c;
    atomicAdd(&count[(size_t)b * (GRID_CELL_CAP + 1) + c], 1);
}

// start[b, 0..ncell] = exclusive scan of the counts, in place; cursor = copy
__global__ void __launch_bounds__(1024)
grid_scan_kernel(int32_t* __restrict__ count, int32_t* __restrict__ cursor, const GridParams* __restrict__ gp) {
    __shared__ int s_warp[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nc = gp[b].ncell;
    int32_t* __restrict__ c = count + (size_t)b * (GRID_CELL_CAP + 1);
    int32_t* __restrict__ cur = cursor + (size_t)b * GRID_CELL_CAP;
    const int items = (nc + 1 + 1023) / 1024;
    const int i0 = min(tid * items, nc + 1), i1 = min(nc + 1, i0 + items);
    int sum = 0;
    for (int i = i0; i < i1; ++i) sum += (i < nc) ? c[i] : 0;
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
        const int v = (i < nc) ? c[i] : 0;
        c[i] = run;
        if (i < nc) cur[i] = run;
        run += v;
    }
}

// points grouped by cell: (x, y, z, index) as one float4 each (the order inside a cell is arbitrary: every consumer sorts
// its candidates by (d2, index))
__global__ void __launch_bounds__(256)
grid_fill_kernel(const float* __restrict__ p, int N, const int32_t* __restrict__ n_src, const int32_t* __restrict__ cellid,
                 int32_t* __restrict__ cursor, float4* __restrict__ sorted) {
    const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
    if (n >= len_valid(n_src, b, N)) return;
    const float* __restrict__ s = p + ((size_t)b * N + n) * 3;
    const int c = cellid[(size_t)b * N + n];
    const int pos = atomicAdd(&cursor[(size_t)b * GRID_CELL_CAP + c], 1);
    sorted[(size_t)b * N + pos] = make_float4(s[0], s[1], s[2], __int_as_float(n));
}

// insert the candidates of one contiguous run of cells into the warp's sorted list
template <int NSLOT, bool RADIUS>
__device__ __forceinline__ void grid_scan_range(const float4* __restrict__ pts, int beg, int end, float qx, float qy, float qz,
                                                float r2, int K, WarpList<NSLOT>& list, u64& thr, int lane) {
    for (int c0 = beg; c0 < end; c0 += 32) {
        const int j = c0 + lane;
        u64 key = PCNBR_KEY_MAX;
        if (j < end) {
            const float4 s = pts[j];
            const float d2 = d2_direct(s.x, s.y, s.z, qx, qy, qz);
            if (!RADIUS || d2 <= r2) key = pack_key(f2ord(d2), (uint32_t)__float_as_int(s.w));     // common.py:56-59 / :112
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
}

// The 3 x 3 x 3 cells around (cx, cy, cz) as ONE flattened candidate stream: lanes 0..8 fetch the bounds of the nine
// x-adjacent runs at once (one load latency instead of nine dependent ones), an exclusive prefix over the run lengths maps
// a flat candidate number to (run, offset), and the warp walks the stream 32 candidates at a time.
template <int NSLOT, bool RADIUS>
__device__ __forceinline__ void grid_scan_cube1(const float4* __restrict__ pts, const int32_t* __restrict__ st, const GridParams& g,
                                                int cx, int cy, int cz, float qx, float qy, float qz, float r2, int K,
                                                WarpList<NSLOT>& list, u64& thr, int lane) {
    int beg = 0, len = 0;
    if (lane < 9) {
        const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
        if (z >= 0 && z < g.nz && y >= 0 && y < g.ny && x0 <= x1) {
            const int row = (z * g.ny + y) * g.nx;
            beg = st[row + x0];
            len = st[row + x1 + 1] - beg;
        }
    }
    int off = len;                                       // inclusive prefix over lanes 0..8
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const int y = __shfl_up_sync(PCNBR_FULL, off, d);
        if (lane >= d) off += y;
    }
    const int total = __shfl_sync(PCNBR_FULL, off, 8);
    off -= len;                                          // exclusive
    for (int c0 = 0; c0 < total; c0 += 32) {
        const int t = c0 + lane;
        int run = 0;
#pragma unroll
        for (int r = 1; r < 9; ++r) run += (t >= __shfl_sync(PCNBR_FULL, off, r)) ? 1 : 0;      // offsets are non-decreasing
        const int rb = __shfl_sync(PCNBR_FULL, beg, run), ro = __shfl_sync(PCNBR_FULL, off, run);
        u64 key = PCNBR_KEY_MAX;
        if (t < total) {
            const float4 s = pts[rb + (t - ro)];
            const float d2 = d2_direct(s.x, s.y, s.z, qx, qy, qz);
            if (!RADIUS || d2 <= r2) key = pack_key(f2ord(d2), (uint32_t)__float_as_int(s.w));
        }
        if (c0 == 0 && !RADIUS) {
            // the first 32 candidates all enter the empty list: ONE warp sort instead of 32 sequential inserts
            list.v[0] = warp_sort64(key, lane);
            thr = list.at(K - 1);
            continue;
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
}

// Ball query: one warp per query, the 3 x 3 x 3 cells around it (3 x-adjacent cells are one contiguous run).
template <int NSLOT>
__global__ void __launch_bounds__(256)
ball_grid_kernel(const float* __restrict__ q, int M, int N_alloc, float r2, int K, const int32_t* __restrict__ n_qry,
                 const int32_t* __restrict__ n_src, const GridParams* __restrict__ gp,
                 const int32_t* __restrict__ start, const float4* __restrict__ sorted, int32_t* __restrict__ idx) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * 8 + warp;
    if (m >= M) return;
    const int N = len_valid(n_src, b, N_alloc);
    if (m >= len_valid(n_qry, b, M)) { len_fill_row(idx, nullptr, ((size_t)b * M + m) * K, K, lane); return; }
    const GridParams g = gp[b];
    const float* __restrict__ c = q + ((size_t)b * M + m) * 3;
    const float qx = c[0], qy = c[1], qz = c[2];
    const int32_t* __restrict__ st = start + (size_t)b * (GRID_CELL_CAP + 1);
    const float4* __restrict__ pts = sorted + (size_t)b * N_alloc;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.nx), cy = grid_coord(qy, g.oy, g.inv_h, g.ny), cz = grid_coord(qz, g.oz, g.inv_h, g.nz);
    WarpList<NSLOT> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;
    grid_scan_cube1<NSLOT, true>(pts, st, g, cx, cy, cz, qx, qy, qz, r2, K, list, thr, lane);
    // members first (ascending (d2, index)), then the lowest indices that are not members
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) cnt += __popc(__ballot_sync(PCNBR_FULL, list.v[s] != PCNBR_KEY_MAX));
    int32_t* __restrict__ out = idx + ((size_t)b * M + m) * K;
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const int pos = s * 32 + lane;
        if (pos < K && pos < cnt) out[pos] = (int32_t)(uint32_t)list.v[s];
    }
    int filled = cnt;
    for (int j0 = 0; filled < K && j0 < N; j0 += 32) {
        const int j = j0 + lane;
        bool member = false;
        for (int t = 0; t < cnt; ++t) member |= ((int)(uint32_t)list.at(t) == j);
        const uint32_t mask = __ballot_sync(PCNBR_FULL, j < N && !member);
        const int pos = filled + __popc(mask & ((1u << lane) - 1u));
        if (j < N && !member && pos < K) out[pos] = j;
        filled += __popc(mask);
    }
}

// k-NN (k <= 32): rings of cells until the k-th best distance lies strictly inside the scanned cube.
__global__ void __launch_bounds__(256)
knn_grid_kernel(const float* __restrict__ q, int M, int N, int K, const int32_t* __restrict__ n_qry,
                const GridParams* __restrict__ gp,
                const int32_t* __restrict__ start, const float4* __restrict__ sorted, int32_t* __restrict__ idx,
                float* __restrict__ d2out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = blockIdx.x * 8 + warp;
    if (m >= M) return;
    if (m >= len_valid(n_qry, b, M)) { len_fill_row(idx, d2out, ((size_t)b * M + m) * K, K, lane); return; }
    const GridParams g = gp[b];
    const float* __restrict__ c = q + ((size_t)b * M + m) * 3;
    const float qx = c[0], qy = c[1], qz = c[2];
    const int32_t* __restrict__ st = start + (size_t)b * (GRID_CELL_CAP + 1);
    const float4* __restrict__ pts = sorted + (size_t)b * N;
    const int cx = grid_coord(qx, g.ox, g.inv_h, g.nx), cy = grid_coord(qy, g.oy, g.inv_h, g.ny), cz = grid_coord(qz, g.oz, g.inv_h, g.nz);
    WarpList<1> list;
    list.init();
    u64 thr = PCNBR_KEY_MAX;
    const float slack = 2e-3f * g.h;                    // cells are assigned with fp32 rounding: shrink the safe margin
    const int rmax = max(g.nx, max(g.ny, g.nz)) + 2;   // the query cell is within [-1, n]: the cube covers the grid by then
    grid_scan_cube1<1, false>(pts, st, g, cx, cy, cz, qx, qy, qz, 0.f, K, list, thr, lane);      // R = 0 and R = 1 in one stream
    for (int R = 1; R <= rmax; ++R) {
        // shell R >= 2: cells with max(|dx|, |dy|, |dz|) == R, clipped to the grid (the cube above covered R <= 1)
        if (R >= 2)
        for (int z = max(cz - R, 0); z <= min(cz + R, g.nz - 1); ++z)
            for (int y = max(cy - R, 0); y <= min(cy + R, g.ny - 1); ++y) {
                const int row = (z * g.ny + y) * g.nx;
                const bool face = (abs(z - cz) == R) || (abs(y - cy) == R);
                if (face) {
                    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
                    if (x0 <= x1) grid_scan_range<1, false>(pts, st[row + x0], st[row + x1 + 1], qx, qy, qz, 0.f, K, list, thr, lane);
                } else {
                    const int xa = cx - R, xb = cx + R;
                    if (xa >= 0 && xa < g.nx) grid_scan_range<1, false>(pts, st[row + xa], st[row + xa + 1], qx, qy, qz, 0.f, K, list, thr, lane);
                    if (xb >= 0 && xb < g.nx && xb != xa) grid_scan_range<1, false>(pts, st[row + xb], st[row + xb + 1], qx, qy, qz, 0.f, K, list, thr, lane);
                }
            }
        // every point lies in a cell of the grid, so a side of the cube that reaches the grid's border is unbounded
        const bool all = cx - R <= 0 && cx + R >= g.nx - 1 && cy - R <= 0 && cy + R >= g.ny - 1 && cz - R <= 0 && cz + R >= g.nz - 1;
        if (all) break;
        if (thr != PCNBR_KEY_MAX) {
            float margin = 3.0e38f;
            if (cx - R > 0)        margin = fminf(margin, qx - (g.ox + (float)(cx - R) * g.h));
            if (cx + R < g.nx - 1) margin = fminf(margin, (g.ox + (float)(cx + R + 1) * g.h) - qx);
            if (cy - R > 0)        margin = fminf(margin, qy - (g.oy + (float)(cy - R) * g.h));
            if (cy + R < g.ny - 1) margin = fminf(margin, (g.oy + (float)(cy + R + 1) * g.h) - qy);
            if (cz - R > 0)        margin = fminf(margin, qz - (g.oz + (float)(cz - R) * g.h));
            if (cz + R < g.nz - 1) margin = fminf(margin, (g.oz + (float)(cz + R + 1) * g.h) - qz);
            margin -= slack;
            const float dk = ord2f((uint32_t)(thr >> 32));
            if (margin > 0.f && dk < margin * margin * 0.9999f) break;      // strictly inside: no unseen point can tie or beat it
        }
    }
    if (lane < K) {
        const size_t o = ((size_t)b * M + m) * K + lane;
        idx[o] = (int32_t)min((uint32_t)list.v[0], (uint32_t)(N - 1));     // (an empty slot -- K > valid sources -- stays in range)
        if (d2out) d2out[o] = ord2f((uint32_t)(list.v[0] >> 32));
    }
}

struct GridWs {
    GridParams* gp;
    int32_t *cellid, *count, *cursor;
    float4* sorted;
    size_t bytes;
};

static size_t grid_align(size_t x) { return (x + 255) & ~(size_t)255; }

static GridWs grid_carve(void* ws, int B, int N) {
    GridWs w;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += grid_align(n); return (uint8_t*)ws + o; };
    w.gp = (GridParams*)take(sizeof(GridParams) * (size_t)B);
    w.count = (int32_t*)take(4 * (size_t)B * (GRID_CELL_CAP + 1));
    w.cursor = (int32_t*)take(4 * (size_t)B * GRID_CELL_CAP);
    w.cellid = (int32_t*)take(4 * (size_t)B * N);
    w.sorted = (float4*)take(16 * (size_t)B * N);
    w.bytes = off;
    return w;
}

// bin the B clouds p (B,N,3): edge h_req (ball query) or from the density (k-NN, ~ppc points per cell)
static int grid_build(const float* p, int B, int N, const int32_t* n_src, float h_req, float ppc, const GridWs& w, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(w.count, 0, 4 * (size_t)B * (GRID_CELL_CAP + 1), s);
    if (e != cudaSuccess) return (int)e;
    PCNBR_TIMED("grid_bbox_kernel", s, 12.0 * B * N, 0.0, (grid_bbox_kernel<<<B, 256, 0, s>>>(p, N, n_src, h_req, ppc, w.gp)));
    PCNBR_CHECK_LAUNCH();
    const dim3 gn((N + 255) / 256, B);
    PCNBR_TIMED("grid_count_kernel", s, 16.0 * B * N, 0.0, (grid_count_kernel<<<gn, 256, 0, s>>>(p, N, n_src, w.gp, w.cellid, w.count)));
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("grid_scan_kernel", s, 12.0 * B * GRID_CELL_CAP, 0.0, (grid_scan_kernel<<<B, 1024, 0, s>>>(w.count, w.cursor, w.gp)));
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("grid_fill_kernel", s, 32.0 * B * N, 0.0, (grid_fill_kernel<<<gn, 256, 0, s>>>(p, N, n_src, w.cellid, w.cursor, w.sorted)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" size_t pcnbr_grid_ws_bytes(int B, int N) { return grid_carve(nullptr, B, N).bytes; }

// Same result as pcnbr_ball_query_f32, bit for bit.  ws: pcnbr_grid_ws_bytes(B, N).
extern "C" int pcnbr_ball_query_grid_f32(const float* q, const float* p, int B, int M, int N, float r2, int K, int32_t* idx,
                                         void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    return pcnbr_ball_query_len_f32(q, p, B, M, N, r2, K, nullptr, nullptr, idx, ws, ws_bytes, stream);
}

// Length-aware ball query: cloud b has n_src[b] real points and n_qry[b] real centroids (NULL: all).  ws == NULL: M x N scan.
extern "C" int pcnbr_ball_query_len_f32(const float* q, const float* p, int B, int M, int N, float r2, int K, const int32_t* n_qry,
                                        const int32_t* n_src, int32_t* idx, void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    if (!ws) return select_ball_len(q, p, B, M, N, r2, K, n_qry, n_src, idx, (cudaStream_t)stream);
    if (!q || !p || !idx || B <= 0 || M <= 0 || N <= 0 || K <= 0 || K > N || !(r2 >= 0.f)) return PCNBR_E_BADARG;
    if (K > 128) return PCNBR_E_TOOLARGE;
    if (!ws || ws_bytes < pcnbr_grid_ws_bytes(B, N)) return PCNBR_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const GridWs w = grid_carve(ws, B, N);
    // cell edge: 0.1 % above the radius -- the fp32 test d2 <= r2 can accept a point a few ulps beyond r, and the cell
    // coordinates carry ~1e-5 cells of rounding
    const float h = sqrtf(r2) * 1.001f + 1e-30f;
    int rc = grid_build(p, B, N, n_src, h, 0.f, w, s);
    if (rc) return rc;
    const dim3 grid((M + 7) / 8, B);
    // K2: compulsory bytes 12 (N + M) + 4 M K per cloud; the flops are what the scan of 27 cells costs
    // (~27 * N / ncell candidates per query), stated as 8 flop per candidate pair at the mean cell occupancy
    const double wb = (double)B * (12.0 * (N + M) + 4.0 * M * K), wf = 8.0 * B * (double)M * 27.0 * 2.0;
    if (K <= 32)      PCNBR_TIMED("ball_grid_kernel", s, wb, wf, (ball_grid_kernel<1><<<grid, 256, 0, s>>>(q, M, N, r2, K, n_qry, n_src, w.gp, w.count, w.sorted, idx)));
    else if (K <= 64) PCNBR_TIMED("ball_grid_kernel", s, wb, wf, (ball_grid_kernel<2><<<grid, 256, 0, s>>>(q, M, N, r2, K, n_qry, n_src, w.gp, w.count, w.sorted, idx)));
    else              PCNBR_TIMED("ball_grid_kernel", s, wb, wf, (ball_grid_kernel<4><<<grid, 256, 0, s>>>(q, M, N, r2, K, n_qry, n_src, w.gp, w.count, w.sorted, idx)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

// Same result as pcnbr_knn_direct_f32 (k <= 32), bit for bit.  ws: pcnbr_grid_ws_bytes(B, N).
extern "C" int pcnbr_knn_direct_grid_f32(const float* q, const float* p, int B, int M, int N, int K, int32_t* idx, float* d2,
                                         void* ws, size_t ws_bytes, pcnbr_stream_t stream) {
    return pcnbr_knn_direct_len_f32(q, p, B, M, N, K, nullptr, nullptr, idx, d2, ws, ws_bytes, stream);
}

// Length-aware k-NN (direct distances): n_qry[b] real queries, n_src[b] real sources (NULL: all).  ws == NULL: M x N scan.
extern "C" int pcnbr_knn_direct_len_f32(const float* q, const float* p, int B, int M, int N, int K, const int32_t* n_qry,
                                        const int32_t* n_src, int32_t* idx, float* d2, void* ws, size_t ws_bytes,
                                        pcnbr_stream_t stream) {
    if (!ws) return select_knn_len(q, p, B, M, N, K, n_qry, n_src, idx, d2, (cudaStream_t)stream);
    if (!q || !p || !idx || B <= 0 || M <= 0 || N <= 0 || K <= 0 || K > N) return PCNBR_E_BADARG;
    if (K > 32) return PCNBR_E_TOOLARGE;
    if (!ws || ws_bytes < pcnbr_grid_ws_bytes(B, N)) return PCNBR_E_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const GridWs w = grid_carve(ws, B, N);
    int rc = grid_build(p, B, N, n_src, 0.f, 2.0f, w, s);
    if (rc) return rc;
    const double wb = (double)B * (12.0 * (N + M) + 8.0 * M * K), wf = 8.0 * B * (double)M * 27.0 * 2.0;
    PCNBR_TIMED("knn_grid_kernel", s, wb, wf,
                (knn_grid_kernel<<<dim3((M + 7) / 8, B), 256, 0, s>>>(q, M, N, K, n_qry, w.gp, w.count, w.sorted, idx, d2)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
