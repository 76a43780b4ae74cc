// blocks.cu -- the data formats either side of the hot path.
//
// (1) block_batch_kernel: the S3DIS block dataloader's per-step work done in HBM.
//     Reference: data_processing/block_datasets.py:117-128 (`torch.load` of one block file per sample, random row
//     selection `points[sampled_indices]`, `labels[sampled_indices]`) and :5-29 (`collate_blocks`: zero-padded
//     (B,N,9) f32 / (B,N,14) u8 batch + lengths).  Here every block of the split lives packed back to back in HBM
//     ((T,9) f32 + (T,L) u8: the whole S3DIS set is ~14 GB of a 180 GB device) and a batch is ONE gather launch: no
//     file reads, no host collate, no H2D copy of the batch.  Lanes own consecutive output words (coalesced stores);
//     the 36-byte source rows come from HBM/L2 as they are.
//
// (2) window_merge_kernel: the tail of sliding-window scene inference.
//     Reference: models/dgcnn/utils.py:102-131 (`predict_single_scene`): windows of `window` points every `step`
//     points, `all_logits[start:end] += logits; point_counts[start:end] += 1` per window, then
//     `all_logits / point_counts`, `argmax`, `softmax(..).max()`.  Here the windows run through the model as ONE batch
//     and this kernel does the overlap-add (ascending window order = the reference's summation order, so the mean
//     logits are bit-identical), the division, the argmax (first maximum) and the confidence in one pass.
#include "common.cuh"

namespace pcnbr {

constexpr int BLK_PC = 9;                                    // floats per point record (xyz, rgb, normalised xyz)

// VEC: every thread produces 4 consecutive output words (points) / bytes (labels) from up to two source rows -- four
// independent loads in flight per thread and one 16-byte / 4-byte store (S % 4 == 0 keeps every slot's base aligned).
template <bool VEC>
__global__ void __launch_bounds__(256)
block_batch_kernel(const float* __restrict__ points, const uint8_t* __restrict__ labels,
                   const long long* __restrict__ block_start, const int* __restrict__ block_ids,
                   const int* __restrict__ sel, int S, int L, float* __restrict__ out_points,
                   uint8_t* __restrict__ out_labels, long long* __restrict__ out_len) {
    const int b = blockIdx.y;
    const int blk = block_ids[b];
    const long long base = block_start[blk];
    const long long nb = block_start[blk + 1] - base;
    const int* __restrict__ selb = sel ? sel + (size_t)b * S : nullptr;
    if (blockIdx.x == 0 && threadIdx.x == 0) out_len[b] = sel ? (long long)S : (nb < S ? nb : (long long)S);
    const float* __restrict__ pb = points + (size_t)base * BLK_PC;
    const uint8_t* __restrict__ lb = labels + (size_t)base * L;
    float* __restrict__ op = out_points + (size_t)b * S * BLK_PC;
    uint8_t* __restrict__ ol = out_labels + (size_t)b * S * L;
    const int nw = S * BLK_PC, nl = S * L;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    // source row of batch row r, or -1 for zero padding (collate_blocks :19-25) / an out-of-range selection
    auto src_row = [&](int r) -> long long {
        const long long s = selb ? (long long)selb[r] : (long long)r;
        return (s >= 0 && s < nb) ? s : -1;
    };
    if (VEC) {
        for (int q = t0; q < nw / 4; q += nt) {
            const int o = q * 4, r0 = o / BLK_PC, c0 = o - r0 * BLK_PC;
            const long long s0 = src_row(r0), s1 = (c0 + 3 >= BLK_PC) ? src_row(r0 + 1) : -1;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = c0 + j;
                const long long s = c < BLK_PC ? s0 : s1;
                v[j] = s >= 0 ? pb[s * BLK_PC + (c < BLK_PC ? c : c - BLK_PC)] : 0.f;
            }
            *reinterpret_cast<float4*>(op + o) = make_float4(v[0], v[1], v[2], v[3]);
        }
        for (int q = t0; q < nl / 4; q += nt) {
            const int o = q * 4;
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = (o + j) / L, c = (o + j) - r * L;
                const long long s = src_row(r);
                w |= (uint32_t)(s >= 0 ? lb[s * L + c] : (uint8_t)0) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(ol + o) = w;
        }
    } else {
        for (int o = t0; o < nw; o += nt) {
            const int r = o / BLK_PC, c = o - r * BLK_PC;
            const long long s = src_row(r);
            op[o] = s >= 0 ? pb[s * BLK_PC + c] : 0.f;
        }
        for (int o = t0; o < nl; o += nt) {
            const int r = o / L, c = o - r * L;
            const long long s = src_row(r);
            ol[o] = s >= 0 ? lb[s * L + c] : (uint8_t)0;
        }
    }
}

constexpr int WM_MAXC = 64;

// One thread per scene point.  Window w covers points [w*step, min(n, w*step + window)); its logits are rows
// win_off[w] .. of `logits` ((sum of window lengths), C).
template <int CMAX>
__global__ void __launch_bounds__(256)
window_merge_kernel(const float* __restrict__ logits, const long long* __restrict__ win_off, int W, long long n_points,
                    int window, int step, int C, float* __restrict__ mean_logits, long long* __restrict__ pred,
                    float* __restrict__ conf) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_points) return;
    // windows with w*step <= n < w*step + window
    long long w1 = n / step;
    if (w1 > W - 1) w1 = W - 1;
    long long w0 = n - window + 1 <= 0 ? 0 : (n - window + 1 + step - 1) / step;
    float acc[CMAX];                                                        // registers: every loop below is fully unrolled
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    int count = 0;
    for (long long w = w0; w <= w1; ++w) {
        const float* __restrict__ row = logits + (size_t)(win_off[w] + (n - w * step)) * C;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
            if (c < C) acc[c] = __fadd_rn(acc[c], row[c]);                   // all_logits[start:end] += logits, utils.py:120
        ++count;
    }
    const float cnt = (float)count;
    float best = 0.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
            const float v = __fdiv_rn(acc[c], cnt);                          // utils.py:124
            acc[c] = v;
            if (mean_logits) mean_logits[(size_t)n * C + c] = v;
            if (c == 0 || v > best) { best = v; arg = c; }                   // first maximum wins (torch.argmax)
        }
    }
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
        if (c < C) z += expf(acc[c] - best);
    pred[n] = arg;
    conf[n] = 1.f / z;                                                       // max of the softmax = exp(0) / sum
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_block_batch(const float* points, const uint8_t* labels, const long long* block_start,
                                 const int* block_ids, const int* sel, int B, int S, int L, float* out_points,
                                 uint8_t* out_labels, long long* out_len, pcnbr_stream_t stream) {
    if (!points || !labels || !block_start || !block_ids || !out_points || !out_labels || !out_len || B <= 0 || S <= 0 ||
        L <= 0)
        return PCNBR_E_BADARG;
    if ((long long)S * (BLK_PC > L ? BLK_PC : L) > 0x7fffffffLL || B > 65535) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const bool vec = (S % 4) == 0;                           // every slot's output base stays 16-byte / 4-byte aligned
    int gx = (S * BLK_PC / (vec ? 4 : 1) + 255) / 256;      // one pass: 4 words per thread, all loads of a thread in flight
    if (gx > 1024) gx = 1024;
    if (gx < 1) gx = 1;
    const double bytes = (double)B * S * (2.0 * (BLK_PC * 4 + L) + (sel ? 4.0 : 0.0)) + 8.0 * B;
    if (vec)
        PCNBR_TIMED("block_batch_kernel", s, bytes, 0.0,
                    (block_batch_kernel<true><<<dim3(gx, B), 256, 0, s>>>(points, labels, block_start, block_ids, sel, S, L,
                                                                           out_points, out_labels, out_len)));
    else
        PCNBR_TIMED("block_batch_kernel", s, bytes, 0.0,
                    (block_batch_kernel<false><<<dim3(gx, B), 256, 0, s>>>(points, labels, block_start, block_ids, sel, S, L,
                                                                            out_points, out_labels, out_len)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}

extern "C" int pcnbr_window_merge_f32(const float* logits, const long long* win_off, int W, long long n_points, int window,
                                      int step, int C, float* mean_logits, long long* pred, float* conf,
                                      pcnbr_stream_t stream) {
    if (!logits || !win_off || !pred || !conf || W <= 0 || n_points <= 0 || window <= 0 || step <= 0 || step > window || C <= 0)
        return PCNBR_E_BADARG;
    if (C > WM_MAXC) return PCNBR_E_TOOLARGE;
    if ((long long)(W - 1) * step >= n_points || (long long)(W - 1) * step + window < n_points) return PCNBR_E_BADARG;
    cudaStream_t s = (cudaStream_t)stream;
    const long long gx = (n_points + 255) / 256;
    if (gx > 0x7fffffffLL) return PCNBR_E_TOOLARGE;
    const double cover = (double)window / step;              // average windows per point
    const double bytes = (double)n_points * C * 4.0 * (cover + (mean_logits ? 1.0 : 0.0)) + 12.0 * (double)n_points;
    if (C <= 16)
        PCNBR_TIMED("window_merge_kernel", s, bytes, (double)n_points * C * (cover + 4.0),
                    (window_merge_kernel<16><<<(unsigned)gx, 256, 0, s>>>(logits, win_off, W, n_points, window, step, C,
                                                                          mean_logits, pred, conf)));
    else
        PCNBR_TIMED("window_merge_kernel", s, bytes, (double)n_points * C * (cover + 4.0),
                    (window_merge_kernel<WM_MAXC><<<(unsigned)gx, 256, 0, s>>>(logits, win_off, W, n_points, window, step, C,
                                                                               mean_logits, pred, conf)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
