// metrics.cu -- SURVEY.md 8f-1: the evaluation metrics of Training/metrics.py as ONE kernel and no host round trips.
//
// Reference: Training/metrics.py:3-146 -- overall_accuracy / update_accuracy / confusion_matrix /
// intersection_over_union / update_intersection_over_union each loop over the batch (and over C or C^2 class pairs) in
// Python with an .item() sync per iteration (B*C^2 = 5408 syncs per batch for the confusion matrix at B=32, C=13).
// All five are functions of the confusion matrix M[label, predicted] over the unpadded points:
//     correct = trace M,  intersection_c = M[c,c],  union_c = row_c + col_c - M[c,c].
// One thread per point: argmax of the C scores (first maximum wins, as torch.argmax), argmax of the one-hot label row,
// a shared-memory C x C tile of int counters per CTA, flushed with 64-bit integer atomics (integer sums are
// order-independent: the result is deterministic).
#include "common.cuh"

namespace pcnbr {

constexpr int MT_MAXC = 64;

__global__ void __launch_bounds__(256)
confusion_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ onehot, const long long* __restrict__ lengths,
                 int N, int C, unsigned long long* __restrict__ matrix, unsigned long long* __restrict__ unlabeled) {
    extern __shared__ int mt_tile[];                     // C * C (+ C: predictions of the rows without any label)
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < C * C + C; i += blockDim.x) mt_tile[i] = 0;
    __syncthreads();
    long long len = lengths ? lengths[b] : (long long)N;
    if (len > N) len = N;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < len; n += (long long)gridDim.x * blockDim.x) {
        const float* __restrict__ p = pred + ((size_t)b * N + n) * C;
        const uint8_t* __restrict__ l = onehot + ((size_t)b * N + n) * C;
        int pc = 0, lc = 0;
        float best = p[0];
        uint8_t lb = l[0];
        for (int c = 1; c < C; ++c) {
            const float v = p[c];
            if (v > best) { best = v; pc = c; }           // strict: first maximum wins (torch.argmax)
            const uint8_t w = l[c];
            if (w > lb) { lb = w; lc = c; }
        }
        atomicAdd(&mt_tile[lc * C + pc], 1);              // an all-zero label row counts as class 0: labels.argmax(-1), metrics.py:20,72
        if (unlabeled && lb == 0) atomicAdd(&mt_tile[C * C + pc], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (mt_tile[i]) atomicAdd(&matrix[i], (unsigned long long)mt_tile[i]);
    if (unlabeled)
        for (int i = threadIdx.x; i < C; i += blockDim.x)
            if (mt_tile[C * C + i]) atomicAdd(&unlabeled[i], (unsigned long long)mt_tile[C * C + i]);
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_confusion_f32(const float* pred, const uint8_t* onehot, const long long* lengths, int B, int N, int C,
                                   long long* matrix, pcnbr_stream_t stream) {
    return pcnbr_confusion_ex_f32(pred, onehot, lengths, B, N, C, matrix, nullptr, stream);
}

extern "C" int pcnbr_confusion_ex_f32(const float* pred, const uint8_t* onehot, const long long* lengths, int B, int N, int C,
                                      long long* matrix, long long* unlabeled, pcnbr_stream_t stream) {
    if (!pred || !onehot || !matrix || B <= 0 || N <= 0 || C <= 0) return PCNBR_E_BADARG;
    if (C > MT_MAXC) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    int gx = (N + 255) / 256;
    if (gx > 148 * 4) gx = 148 * 4;
    PCNBR_TIMED("confusion_kernel", s, (double)B * N * (5.0 * C) + 8.0 * C * C, 2.0 * B * (double)N * C,
                (confusion_kernel<<<dim3(gx, B), 256, (size_t)(C * C + C) * sizeof(int), s>>>(pred, onehot, lengths, N, C,
                                                                                     (unsigned long long*)matrix, (unsigned long long*)unlabeled)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
