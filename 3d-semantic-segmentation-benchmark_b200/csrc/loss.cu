// loss.cu -- SURVEY.md 8f-1: the training loss of Training/train_model.py:15-57 (masked one-hot cross entropy) as two
// kernels, forward and gradient together, no host synchronisation.
//
// Reference: log_softmax over the classes, -sum(onehot * logp) per point, a position mask built from pad_starts, the
// masked mean -- about ten library launches forward, as many backward, and a `.item()` sync on the number of unpadded
// points (train_model.py:53).  Here one thread per point computes the log-sum-exp, its loss term and -- because the
// number of unpadded points depends on the lengths only -- the FINAL gradient (softmax * sum(y) - y) / count in the same
// pass; the per-block loss partials are summed in block order by a second, single-CTA kernel (deterministic).
#include "common.cuh"

namespace pcnbr {

constexpr int LS_MAXC = 64;

__global__ void __launch_bounds__(256)
masked_ce_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ onehot, const long long* __restrict__ lengths,
                 int B, int L, int C, float* __restrict__ partial, float* __restrict__ dlogits) {
    __shared__ float red[8];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long count = 0;                                   // unpadded points of the batch (train_model.py:50)
    for (int i = 0; i < B; ++i) {
        long long l = lengths ? lengths[i] : (long long)L;
        count += l < 0 ? 0 : (l > L ? L : l);
    }
    const float inv = count > 0 ? 1.0f / (float)count : 0.f;
    long long len = lengths ? lengths[b] : (long long)L;
    if (len > L) len = L;
    float acc = 0.f;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < L; n += gridDim.x * blockDim.x) {
        const float* __restrict__ x = logits + ((size_t)b * L + n) * C;
        const uint8_t* __restrict__ y = onehot + ((size_t)b * L + n) * C;
        float* __restrict__ g = dlogits ? dlogits + ((size_t)b * L + n) * C : nullptr;
        if (n >= len) {                                    // padding: no loss, no gradient
            if (g) for (int c = 0; c < C; ++c) g[c] = 0.f;
            continue;
        }
        float m = x[0];
        for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
        float se = 0.f, ysum = 0.f, yx = 0.f;
        for (int c = 0; c < C; ++c) {
            se += expf(x[c] - m);
            const float yc = (float)y[c];
            ysum += yc;
            yx += yc * x[c];
        }
        const float lse = m + logf(se);
        acc += ysum * lse - yx;                            // -sum_c y_c (x_c - lse)
        if (g) {
            const float rs = 1.0f / se;
            for (int c = 0; c < C; ++c) g[c] = (expf(x[c] - m) * rs * ysum - (float)y[c]) * inv;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(PCNBR_FULL, acc, d);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int k = 1; k < 8; ++k) t += red[k];
        partial[(size_t)b * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(32)
masked_ce_finalize_kernel(const float* __restrict__ partial, int n, const long long* __restrict__ lengths, int B, int L,
                          float* __restrict__ loss) {
    if (threadIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)partial[i];
    long long count = 0;
    for (int i = 0; i < B; ++i) {
        long long l = lengths ? lengths[i] : (long long)L;
        count += l < 0 ? 0 : (l > L ? L : l);
    }
    loss[0] = count > 0 ? (float)(s / (double)count) : 0.f;
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" int pcnbr_masked_ce_blocks(int L) {
    int gx = (L + 255) / 256;
    return gx > 32 ? 32 : gx;
}

extern "C" int pcnbr_masked_ce_f32(const float* logits, const uint8_t* onehot, const long long* lengths, int B, int L, int C,
                                   float* loss, float* dlogits, float* partial, pcnbr_stream_t stream) {
    if (!logits || !onehot || !loss || !partial || B <= 0 || L <= 0 || C <= 0) return PCNBR_E_BADARG;
    if (C > LS_MAXC || B > 4096) return PCNBR_E_TOOLARGE;
    cudaStream_t s = (cudaStream_t)stream;
    const int gx = pcnbr_masked_ce_blocks(L);
    PCNBR_TIMED("masked_ce_kernel", s, (double)B * L * C * (dlogits ? 9.0 : 5.0), 8.0 * B * (double)L * C,
                (masked_ce_kernel<<<dim3(gx, B), 256, 0, s>>>(logits, onehot, lengths, B, L, C, partial, dlogits)));
    PCNBR_CHECK_LAUNCH();
    PCNBR_TIMED("masked_ce_finalize_kernel", s, 4.0 * gx * B, (double)gx * B,
                (masked_ce_finalize_kernel<<<1, 32, 0, s>>>(partial, gx * B, lengths, B, L, loss)));
    PCNBR_CHECK_LAUNCH();
    return 0;
}
