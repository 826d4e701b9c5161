// Callers on either side of the hot path:
//
//   lct_si_sdr         batched scale-invariant SDR on the device, one launch for a whole validation batch with per-row
//                      valid lengths (reference train.py:261-282 `_si_sdr_torch`, called once per utterance with a host
//                      sync each, train.py:313-334)
//   lct_crop_segments  the training loader's segment cropping + zero-padding collate on the GPU: B segments of T
//                      samples cut out of utterances that live in one packed device buffer (reference
//                      datasets/datasets.py:131-156 `_crop_pair`, :187-230 `collate_fn`) - integer indexing, bit exact;
//                      with start = 0 and T = the longest utterance it is collate_fn's zero padding of a ragged batch
#include "common.cuh"

namespace {

constexpr int kSdrThreads = 512;

// moments of the first L samples of a row in double: sum r, sum e, sum r^2, sum e^2, sum r e
__global__ void __launch_bounds__(kSdrThreads) si_sdr_kernel(const float* __restrict__ ref, const float* __restrict__ est,
                                                             const int64_t* __restrict__ lengths, float* __restrict__ out,
                                                             int64_t T_ref, int64_t T_est, float eps) {
    __shared__ double red[5][kSdrThreads / 32];
    const int b = blockIdx.x;
    int64_t L = T_ref < T_est ? T_ref : T_est;                 // "align length" (train.py:267-269)
    if (lengths) L = lengths[b] < L ? lengths[b] : L;
    if (L < 0) L = 0;
    const float* r = ref + (int64_t)b * T_ref;
    const float* e = est + (int64_t)b * T_est;
    double s[5] = {0, 0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < L; i += kSdrThreads) {
        const double rv = r[i], ev = e[i];
        s[0] += rv; s[1] += ev; s[2] += rv * rv; s[3] += ev * ev; s[4] += rv * ev;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (lane == 0) red[k][w] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < 5; ++k)
            for (int i = 0; i < kSdrThreads / 32; ++i) t[k] += red[k][i];
        const double n = L > 0 ? (double)L : 1.0;
        // zero-mean signals (train.py:272-273): central second moments
        const double srr = t[2] - t[0] * t[0] / n, see = t[3] - t[1] * t[1] / n, sre = t[4] - t[0] * t[1] / n;
        const double scale = sre / (srr + (double)eps);                      // train.py:275-276
        const double target = scale * scale * srr;                           // sum s_target^2
        double noise = see - 2.0 * scale * sre + target;                     // sum (e - scale r)^2
        if (noise < 0.0) noise = 0.0;
        out[b] = (float)(10.0 * log10((target + (double)eps) / (noise + (double)eps)));    // train.py:280-281
    }
}

// out_*[b, t] = src_*[offset[b] + start[b] + t] for t < min(T, length[b] - start[b]) else 0
__global__ void crop_segments_kernel(const float* __restrict__ noisy, const float* __restrict__ clean,
                                     const int64_t* __restrict__ offset_n, const int64_t* __restrict__ offset_c,
                                     const int64_t* __restrict__ len_n, const int64_t* __restrict__ len_c,
                                     const int64_t* __restrict__ start, float* __restrict__ out_n,
                                     float* __restrict__ out_c, int64_t T) {
    const int b = blockIdx.y;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t s = start[b];
    const int64_t i = s + t;
    out_n[(int64_t)b * T + t] = i < len_n[b] ? noisy[offset_n[b] + i] : 0.f;
    if (out_c) out_c[(int64_t)b * T + t] = i < len_c[b] ? clean[offset_c[b] + i] : 0.f;
}

}  // namespace

// out[b] = SI-SDR in dB of est[b, :L_b] against ref[b, :L_b], L_b = min(lengths[b], T_ref, T_est) (lengths optional).
LCT_API int lct_si_sdr(const float* ref, const float* est, const int64_t* lengths, float* out, int64_t B, int64_t T_ref,
                       int64_t T_est, float eps, cudaStream_t st) {
    if (!ref || !est || !out || B <= 0 || B >= (1LL << 31) || T_ref < 0 || T_est < 0) return LCT_EINVAL;
    si_sdr_kernel<<<(unsigned)B, kSdrThreads, 0, st>>>(ref, est, lengths, out, T_ref, T_est, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// B segments of T samples cut from packed utterance buffers (device arrays offset_*, len_*, start of B int64 each);
// clean / out_c optional (inference: noisy only).  Samples past the end of an utterance are zero (collate_fn padding).
LCT_API int lct_crop_segments(const float* noisy, const float* clean, const int64_t* offset_n, const int64_t* offset_c,
                              const int64_t* len_n, const int64_t* len_c, const int64_t* start, float* out_n, float* out_c,
                              int64_t B, int64_t T, cudaStream_t st) {
    if (!noisy || !offset_n || !len_n || !start || !out_n || B <= 0 || B >= 65536 || T <= 0) return LCT_EINVAL;
    if (out_c && (!clean || !offset_c || !len_c)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(T, 256), (unsigned)B);
    crop_segments_kernel<<<grid, 256, 0, st>>>(noisy, clean, offset_n, offset_c, len_n, len_c, start, out_n, out_c, T);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
