// conv_post of the waveform discriminators: C -> 1 channel, kernel K (3), stride 1, "same" padding
// (reference models/discriminators.py:59-66 and :188-196; called at :100 and :222).
//
// With a single output channel the generic [positions x out-channels] register tile has nothing to
// tile over and the grid collapses to a handful of CTAs, so this layer gets its own kernels: it is
// a pure HBM-bound channel reduction (AI ~ 1.5 flop/B).  Forward splits the 1024 input channels
// over CTAs and finishes with one atomic per (position, channel chunk); wgrad gives every warp one
// input channel and reduces over positions with shuffles; dgrad is an elementwise outer product
// with the fused (feature-matching gradient + LeakyReLU') epilogue of the generic dgrad kernel.
#include "common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kCC = 16;     // input channels per CTA (forward): 64 chunks of the 1024 channels keep 1000+ CTAs in flight
constexpr int kMaxK = 8;

// y[b, j] += sum_{ci in chunk} sum_k w[ci][k] * x[b, ci, l + k - pad, p]   (+ bias once); y pre-zeroed
__global__ void __launch_bounds__(kT) post_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int C,
                                                      int L, int P, int K, int pad) {
    __shared__ float ws[kCC * kMaxK];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kCC;
    const int cc = min(kCC, C - c0);
    for (int i = threadIdx.x; i < cc * K; i += kT) ws[i] = w[(size_t)c0 * K + i];
    __syncthreads();
    const int jtot = L * P;
    const int j = blockIdx.x * kT + threadIdx.x;
    if (j >= jtot) return;
    const int l = j / P;
    const float* xb = x + ((size_t)b * C + c0) * jtot + j;
    float acc = (blockIdx.y == 0 && bias) ? bias[0] : 0.f;
    bool ok[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) ok[k] = k < K && (l + k - pad) >= 0 && (l + k - pad) < L;
#pragma unroll 4
    for (int c = 0; c < cc; ++c) {
        const float* xc = xb + (size_t)c * jtot;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k)
            if (ok[k]) acc = fmaf(ws[c * K + k], __ldg(xc + (k - pad) * P), acc);
    }
    atomicAdd(&y[(size_t)b * jtot + j], acc);
}

// dw[ci][k] += sum_{b, j} dy[b, j] * x[b, ci, l + k - pad, p];  db += sum dy    (one warp per channel)
__global__ void __launch_bounds__(kT) post_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dw, float* __restrict__ db, int C, int L,
                                                        int P, int K, int pad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int ci = blockIdx.x * (kT / 32) + warp;
    const int jtot = L * P;
    const float* dyb = dy + (size_t)b * jtot;
    if (blockIdx.x == 0 && warp == 0 && db) {
        float s = 0.f;
        for (int j = lane; j < jtot; j += 32) s += dyb[j];
        s = warp_sum(s);
        if (lane == 0) atomicAdd(db, s);
    }
    if (ci >= C) return;
    const float* xc = x + ((size_t)b * C + ci) * jtot;
    float acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = 0.f;
    for (int j = lane; j < jtot; j += 32) {
        const float g = dyb[j];
        const int l = j / P;
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) {
            const int li = l + k - pad;
            if (k < K && li >= 0 && li < L) acc[k] = fmaf(g, xc[j + (k - pad) * P], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
        if (k < K) {
            float s = warp_sum(acc[k]);
            if (lane == 0) atomicAdd(&dw[(size_t)ci * K + k], s);
        }
    }
}

// dx[b, ci, l, p] = (sum_k dy[b, l + pad - k, p] * w[ci][k] + gextra) * act'(xact)
__global__ void __launch_bounds__(kT) post_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                        float* __restrict__ dx, const float* __restrict__ gextra,
                                                        const float* __restrict__ xact, int C, int L, int P, int K,
                                                        int pad, int act, float slope) {
    const int b = blockIdx.z, ci = blockIdx.y;
    const int jtot = L * P;
    const int j = blockIdx.x * kT + threadIdx.x;
    if (j >= jtot) return;
    const int l = j / P;
    const float* dyb = dy + (size_t)b * jtot;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
        const int lo = l + pad - k;
        if (k < K && lo >= 0 && lo < L) acc = fmaf(dyb[j + (pad - k) * P], __ldg(&w[(size_t)ci * K + k]), acc);
    }
    const size_t idx = ((size_t)b * C + ci) * jtot + j;
    if (gextra) acc += gextra[idx];
    if (xact) acc *= act_grad_from_out(xact[idx], act, slope);
    dx[idx] = acc;
}

bool ok_shape(int64_t B, int64_t C, int64_t L, int64_t P, int64_t K) {
    return B > 0 && B < 65536 && C > 0 && C < 65536 * 8 && L > 0 && P > 0 && K > 0 && K <= kMaxK && (K & 1) &&
           L * P < (1LL << 30);
}

}  // namespace

// y [B,1,L,P] (pre-zeroed) += conv(x [B,C,L,P], w [1,C,K]) + bias, stride 1, pad K/2
LCT_API int lct_conv_post_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t C,
                              int64_t L, int64_t P, int64_t K, cudaStream_t st) {
    if (!x || !w || !y || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(L * P, kT), (unsigned)ceil_div64(C, kCC), (unsigned)B);
    post_fwd_kernel<<<grid, kT, 0, st>>>(x, w, bias, y, (int)C, (int)L, (int)P, (int)K, (int)(K / 2));
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dw [1,C,K] and db [1] accumulated
LCT_API int lct_conv_post_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t C, int64_t L,
                                int64_t P, int64_t K, cudaStream_t st) {
    if (!x || !dy || !dw || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(C, kT / 32), (unsigned)B);
    post_wgrad_kernel<<<grid, kT, 0, st>>>(x, dy, dw, db, (int)C, (int)L, (int)P, (int)K, (int)(K / 2));
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_conv_post_dgrad(const float* dy, const float* w, float* dx, const float* gextra, const float* xact,
                                int64_t B, int64_t C, int64_t L, int64_t P, int64_t K, int act, float slope,
                                cudaStream_t st) {
    if (!dy || !w || !dx || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(L * P, kT), (unsigned)C, (unsigned)B);
    post_dgrad_kernel<<<grid, kT, 0, st>>>(dy, w, dx, gextra, xact, (int)C, (int)L, (int)P, (int)K, (int)(K / 2), act,
                                           slope);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
