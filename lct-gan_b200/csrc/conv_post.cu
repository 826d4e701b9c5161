// conv_post of the waveform discriminators: C -> 1 channel, kernel K (3), stride 1, "same" padding
// (reference models/discriminators.py:59-66 and :188-196; called at :100 and :222).
//
// With a single output channel the generic [positions x out-channels] register tile has nothing to
// tile over and the grid collapses to a handful of CTAs, so this layer gets its own kernels: it is
// a pure HBM-bound channel reduction (AI ~ 1.5 flop/B).  Forward gives every CTA 16 positions and all
// input channels (partial sums over 16 channel groups meet in shared memory); wgrad gives every warp one
// input channel and reduces over positions with shuffles; dgrad is an elementwise outer product
// with the fused (feature-matching gradient + LeakyReLU') epilogue of the generic dgrad kernel.
#include "common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kMaxK = 8;
// forward: FP positions per CTA (8 / 4 / 2, chosen so that even the short maps give every SM a CTA), kT / FP channel
// groups per CTA (channel c goes to group c % FG)

// y[b, j] = bias + sum_{ci} sum_k w[ci][k] * x[b, ci, l + k - pad, p].
// One CTA owns FP consecutive positions of one batch row and ALL input channels: thread (g, pos) walks the channels
// c = g, g + FG, ... with 8 channels x K taps of independent loads in flight, the FG partial sums meet in shared
// memory, and y is written exactly once (the first version split the channels over CTAs and finished with 64 atomics
// per output element: 49 us for a 26 MB read; this form needs neither the atomics nor a zero-filled output).
// (KT = compile-time tap count, 3 for conv_post; 0 = run-time K up to kMaxK with predicated taps)
template <int kFP, int KT>
__global__ void __launch_bounds__(kT) post_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int C,
                                                      int L, int P, int K, int pad) {
    constexpr int kFG = kT / kFP;
    constexpr int KK = KT > 0 ? KT : kMaxK;
    extern __shared__ float ws[];            // [C * K] weights, then [kFG][kFP] partial sums
    float* red = ws + C * K;
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < C * K; i += kT) ws[i] = w[i];
    __syncthreads();
    const int jtot = L * P;
    const int pos = threadIdx.x % kFP, g = threadIdx.x / kFP;
    const int j = blockIdx.x * kFP + pos;
    float acc = 0.f;
    if (j < jtot) {
        const int l = j / P;
        bool ok[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) ok[k] = k < K && (l + k - pad) >= 0 && (l + k - pad) < L;
        const float* xb = x + (size_t)b * C * jtot + j;
#pragma unroll 16
        for (int c = g; c < C; c += kFG) {
            const float* xc = xb + (size_t)c * jtot;
#pragma unroll
            for (int k = 0; k < KK; ++k)
                if (ok[k]) acc = fmaf(ws[c * K + k], __ldg(xc + (k - pad) * P), acc);
        }
    }
    red[g * kFP + pos] = acc;
    __syncthreads();
    if (threadIdx.x < kFP && blockIdx.x * kFP + threadIdx.x < jtot) {
        float s = bias ? bias[0] : 0.f;
#pragma unroll
        for (int i = 0; i < kFG; ++i) s += red[i * kFP + threadIdx.x];
        y[(size_t)b * jtot + blockIdx.x * kFP + threadIdx.x] = s;
    }
}

// dw[ci][k] += sum_{b, j} dy[b, j] * x[b, ci, l + k - pad, p];  db += sum dy    (one warp per channel)
template <int KT>
__global__ void __launch_bounds__(kT) post_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        float* __restrict__ dw, float* __restrict__ db, int C, int L,
                                                        int P, int K, int pad) {
    constexpr int KK = KT > 0 ? KT : kMaxK;
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
    float acc[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[k] = 0.f;
    for (int j = lane; j < jtot; j += 32) {
        const float g = dyb[j];
        const int l = j / P;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            const int li = l + k - pad;
            if (k < K && li >= 0 && li < L) acc[k] = fmaf(g, xc[j + (k - pad) * P], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < KK; ++k) {
        if (k < K) {
            float s = warp_sum(acc[k]);
            if (lane == 0) atomicAdd(&dw[(size_t)ci * K + k], s);
        }
    }
}

// dx[b, ci, l, p] = (sum_k dy[b, l + pad - k, p] * w[ci][k] + gextra) * act'(xact)
// thread = position; the K taps of dy stay in registers while the CTA walks kDC input channels
constexpr int kDT = 128, kDC = 32;
template <int KT>
__global__ void __launch_bounds__(kDT) post_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                         float* __restrict__ dx, const float* __restrict__ gextra,
                                                         const float* __restrict__ xact, int C, int L, int P, int K,
                                                         int pad, int act, float slope) {
    constexpr int KK = KT > 0 ? KT : kMaxK;
    __shared__ float ws[kDC * kMaxK];
    const int b = blockIdx.z, c0 = blockIdx.y * kDC;
    const int cc = min(kDC, C - c0);
    for (int i = threadIdx.x; i < cc * K; i += kDT) ws[i] = w[(size_t)c0 * K + i];
    __syncthreads();
    const int jtot = L * P;
    const int j = blockIdx.x * kDT + threadIdx.x;
    if (j >= jtot) return;
    const int l = j / P;
    const float* dyb = dy + (size_t)b * jtot;
    float g[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
        const int lo = l + pad - k;
        g[k] = (k < K && lo >= 0 && lo < L) ? dyb[j + (pad - k) * P] : 0.f;
    }
    size_t idx = ((size_t)b * C + c0) * jtot + j;
#pragma unroll 4
    for (int c = 0; c < cc; ++c, idx += jtot) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < KK; ++k)
            if (k < K) acc = fmaf(g[k], ws[c * K + k], acc);
        if (gextra) acc += gextra[idx];
        if (xact) acc *= act_grad_from_out(xact[idx], act, slope);
        dx[idx] = acc;
    }
}

bool ok_shape(int64_t B, int64_t C, int64_t L, int64_t P, int64_t K) {
    return B > 0 && B < 65536 && C > 0 && C < 65536 * 8 && L > 0 && P > 0 && K > 0 && K <= kMaxK && (K & 1) &&
           L * P < (1LL << 30);
}

}  // namespace

// y [B,1,L,P] = conv(x [B,C,L,P], w [1,C,K]) + bias, stride 1, pad K/2   (y is overwritten: no zero fill needed)
LCT_API int lct_conv_post_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t C,
                              int64_t L, int64_t P, int64_t K, cudaStream_t st) {
    if (!x || !w || !y || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    const size_t smem = ((size_t)C * K + kT) * sizeof(float);
    if (smem > 40 * 1024) return LCT_EUNSUPPORTED;
    const int64_t jtot = L * P;
#define LCT_POST_FWD(FP)                                                                                        \
    do {                                                                                                        \
        dim3 grid((unsigned)ceil_div64(jtot, FP), (unsigned)B);                                                 \
        if (K == 3) post_fwd_kernel<FP, 3><<<grid, kT, smem, st>>>(x, w, bias, y, (int)C, (int)L, (int)P, 3, 1); \
        else post_fwd_kernel<FP, 0><<<grid, kT, smem, st>>>(x, w, bias, y, (int)C, (int)L, (int)P, (int)K, (int)(K / 2)); \
    } while (0)
    // 8 positions = one 32-byte sector per (channel, group); fewer positions per CTA on the short maps so that every SM
    // still gets a few CTAs (the kernel is bound by load latency: per-thread chains must be short and numerous)
    if (ceil_div64(jtot, 8) * B >= 296) LCT_POST_FWD(8);
    else if (ceil_div64(jtot, 4) * B >= 296) LCT_POST_FWD(4);
    else LCT_POST_FWD(2);
#undef LCT_POST_FWD
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dw [1,C,K] and db [1] accumulated
LCT_API int lct_conv_post_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t C, int64_t L,
                                int64_t P, int64_t K, cudaStream_t st) {
    if (!x || !dy || !dw || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(C, kT / 32), (unsigned)B);
    if (K == 3) post_wgrad_kernel<3><<<grid, kT, 0, st>>>(x, dy, dw, db, (int)C, (int)L, (int)P, 3, 1);
    else post_wgrad_kernel<0><<<grid, kT, 0, st>>>(x, dy, dw, db, (int)C, (int)L, (int)P, (int)K, (int)(K / 2));
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_conv_post_dgrad(const float* dy, const float* w, float* dx, const float* gextra, const float* xact,
                                int64_t B, int64_t C, int64_t L, int64_t P, int64_t K, int act, float slope,
                                cudaStream_t st) {
    if (!dy || !w || !dx || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(L * P, kDT), (unsigned)ceil_div64(C, kDC), (unsigned)B);
    if (K == 3) post_dgrad_kernel<3><<<grid, kDT, 0, st>>>(dy, w, dx, gextra, xact, (int)C, (int)L, (int)P, 3, 1, act, slope);
    else post_dgrad_kernel<0><<<grid, kDT, 0, st>>>(dy, w, dx, gextra, xact, (int)C, (int)L, (int)P, (int)K, (int)(K / 2), act,
                                                    slope);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
