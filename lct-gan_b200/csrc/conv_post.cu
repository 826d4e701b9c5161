// conv_post of the waveform discriminators: C -> 1 channel, kernel K (3), stride 1, "same" padding
// (reference models/discriminators.py:59-66 and :188-196; called at :100 and :222).
//
// With a single output channel the generic [positions x out-channels] register tile has nothing to
// tile over and the grid collapses to a handful of CTAs, so this layer gets its own kernels: it is
// a pure HBM-bound channel reduction (AI ~ 1.5 flop/B).  Forward gives every CTA up to 512 positions and a
// slice of the input channels (a warp reads whole channel rows, lane = position); wgrad gives every warp one
// input channel and reduces over positions with shuffles; dgrad is an elementwise outer product
// with the fused (feature-matching gradient + LeakyReLU') epilogue of the generic dgrad kernel.
#include "common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kMaxK = 8;
constexpr int kPT = 512;            // forward: positions per CTA (16 per lane)

// y[b, j] = bias + sum_{ci} sum_k w[ci][k] * x[b, ci, l + k - pad, p],  j = l P + p.
// The maps are short and deep (125 .. 400 positions x 1024 channels at 2 s), and contiguous along j only.  A CTA owns
// up to 512 positions of one batch row and a SLICE of the channels; a warp walks its channels row by row, lane = position,
// so every load instruction reads 128 consecutive bytes of one channel row (tap k is the same row shifted by (k - pad) P
// floats: j' = j + (k - pad) P keeps p, and l' is inside [0, L) exactly when j' is inside [0, L P)).  The 8 warps' sums
// meet in shared memory and one atomicAdd per position and CTA adds the slice to y (zero-filled by the launcher).
// (Round 1 gave a CTA 8 positions and all channels: 32-byte pieces of 1024 different rows per CTA - 21 us for 13 MB.)
// (KT = compile-time tap count, 3 for conv_post; 0 = run-time K up to kMaxK with predicated taps)
__global__ void post_zero_kernel(float* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = 0.f;
}

template <int KT>
__global__ void __launch_bounds__(kT) post_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ y, int C,
                                                      int L, int P, int K, int pad, int cs) {
    constexpr int KK = KT > 0 ? KT : kMaxK;
    constexpr int NI = kPT / 32;
    __shared__ float red[kT / 32][kPT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z;
    const int jtot = L * P;
    const int j0 = blockIdx.x * kPT;
    const int c0 = blockIdx.y * cs, c1 = min(C, c0 + cs);
    float acc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[i] = 0.f;
    const float* xb = x + (size_t)b * C * jtot;
    for (int c = c0 + warp; c < c1; c += kT / 32) {
        const float* xc = xb + (size_t)c * jtot;
        float wk[KK];
#pragma unroll
        for (int k = 0; k < KK; ++k) wk[k] = k < K ? __ldg(w + (size_t)c * K + k) : 0.f;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int j = j0 + lane + 32 * i;
#pragma unroll
            for (int k = 0; k < KK; ++k) {
                const int jj = j + (k - pad) * P;
                if (k < K && j < jtot && (unsigned)jj < (unsigned)jtot) acc[i] = fmaf(wk[k], __ldg(xc + jj), acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) red[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    for (int t = threadIdx.x; t < kPT; t += kT) {
        if (j0 + t < jtot) {
            float s = (blockIdx.y == 0 && bias) ? bias[0] : 0.f;
#pragma unroll
            for (int q = 0; q < kT / 32; ++q) s += red[q][t];
            atomicAdd(&y[(size_t)b * jtot + j0 + t], s);
        }
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

// y [B,1,L,P] = conv(x [B,C,L,P], w [1,C,K]) + bias, stride 1, pad K/2   (y is overwritten: cleared here, then the
// channel slices are added with one atomic per position and slice)
LCT_API int lct_conv_post_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t C,
                              int64_t L, int64_t P, int64_t K, cudaStream_t st) {
    if (!x || !w || !y || !ok_shape(B, C, L, P, K)) return LCT_EINVAL;
    const int64_t jtot = L * P;
    const int64_t ntile = ceil_div64(jtot, kPT);
    // enough channel slices for ~2 CTAs per SM, at least one channel per warp
    int64_t slices = ceil_div64(296, B * ntile);
    if (slices > C / (kT / 32)) slices = C / (kT / 32);
    if (slices < 1) slices = 1;
    int64_t cs = ceil_div64(C, slices);
    cs = ceil_div64(cs, kT / 32) * (kT / 32);
    slices = ceil_div64(C, cs);
    if (slices > 65535 || B > 65535) return LCT_EUNSUPPORTED;
    post_zero_kernel<<<(unsigned)ceil_div64(B * jtot, 256), 256, 0, st>>>(y, B * jtot);
    LCT_RETURN_IF_LAUNCH_FAILED();
    dim3 grid((unsigned)ntile, (unsigned)slices, (unsigned)B);
    if (K == 3) post_fwd_kernel<3><<<grid, kT, 0, st>>>(x, w, bias, y, (int)C, (int)L, (int)P, 3, 1, (int)cs);
    else post_fwd_kernel<0><<<grid, kT, 0, st>>>(x, w, bias, y, (int)C, (int)L, (int)P, (int)K, (int)(K / 2), (int)cs);
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
