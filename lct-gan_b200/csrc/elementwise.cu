// Elementwise spectral helpers and the small integer-indexed waveform ops.
//
// Replaces (reference file:line):
//   magnitude            datasets/stft.py:138-160
//   compress/decompress  datasets/stft.py:163-178, :221-240
//   compute_compressed_irm datasets/stft.py:184-218
//   apply_mask           datasets/stft.py:243-290
//   right reflect pad of PeriodDiscriminator   models/discriminators.py:84-88  (integer indexing, bit exact)
//   AvgPool1d(4,2,2,count_include_pad=False)   models/discriminators.py:252-255, :284
// All are pure streaming kernels (one read + one write per element).
#include "common.cuh"

namespace {

constexpr int kT = 256;
inline unsigned nblk(int64_t n) { return (unsigned)ceil_div64(n, kT); }

__global__ void magnitude_fwd_k(const float2* __restrict__ s, float* __restrict__ m, int64_t n, float power,
                                float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float2 v = s[i];
    float r = fmaxf(hypotf(v.x, v.y), eps);
    m[i] = (power == 1.f) ? r : powf(r, power);
}

__global__ void magnitude_bwd_k(const float2* __restrict__ s, const float* __restrict__ gm,
                                float2* __restrict__ gs, int64_t n, float power, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float2 v = s[i];
    float r = hypotf(v.x, v.y);
    float g = gm[i];
    float rc = fmaxf(r, eps);
    if (power != 1.f) g *= power * powf(rc, power - 1.f);
    // clamp_min passes the gradient where r >= eps; d|z| = z/|z| (0 at z = 0)
    float k = (r >= eps && r > 0.f) ? g / r : 0.f;
    gs[i] = make_float2(k * v.x, k * v.y);
}

__global__ void powclamp_fwd_k(const float* __restrict__ x, float* __restrict__ y, int64_t n, float e, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i < n) y[i] = powf(fmaxf(x[i], eps), e);
}

__global__ void powclamp_bwd_k(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                               int64_t n, float e, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float v = x[i];
    gx[i] = (v >= eps) ? gy[i] * e * powf(v, e - 1.f) : 0.f;
}

__global__ void irm_fwd_k(const float2* __restrict__ cs, const float2* __restrict__ ns, float* __restrict__ o,
                          int64_t n, float c, float gamma, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float2 a = cs[i], b = ns[i];
    float cm = powf(fmaxf(hypotf(a.x, a.y), eps), c);
    float nm = powf(fmaxf(hypotf(b.x, b.y), eps), c);
    o[i] = cm / (nm + gamma);
}

__global__ void apply_mask_fwd_k(const float2* __restrict__ s, const float* __restrict__ m, float2* __restrict__ o,
                                 int64_t n, int compressed, float c, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float mk = m[i];
    if (compressed) mk = powf(fmaxf(mk, eps), 1.f / c);
    mk = fmaxf(mk, 0.f);
    float2 v = s[i];
    o[i] = make_float2(v.x * mk, v.y * mk);
}

__global__ void apply_mask_bwd_k(const float2* __restrict__ s, const float* __restrict__ m,
                                 const float2* __restrict__ go, float* __restrict__ gm, float2* __restrict__ gs,
                                 int64_t n, int compressed, float c, float eps) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    float mk = m[i];
    float lin = mk, dlin = 1.f;
    if (compressed) {
        float mc = fmaxf(mk, eps);
        float ic = 1.f / c;
        lin = powf(mc, ic);
        dlin = (mk >= eps) ? ic * powf(mc, ic - 1.f) : 0.f;
    }
    if (lin < 0.f) { lin = 0.f; dlin = 0.f; }
    float2 g = go[i];
    if (gm) {
        float2 v = s[i];
        gm[i] = (v.x * g.x + v.y * g.y) * dlin;
    }
    if (gs) gs[i] = make_float2(g.x * lin, g.y * lin);
}

// y[b, j] = x[b, j] for j < T, x[b, 2(T-1)-j] for T <= j < T+pad   (F.pad(..., mode="reflect") on the right)
__global__ void reflect_pad_right_fwd_k(const float* __restrict__ x, float* __restrict__ y, int T, int pad) {
    int j = blockIdx.x * kT + threadIdx.x;
    int b = blockIdx.y;
    if (j >= T + pad) return;
    int src = j < T ? j : 2 * (T - 1) - j;
    y[(size_t)b * (T + pad) + j] = x[(size_t)b * T + src];
}

__global__ void reflect_pad_right_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int T, int pad) {
    int t = blockIdx.x * kT + threadIdx.x;
    int b = blockIdx.y;
    if (t >= T) return;
    const float* g = gy + (size_t)b * (T + pad);
    float s = g[t];
    int j = 2 * (T - 1) - t;   // padded position that mirrors t
    if (j >= T && j < T + pad) s += g[j];
    gx[(size_t)b * T + t] = s;
}

// AvgPool1d(kernel 4, stride 2, padding 2, count_include_pad=False): Lout = L/2 + 1
__global__ void avgpool4_fwd_k(const float* __restrict__ x, float* __restrict__ y, int L, int Lout) {
    int o = blockIdx.x * kT + threadIdx.x;
    int b = blockIdx.y;
    if (o >= Lout) return;
    int lo = max(2 * o - 2, 0), hi = min(2 * o + 2, L);   // [lo, hi)
    const float* xr = x + (size_t)b * L;
    float s = 0.f;
    for (int i = lo; i < hi; ++i) s += xr[i];
    y[(size_t)b * Lout + o] = s / (float)(hi - lo);
}

__global__ void avgpool4_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int L, int Lout) {
    int i = blockIdx.x * kT + threadIdx.x;
    int b = blockIdx.y;
    if (i >= L) return;
    // windows o with 2o-2 <= i < 2o+2  ->  o in [ceil((i-1)/2), floor((i+2)/2)]
    int o_lo = max((i - 1 + 1) / 2, 0);   // ceil((i-1)/2) for i >= 1; 0 for i == 0
    if (i == 0) o_lo = 0;
    int o_hi = min((i + 2) / 2, Lout - 1);
    const float* g = gy + (size_t)b * Lout;
    float s = 0.f;
    for (int o = o_lo; o <= o_hi; ++o) {
        int lo = max(2 * o - 2, 0), hi = min(2 * o + 2, L);
        if (i >= lo && i < hi) s += g[o] / (float)(hi - lo);
    }
    gx[(size_t)b * L + i] = s;
}

__global__ void axpby_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n,
                        float ka, float kb) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i < n) o[i] = ka * a[i] + (b ? kb * b[i] : 0.f);
}
// 16-byte form (all pointers 16-byte aligned): thread t handles float4 t, t + stride, ...; the last n % 4 elements scalar.
// (o may alias a or b: every element is read before it is written by the same thread)
__global__ void axpby4_k(const float* a, const float* b, float* o, int64_t n, float ka, float kb) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kT;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n4; i += stride) {
        const float4 x = reinterpret_cast<const float4*>(a)[i];
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b) y = reinterpret_cast<const float4*>(b)[i];
        reinterpret_cast<float4*>(o)[i] = make_float4(ka * x.x + kb * y.x, ka * x.y + kb * y.y, ka * x.z + kb * y.z,
                                                      ka * x.w + kb * y.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        o[i] = ka * a[i] + (b ? kb * b[i] : 0.f);
    }
}

__global__ void add2d_k(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                        float* __restrict__ o, int ldo, int64_t M, int N) {
    int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
    if (i >= M * N) return;
    int64_t m = i / N;
    int n = (int)(i - m * N);
    o[m * ldo + n] = a[m * lda + n] + b[m * ldb + n];
}

}  // namespace

// out[m, n] = a[m, n] + b[m, n] on row-strided [M, N] views
LCT_API int lct_add2d(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo, int64_t M,
                      int64_t N, cudaStream_t st) {
    if (!a || !b || !out || M <= 0 || N <= 0) return LCT_EINVAL;
    add2d_k<<<nblk(M * N), kT, 0, st>>>(a, (int)lda, b, (int)ldb, out, (int)ldo, M, (int)N);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_magnitude_fwd(const float* spec, float* mag, int64_t n, float power, float eps, cudaStream_t st) {
    if (!spec || !mag || n <= 0) return LCT_EINVAL;
    magnitude_fwd_k<<<nblk(n), kT, 0, st>>>(reinterpret_cast<const float2*>(spec), mag, n, power, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_magnitude_bwd(const float* spec, const float* gmag, float* gspec, int64_t n, float power, float eps,
                              cudaStream_t st) {
    if (!spec || !gmag || !gspec || n <= 0) return LCT_EINVAL;
    magnitude_bwd_k<<<nblk(n), kT, 0, st>>>(reinterpret_cast<const float2*>(spec), gmag,
                                            reinterpret_cast<float2*>(gspec), n, power, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// y = max(x, eps)^e  (compress: e = c; decompress: e = 1/c)
LCT_API int lct_powclamp_fwd(const float* x, float* y, int64_t n, float e, float eps, cudaStream_t st) {
    if (!x || !y || n <= 0) return LCT_EINVAL;
    powclamp_fwd_k<<<nblk(n), kT, 0, st>>>(x, y, n, e, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_powclamp_bwd(const float* x, const float* gy, float* gx, int64_t n, float e, float eps,
                             cudaStream_t st) {
    if (!x || !gy || !gx || n <= 0) return LCT_EINVAL;
    powclamp_bwd_k<<<nblk(n), kT, 0, st>>>(x, gy, gx, n, e, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_irm_fwd(const float* clean_spec, const float* noisy_spec, float* irm_c, int64_t n, float c,
                        float gamma, float eps, cudaStream_t st) {
    if (!clean_spec || !noisy_spec || !irm_c || n <= 0) return LCT_EINVAL;
    irm_fwd_k<<<nblk(n), kT, 0, st>>>(reinterpret_cast<const float2*>(clean_spec),
                                      reinterpret_cast<const float2*>(noisy_spec), irm_c, n, c, gamma, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_apply_mask_fwd(const float* spec, const float* mask, float* out, int64_t n, int compressed, float c,
                               float eps, cudaStream_t st) {
    if (!spec || !mask || !out || n <= 0) return LCT_EINVAL;
    apply_mask_fwd_k<<<nblk(n), kT, 0, st>>>(reinterpret_cast<const float2*>(spec), mask,
                                             reinterpret_cast<float2*>(out), n, compressed, c, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_apply_mask_bwd(const float* spec, const float* mask, const float* gout, float* gmask, float* gspec,
                               int64_t n, int compressed, float c, float eps, cudaStream_t st) {
    if (!spec || !mask || !gout || (!gmask && !gspec) || n <= 0) return LCT_EINVAL;
    apply_mask_bwd_k<<<nblk(n), kT, 0, st>>>(reinterpret_cast<const float2*>(spec), mask,
                                             reinterpret_cast<const float2*>(gout), gmask,
                                             reinterpret_cast<float2*>(gspec), n, compressed, c, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_reflect_pad_right_fwd(const float* x, float* y, int64_t B, int64_t T, int64_t pad, cudaStream_t st) {
    if (!x || !y || B <= 0 || B >= 65536 || T < 2 || pad < 0 || pad >= T) return LCT_EINVAL;
    dim3 grid(nblk(T + pad), (unsigned)B);
    reflect_pad_right_fwd_k<<<grid, kT, 0, st>>>(x, y, (int)T, (int)pad);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_reflect_pad_right_bwd(const float* gy, float* gx, int64_t B, int64_t T, int64_t pad, cudaStream_t st) {
    if (!gy || !gx || B <= 0 || B >= 65536 || T < 2 || pad < 0 || pad >= T) return LCT_EINVAL;
    dim3 grid(nblk(T), (unsigned)B);
    reflect_pad_right_bwd_k<<<grid, kT, 0, st>>>(gy, gx, (int)T, (int)pad);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_avgpool4_fwd(const float* x, float* y, int64_t B, int64_t L, cudaStream_t st) {
    if (!x || !y || B <= 0 || B >= 65536 || L <= 0) return LCT_EINVAL;
    int Lout = (int)(L / 2 + 1);
    dim3 grid(nblk(Lout), (unsigned)B);
    avgpool4_fwd_k<<<grid, kT, 0, st>>>(x, y, (int)L, Lout);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_avgpool4_bwd(const float* gy, float* gx, int64_t B, int64_t L, cudaStream_t st) {
    if (!gy || !gx || B <= 0 || B >= 65536 || L <= 0) return LCT_EINVAL;
    int Lout = (int)(L / 2 + 1);
    dim3 grid(nblk(L), (unsigned)B);
    avgpool4_bwd_k<<<grid, kT, 0, st>>>(gy, gx, (int)L, Lout);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// out = ka * a + kb * b   (b may be null)
LCT_API int lct_axpby(const float* a, const float* b, float* out, int64_t n, float ka, float kb, cudaStream_t st) {
    if (!a || !out || n <= 0) return LCT_EINVAL;
    if (n >= 4096 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0) {
        int64_t blocks = ((n >> 2) + kT - 1) / kT;
        if (blocks > 148 * 16) blocks = 148 * 16;
        axpby4_k<<<(unsigned)blocks, kT, 0, st>>>(a, b, out, n, ka, kb);
    } else {
        axpby_k<<<nblk(n), kT, 0, st>>>(a, b, out, n, ka, kb);
    }
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
