// Generator encoder / decoder convolutions, channels-last [B, T, F, C] fp32.
//
// Replaces (reference file:line):
//   conv1..3   nn.Conv2d(k=(2,3), s=(1,2), p=(1,1)) + LeakyReLU(0.2)     models/generator.py:461-481, :570-572
//   deconv2..4 nn.ConvTranspose2d(k=(2,3), s=(1,2), p=(1,1), op=(0,1))   models/generator.py:506-529, :587-599
//   skip2..4   1x1 Conv2d(1 -> C) on the raw magnitude + cropped add      models/generator.py:484-498, :565-567, :587-598
//   final crop / zero-pad / sigmoid                                       models/generator.py:601-630
//
// The convolution and its transpose are each written once in gather form; the forward of one is
// the data-gradient of the other (same weight indexing w[dst][src] / w[src][dst]).  Channel
// counts are 1..64 with K = 6..384: HBM bound, so each CTA keeps the whole filter bank in shared
// memory (transposed to [tap][src][dst]) and walks output positions with a 4-wide register tile.
// The 1x1 skip convolutions are never materialised at full resolution (the reference writes
// 66+33+17 MB of which most is cropped away): they are an outer product fused into the add.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPT = 4;   // output positions (consecutive in f) per thread

struct GConvParams {
    const float* in; const float* w; const float* bias; float* out;
    const float* gmul;   // optional: multiply by act'(gmul) (dgrad through the previous activation)
    int B, Ti, Fi, Cs, To, Fo, Cd;
    int act; float slope; int gact; float gslope;
    int nwork;           // B * To * ceil(Fo / kPT)
    int fgroups;
};

// stage w into smem as wt[tap][s][d]; SRC_MAJOR: w[s][d][kt][kf] else w[d][s][kt][kf]
template <bool SRC_MAJOR>
__device__ __forceinline__ void stage_weights(float* wt, const float* __restrict__ w, int Cs, int Cd) {
    const int n = 6 * Cs * Cd;
    for (int idx = threadIdx.x; idx < n; idx += kThreads) {
        int d = idx % Cd;
        int r = idx / Cd;
        int s = r % Cs;
        int tap = r / Cs;
        wt[idx] = SRC_MAJOR ? w[((size_t)s * Cd + d) * 6 + tap] : w[((size_t)d * Cs + s) * 6 + tap];
    }
}

__device__ __forceinline__ void fma_src(float (&acc)[kPT], int q, const float* __restrict__ src, int Cs,
                                        const float* __restrict__ wtap, int Cd, int d) {
    if ((Cs & 3) == 0) {
        for (int s = 0; s < Cs; s += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(src + s));
            const float* wp = wtap + s * Cd + d;
            acc[q] = fmaf(x.x, wp[0], acc[q]);
            acc[q] = fmaf(x.y, wp[Cd], acc[q]);
            acc[q] = fmaf(x.z, wp[2 * Cd], acc[q]);
            acc[q] = fmaf(x.w, wp[3 * Cd], acc[q]);
        }
    } else {
        for (int s = 0; s < Cs; ++s) acc[q] = fmaf(__ldg(src + s), wtap[s * Cd + d], acc[q]);
    }
}

// CONV form:  out[b,t,f,d] = sum_{kt,kf,s} in[b, t+kt-1, 2f+kf-1, s] * w[d][s][kt][kf]
// DECONV form: out[b,t,f,d] = sum_{kt,kf,s} in[b, t+1-kt, (f+1-kf)/2, s] * w[s][d][kt][kf]   ((f+1-kf) even)
template <bool DECONV>
__global__ void __launch_bounds__(kThreads) gconv_kernel(const GConvParams p) {
    extern __shared__ __align__(16) float wt[];
    stage_weights<DECONV>(wt, p.w, p.Cs, p.Cd);
    __syncthreads();
    const int d = threadIdx.x % p.Cd;
    const int py = threadIdx.x / p.Cd;
    const int per_cta = kThreads / p.Cd;
    for (int wi = blockIdx.x * per_cta + py; wi < p.nwork; wi += gridDim.x * per_cta) {
        const int fg = wi % p.fgroups;
        const int bt = wi / p.fgroups;
        const int t = bt % p.To, b = bt / p.To;
        const int f0 = fg * kPT;
        float acc[kPT];
#pragma unroll
        for (int q = 0; q < kPT; ++q) acc[q] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            const int ti = DECONV ? (t + 1 - kt) : (t + kt - 1);
            if (ti < 0 || ti >= p.Ti) continue;
            const float* rowp = p.in + ((size_t)b * p.Ti + ti) * p.Fi * p.Cs;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const float* wtap = wt + (size_t)(kt * 3 + kf) * p.Cs * p.Cd;
#pragma unroll
                for (int q = 0; q < kPT; ++q) {
                    const int f = f0 + q;
                    if (f >= p.Fo) continue;
                    int fi;
                    if (DECONV) {
                        const int num = f + 1 - kf;
                        if (num < 0 || (num & 1)) continue;
                        fi = num >> 1;
                    } else {
                        fi = 2 * f + kf - 1;
                    }
                    if (fi < 0 || fi >= p.Fi) continue;
                    fma_src(acc, q, rowp + (size_t)fi * p.Cs, p.Cs, wtap, p.Cd, d);
                }
            }
        }
        const float bv = p.bias ? p.bias[d] : 0.f;
#pragma unroll
        for (int q = 0; q < kPT; ++q) {
            const int f = f0 + q;
            if (f >= p.Fo) continue;
            const size_t o = (((size_t)b * p.To + t) * p.Fo + f) * p.Cd + d;
            float v = apply_act(acc[q] + bv, p.act, p.slope);
            if (p.gmul) v *= act_grad_from_out(p.gmul[o], p.gact, p.gslope);
            p.out[o] = v;
        }
    }
}

// dW[a][c][kt][kf] += sum_{b,t,f} S[b,t,f,a] * Lg[b, t+kt-1, 2f+kf-1, c]
//   conv:   S = dOut (a = out channel),  Lg = input  (c = in channel)   -> w[co][ci][2][3]
//   deconv: S = input (a = in channel),  Lg = dOut   (c = out channel)  -> w[ci][co][2][3]
struct GWgradParams {
    const float* S; const float* Lg; float* dW;
    int B, Ts, Fs, Ca, Tl, Fl, Cc;
    int rows_per_cta;
};

constexpr int kWgMaxPairs = 8;   // (a,c) pairs per thread

__global__ void __launch_bounds__(kThreads) gconv_wgrad_kernel(const GWgradParams p) {
    extern __shared__ __align__(16) float sm[];
    float* Ssm = sm;                                   // [Fs][Ca]
    float* Lsm = sm + (size_t)p.Fs * p.Ca;             // [2][Fl + 2][Cc]   (one zero column on each side)
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * p.rows_per_cta;
    const int npairs = p.Ca * p.Cc;
    float acc[kWgMaxPairs][6];
#pragma unroll
    for (int i = 0; i < kWgMaxPairs; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[i][k] = 0.f;
    const int lrow = (p.Fl + 2) * p.Cc;
    for (int t = t0; t < min(t0 + p.rows_per_cta, p.Ts); ++t) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < p.Fs * p.Ca; idx += kThreads)
            Ssm[idx] = p.S[((size_t)b * p.Ts + t) * p.Fs * p.Ca + idx];
        for (int idx = threadIdx.x; idx < 2 * lrow; idx += kThreads) {
            int kt = idx / lrow;
            int rem = idx - kt * lrow;
            int fc = rem / p.Cc, c = rem - fc * p.Cc;
            int fl = fc - 1, tl = t + kt - 1;
            float v = 0.f;
            if (fl >= 0 && fl < p.Fl && tl >= 0 && tl < p.Tl) v = p.Lg[(((size_t)b * p.Tl + tl) * p.Fl + fl) * p.Cc + c];
            Lsm[idx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kWgMaxPairs; ++i) {
            const int pr = threadIdx.x + i * kThreads;
            if (pr >= npairs) break;
            const int a = pr / p.Cc, c = pr - a * p.Cc;
            for (int f = 0; f < p.Fs; ++f) {
                const float sv = Ssm[f * p.Ca + a];
                const int fb = 2 * f;   // smem column of Lg index 2f-1  (shifted by the zero column)
                if (fb + 2 >= p.Fl + 2) {
                    // guarded tail (Lg narrower than 2*Fs+1)
#pragma unroll
                    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                        for (int kf = 0; kf < 3; ++kf) {
                            int col = fb + kf;
                            if (col < p.Fl + 2) acc[i][kt * 3 + kf] = fmaf(sv, Lsm[kt * lrow + col * p.Cc + c], acc[i][kt * 3 + kf]);
                        }
                } else {
#pragma unroll
                    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                        for (int kf = 0; kf < 3; ++kf)
                            acc[i][kt * 3 + kf] = fmaf(sv, Lsm[kt * lrow + (fb + kf) * p.Cc + c], acc[i][kt * 3 + kf]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kWgMaxPairs; ++i) {
        const int pr = threadIdx.x + i * kThreads;
        if (pr >= npairs) break;
#pragma unroll
        for (int k = 0; k < 6; ++k) atomicAdd(&p.dW[(size_t)pr * 6 + k], acc[i][k]);
    }
}

// Register-blocked form of the same reduction: a thread owns a 4 (a) x TC (c) block of (a, c) pairs, i.e. 24 TC
// accumulators, and per position reads 4 S values and 6 TC Lg values for 24 TC FMAs (the kernel above reads 7 words
// per 6 FMAs and was bound by shared-memory loads).  Threads that do not fit the pair grid split the f range; the
// slices meet in shared memory, then one atomic per weight per CTA.
template <int TC>
__global__ void __launch_bounds__(kThreads) gconv_wgrad_blk_kernel(const GWgradParams p) {
    extern __shared__ __align__(16) float sm[];
    float* Ssm = sm;                                   // [Fs][Ca]
    float* Lsm = sm + (size_t)p.Fs * p.Ca;             // [2][Fl + 2][Cc]   (one zero column on each side)
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * p.rows_per_cta;
    const int ntc = p.Cc / TC, ntiles = (p.Ca / 4) * ntc, nslices = kThreads / ntiles;
    const int tile = threadIdx.x % ntiles, slice = threadIdx.x / ntiles;
    const int a0 = (tile / ntc) * 4, c0 = (tile % ntc) * TC;
    float acc[4][TC][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j)
#pragma unroll
            for (int k = 0; k < 6; ++k) acc[i][j][k] = 0.f;
    const int lrow = (p.Fl + 2) * p.Cc;
    for (int t = t0; t < min(t0 + p.rows_per_cta, p.Ts); ++t) {
        __syncthreads();
        const float* srow = p.S + ((size_t)b * p.Ts + t) * p.Fs * p.Ca;
        for (int idx = threadIdx.x * 4; idx < p.Fs * p.Ca; idx += kThreads * 4)
            *reinterpret_cast<float4*>(Ssm + idx) = *reinterpret_cast<const float4*>(srow + idx);
        for (int idx = threadIdx.x; idx < 2 * lrow; idx += kThreads) {
            int kt = idx / lrow;
            int rem = idx - kt * lrow;
            int fc = rem / p.Cc, c = rem - fc * p.Cc;
            int fl = fc - 1, tl = t + kt - 1;
            float v = 0.f;
            if (fl >= 0 && fl < p.Fl && tl >= 0 && tl < p.Tl) v = p.Lg[(((size_t)b * p.Tl + tl) * p.Fl + fl) * p.Cc + c];
            Lsm[idx] = v;
        }
        __syncthreads();
        if (slice < nslices) {
            for (int f = slice; f < p.Fs; f += nslices) {
                const float4 s4 = *reinterpret_cast<const float4*>(Ssm + f * p.Ca + a0);
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int kt = 0; kt < 2; ++kt)
#pragma unroll
                    for (int kf = 0; kf < 3; ++kf) {
                        const int col = 2 * f + kf;          // smem column of Lg index 2f + kf - 1
                        if (col >= p.Fl + 2) continue;
                        float lv[TC];
                        const float* lp = Lsm + kt * lrow + col * p.Cc + c0;
                        if (TC == 4) {
                            const float4 l4 = *reinterpret_cast<const float4*>(lp);
                            lv[0] = l4.x; lv[1 % TC] = l4.y; lv[2 % TC] = l4.z; lv[3 % TC] = l4.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < TC; ++j) lv[j] = lp[j];
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) acc[i][j][kt * 3 + kf] = fmaf(sv[i], lv[j], acc[i][j][kt * 3 + kf]);
                    }
            }
        }
    }
    __syncthreads();
    float* red = sm;                                    // [Ca * Cc * 6]
    const int nred = p.Ca * p.Cc * 6;
    for (int i = threadIdx.x; i < nred; i += kThreads) red[i] = 0.f;
    __syncthreads();
    if (slice < nslices) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j)
#pragma unroll
                for (int k = 0; k < 6; ++k) atomicAdd(&red[((a0 + i) * p.Cc + c0 + j) * 6 + k], acc[i][j][k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nred; i += kThreads) atomicAdd(&p.dW[i], red[i]);
}

// out[b,t,f,c] = h[b,t,f,c] + mag[b,t,f] * w[c] + bias[c]   on the common low-index corner
__global__ void skip_add_fwd_kernel(const float* __restrict__ h, const float* __restrict__ mag,
                                    const float* __restrict__ w, const float* __restrict__ bias,
                                    float* __restrict__ out, int B, int Th, int Fh, int Tm, int Fm, int To, int Fo,
                                    int C) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)B * To * Fo * C;
    if (i >= n) return;
    int c = (int)(i % C);
    int64_t r = i / C;
    int f = (int)(r % Fo);
    r /= Fo;
    int t = (int)(r % To);
    int b = (int)(r / To);
    float hv = h[(((size_t)b * Th + t) * Fh + f) * C + c];
    float mv = mag[((size_t)b * Tm + t) * Fm + f];
    out[i] = hv + mv * w[c] + bias[c];
}

// g: [B,To,Fo,C] grad of the sum.  dh (optional): [B,Th,Fh,C] = g zero-extended; dw[c] += sum g*mag; db[c] += sum g
__global__ void skip_add_bwd_kernel(const float* __restrict__ g, const float* __restrict__ mag,
                                    float* __restrict__ dh, float* __restrict__ dw, float* __restrict__ db, int B,
                                    int Th, int Fh, int Tm, int Fm, int To, int Fo, int C, int rows_per_cta) {
    // one thread per channel slot; CTA walks `rows_per_cta` (b,t) rows
    extern __shared__ float red[];   // [2][blockDim]
    const int c = threadIdx.x % C;
    const int lanes = blockDim.x / C;      // positions processed in parallel
    const int pl = threadIdx.x / C;
    float aw = 0.f, ab = 0.f;
    const int64_t rows = (int64_t)B * To;
    for (int64_t r = (int64_t)blockIdx.x * rows_per_cta; r < min(rows, (int64_t)(blockIdx.x + 1) * rows_per_cta); ++r) {
        const int b = (int)(r / To), t = (int)(r % To);
        for (int f = pl; f < Fo; f += lanes) {
            float gv = g[(((size_t)b * To + t) * Fo + f) * C + c];
            float mv = mag[((size_t)b * Tm + t) * Fm + f];
            aw += gv * mv;
            ab += gv;
        }
    }
    red[threadIdx.x] = aw;
    red[blockDim.x + threadIdx.x] = ab;
    __syncthreads();
    if (threadIdx.x < C) {
        float sw = 0.f, sb = 0.f;
        for (int l = 0; l < lanes; ++l) {
            sw += red[l * C + threadIdx.x];
            sb += red[blockDim.x + l * C + threadIdx.x];
        }
        atomicAdd(&dw[threadIdx.x], sw);
        atomicAdd(&db[threadIdx.x], sb);
    }
    (void)dh; (void)Th; (void)Fh;
}

__global__ void crop_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int Ts, int Fs,
                                int Td, int Fd, int C) {
    // dst[b,t,f,c] = (t < Ts && f < Fs) ? src[b,t,f,c] : 0
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)B * Td * Fd * C;
    if (i >= n) return;
    int c = (int)(i % C);
    int64_t r = i / C;
    int f = (int)(r % Fd);
    r /= Fd;
    int t = (int)(r % Td);
    int b = (int)(r / Td);
    dst[i] = (t < Ts && f < Fs) ? src[(((size_t)b * Ts + t) * Fs + f) * C + c] : 0.f;
}

// mask[b,t,f] = sigmoid( (t<Ty && f<Fy) ? y[b,t,f] : 0 )  for t<T, f<F   (sigmoid optional)
__global__ void final_mask_fwd_kernel(const float* __restrict__ y, float* __restrict__ mask, int B, int Ty, int Fy,
                                      int T, int F, int use_sigmoid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)B * T * F;
    if (i >= n) return;
    int f = (int)(i % F);
    int64_t r = i / F;
    int t = (int)(r % T);
    int b = (int)(r / T);
    float v = (t < Ty && f < Fy) ? y[((size_t)b * Ty + t) * Fy + f] : 0.f;
    mask[i] = use_sigmoid ? 1.f / (1.f + expf(-v)) : v;
}

// dpre[b,t,f] (t<Ty,f<Fy) = (t<T && f<F) ? gmask * sigmoid' * relu'(y) : 0     (y is post-ReLU)
__global__ void final_mask_bwd_kernel(const float* __restrict__ y, const float* __restrict__ mask,
                                      const float* __restrict__ gmask, float* __restrict__ dpre, int B, int Ty,
                                      int Fy, int T, int F, int use_sigmoid, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = (int64_t)B * Ty * Fy;
    if (i >= n) return;
    int f = (int)(i % Fy);
    int64_t r = i / Fy;
    int t = (int)(r % Ty);
    int b = (int)(r / Ty);
    float v = 0.f;
    if (t < T && f < F) {
        size_t o = ((size_t)b * T + t) * F + f;
        float m = mask[o];
        v = gmask[o] * (use_sigmoid ? m * (1.f - m) : 1.f) * act_grad_from_out(y[i], act, slope);
    }
    dpre[i] = v;
}

bool pow2_le256(int64_t c) { return c >= 1 && c <= 256 && (c & (c - 1)) == 0; }

template <bool DECONV>
int launch_gconv(GConvParams& p, cudaStream_t st) {
    if (!pow2_le256(p.Cd)) return LCT_EUNSUPPORTED;
    p.fgroups = (p.Fo + kPT - 1) / kPT;
    int64_t nwork = (int64_t)p.B * p.To * p.fgroups;
    if (nwork >= (1LL << 31)) return LCT_EINVAL;
    p.nwork = (int)nwork;
    size_t smem = (size_t)6 * p.Cs * p.Cd * sizeof(float);
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gconv_kernel<DECONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int per_cta = kThreads / p.Cd;
    int64_t blocks = ceil_div64(nwork, per_cta);
    const int64_t cap = 148 * 8;
    if (blocks > cap) blocks = cap;
    gconv_kernel<DECONV><<<(unsigned)blocks, kThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

}  // namespace

// transposed = 0: out = act(conv(in, w[Cd][Cs][2][3]) + bias);  1: out = act(conv_transpose(in, w[Cs][Cd][2][3]) + bias)
// optionally multiplied by act'(gmul) (gact/gslope) so the same kernels serve as data-gradients.
LCT_API int lct_gconv(const float* in, const float* w, const float* bias, float* out, const float* gmul,
                      int transposed, int64_t B, int64_t Ti, int64_t Fi, int64_t Cs, int64_t To, int64_t Fo,
                      int64_t Cd, int act, float slope, int gact, float gslope, cudaStream_t st) {
    if (!in || !w || !out || B <= 0 || Ti <= 0 || Fi <= 0 || Cs <= 0 || To <= 0 || Fo <= 0 || Cd <= 0) return LCT_EINVAL;
    GConvParams p = {};
    p.in = in; p.w = w; p.bias = bias; p.out = out; p.gmul = gmul;
    p.B = (int)B; p.Ti = (int)Ti; p.Fi = (int)Fi; p.Cs = (int)Cs; p.To = (int)To; p.Fo = (int)Fo; p.Cd = (int)Cd;
    p.act = act; p.slope = slope; p.gact = gact; p.gslope = gslope;
    return transposed ? launch_gconv<true>(p, st) : launch_gconv<false>(p, st);
}

// rowgemm.cu: the same reduction as a 3xTF32 tensor-core TN GEMM with a gathered operand (tensor-core mode)
int lct_gconv_wgrad_try(const float* S, const float* Lg, float* dW, int64_t B, int64_t Ts, int64_t Fs, int64_t Ca,
                        int64_t Tl, int64_t Fl, int64_t Cc, cudaStream_t st, int* rc);
int lct_rowgemm_enabled();      // gemm.cu: lct_set_rowgemm switch (tensor-core mode)

// dW[Ca][Cc][2][3] += sum S[b,t,f,a] * Lg[b,t+kt-1,2f+kf-1,c]   (caller zeroes dW)
LCT_API int lct_gconv_wgrad(const float* S, const float* Lg, float* dW, int64_t B, int64_t Ts, int64_t Fs, int64_t Ca,
                            int64_t Tl, int64_t Fl, int64_t Cc, cudaStream_t st) {
    if (!S || !Lg || !dW || B <= 0 || B >= 65536 || Ts <= 0 || Fs <= 0 || Ca <= 0 || Tl <= 0 || Fl <= 0 || Cc <= 0)
        return LCT_EINVAL;
    if (Ca * Cc > (int64_t)kWgMaxPairs * kThreads) return LCT_EUNSUPPORTED;
    if (lct_rowgemm_enabled()) {
        int rc = 0;
        if (lct_gconv_wgrad_try(S, Lg, dW, B, Ts, Fs, Ca, Tl, Fl, Cc, st, &rc)) return rc;
    }
    GWgradParams p;
    p.S = S; p.Lg = Lg; p.dW = dW;
    p.B = (int)B; p.Ts = (int)Ts; p.Fs = (int)Fs; p.Ca = (int)Ca; p.Tl = (int)Tl; p.Fl = (int)Fl; p.Cc = (int)Cc;
    {
        // register-blocked kernel: 4 x TC pair blocks must tile the 256 threads
        const int tc = (Cc % 4 == 0) ? 4 : 1;
        const int64_t ntiles = (Ca / 4) * (Cc / tc);
        const bool aligned = ((uintptr_t)S & 15) == 0 && (Ca % 4) == 0 && (Cc == 1 || Cc % 4 == 0);
        if (aligned && ntiles >= 1 && ntiles <= kThreads && kThreads % ntiles == 0) {
            int rows = (int)ceil_div64(B * Ts, 296);          // ~2 CTAs per SM: every CTA ends with Ca*Cc*6 atomics
            if (rows < 1) rows = 1;
            p.rows_per_cta = rows;
            size_t smem = ((size_t)Fs * Ca + (size_t)2 * (Fl + 2) * Cc) * sizeof(float);
            const size_t red = (size_t)Ca * Cc * 6 * sizeof(float);
            if (smem < red) smem = red;
            if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
            dim3 grid((unsigned)ceil_div64(Ts, rows), (unsigned)B);
            if (tc == 4) {
                if (smem > 48 * 1024) {
                    cudaError_t e = cudaFuncSetAttribute(gconv_wgrad_blk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e != cudaSuccess) return (int)e;
                }
                gconv_wgrad_blk_kernel<4><<<grid, kThreads, smem, st>>>(p);
            } else {
                if (smem > 48 * 1024) {
                    cudaError_t e = cudaFuncSetAttribute(gconv_wgrad_blk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    if (e != cudaSuccess) return (int)e;
                }
                gconv_wgrad_blk_kernel<1><<<grid, kThreads, smem, st>>>(p);
            }
            LCT_RETURN_IF_LAUNCH_FAILED();
            return 0;
        }
    }
    p.rows_per_cta = 2;
    size_t smem = ((size_t)Fs * Ca + (size_t)2 * (Fl + 2) * Cc) * sizeof(float);
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(gconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)ceil_div64(Ts, p.rows_per_cta), (unsigned)B);
    gconv_wgrad_kernel<<<grid, kThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_skip_add_fwd(const float* h, const float* mag, const float* w, const float* bias, float* out,
                             int64_t B, int64_t Th, int64_t Fh, int64_t Tm, int64_t Fm, int64_t C, cudaStream_t st) {
    if (!h || !mag || !w || !bias || !out || B <= 0 || Th <= 0 || Fh <= 0 || Tm <= 0 || Fm <= 0 || C <= 0) return LCT_EINVAL;
    int To = (int)(Th < Tm ? Th : Tm), Fo = (int)(Fh < Fm ? Fh : Fm);
    int64_t n = B * To * Fo * C;
    skip_add_fwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(h, mag, w, bias, out, (int)B, (int)Th, (int)Fh,
                                                                      (int)Tm, (int)Fm, To, Fo, (int)C);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// g [B,min(Th,Tm),min(Fh,Fm),C] -> dw[C], db[C] accumulated; dh [B,Th,Fh,C] (optional) = g zero-extended
LCT_API int lct_skip_add_bwd(const float* g, const float* mag, float* dh, float* dw, float* db, int64_t B, int64_t Th,
                             int64_t Fh, int64_t Tm, int64_t Fm, int64_t C, cudaStream_t st) {
    if (!g || !mag || !dw || !db || B <= 0 || Th <= 0 || Fh <= 0 || Tm <= 0 || Fm <= 0 || !pow2_le256(C)) return LCT_EINVAL;
    int To = (int)(Th < Tm ? Th : Tm), Fo = (int)(Fh < Fm ? Fh : Fm);
    const int rows_per_cta = 4;
    int64_t rows = B * To;
    skip_add_bwd_kernel<<<(unsigned)ceil_div64(rows, rows_per_cta), 256, 2 * 256 * sizeof(float), st>>>(
        g, mag, nullptr, dw, db, (int)B, (int)Th, (int)Fh, (int)Tm, (int)Fm, To, Fo, (int)C, rows_per_cta);
    LCT_RETURN_IF_LAUNCH_FAILED();
    if (dh) {
        int64_t n = B * Th * Fh * C;
        crop_pad_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(g, dh, (int)B, To, Fo, (int)Th, (int)Fh, (int)C);
        LCT_RETURN_IF_LAUNCH_FAILED();
    }
    return 0;
}

// dst[B,Td,Fd,C] = src[B,Ts,Fs,C] cropped / zero-extended at the high-index side
LCT_API int lct_crop_pad(const float* src, float* dst, int64_t B, int64_t Ts, int64_t Fs, int64_t Td, int64_t Fd,
                         int64_t C, cudaStream_t st) {
    if (!src || !dst || B <= 0 || Ts <= 0 || Fs <= 0 || Td <= 0 || Fd <= 0 || C <= 0) return LCT_EINVAL;
    int64_t n = B * Td * Fd * C;
    crop_pad_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(src, dst, (int)B, (int)Ts, (int)Fs, (int)Td, (int)Fd,
                                                                  (int)C);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_final_mask_fwd(const float* y, float* mask, int64_t B, int64_t Ty, int64_t Fy, int64_t T, int64_t F,
                               int use_sigmoid, cudaStream_t st) {
    if (!y || !mask || B <= 0 || Ty <= 0 || Fy <= 0 || T <= 0 || F <= 0) return LCT_EINVAL;
    int64_t n = B * T * F;
    final_mask_fwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(y, mask, (int)B, (int)Ty, (int)Fy, (int)T,
                                                                        (int)F, use_sigmoid);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_final_mask_bwd(const float* y, const float* mask, const float* gmask, float* dpre, int64_t B,
                               int64_t Ty, int64_t Fy, int64_t T, int64_t F, int use_sigmoid, int act, float slope,
                               cudaStream_t st) {
    if (!y || !mask || !gmask || !dpre || B <= 0 || Ty <= 0 || Fy <= 0 || T <= 0 || F <= 0) return LCT_EINVAL;
    int64_t n = B * Ty * Fy;
    final_mask_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(y, mask, gmask, dpre, (int)B, (int)Ty, (int)Fy,
                                                                        (int)T, (int)F, use_sigmoid, act, slope);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
