// Grouped strided 1-D convolutions of the waveform discriminators, fp32 SIMT path.
//
// Replaces (reference file:line):
//   PeriodDiscriminator convs: weight_norm(Conv2d(k=(K,1), stride=(S,1), pad=(K//2,0), groups=G))
//                              models/discriminators.py:37-67, :93-98
//   ScaleDiscriminator convs:  weight_norm(Conv1d(k=K, stride=S, pad=K//2, groups=G))
//                              models/discriminators.py:166-196, :215-220
//   weight_norm reparametrisation w = g * v / ||v||   (torch.nn.utils.weight_norm, dim=0)
//
// Both families are one operator: a convolution along L of a tensor [B, C, L, P] whose innermost
// axis P (the period; 1 for the scale discriminators) is carried along untouched.  The layout is
// the reference's own NCHW / NCL, so feature maps are returned to Python without a copy.
//
// Per-group GEMMs are tiny (N/group 4..32, K/group 5..164): AI 2..40 flop/B, i.e. HBM/issue bound.
// Each CTA stages the input window of its output tile in shared memory, de-interleaved by stride
// phase so that threads walking consecutive output positions read consecutive banks, and keeps a
// [positions x out-channels] register tile.  dgrad is evaluated in polyphase form (S stride-1
// sub-filters), wgrad reduces position tiles in registers and finishes with one atomic per
// (weight, CTA).  The dense 1024->1024 layer (MSD convs.5) has its own tcgen05 path (dense_conv.cu).
#include "common.cuh"

namespace {

constexpr int kFwdThreads = 128;
constexpr int kSmemBudget = 40 * 1024;

struct ConvShape {
    int B, Cin, Cout, G, K, S, pad, Lin, Lout, P;
    int Cin_g, Cout_g;
};

bool fill_shape(ConvShape& s, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad,
                int64_t Lin, int64_t P) {
    if (B <= 0 || B >= 65536 || Cin <= 0 || Cout <= 0 || G <= 0 || K <= 0 || S <= 0 || pad < 0 || Lin <= 0 || P <= 0)
        return false;
    if (Cin % G || Cout % G) return false;
    int64_t Lout = (Lin + 2 * pad - K) / S + 1;
    if (Lin + 2 * pad < K || Lout <= 0) return false;
    if (Lin * P >= (1LL << 30) || Lout * P >= (1LL << 30)) return false;
    s.B = (int)B; s.Cin = (int)Cin; s.Cout = (int)Cout; s.G = (int)G; s.K = (int)K; s.S = (int)S; s.pad = (int)pad;
    s.Lin = (int)Lin; s.Lout = (int)Lout; s.P = (int)P;
    s.Cin_g = (int)(Cin / G); s.Cout_g = (int)(Cout / G);
    return true;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct FwdParams {
    const float* x; const float* w; const float* bias; float* y;
    ConvShape s;
    int act; float slope;
    int CIC;      // input channels staged per pass
    int WS;       // floats per channel in the staged window
    int Q;        // window rows per phase
    int rows_max; // max output rows spanned by a tile
};

template <int OCT, int NPT>
__global__ void __launch_bounds__(kFwdThreads) conv_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(16) float sm[];
    const ConvShape& s = p.s;
    constexpr int TJ = kFwdThreads * NPT;
    float* win = sm;                       // [CIC][WS]
    float* ws = sm + (size_t)p.CIC * p.WS; // [CIC][K][OCT]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int co0 = blockIdx.y * OCT;
    const int g = co0 / s.Cout_g;
    const int jtot = s.Lout * s.P;
    const int j0 = blockIdx.x * TJ;
    const int l_first = j0 / s.P;
    const int l_last = (min(j0 + TJ, jtot) - 1) / s.P;
    const int row0 = l_first * s.S - s.pad;
    const int NR = (l_last - l_first) * s.S + s.K;
    const int QP = p.Q * s.P;

    int off[NPT];
    bool valid[NPT];
#pragma unroll
    for (int n = 0; n < NPT; ++n) {
        int j = j0 + tid + n * kFwdThreads;
        valid[n] = j < jtot;
        off[n] = valid[n] ? (j - l_first * s.P) : 0;
    }
    float acc[NPT][OCT];
#pragma unroll
    for (int n = 0; n < NPT; ++n)
#pragma unroll
        for (int o = 0; o < OCT; ++o) acc[n][o] = 0.f;

    for (int c0 = 0; c0 < s.Cin_g; c0 += p.CIC) {
        const int cc = min(p.CIC, s.Cin_g - c0);
        __syncthreads();
        // stage the input window, de-interleaved by stride phase
        const int per = NR * s.P;
        for (int idx = tid; idx < cc * per; idx += kFwdThreads) {
            int ci = idx / per;
            int rem = idx - ci * per;
            int rw = rem / s.P;
            int pp = rem - rw * s.P;
            int row = row0 + rw;
            float v = 0.f;
            if (row >= 0 && row < s.Lin)
                v = p.x[(((size_t)b * s.Cin + (g * s.Cin_g + c0 + ci)) * s.Lin + row) * s.P + pp];
            win[ci * p.WS + ((rw % s.S) * p.Q + rw / s.S) * s.P + pp] = v;
        }
        for (int idx = tid; idx < cc * s.K * OCT; idx += kFwdThreads) {
            int o = idx % OCT;
            int t = idx / OCT;
            int k = t % s.K;
            int ci = t / s.K;
            ws[idx] = p.w[((size_t)(co0 + o) * s.Cin_g + (c0 + ci)) * s.K + k];
        }
        __syncthreads();
        for (int ci = 0; ci < cc; ++ci) {
            const float* wc = win + ci * p.WS;
            int kr = 0, kq = 0;   // k % S, k / S
            for (int k = 0; k < s.K; ++k) {
                const int koff = kr * QP + kq * s.P;
                float xv[NPT];
#pragma unroll
                for (int n = 0; n < NPT; ++n) xv[n] = wc[koff + off[n]];
                const float* wk = ws + (ci * s.K + k) * OCT;
                float wv[OCT];
                if (OCT % 4 == 0) {
#pragma unroll
                    for (int o = 0; o < OCT; o += 4) {
                        float4 t4 = *reinterpret_cast<const float4*>(wk + o);
                        wv[o] = t4.x; wv[o + 1] = t4.y; wv[o + 2] = t4.z; wv[o + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < OCT; ++o) wv[o] = wk[o];
                }
#pragma unroll
                for (int n = 0; n < NPT; ++n)
#pragma unroll
                    for (int o = 0; o < OCT; ++o) acc[n][o] = fmaf(xv[n], wv[o], acc[n][o]);
                if (++kr == s.S) { kr = 0; ++kq; }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < OCT; ++o) {
        const float bv = p.bias ? p.bias[co0 + o] : 0.f;
        float* yo = p.y + ((size_t)b * s.Cout + co0 + o) * jtot;
#pragma unroll
        for (int n = 0; n < NPT; ++n) {
            if (valid[n]) yo[j0 + tid + n * kFwdThreads] = apply_act(acc[n][o] + bv, p.act, p.slope);
        }
    }
}

template <int OCT, int NPT>
int launch_fwd(FwdParams& p, cudaStream_t st) {
    const ConvShape& s = p.s;
    constexpr int TJ = kFwdThreads * NPT;
    p.rows_max = TJ / s.P + 2;
    const int NRmax = (p.rows_max - 1) * s.S + s.K;
    p.Q = (NRmax + s.S - 1) / s.S + 1;
    p.WS = (p.Q * s.S * s.P + 3) & ~3;
    size_t per_ci = (size_t)(p.WS + s.K * OCT) * sizeof(float);
    int cic = (int)(kSmemBudget / per_ci);
    if (cic < 1) cic = 1;
    if (cic > s.Cin_g) cic = s.Cin_g;
    p.CIC = cic;
    size_t smem = per_ci * cic;
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel<OCT, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)ceil_div64((int64_t)s.Lout * s.P, TJ), (unsigned)(s.Cout / OCT), (unsigned)s.B);
    conv_fwd_kernel<OCT, NPT><<<grid, kFwdThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// dgrad (polyphase):  u = m + pad = S*q + r ;  dX[u] = sum_t sum_co dY[q - t] * w[co][ci][r + S*t]
// ---------------------------------------------------------------------------------------------
struct DgradParams {
    const float* dy; const float* w; float* dx;
    const float* gextra;   // optional, added to dX before the activation derivative
    const float* xact;     // optional: the (post-activation) tensor dX is the gradient of
    ConvShape s;
    int act; float slope;
    int COC, WS, Tmax, Qtot;
};

template <int S, int CIT, int NQ>
__global__ void __launch_bounds__(kFwdThreads) conv_dgrad_kernel(const DgradParams p) {
    extern __shared__ __align__(16) float sm[];
    const ConvShape& s = p.s;
    constexpr int TJ = kFwdThreads * NQ;
    float* win = sm;                          // [COC][WS]
    float* wd = sm + (size_t)p.COC * p.WS;    // [COC][Tmax][S][CIT]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int ci0 = blockIdx.y * CIT;
    const int g = ci0 / s.Cin_g;
    const int cil0 = ci0 - g * s.Cin_g;       // channel offset inside the group
    const int jtot = p.Qtot * s.P;
    const int j0 = blockIdx.x * TJ;
    const int q_first = j0 / s.P;
    const int q_last = (min(j0 + TJ, jtot) - 1) / s.P;
    const int lrow0 = q_first - p.Tmax + 1;
    const int NRy = q_last - q_first + p.Tmax;

    int off[NQ];
    bool valid[NQ];
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
        int j = j0 + tid + n * kFwdThreads;
        valid[n] = j < jtot;
        off[n] = valid[n] ? (j - q_first * s.P) : 0;
    }
    float acc[NQ][S][CIT];
#pragma unroll
    for (int n = 0; n < NQ; ++n)
#pragma unroll
        for (int r = 0; r < S; ++r)
#pragma unroll
            for (int c = 0; c < CIT; ++c) acc[n][r][c] = 0.f;

    const int jy = s.Lout * s.P;
    for (int c0 = 0; c0 < s.Cout_g; c0 += p.COC) {
        const int cc = min(p.COC, s.Cout_g - c0);
        __syncthreads();
        const int per = NRy * s.P;
        for (int idx = tid; idx < cc * per; idx += kFwdThreads) {
            int co = idx / per;
            int rem = idx - co * per;
            int rw = rem / s.P;
            int pp = rem - rw * s.P;
            int l = lrow0 + rw;
            float v = 0.f;
            if (l >= 0 && l < s.Lout)
                v = p.dy[((size_t)b * s.Cout + (g * s.Cout_g + c0 + co)) * jy + (size_t)l * s.P + pp];
            win[co * p.WS + rem] = v;
        }
        const int wper = p.Tmax * S * CIT;
        for (int idx = tid; idx < cc * wper; idx += kFwdThreads) {
            int co = idx / wper;
            int rem = idx - co * wper;
            int t = rem / (S * CIT);
            int rc = rem - t * (S * CIT);
            int r = rc / CIT;
            int c = rc - r * CIT;
            int k = r + S * t;
            float v = 0.f;
            if (k < s.K)
                v = p.w[((size_t)(g * s.Cout_g + c0 + co) * s.Cin_g + (cil0 + c)) * s.K + k];
            wd[idx] = v;
        }
        __syncthreads();
        for (int co = 0; co < cc; ++co) {
            const float* wc = win + co * p.WS;
            for (int t = 0; t < p.Tmax; ++t) {
                const int base = (p.Tmax - 1 - t) * s.P;
                float dv[NQ];
#pragma unroll
                for (int n = 0; n < NQ; ++n) dv[n] = wc[base + off[n]];
                const float* wk = wd + (co * p.Tmax + t) * (S * CIT);
                float wv[S * CIT];
                if constexpr ((S * CIT) % 4 == 0) {
#pragma unroll
                    for (int i = 0; i < S * CIT; i += 4) {
                        float4 t4 = *reinterpret_cast<const float4*>(wk + i);
                        wv[i] = t4.x; wv[i + 1] = t4.y; wv[i + 2] = t4.z; wv[i + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < S * CIT; ++i) wv[i] = wk[i];
                }
#pragma unroll
                for (int n = 0; n < NQ; ++n)
#pragma unroll
                    for (int r = 0; r < S; ++r)
#pragma unroll
                        for (int c = 0; c < CIT; ++c) acc[n][r][c] = fmaf(dv[n], wv[r * CIT + c], acc[n][r][c]);
            }
        }
    }
    const int jx = s.Lin * s.P;
#pragma unroll
    for (int n = 0; n < NQ; ++n) {
        if (!valid[n]) continue;
        int j = j0 + tid + n * kFwdThreads;
        int q = j / s.P, pp = j - q * s.P;
#pragma unroll
        for (int r = 0; r < S; ++r) {
            int m = S * q + r - s.pad;
            if (m < 0 || m >= s.Lin) continue;
#pragma unroll
            for (int c = 0; c < CIT; ++c) {
                size_t idx = ((size_t)b * s.Cin + ci0 + c) * jx + (size_t)m * s.P + pp;
                float v = acc[n][r][c];
                if (p.gextra) v += p.gextra[idx];
                if (p.xact) v *= act_grad_from_out(p.xact[idx], p.act, p.slope);
                p.dx[idx] = v;
            }
        }
    }
}

template <int S, int CIT, int NQ>
int launch_dgrad(DgradParams& p, cudaStream_t st) {
    const ConvShape& s = p.s;
    constexpr int TJ = kFwdThreads * NQ;
    p.Tmax = (s.K + S - 1) / S;
    p.Qtot = (s.Lin + s.pad + S - 1) / S;
    const int rows = TJ / s.P + 2 + p.Tmax;
    p.WS = (rows * s.P + 3) & ~3;
    size_t per_co = (size_t)(p.WS + p.Tmax * S * CIT) * sizeof(float);
    int coc = (int)(kSmemBudget / per_co);
    if (coc < 1) coc = 1;
    if (coc > s.Cout_g) coc = s.Cout_g;
    p.COC = coc;
    size_t smem = per_co * coc;
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv_dgrad_kernel<S, CIT, NQ>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)ceil_div64((int64_t)p.Qtot * s.P, TJ), (unsigned)(s.Cin / CIT), (unsigned)s.B);
    conv_dgrad_kernel<S, CIT, NQ><<<grid, kFwdThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int S>
int dispatch_dgrad_cit(DgradParams& p, cudaStream_t st) {
    const int cg = p.s.Cin_g;
    if constexpr (S == 1) {
        if (cg % 16 == 0) return launch_dgrad<S, 16, 4>(p, st);
    }
    if constexpr (S <= 3) {
        if (cg % 8 == 0) return launch_dgrad<S, 8, 2>(p, st);
    }
    if (cg % 4 == 0) return launch_dgrad<S, 4, (S <= 2 ? 4 : 2)>(p, st);
    return launch_dgrad<S, 1, 4>(p, st);
}

// ---------------------------------------------------------------------------------------------
// wgrad:  dW[co][ci][k] += sum_{b, l, p} dY[b][co][l][p] * X[b][ci][l*S + k - pad][p];  db[co] += sum dY
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 256;
constexpr int kWgTJ = 256;   // output positions per CTA

struct WgradParams {
    const float* x; const float* dy; float* dw; float* db;
    ConvShape s;
    int CIC, WS, Q;
};

template <int OCT>
__global__ void __launch_bounds__(kWgThreads) conv_wgrad_kernel(const WgradParams p) {
    extern __shared__ __align__(16) float sm[];
    const ConvShape& s = p.s;
    float* dyt = sm;                                   // [kWgTJ][OCT]
    float* win = sm + (size_t)kWgTJ * OCT;             // [CIC][WS]
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int co0 = blockIdx.y * OCT;
    const int g = co0 / s.Cout_g;
    const int jtot = s.Lout * s.P;
    const int j0 = blockIdx.x * kWgTJ;
    const int nj = min(kWgTJ, jtot - j0);
    const int l_first = j0 / s.P;
    const int l_last = (j0 + nj - 1) / s.P;
    const int row0 = l_first * s.S - s.pad;
    const int NR = (l_last - l_first) * s.S + s.K;
    const int QP = p.Q * s.P;
    const int joff = j0 - l_first * s.P;   // offset of the first tile position inside its row block

    for (int idx = tid; idx < kWgTJ * OCT; idx += kWgThreads) {
        int o = idx % OCT;
        int jr = idx / OCT;
        float v = 0.f;
        if (jr < nj) v = p.dy[((size_t)b * s.Cout + co0 + o) * jtot + j0 + jr];
        dyt[idx] = v;
    }
    __syncthreads();
    if (p.db) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int o = warp; o < OCT; o += kWgThreads / 32) {
            float a = 0.f;
            for (int jr = lane; jr < nj; jr += 32) a += dyt[jr * OCT + o];
            a = warp_sum(a);
            if (lane == 0) atomicAdd(&p.db[co0 + o], a);
        }
    }
    for (int c0 = 0; c0 < s.Cin_g; c0 += p.CIC) {
        const int cc = min(p.CIC, s.Cin_g - c0);
        __syncthreads();
        const int per = NR * s.P;
        for (int idx = tid; idx < cc * per; idx += kWgThreads) {
            int ci = idx / per;
            int rem = idx - ci * per;
            int rw = rem / s.P;
            int pp = rem - rw * s.P;
            int row = row0 + rw;
            float v = 0.f;
            if (row >= 0 && row < s.Lin)
                v = p.x[(((size_t)b * s.Cin + (g * s.Cin_g + c0 + ci)) * s.Lin + row) * s.P + pp];
            win[ci * p.WS + ((rw % s.S) * p.Q + rw / s.S) * s.P + pp] = v;
        }
        __syncthreads();
        const int np = cc * s.K;
        const int nslice = kWgThreads / np;     // host guarantees CIC*K <= kWgThreads
        const int pi = tid % np, sl = tid / np;
        if (sl < nslice) {
            const int ci = pi / s.K, k = pi - ci * s.K;
            const float* wc = win + ci * p.WS + (k % s.S) * QP + (k / s.S) * s.P + joff;
            float acc[OCT];
#pragma unroll
            for (int o = 0; o < OCT; ++o) acc[o] = 0.f;
            for (int jr = sl; jr < nj; jr += nslice) {
                const float xv = wc[jr];
                const float* d = dyt + jr * OCT;
                if (OCT % 4 == 0) {
#pragma unroll
                    for (int o = 0; o < OCT; o += 4) {
                        float4 t4 = *reinterpret_cast<const float4*>(d + o);
                        acc[o] = fmaf(xv, t4.x, acc[o]);
                        acc[o + 1] = fmaf(xv, t4.y, acc[o + 1]);
                        acc[o + 2] = fmaf(xv, t4.z, acc[o + 2]);
                        acc[o + 3] = fmaf(xv, t4.w, acc[o + 3]);
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < OCT; ++o) acc[o] = fmaf(xv, d[o], acc[o]);
                }
            }
#pragma unroll
            for (int o = 0; o < OCT; ++o)
                atomicAdd(&p.dw[((size_t)(co0 + o) * s.Cin_g + (c0 + ci)) * s.K + k], acc[o]);
        }
    }
}

template <int OCT>
int launch_wgrad(WgradParams& p, cudaStream_t st) {
    const ConvShape& s = p.s;
    const int rows_max = kWgTJ / s.P + 2;
    const int NRmax = (rows_max - 1) * s.S + s.K;
    p.Q = (NRmax + s.S - 1) / s.S + 1;
    p.WS = p.Q * s.S * s.P;
    if (s.K > kWgThreads) return LCT_EUNSUPPORTED;
    size_t per_ci = (size_t)p.WS * sizeof(float);
    int cic = (int)(kSmemBudget / per_ci);
    if (cic < 1) cic = 1;
    if (cic > s.Cin_g) cic = s.Cin_g;
    if (cic * s.K > kWgThreads) cic = kWgThreads / s.K;
    p.CIC = cic;
    size_t smem = per_ci * cic + (size_t)kWgTJ * OCT * sizeof(float);
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<OCT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)ceil_div64((int64_t)s.Lout * s.P, kWgTJ), (unsigned)(s.Cout / OCT), (unsigned)s.B);
    conv_wgrad_kernel<OCT><<<grid, kWgThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// weight norm:  w[co,:] = g[co] * v[co,:] / ||v[co,:]||
// ---------------------------------------------------------------------------------------------
__global__ void wnorm_fwd_kernel(const float* __restrict__ g, const float* __restrict__ v, float* __restrict__ w,
                                 float* __restrict__ norm, int row) {
    __shared__ float red[32];
    const int co = blockIdx.x;
    const float* vr = v + (size_t)co * row;
    float a = 0.f;
    for (int i = threadIdx.x; i < row; i += blockDim.x) a += vr[i] * vr[i];
    float nrm = sqrtf(block_sum(a, red));
    float sc = g[co] / nrm;
    for (int i = threadIdx.x; i < row; i += blockDim.x) w[(size_t)co * row + i] = vr[i] * sc;
    if (threadIdx.x == 0 && norm) norm[co] = nrm;
}

// dg = <dW, v> / ||v||;  dv = g/||v|| * (dW - <dW, v> / ||v||^2 * v)
__global__ void wnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v,
                                 const float* __restrict__ dw, float* __restrict__ dg, float* __restrict__ dv,
                                 int row) {
    __shared__ float red[32];
    const int co = blockIdx.x;
    const float* vr = v + (size_t)co * row;
    const float* dr = dw + (size_t)co * row;
    float a = 0.f, d = 0.f;
    for (int i = threadIdx.x; i < row; i += blockDim.x) {
        float vv = vr[i];
        a += vv * vv;
        d += vv * dr[i];
    }
    float n2 = block_sum(a, red);
    float dot = block_sum(d, red);
    float nrm = sqrtf(n2);
    float gg = g[co];
    if (threadIdx.x == 0) dg[co] = dot / nrm;
    float k1 = gg / nrm, k2 = dot / n2;
    for (int i = threadIdx.x; i < row; i += blockDim.x) dv[(size_t)co * row + i] = k1 * (dr[i] - k2 * vr[i]);
}

// ---- all layers of a stack in one launch (the per-layer launches sat on the critical path of every chain) ----
constexpr int kMaxWn = 16;
struct WnSegs {
    const float* g[kMaxWn];
    const float* v[kMaxWn];
    const float* dw[kMaxWn];
    float* w[kMaxWn];      // fwd: w;  bwd: dv
    float* dg[kMaxWn];
    int row0[kMaxWn + 1];  // prefix sum of rows (out channels)
    int cta0[kMaxWn + 1];  // prefix sum of CTAs: short rows (<= 128 elements) go four to a CTA, one per warp
    int rowlen[kMaxWn];
    // optional staged weight images for the tensor-core conv kernels (conv_mma.cu); geo = {Cog, K, S, Tmax,
    // KKpad_f, NS_f, KKpad_d, NS_d}
    float* img_f[kMaxWn];
    float* img_d[kMaxWn];
    int geo[kMaxWn][8];
    int nseg;
};

__device__ __forceinline__ float round_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// One CTA (4 warps) per output channel, or - rows of <= 128 elements: all MPD layers, 5..80 elements - one WARP per
// output channel, four to a CTA, reduced by shuffles without a barrier (the one-CTA-per-5-element-row form took 17 us
// for 0.6 MB at the head of every period-discriminator chain).
template <bool BWD>
__global__ void __launch_bounds__(128) mt_wnorm_kernel(const WnSegs S) {
    __shared__ float red[32];
    int seg = 0;
    while (seg + 1 < S.nseg && (int)blockIdx.x >= S.cta0[seg + 1]) ++seg;
    const int row = S.rowlen[seg];
    const bool per_warp = row <= 128;
    const int nrows = S.row0[seg + 1] - S.row0[seg];
    const int co = per_warp ? ((int)blockIdx.x - S.cta0[seg]) * 4 + ((int)threadIdx.x >> 5) : (int)blockIdx.x - S.cta0[seg];
    if (co >= nrows) return;                       // (per-warp form only: whole warps leave, nobody waits at a barrier)
    const int t0 = per_warp ? ((int)threadIdx.x & 31) : (int)threadIdx.x;
    const int tstep = per_warp ? 32 : (int)blockDim.x;
    const float* vr = S.v[seg] + (size_t)co * row;
    const float gg = S.g[seg][co];
    float a = 0.f, d = 0.f;
    if (BWD) {
        const float* dr = S.dw[seg] + (size_t)co * row;
        for (int i = t0; i < row; i += tstep) {
            float vv = vr[i];
            a += vv * vv;
            d += vv * dr[i];
        }
        float n2, dot;
        if (per_warp) {
            n2 = warp_sum(a);
            dot = warp_sum(d);
            n2 = __shfl_sync(0xffffffffu, n2, 0);
            dot = __shfl_sync(0xffffffffu, dot, 0);
        } else {
            n2 = block_sum(a, red);
            dot = block_sum(d, red);
        }
        const float nrm = sqrtf(n2);
        if (t0 == 0) S.dg[seg][co] = dot / nrm;
        const float k1 = gg / nrm, k2 = dot / n2;
        float* dv = S.w[seg] + (size_t)co * row;
        for (int i = t0; i < row; i += tstep) dv[i] = k1 * (dr[i] - k2 * vr[i]);
    } else {
        for (int i = t0; i < row; i += tstep) a += vr[i] * vr[i];
        float n2;
        if (per_warp) n2 = __shfl_sync(0xffffffffu, warp_sum(a), 0);
        else n2 = block_sum(a, red);
        const float sc = gg / sqrtf(n2);
        float* w = S.w[seg] + (size_t)co * row;
        float* imf = S.img_f[seg];
        float* imd = S.img_d[seg];
        const int Cog = S.geo[seg][0], K = S.geo[seg][1], St = S.geo[seg][2], Tmax = S.geo[seg][3];
        const int KKf = S.geo[seg][4], NSf = S.geo[seg][5], KKd = S.geo[seg][6], NSd = S.geo[seg][7];
        const int g = imf ? co / Cog : 0, col = imf ? co - g * Cog : 0;
        for (int i = t0; i < row; i += tstep) {
            const float v = vr[i] * sc;
            w[i] = v;
            if (imf) {
                // forward image  [g][kk = ci*K + k][n = co within group]
                const float q = round_tf32(v);
                imf[((size_t)g * KKf + i) * NSf + col] = q;
                // data-gradient image  [g][kk = co*Tmax + t][n = ci*S + r],  k = r + S*t
                const int ci = i / K, k = i - ci * K;
                const int t = k / St, r = k - t * St;
                imd[((size_t)g * KKd + col * Tmax + t) * NSd + ci * St + r] = q;
            }
        }
    }
}

int fill_wn(WnSegs& S, const void* const* g, const void* const* v, const void* const* dw, void* const* w,
            void* const* dg, const int64_t* rows, const int64_t* rowlen, int64_t nseg, bool bwd) {
    for (int i = 0; i < kMaxWn; ++i) S.img_f[i] = S.img_d[i] = nullptr;
    if (!g || !v || !w || !rows || !rowlen || nseg <= 0 || nseg > kMaxWn || (bwd && (!dw || !dg))) return LCT_EINVAL;
    int tot = 0, ctas = 0;
    for (int i = 0; i < nseg; ++i) {
        if (!g[i] || !v[i] || !w[i] || rows[i] <= 0 || rowlen[i] <= 0 || (bwd && (!dw[i] || !dg[i]))) return LCT_EINVAL;
        S.g[i] = (const float*)g[i]; S.v[i] = (const float*)v[i]; S.w[i] = (float*)w[i];
        S.dw[i] = bwd ? (const float*)dw[i] : nullptr; S.dg[i] = bwd ? (float*)dg[i] : nullptr;
        S.row0[i] = tot; S.rowlen[i] = (int)rowlen[i];
        S.cta0[i] = ctas;
        tot += (int)rows[i];
        ctas += rowlen[i] <= 128 ? (int)((rows[i] + 3) / 4) : (int)rows[i];
    }
    S.row0[nseg] = tot;
    S.cta0[nseg] = ctas;
    S.nseg = (int)nseg;
    return ctas;
}

}  // namespace

// weight norm of up to 16 layers in one launch; g/v/w: HOST arrays of device pointers, rows/rowlen: HOST arrays
LCT_API int lct_mt_weight_norm_fwd(const void* const* g, const void* const* v, void* const* w, const int64_t* rows,
                                   const int64_t* rowlen, void* const* img_f, void* const* img_d, const int64_t* geo,
                                   int64_t nseg, cudaStream_t st) {
    WnSegs S;
    int tot = fill_wn(S, g, v, nullptr, w, nullptr, rows, rowlen, nseg, false);
    if (tot < 0) return tot;
    if (img_f && img_d && geo) {
        for (int i = 0; i < nseg; ++i) {
            S.img_f[i] = (float*)img_f[i];
            S.img_d[i] = (float*)img_d[i];
            if ((S.img_f[i] == nullptr) != (S.img_d[i] == nullptr)) return LCT_EINVAL;
            for (int j = 0; j < 8; ++j) S.geo[i][j] = (int)geo[i * 8 + j];
        }
    }
    mt_wnorm_kernel<false><<<tot, 128, 0, st>>>(S);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_mt_weight_norm_bwd(const void* const* g, const void* const* v, const void* const* dw, void* const* dg,
                                   void* const* dv, const int64_t* rows, const int64_t* rowlen, int64_t nseg,
                                   cudaStream_t st) {
    WnSegs S;
    int tot = fill_wn(S, g, v, dw, dv, dg, rows, rowlen, nseg, true);
    if (tot < 0) return tot;
    mt_wnorm_kernel<true><<<tot, 128, 0, st>>>(S);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// y[B,Cout,Lout,P] = act(conv_L(x[B,Cin,Lin,P], w[Cout,Cin/G,K]) + bias)
LCT_API int lct_conv1d_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t Cin,
                           int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P,
                           int act, float slope, cudaStream_t st) {
    FwdParams p = {};
    if (!x || !w || !y || !fill_shape(p.s, B, Cin, Cout, G, K, S, pad, Lin, P)) return LCT_EINVAL;
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.act = act; p.slope = slope;
    const int cg = p.s.Cout_g;
    const bool small = (int64_t)p.s.Lout * p.s.P <= 256;   // short rows: one position per thread
    if (cg % 16 == 0) return small ? launch_fwd<16, 1>(p, st) : launch_fwd<16, 4>(p, st);
    if (cg % 4 == 0) return small ? launch_fwd<4, 1>(p, st) : launch_fwd<4, 4>(p, st);
    return small ? launch_fwd<1, 1>(p, st) : launch_fwd<1, 4>(p, st);
}

// dx[B,Cin,Lin,P] = (conv^T(dy, w) + gextra) * act'(xact)      (gextra, xact optional)
LCT_API int lct_conv1d_dgrad(const float* dy, const float* w, float* dx, const float* gextra, const float* xact,
                             int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad,
                             int64_t Lin, int64_t P, int act, float slope, cudaStream_t st) {
    DgradParams p = {};
    if (!dy || !w || !dx || !fill_shape(p.s, B, Cin, Cout, G, K, S, pad, Lin, P)) return LCT_EINVAL;
    p.dy = dy; p.w = w; p.dx = dx; p.gextra = gextra; p.xact = xact; p.act = act; p.slope = slope;
    switch (p.s.S) {
        case 1: return dispatch_dgrad_cit<1>(p, st);
        case 2: return dispatch_dgrad_cit<2>(p, st);
        case 3: return dispatch_dgrad_cit<3>(p, st);
        case 4: return dispatch_dgrad_cit<4>(p, st);
        default: return LCT_EUNSUPPORTED;
    }
}

// dw[Cout,Cin/G,K] += ..., db[Cout] += ...  (accumulating: the caller zeroes them)
LCT_API int lct_conv1d_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t Cin,
                             int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P,
                             cudaStream_t st) {
    WgradParams p = {};
    if (!x || !dy || !dw || !fill_shape(p.s, B, Cin, Cout, G, K, S, pad, Lin, P)) return LCT_EINVAL;
    p.x = x; p.dy = dy; p.dw = dw; p.db = db;
    const int cg = p.s.Cout_g;
    if (cg % 16 == 0) return launch_wgrad<16>(p, st);
    if (cg % 4 == 0) return launch_wgrad<4>(p, st);
    return launch_wgrad<1>(p, st);
}

LCT_API int lct_weight_norm_fwd(const float* g, const float* v, float* w, float* norm, int64_t Cout, int64_t row,
                                cudaStream_t st) {
    if (!g || !v || !w || Cout <= 0 || row <= 0) return LCT_EINVAL;
    int threads = row >= 1024 ? 256 : (row >= 128 ? 128 : 32);
    wnorm_fwd_kernel<<<(unsigned)Cout, threads, 0, st>>>(g, v, w, norm, (int)row);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_weight_norm_bwd(const float* g, const float* v, const float* dw, float* dg, float* dv, int64_t Cout,
                                int64_t row, cudaStream_t st) {
    if (!g || !v || !dw || !dg || !dv || Cout <= 0 || row <= 0) return LCT_EINVAL;
    int threads = row >= 1024 ? 256 : (row >= 128 ? 128 : 32);
    wnorm_bwd_kernel<<<(unsigned)Cout, threads, 0, st>>>(g, v, dw, dg, dv, (int)row);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
