// STFT / iSTFT front end.
//
// Replaces: ComplexSTFT.forward (reference datasets/stft.py:59-88, torch.stft -> cuFFT),
//           ComplexSTFT.istft   (datasets/stft.py:90-132, torch.istft),
//           magnitude/compress/compute_compressed_irm/apply_mask (datasets/stft.py:138-290),
//           TFFeatures.forward  (datasets/tf_features.py:85-146).
//
// Design: these kernels are HBM/latency bound (387 KB per 2 s sample for an STFT-512), so the
// FFT is a shared-memory Stockham autosort (radix 8/4/2/3/5, fp32, double-generated twiddles staged in smem)
// with the real-input "two for one" trick: two real sequences ride in the real and imaginary
// lane of one complex transform.  Framing, reflect padding, windowing, |.|, power-law
// compression, the compressed IRM, mask application, overlap-add and the window-envelope
// normalisation are fused into the prologue/epilogue of the transform, so a spectrogram is
// read or written at most once.
//
// Physical layout of every spectrogram: [B, Tf, F] (frequency innermost), complex interleaved.
// That is the memory layout torch.stft itself produces (it returns a transposed view);
// the Python boundary hands out the same [B, F, Tf] view.
#include "common.cuh"

namespace {

constexpr int kSlots = 4;     // concurrent complex FFTs per CTA
constexpr int kGroup = 64;    // threads cooperating on one FFT
constexpr int kThreads = kSlots * kGroup;
constexpr int kMaxPass = 12;
// FFT buffers are padded by one element per 8 (index i lives at i + i/8): the radix-8 / radix-4 Stockham scatter
// dst[j0 + r*Ns] has stride 8 (or 4) across lanes for the first passes, which without padding is a 16-way bank conflict
#define FIDX(i) ((i) + ((i) >> 3))
__host__ __device__ constexpr int fft_padded(int n) { return n + (n >> 3) + 1; }
constexpr int kMaxN = 2048;

struct FftPlan {
    int n;
    int npass;
    int radix[kMaxPass];
};

bool make_plan(int64_t n, FftPlan* p) {
    if (n < 8 || n > kMaxN || (n & 1)) return false;
    p->n = (int)n;
    p->npass = 0;
    int m = (int)n;
    const int cand[5] = {8, 4, 2, 3, 5};   // power-of-two radices first: Ns is a power of two while they run
    for (int c = 0; c < 5; ++c) {
        while (m % cand[c] == 0) {
            if (p->npass == kMaxPass) return false;
            p->radix[p->npass++] = cand[c];
            m /= cand[c];
        }
    }
    return m == 1;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void dft_small(float2* v);

template <>
__device__ __forceinline__ void dft_small<2>(float2* v) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft_small<4>(float2* v) {
    float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    float2 c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(a, c);
    v[1] = cadd(b, d);
    v[2] = csub(a, c);
    v[3] = csub(b, d);
}
template <>
__device__ __forceinline__ void dft_small<8>(float2* v) {
    // decimation in frequency: X[2m] = DFT4(v[j] + v[j+4]), X[2m+1] = DFT4((v[j] - v[j+4]) * W8^j)
    const float h = 0.70710678118654752440f;
    float2 e[4], o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        e[j] = cadd(v[j], v[j + 4]);
        o[j] = csub(v[j], v[j + 4]);
    }
    o[1] = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * (1 - i)/sqrt2
    o[2] = mul_mi(o[2]);                                                   // * (-i)
    o[3] = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * (-1 - i)/sqrt2
    dft_small<4>(e);
    dft_small<4>(o);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        v[2 * m] = e[m];
        v[2 * m + 1] = o[m];
    }
}
template <>
__device__ __forceinline__ void dft_small<3>(float2* v) {
    const float s = 0.86602540378443864676f;
    float2 t1 = cadd(v[1], v[2]), t2 = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    float2 r = make_float2(s * t2.y, -s * t2.x);   // -i * s * t2
    v[0] = cadd(v[0], t1);
    v[1] = cadd(m, r);
    v[2] = csub(m, r);
}
template <>
__device__ __forceinline__ void dft_small<5>(float2* v) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 p1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    float2 p2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    float2 m1 = mul_mi(q1), m2 = mul_mi(q2);
    v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = cadd(p1, m1);
    v[4] = csub(p1, m1);
    v[2] = cadd(p2, m2);
    v[3] = csub(p2, m2);
}

// One Stockham pass of radix R over a length-N sequence, src -> dst, executed by kGroup threads.
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ src, float2* __restrict__ dst,
                                         const float2* __restrict__ tw, int N, int Ns, int g) {
    const int nb = N / R;
    const int tscale = N / (Ns * R);
    const bool pow2 = (Ns & (Ns - 1)) == 0;
    for (int j = g; j < nb; j += kGroup) {
        const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[FIDX(j + r * nb)];
        if (Ns > 1) {
            const int ts = k * tscale;
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[r * ts]);
        }
        dft_small<R>(v);
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[FIDX(j0 + r * Ns)] = v[r];
    }
}

// Forward complex FFT (sign -) of the sequence in `a` (ping-pong with `b`); returns the buffer
// holding the result.  All threads of the CTA must call it (it uses __syncthreads()).
__device__ __forceinline__ float2* fft_forward(float2* a, float2* b, const float2* __restrict__ tw,
                                               const FftPlan& plan, int g) {
    int Ns = 1;
    for (int p = 0; p < plan.npass; ++p) {
        const int R = plan.radix[p];
        if (R == 8) fft_pass<8>(a, b, tw, plan.n, Ns, g);
        else if (R == 4) fft_pass<4>(a, b, tw, plan.n, Ns, g);
        else if (R == 2) fft_pass<2>(a, b, tw, plan.n, Ns, g);
        else if (R == 3) fft_pass<3>(a, b, tw, plan.n, Ns, g);
        else fft_pass<5>(a, b, tw, plan.n, Ns, g);
        __syncthreads();
        float2* t = a; a = b; b = t;
        Ns *= R;
    }
    return a;
}

// tw: [0, n) W_n^m | [n, n + 64) W_64^{m0 k1} at m0 * 8 + k1 | [n + 64, 2n + 64) W_n^{q j} at q * 64 + j  (stft_warp.cuh)
__global__ void twiddle_kernel(float2* tw, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n + 64) return;
    double num, den = (double)n;
    if (i < n) {
        num = (double)i;
    } else if (i < n + 64) {
        const int e = i - n;
        num = (double)((e >> 3) * (e & 7));
        den = 64.0;
    } else {
        const int e = i - n - 64;
        num = (double)((e >> 6) * (e & 63));
    }
    double s, c;
    sincospi(-2.0 * num / den, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
}

__global__ void envelope_kernel(const float* __restrict__ w, float* __restrict__ env, int N, int hop,
                                int Tf, int Ltot) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= Ltot) return;
    int m_hi = min(t / hop, Tf - 1);
    int m_lo = max(0, (t - N + hop) / hop);   // ceil((t-N+1)/hop) for t-N+1 > 0
    if (t - N + 1 <= 0) m_lo = 0;
    float s = 0.f;
    for (int m = m_lo; m <= m_hi; ++m) {
        float v = w[t - m * hop];
        s += v * v;
    }
    env[t] = s;
}

// ------------------------------------------------------------------------------------
// real frames -> one-sided spectrum
// ------------------------------------------------------------------------------------
// |z| and x^e for the fused epilogues: sqrt(fma) and exp2(e * lg2.approx(x)) (x > 0; relative error ~1e-7, far inside
// the 1e-5 front-end tolerance) instead of hypotf / powf, whose ~100-instruction slow paths dominated these kernels
__device__ __forceinline__ float cabs_fast(float x, float y) { return sqrtf(fmaf(x, x, y * y)); }
__device__ __forceinline__ float pow_fast(float x, float e) { return exp2f(e * __log2f(x)); }

enum { R2C_STFT = 0, R2C_TFF = 1, R2C_ISTFT_BWD = 2, R2C_MRLOSS = 3 };

struct R2CParams {
    const float* a;        // x | noisy | gy | y_hat
    const float* b;        // - | clean | -  | y
    const float* window;   // [N]
    const float2* tw;      // [N]
    const float* env;      // [Ltot] (ISTFT_BWD)
    float2* spec_a;        // STFT: spec; TFF: noisy spec (opt); ISTFT_BWD: gspec (opt)
    float2* spec_b;        // TFF: clean spec (opt)
    float* o0;             // STFT: mag (opt); TFF: noisy_mag
    float* o1;             // TFF: irm_c
    float* o2;             // TFF: noisy_mag_c
    const float2* xspec;   // ISTFT_BWD with mask: the spectrum the mask was applied to
    const float* mask;     // ISTFT_BWD with mask: mask_c
    float* gmask;          // ISTFT_BWD with mask: grad of mask_c
    float* acc;            // MRLOSS: [2][64] partial sums (squared magnitude error, |diff|^2), 64 slots each
    int B, T, N, hop, Tf, F;
    float c, gamma, eps, sc_int, sc_edge;
    FftPlan plan;
};

__device__ __forceinline__ float reflect_load(const float* __restrict__ x, int T, int p) {
    if (p < 0) p = -p;
    if (p >= T) p = 2 * (T - 1) - p;
    return x[p];
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) r2c_kernel(const R2CParams P) {
    extern __shared__ float2 smem[];
    const int tid = threadIdx.x;
    const int slot = tid / kGroup, g = tid % kGroup;
    const int N = P.N, F = P.F, half = N / 2;
    const int NP = fft_padded(N);
    float2* buf0 = smem + (size_t)slot * 2 * NP;
    float2* buf1 = buf0 + NP;
    float2* stw = smem + (size_t)kSlots * 2 * NP;    // twiddle table staged once per CTA
    for (int i = tid; i < N; i += kThreads) stw[i] = __ldg(&P.tw[i]);
    const int b = blockIdx.y;
    constexpr bool kPairSig = (MODE == R2C_TFF || MODE == R2C_MRLOSS);
    int ma, mb;
    if (kPairSig) {
        ma = mb = blockIdx.x * kSlots + slot;
    } else {
        ma = blockIdx.x * 2 * kSlots + 2 * slot;
        mb = ma + 1;
    }
    const bool va = ma < P.Tf, vb = mb < P.Tf;

    // ---- prologue: framing (+ padding) + window
    {
        const float* xa = P.a + (size_t)b * P.T;
        const float* xb = kPairSig ? P.b + (size_t)b * P.T : xa;
        for (int n = g; n < N; n += kGroup) {
            const float w = __ldg(&P.window[n]);
            float fa = 0.f, fb = 0.f;
            if (MODE == R2C_ISTFT_BWD) {
                // adjoint of "slice [N/2, N/2+T) of the OLA buffer, divided by the envelope"
                const int Ltot = N + P.hop * (P.Tf - 1);
                if (va) {
                    int tp = ma * P.hop + n, t = tp - half;
                    if (t >= 0 && t < P.T && tp < Ltot) fa = xa[t] / __ldg(&P.env[tp]);
                }
                if (vb) {
                    int tp = mb * P.hop + n, t = tp - half;
                    if (t >= 0 && t < P.T && tp < Ltot) fb = xa[t] / __ldg(&P.env[tp]);
                }
            } else {
                if (va) fa = reflect_load(xa, P.T, ma * P.hop + n - half);
                if (vb) fb = reflect_load(xb, P.T, mb * P.hop + n - half);
            }
            buf0[FIDX(n)] = make_float2(fa * w, fb * w);
        }
    }
    __syncthreads();
    const float2* Z = fft_forward(buf0, buf1, stw, P.plan, g);

    // ---- epilogue: split the two real transforms, fuse the elementwise consumers
    float acc_mag = 0.f, acc_cplx = 0.f;
    for (int k = g; k <= half; k += kGroup) {
        const float2 zk = Z[FIDX(k)];
        const float2 zn = Z[k == 0 ? 0 : FIDX(N - k)];
        float2 A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        float2 Bv = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        if (MODE == R2C_STFT) {
            if (va) {
                size_t o = ((size_t)b * P.Tf + ma) * F + k;
                P.spec_a[o] = A;
                if (P.o0) P.o0[o] = fmaxf(cabs_fast(A.x, A.y), P.eps);
            }
            if (vb) {
                size_t o = ((size_t)b * P.Tf + mb) * F + k;
                P.spec_a[o] = Bv;
                if (P.o0) P.o0[o] = fmaxf(cabs_fast(Bv.x, Bv.y), P.eps);
            }
        } else if (MODE == R2C_TFF) {
            if (va) {
                size_t o = ((size_t)b * P.Tf + ma) * F + k;
                const float nm = fmaxf(cabs_fast(A.x, A.y), P.eps);
                const float cm = fmaxf(cabs_fast(Bv.x, Bv.y), P.eps);
                const float nmc = pow_fast(nm, P.c);
                P.o0[o] = nm;
                P.o1[o] = pow_fast(cm, P.c) / (nmc + P.gamma);
                P.o2[o] = nmc;
                if (P.spec_a) P.spec_a[o] = A;
                if (P.spec_b) P.spec_b[o] = Bv;
            }
        } else if (MODE == R2C_MRLOSS) {
            if (va) {
                const float ma_ = fmaxf(cabs_fast(A.x, A.y), P.eps);
                const float mb_ = fmaxf(cabs_fast(Bv.x, Bv.y), P.eps);
                const float dm = ma_ - mb_;
                const float dx = A.x - Bv.x, dy = A.y - Bv.y;
                acc_mag += dm * dm;
                acc_cplx += dx * dx + dy * dy;
            }
        } else {  // R2C_ISTFT_BWD: grad of irfft = (c_k / N) * FFT, c_k = 1 at DC/Nyquist else 2
            const float sc = (k == 0 || k == half) ? P.sc_edge : P.sc_int;
            A.x *= sc; A.y *= sc; Bv.x *= sc; Bv.y *= sc;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool v = q ? vb : va;
                if (!v) continue;
                const float2 G = q ? Bv : A;
                size_t o = ((size_t)b * P.Tf + (q ? mb : ma)) * F + k;
                if (P.mask) {
                    // E = X * max(mask, eps)^(1/c)  (apply_mask(compressed=True), stft.py:282-289)
                    const float2 X = P.xspec[o];
                    const float mk = P.mask[o];
                    const float mc = fmaxf(mk, P.eps);
                    const float inv_c = 1.f / P.c;
                    const float lin = pow_fast(mc, inv_c);
                    const float dlin = (mk >= P.eps) ? inv_c * pow_fast(mc, inv_c - 1.f) : 0.f;
                    P.gmask[o] = (X.x * G.x + X.y * G.y) * dlin;
                    if (P.spec_a) P.spec_a[o] = make_float2(G.x * lin, G.y * lin);
                } else {
                    P.spec_a[o] = G;
                }
            }
        }
    }
    if (MODE == R2C_MRLOSS) {
        __shared__ float red[32];
        float s0 = block_sum(acc_mag, red);
        float s1 = block_sum(acc_cplx, red);
        if (tid == 0) {   // 64 accumulator slots per sum: thousands of CTAs on one address serialise in L2
            const int slot_id = (blockIdx.x + blockIdx.y * gridDim.x) & 63;
            atomicAdd(&P.acc[slot_id], s0);
            atomicAdd(&P.acc[64 + slot_id], s1);
        }
    }
}

// ------------------------------------------------------------------------------------
// one-sided spectrum -> windowed real frames -> overlap-add
// ------------------------------------------------------------------------------------
enum { C2R_ISTFT = 0, C2R_STFT_BWD = 1 };

struct C2RParams {
    const float2* spec;    // [B, Tf, F]
    const float* mask;     // optional [B, Tf, F] compressed mask (ISTFT)
    const float* window;
    const float2* tw;
    const float* env;      // [Ltot] (ISTFT)
    float* out;            // ISTFT: y [B, length]; STFT_BWD: padded grad [B, Ltot]
    int B, N, hop, Tf, F, R, FR, length, Ltot;
    float c, eps, sc_int, sc_edge;
    FftPlan plan;
};

__device__ __forceinline__ float2 load_bin(const C2RParams& P, size_t base, int k, int half, bool valid) {
    if (!valid) return make_float2(0.f, 0.f);
    float2 v = P.spec[base + k];
    float sc = (k == 0 || k == half) ? P.sc_edge : P.sc_int;
    if (P.mask) {
        float mk = fmaxf(P.mask[base + k], P.eps);
        sc *= fmaxf(pow_fast(mk, 1.f / P.c), 0.f);
    }
    v.x *= sc;
    v.y *= sc;
    if (k == 0 || k == half) v.y = 0.f;   // c2r ignores the imaginary part of DC / Nyquist
    return v;
}

}  // namespace
#include "stft_warp.cuh"
namespace {

template <int MODE>
__global__ void __launch_bounds__(kThreads) c2r_kernel(const C2RParams P) {
    extern __shared__ float2 smem[];
    const int tid = threadIdx.x;
    const int slot = tid / kGroup, g = tid % kGroup;
    const int N = P.N, F = P.F, half = N / 2;
    constexpr int NF = 2 * kSlots;
    const int NP = fft_padded(N);
    float2* buf0 = smem + (size_t)slot * 2 * NP;
    float2* buf1 = buf0 + NP;
    float* tfr = reinterpret_cast<float*>(smem + (size_t)kSlots * 2 * NP);  // [NF][N]
    float2* stw = reinterpret_cast<float2*>(tfr + (size_t)NF * N);          // twiddle table staged once per CTA
    for (int i = tid; i < N; i += kThreads) stw[i] = __ldg(&P.tw[i]);
    const int b = blockIdx.y;
    const int m0 = blockIdx.x * P.FR;
    const int mfirst = m0 - (P.R - 1);
    const int ma = mfirst + 2 * slot, mb = ma + 1;
    const bool va = ma >= 0 && ma < P.Tf, vb = mb >= 0 && mb < P.Tf;
    const size_t base_a = ((size_t)b * P.Tf + (va ? ma : 0)) * F;
    const size_t base_b = ((size_t)b * P.Tf + (vb ? mb : 0)) * F;

    // Z = P + iQ with P, Q the Hermitian extensions of frames a, b; we transform conj(Z)
    // with the forward FFT and conjugate the result:  ifft(Z) = conj(fft(conj(Z))).
    for (int k = g; k <= half; k += kGroup) {
        const float2 p = load_bin(P, base_a, k, half, va);
        const float2 q = load_bin(P, base_b, k, half, vb);
        // Z[k] = (p.x - q.y, p.y + q.x);  Z[N-k] = (p.x + q.y, -p.y + q.x)
        buf0[FIDX(k)] = make_float2(p.x - q.y, -(p.y + q.x));
        if (k != 0 && k != half) buf0[FIDX(N - k)] = make_float2(p.x + q.y, -(-p.y + q.x));
    }
    __syncthreads();
    const float2* Z = fft_forward(buf0, buf1, stw, P.plan, g);
    for (int n = g; n < N; n += kGroup) {
        const float2 r = Z[FIDX(n)];
        const float w = __ldg(&P.window[n]);
        tfr[(2 * slot) * N + n] = r.x * w;
        tfr[(2 * slot + 1) * N + n] = -r.y * w;
    }
    __syncthreads();

    const bool last = (m0 + P.FR >= P.Tf);
    const int t_begin = m0 * P.hop;
    const int t_end = last ? P.Ltot : (m0 + P.FR) * P.hop;
    for (int tp = t_begin + tid; tp < t_end; tp += kThreads) {
        const int m_hi = min(tp / P.hop, P.Tf - 1);
        int m_lo = (tp - N + 1 <= 0) ? 0 : (tp - N + P.hop) / P.hop;
        float s = 0.f;
        for (int m = m_lo; m <= m_hi; ++m) {
            const int j = m - mfirst;
            if (j >= 0 && j < NF) s += tfr[j * N + (tp - m * P.hop)];
        }
        if (MODE == C2R_ISTFT) {
            const int t = tp - half;
            if (t >= 0 && t < P.length) P.out[(size_t)b * P.length + t] = s / __ldg(&P.env[tp]);
        } else {
            P.out[(size_t)b * P.Ltot + tp] = s;
        }
    }
    if (MODE == C2R_ISTFT && last) {
        // torch.istft zero-fills when `length` runs past the overlap-add buffer
        for (int t = P.Ltot - half + tid; t < P.length; t += kThreads)
            if (t >= 0) P.out[(size_t)b * P.length + t] = 0.f;
    }
}

// adjoint of the reflect padding: fold the padded gradient back onto the signal
__global__ void reflect_fold_kernel(const float* __restrict__ gxp, float* __restrict__ gx, int B, int T,
                                    int half, int Ltot) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int b = blockIdx.y;
    if (t >= T) return;
    const float* gp = gxp + (size_t)b * Ltot;
    float s = 0.f;
    int p = t + half;
    if (p < Ltot) s += gp[p];
    if (t >= 1 && t <= half) {            // left pad: padded index half - t  <->  x[t]
        int q = half - t;
        if (q < Ltot) s += gp[q];
    }
    if (t <= T - 2 && t >= T - 1 - half) {   // right pad: padded index half + T + j  <->  x[T-2-j]
        int q = half + T + (T - 2 - t);
        if (q < Ltot) s += gp[q];
    }
    gx[(size_t)b * T + t] = s;
}

// dL/dY_hat for the multi-resolution STFT loss of one resolution (losses.py:66-80):
//   l = wm * mean((|A|_eps - |B|_eps)^2) + wc * mean(|A - B|^2);  g = upstream * weight / wsum
__global__ void mrloss_grad_kernel(const float2* __restrict__ A, const float2* __restrict__ Bs,
                                   float2* __restrict__ gA, int64_t n, float eps, float km, float kc,
                                   const float* __restrict__ upstream) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float up = upstream ? upstream[0] : 1.f;
    float2 a = A[i], b = Bs[i];
    float ra = hypotf(a.x, a.y), rb = hypotf(b.x, b.y);
    float ma = fmaxf(ra, eps), mb = fmaxf(rb, eps);
    float gm = (ra >= eps && ra > 0.f) ? km * 2.f * (ma - mb) / ra : 0.f;   // d|A|/dA = A/|A|, sgn(0)=0
    float gx = gm * a.x + kc * 2.f * (a.x - b.x);
    float gy = gm * a.y + kc * 2.f * (a.y - b.y);
    gA[i] = make_float2(up * gx, up * gy);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 40 * 1024) {   // static __shared__ of the kernel counts against the 48 KB default too
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

// register-resident warp FFT (stft_warp.cuh) for n_fft = 64 R, R in {5, 8, 12}; g_fft_generic forces the generic kernels
int g_fft_generic = 0;

template <int R, int MODE>
int launch_r2c_warp(R2CParams& P, cudaStream_t st) {
    const bool pair_sig = (MODE == R2C_TFF || MODE == R2C_MRLOSS);
    const int units = pair_sig ? P.Tf : (P.Tf + 1) / 2;
    dim3 grid((unsigned)ceil_div64(units, wf::kR2CWarps), (unsigned)P.B);
    const size_t smem = (size_t)wf::kR2CWarps * wf::R2CSmem<R, MODE>::WARP_FLOATS * sizeof(float);
    int rc = set_smem(wf::r2c_warp_kernel<R, MODE>, smem);
    if (rc) return rc;
    wf::r2c_warp_kernel<R, MODE><<<grid, wf::kR2CWarps * 32, smem, st>>>(P);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int R, int MODE, int WARPS>
int launch_c2r_warp_w(C2RParams& P, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div64(P.Tf + 1, 2 * WARPS - 1), (unsigned)P.B);
    const size_t smem = (size_t)WARPS * wf::Cfg<R>::BUF * sizeof(float2) + (size_t)WARPS * wf::Cfg<R>::HALF * sizeof(float);
    int rc = set_smem(wf::c2r_warp_kernel<R, MODE, WARPS>, smem);
    if (rc) return rc;
    wf::c2r_warp_kernel<R, MODE, WARPS><<<grid, WARPS * 32, smem, st>>>(P);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int R, int MODE>
int launch_c2r_warp(C2RParams& P, cudaStream_t st) {
    P.Ltot = P.N + P.hop * (P.Tf - 1);
    // small problems: 4-warp CTAs (7 segments each) spread over more SMs; large ones: 8 warps (1 redundant frame in 16)
    if ((int64_t)P.B * (P.Tf + 1) < 148 * 15 * 2) return launch_c2r_warp_w<R, MODE, 4>(P, st);
    return launch_c2r_warp_w<R, MODE, 8>(P, st);
}

template <int MODE>
int launch_r2c(R2CParams& P, cudaStream_t st) {
    if (!g_fft_generic) {
        if (P.N == 320) return launch_r2c_warp<5, MODE>(P, st);
        if (P.N == 512) return launch_r2c_warp<8, MODE>(P, st);
        if (P.N == 768) return launch_r2c_warp<12, MODE>(P, st);
    }
    const bool pair_sig = (MODE == R2C_TFF || MODE == R2C_MRLOSS);
    const int per = pair_sig ? kSlots : 2 * kSlots;
    dim3 grid((unsigned)ceil_div64(P.Tf, per), (unsigned)P.B);
    size_t smem = ((size_t)kSlots * 2 * fft_padded(P.N) + P.N) * sizeof(float2);
    int rc = set_smem(r2c_kernel<MODE>, smem);
    if (rc) return rc;
    r2c_kernel<MODE><<<grid, kThreads, smem, st>>>(P);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int MODE>
int launch_c2r(C2RParams& P, cudaStream_t st) {
    if (!g_fft_generic && 2 * P.hop == P.N) {
        if (P.N == 320) return launch_c2r_warp<5, MODE>(P, st);
        if (P.N == 512) return launch_c2r_warp<8, MODE>(P, st);
        if (P.N == 768) return launch_c2r_warp<12, MODE>(P, st);
    }
    P.R = (int)ceil_div64(P.N, P.hop);
    P.FR = 2 * kSlots - (P.R - 1);
    if (P.FR < 1) return LCT_EUNSUPPORTED;   // hop < n_fft / 8
    P.Ltot = P.N + P.hop * (P.Tf - 1);
    dim3 grid((unsigned)ceil_div64(P.Tf, P.FR), (unsigned)P.B);
    size_t smem = ((size_t)kSlots * 2 * fft_padded(P.N) + P.N) * sizeof(float2) + (size_t)2 * kSlots * P.N * sizeof(float);
    int rc = set_smem(c2r_kernel<MODE>, smem);
    if (rc) return rc;
    c2r_kernel<MODE><<<grid, kThreads, smem, st>>>(P);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

bool stft_args_ok(int64_t B, int64_t T, int64_t n_fft, int64_t hop) {
    return B > 0 && T > n_fft / 2 && hop > 0 && hop <= n_fft && B < 65536;
}

}  // namespace

LCT_API int lct_fft_supported(int64_t n_fft) {
    FftPlan p;
    return make_plan(n_fft, &p) ? 1 : 0;
}

// number of complex (float2) entries of the twiddle buffer lct_fft_twiddles fills for n_fft
LCT_API int lct_fft_twiddle_len(int64_t n_fft) { return (int)(2 * n_fft + 64); }

// 1: use the generic shared-memory Stockham kernels for every size (bring-up / A-B measurements); 0 (default): the
// register-resident warp FFT for n_fft in {320, 512, 768}
LCT_API int lct_fft_force_generic(int on) {
    g_fft_generic = on ? 1 : 0;
    return 0;
}

LCT_API int lct_fft_twiddles(float* tw, int64_t n_fft, cudaStream_t stream) {
    FftPlan p;
    if (!tw || !make_plan(n_fft, &p)) return LCT_EINVAL;
    twiddle_kernel<<<(unsigned)ceil_div64(2 * n_fft + 64, 128), 128, 0, stream>>>(reinterpret_cast<float2*>(tw), (int)n_fft);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_ola_envelope(const float* window, float* env, int64_t n_fft, int64_t hop, int64_t n_frames,
                             cudaStream_t stream) {
    if (!window || !env || n_fft <= 0 || hop <= 0 || n_frames <= 0) return LCT_EINVAL;
    int64_t ltot = n_fft + hop * (n_frames - 1);
    envelope_kernel<<<(unsigned)ceil_div64(ltot, 256), 256, 0, stream>>>(window, env, (int)n_fft, (int)hop,
                                                                          (int)n_frames, (int)ltot);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_stft_fwd(const float* x, const float* window, const float* tw, float* spec, float* mag,
                         int64_t B, int64_t T, int64_t n_fft, int64_t hop, float eps, cudaStream_t stream) {
    R2CParams P = {};
    if (!x || !window || !tw || !spec || !stft_args_ok(B, T, n_fft, hop) || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    P.a = x; P.window = window; P.tw = reinterpret_cast<const float2*>(tw);
    P.spec_a = reinterpret_cast<float2*>(spec); P.o0 = mag;
    P.B = (int)B; P.T = (int)T; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)(1 + T / hop); P.F = (int)(n_fft / 2 + 1);
    P.eps = eps;
    return launch_r2c<R2C_STFT>(P, stream);
}

LCT_API int lct_tf_features_fwd(const float* noisy, const float* clean, const float* window, const float* tw,
                                float* noisy_mag, float* irm_c, float* noisy_mag_c, float* noisy_spec,
                                float* clean_spec, int64_t B, int64_t T, int64_t n_fft, int64_t hop, float c,
                                float gamma, float eps, cudaStream_t stream) {
    R2CParams P = {};
    if (!noisy || !clean || !window || !tw || !noisy_mag || !irm_c || !noisy_mag_c ||
        !stft_args_ok(B, T, n_fft, hop) || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    P.a = noisy; P.b = clean; P.window = window; P.tw = reinterpret_cast<const float2*>(tw);
    P.o0 = noisy_mag; P.o1 = irm_c; P.o2 = noisy_mag_c;
    P.spec_a = reinterpret_cast<float2*>(noisy_spec); P.spec_b = reinterpret_cast<float2*>(clean_spec);
    P.B = (int)B; P.T = (int)T; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)(1 + T / hop); P.F = (int)(n_fft / 2 + 1);
    P.c = c; P.gamma = gamma; P.eps = eps;
    return launch_r2c<R2C_TFF>(P, stream);
}

// acc[0..63] += partial sums of (|A|_eps - |B|_eps)^2, acc[64..127] += partial sums of |A - B|^2 (the caller adds the
// 64 slots of each) over the STFTs of y_hat and y;
// no spectrogram is written.
LCT_API int lct_mrstft_sums(const float* y_hat, const float* y, const float* window, const float* tw, float* acc,
                            int64_t B, int64_t T, int64_t n_fft, int64_t hop, float eps, cudaStream_t stream) {
    R2CParams P = {};
    if (!y_hat || !y || !window || !tw || !acc || !stft_args_ok(B, T, n_fft, hop) || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    P.a = y_hat; P.b = y; P.window = window; P.tw = reinterpret_cast<const float2*>(tw); P.acc = acc;
    P.B = (int)B; P.T = (int)T; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)(1 + T / hop); P.F = (int)(n_fft / 2 + 1);
    P.eps = eps;
    return launch_r2c<R2C_MRLOSS>(P, stream);
}

LCT_API int lct_mrstft_grad_spec(const float* spec_hat, const float* spec_ref, float* gspec, int64_t n_bins,
                                 float eps, float k_mag, float k_cplx, const float* upstream,
                                 cudaStream_t stream) {
    if (!spec_hat || !spec_ref || !gspec || n_bins <= 0) return LCT_EINVAL;
    mrloss_grad_kernel<<<(unsigned)ceil_div64(n_bins, 256), 256, 0, stream>>>(
        reinterpret_cast<const float2*>(spec_hat), reinterpret_cast<const float2*>(spec_ref),
        reinterpret_cast<float2*>(gspec), n_bins, eps, k_mag, k_cplx, upstream);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// gspec [B,Tf,F,2] -> gx [B,T];  work: [B, n_fft + hop*(Tf-1)] floats
LCT_API int lct_stft_bwd(const float* gspec, const float* window, const float* tw, float* work, float* gx,
                         int64_t B, int64_t T, int64_t n_fft, int64_t hop, cudaStream_t stream) {
    C2RParams P = {};
    if (!gspec || !window || !tw || !work || !gx || !stft_args_ok(B, T, n_fft, hop) || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    P.spec = reinterpret_cast<const float2*>(gspec); P.window = window; P.tw = reinterpret_cast<const float2*>(tw);
    P.out = work;
    P.B = (int)B; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)(1 + T / hop); P.F = (int)(n_fft / 2 + 1);
    P.sc_int = 0.5f; P.sc_edge = 1.f; P.eps = 0.f; P.c = 1.f;
    int rc = launch_c2r<C2R_STFT_BWD>(P, stream);
    if (rc) return rc;
    dim3 grid((unsigned)ceil_div64(T, 256), (unsigned)B);
    reflect_fold_kernel<<<grid, 256, 0, stream>>>(work, gx, (int)B, (int)T, (int)(n_fft / 2), P.Ltot);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// spec [B,Tf,F,2] (x optional decompressed mask) -> y [B,length]; env from lct_ola_envelope
LCT_API int lct_istft_fwd(const float* spec, const float* mask_c, const float* window, const float* tw,
                          const float* env, float* y, int64_t B, int64_t n_frames, int64_t n_fft, int64_t hop,
                          int64_t length, float c, float eps, cudaStream_t stream) {
    C2RParams P = {};
    if (!spec || !window || !tw || !env || !y || B <= 0 || B >= 65536 || n_frames <= 0 || hop <= 0 ||
        hop > n_fft || length <= 0 || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    P.spec = reinterpret_cast<const float2*>(spec); P.mask = mask_c; P.window = window;
    P.tw = reinterpret_cast<const float2*>(tw); P.env = env; P.out = y;
    P.B = (int)B; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)n_frames; P.F = (int)(n_fft / 2 + 1);
    P.length = (int)length; P.c = c; P.eps = eps;
    P.sc_int = 1.f / (float)n_fft; P.sc_edge = 1.f / (float)n_fft;
    return launch_c2r<C2R_ISTFT>(P, stream);
}

// gy [B,length] -> gspec [B,Tf,F,2] (grad of the spectrum handed to istft).  With mask_c/xspec
// given, the spectrum was xspec * max(mask_c,eps)^(1/c): gmask receives dL/dmask_c and gspec
// (optional) dL/dxspec.
LCT_API int lct_istft_bwd(const float* gy, const float* window, const float* tw, const float* env, float* gspec,
                          const float* xspec, const float* mask_c, float* gmask, int64_t B, int64_t n_frames,
                          int64_t n_fft, int64_t hop, int64_t length, float c, float eps, cudaStream_t stream) {
    R2CParams P = {};
    if (!gy || !window || !tw || !env || B <= 0 || B >= 65536 || n_frames <= 0 || hop <= 0 || hop > n_fft ||
        length <= 0 || !make_plan(n_fft, &P.plan))
        return LCT_EINVAL;
    if (mask_c ? (!xspec || !gmask) : !gspec) return LCT_EINVAL;
    P.a = gy; P.window = window; P.tw = reinterpret_cast<const float2*>(tw); P.env = env;
    P.spec_a = reinterpret_cast<float2*>(gspec); P.xspec = reinterpret_cast<const float2*>(xspec);
    P.mask = mask_c; P.gmask = gmask;
    P.B = (int)B; P.T = (int)length; P.N = (int)n_fft; P.hop = (int)hop; P.Tf = (int)n_frames; P.F = (int)(n_fft / 2 + 1);
    P.c = c; P.eps = eps;
    P.sc_int = 2.f / (float)n_fft; P.sc_edge = 1.f / (float)n_fft;
    return launch_r2c<R2C_ISTFT_BWD>(P, stream);
}
