// Register-resident warp FFT for the STFT / iSTFT front end (included by stft.cu).
//
// One warp transforms one complex sequence of N = 64 R points (R = 5, 8, 12: n_fft = 320, 512, 768, the three
// resolutions the reference uses, losses.py:11-19 / datasets/stft.py:10-34), i.e. two real frames with the
// "two for one" trick.  The transform is three passes 8 x 8 x R; every thread keeps its butterflies in registers
// and the warp only meets in shared memory twice (2 x N float2 written + read, bank-conflict free, __syncwarp
// only - no CTA barrier).  Framing/windowing is fused into the loads of the first pass and the spectrum
// consumers into the last pass, straight from registers:
//
//   n = q + R (m0 + 8 m1)       k = (k1 + 8 k2) + 64 r                                   W_M = exp(-2 pi i / M)
//   X[k] = sum_q W_R^{q r} W_N^{q j} sum_{m0} W_8^{m0 k2} W_64^{m0 k1} sum_{m1} W_8^{m1 k1} x[n],   j = k1 + 8 k2
//
//   pass 1: thread <-> id = R m0 + q = lane + 32 u   radix-8 over m1 (inputs x[id + 8 R m1]: coalesced loads)
//   pass 2: thread <-> (q, k1) = lane + 32 u         twiddle W_64^{m0 k1}, radix-8 over m0
//   pass 3: thread <-> j = lane and j' = 64 - lane   twiddle W_N^{q j},  radix-R over q  ->  X[j + 64 r]
//
// Bins k and N - k are produced by butterflies j and 64 - j, which pass 3 gives to the SAME thread (lane 0 takes the
// two self-paired butterflies 0 and 32), so the split of the two real spectra is thread local and each lane writes
// consecutive bins (coalesced).  The inverse direction runs the transposed flow graph (radix-R first, natural order
// last) on conj(Z), so the Hermitian extension is thread local as well and the frames come out in coalesced order.
//
// Shared-memory layouts (float2 units, found by exhaustive search, conflict free for the writer and the reader):
//   exchange 1: id + (8 R + 2) k1        exchange 2: 72 q + j
#pragma once

namespace wf {

template <int R> struct Cfg {
    static constexpr int N = 64 * R;
    static constexpr int HALF = 32 * R;
    static constexpr int NB = 8 * R;                  // radix-8 butterflies per pass
    static constexpr int U = (NB + 31) / 32;          // ... per thread
    static constexpr bool FULL = (NB % 32) == 0;      // every (lane, u) slot holds a butterfly
    static constexpr int S1 = 8 * R + 2;
    static constexpr int BUF = 72 * R - 8;            // float2 per warp (covers both exchanges)
    static constexpr int NA = (R % 2 == 0) ? R / 2 : (R + 1) / 2;   // bins j + 64 r <= N/2 of butterfly j = lane (lane >= 1)
    static constexpr int NBN = (R % 2 == 0) ? R / 2 : (R - 1) / 2;  // bins of butterfly 64 - lane
    // lane 0 (butterflies 0 and 32) owns one more bin: the Nyquist bin (A side for even R, B side for odd R)
    static constexpr bool XTRA_A = (R % 2 == 0);
};

// twiddle buffer layout (float2): [0, N) W_N^m | [N, N + 64) T2[m0][k1] = W_64^{m0 k1} | [N + 64, 2N + 64) T3[q][j] = W_N^{q j}
__host__ __device__ constexpr int tw_len(int n) { return 2 * n + 64; }

template <int R> __device__ __forceinline__ void dft_r(float2* v);
template <> __device__ __forceinline__ void dft_r<5>(float2* v) { dft_small<5>(v); }
template <> __device__ __forceinline__ void dft_r<8>(float2* v) { dft_small<8>(v); }
template <> __device__ __forceinline__ void dft_r<12>(float2* v) {
    // 12 = 3 x 4:  X[j + 4 r] = sum_q W_3^{q r} W_12^{q j} sum_m x[q + 3 m] W_4^{m j}
    const float c = 0.86602540378443864676f;
    float2 y[3][4];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int m = 0; m < 4; ++m) y[q][m] = v[q + 3 * m];
        dft_small<4>(y[q]);
    }
    // W_12^{q j}: q = 1: j = 1, 2, 3 -> W^1, W^2, W^3;  q = 2: j = 1, 2, 3 -> W^2, W^4, W^6
    y[1][1] = cmul(y[1][1], make_float2(c, -0.5f));
    y[1][2] = cmul(y[1][2], make_float2(0.5f, -c));
    y[1][3] = mul_mi(y[1][3]);
    y[2][1] = cmul(y[2][1], make_float2(0.5f, -c));
    y[2][2] = cmul(y[2][2], make_float2(-0.5f, -c));
    y[2][3] = make_float2(-y[2][3].x, -y[2][3].y);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 t[3] = {y[0][j], y[1][j], y[2][j]};
        dft_small<3>(t);
#pragma unroll
        for (int r = 0; r < 3; ++r) v[j + 4 * r] = t[r];
    }
}

// v[u][m1] = x[lane + 32 u + 8 R m1]  ->  XA[r] = X[jA + 64 r], XB[r] = X[jB + 64 r]   (jA = lane, jB = lane ? 64 - lane : 32)
template <int R>
__device__ __forceinline__ void fft_dit(float2 (&v)[Cfg<R>::U][8], float2 (&XA)[R], float2 (&XB)[R], float2* buf,
                                        const float2* __restrict__ tw, int lane) {
    using C = Cfg<R>;
    const float2* T2 = tw + C::N;
    const float2* T3 = tw + C::N + 64;
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
            dft_small<8>(v[u]);
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) buf[lane + 32 * u + C::S1 * k1] = v[u][k1];
        }
    }
    __syncwarp();
    const int k1 = lane & 7;
    float2 t2[8];
#pragma unroll
    for (int m0 = 1; m0 < 8; ++m0) t2[m0] = __ldg(&T2[m0 * 8 + k1]);
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
            const int q = (lane + 32 * u) >> 3;
#pragma unroll
            for (int m0 = 0; m0 < 8; ++m0) v[u][m0] = buf[q + C::S1 * k1 + R * m0];
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
            const int q = (lane + 32 * u) >> 3;
#pragma unroll
            for (int m0 = 1; m0 < 8; ++m0) v[u][m0] = cmul(v[u][m0], t2[m0]);
            dft_small<8>(v[u]);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) buf[72 * q + k1 + 8 * k2] = v[u][k2];
        }
    }
    __syncwarp();
    const int jA = lane, jB = lane ? 64 - lane : 32;
#pragma unroll
    for (int q = 0; q < R; ++q) {
        XA[q] = buf[72 * q + jA];
        XB[q] = buf[72 * q + jB];
    }
#pragma unroll
    for (int q = 1; q < R; ++q) {
        XA[q] = cmul(XA[q], __ldg(&T3[64 * q + jA]));
        XB[q] = cmul(XB[q], __ldg(&T3[64 * q + jB]));
    }
    dft_r<R>(XA);
    dft_r<R>(XB);
    __syncwarp();     // buf may be rewritten by the caller's next transform
}

// transposed flow graph: UA[r] = u[jA + 64 r], UB[r] = u[jB + 64 r]  ->  v[u][m1] = FFT(u)[lane + 32 u + 8 R m1]
template <int R>
__device__ __forceinline__ void fft_dif(float2 (&UA)[R], float2 (&UB)[R], float2 (&v)[Cfg<R>::U][8], float2* buf,
                                        const float2* __restrict__ tw, int lane) {
    using C = Cfg<R>;
    const float2* T2 = tw + C::N;
    const float2* T3 = tw + C::N + 64;
    const int jA = lane, jB = lane ? 64 - lane : 32;
    dft_r<R>(UA);
    dft_r<R>(UB);
#pragma unroll
    for (int q = 1; q < R; ++q) {
        UA[q] = cmul(UA[q], __ldg(&T3[64 * q + jA]));
        UB[q] = cmul(UB[q], __ldg(&T3[64 * q + jB]));
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
        buf[72 * q + jA] = UA[q];
        buf[72 * q + jB] = UB[q];
    }
    __syncwarp();
    const int k1 = lane & 7;
    float2 t2[8];
#pragma unroll
    for (int m0 = 1; m0 < 8; ++m0) t2[m0] = __ldg(&T2[m0 * 8 + k1]);
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
            const int q = (lane + 32 * u) >> 3;
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) v[u][k2] = buf[72 * q + k1 + 8 * k2];
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
            const int q = (lane + 32 * u) >> 3;
            dft_small<8>(v[u]);
#pragma unroll
            for (int m0 = 1; m0 < 8; ++m0) v[u][m0] = cmul(v[u][m0], t2[m0]);
#pragma unroll
            for (int m0 = 0; m0 < 8; ++m0) buf[q + C::S1 * k1 + R * m0] = v[u][m0];
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < C::U; ++u) {
        if (C::FULL || lane + 32 * u < C::NB) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) v[u][kk] = buf[lane + 32 * u + C::S1 * kk];
            dft_small<8>(v[u]);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// real frames -> one-sided spectra (STFT, TFFeatures, iSTFT adjoint, MR-STFT loss sums)
// ------------------------------------------------------------------------------------------------
constexpr int kR2CWarps = 4;

__device__ __forceinline__ float sqrt_approx(float x) {     // MUFU.SQRT, <= 2 ulp: far inside the 1e-5 front-end tolerance
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cabs_w(float x, float y) { return sqrt_approx(fmaf(x, x, y * y)); }
// raw MUFU.EX2 / MUFU.LG2 (no denormal fix-up code: every argument here is >= eps = 1e-12 or a moderate exponent)
__device__ __forceinline__ float ex2_w(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_w(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float pow_w(float x, float e) { return ex2_w(e * lg2_w(x)); }

__device__ __forceinline__ void cp_async16_w(float* dst, const float* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_w() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_w() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Warp-level async copy of the floats g[0, n) into st[ph, ph + n), ph = (address of g mod 16) / 4, as 16-byte cp.async
// over the enclosing aligned chunks.  Returns false (nothing issued) when those chunks would leave [lo, hi).
__device__ __forceinline__ bool stage_copy(float* st, const float* g, int n, const float* lo, const float* hi, int lane) {
    const int ph = (int)(((uintptr_t)g & 15) >> 2);
    const float* g0 = g - ph;
    const int nch = (ph + n + 3) >> 2;
    if (g0 < lo || g0 + 4 * nch > hi) return false;
    for (int c = lane; c < nch; c += 32) cp_async16_w(st + 4 * c, g0 + 4 * c);
    return true;
}
__device__ __forceinline__ int stage_phase(const void* g) { return (int)(((uintptr_t)g & 15) >> 2); }

// per-warp shared memory (floats) of the r2c kernel: exchange buffer + (iSTFT adjoint) the staged X / mask rows of the
// warp's two frames, fetched asynchronously at kernel start and consumed by the epilogue
template <int R, int MODE> struct R2CSmem {
    static constexpr int F = Cfg<R>::HALF + 1;
    static constexpr int XS = 4 * F + 8, MK = 2 * F + 10;
    static constexpr int STAGE = (MODE == R2C_ISTFT_BWD) ? XS + MK : 0;
    static constexpr int WARP_FLOATS = 2 * Cfg<R>::BUF + STAGE;
};

template <int R, int MODE>
__global__ void __launch_bounds__(kR2CWarps * 32) r2c_warp_kernel(const R2CParams P) {
    using C = Cfg<R>;
    using SM = R2CSmem<R, MODE>;
    constexpr int N = C::N, HALF = C::HALF, F = HALF + 1;
    extern __shared__ __align__(16) float2 smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = reinterpret_cast<float*>(smem) + (size_t)warp * SM::WARP_FLOATS;
    float2* buf = reinterpret_cast<float2*>(wbase);
    float* st = wbase + 2 * C::BUF;
    const int b = blockIdx.y;
    constexpr bool kPairSig = (MODE == R2C_TFF || MODE == R2C_MRLOSS);
    const int unit = blockIdx.x * kR2CWarps + warp;
    const int ma = kPairSig ? unit : 2 * unit;
    const int mb = kPairSig ? unit : ma + 1;
    if (ma >= P.Tf) return;                      // warp uniform; no CTA barrier below
    const bool vb = mb < P.Tf;
    bool staged = false;
    if (MODE == R2C_ISTFT_BWD && P.mask && vb) {
        // the epilogue's X and mask rows (frames ma, ma + 1: contiguous) start travelling now, behind the transform
        const size_t row = ((size_t)b * P.Tf + ma) * F, tot = (size_t)P.B * P.Tf * F;
        const float* xs = reinterpret_cast<const float*>(P.xspec);
        staged = stage_copy(st, xs + 2 * row, 4 * F, xs, xs + 2 * tot, lane) &&
                 stage_copy(st + SM::XS, P.mask + row, 2 * F, P.mask, P.mask + tot, lane);
        cp_async_commit_w();
    }
    const float* xa = P.a + (size_t)b * P.T;
    const float* xb = kPairSig ? P.b + (size_t)b * P.T : xa;

    // ---- framing + window (0.5 of the two-for-one split folded in: exact)
    float2 v[C::U][8];
    {
        const int sa = ma * P.hop - HALF, sb = mb * P.hop - HALF;
        const bool fast = sa >= 0 && sa + N <= P.T && (!vb || (sb >= 0 && sb + N <= P.T));
        float fa[C::U][8], fb[C::U][8];
        if (MODE == R2C_ISTFT_BWD && vb && P.hop == HALF && sa >= 0 && sa + 3 * HALF <= P.T) {
            // interior pair of frames (every sample inside [0, T), hence inside the overlap-add buffer): 3 half frames of
            // gy / envelope, all loads issued before the first division
            const float* pg = xa + sa + lane;
            const float* pe = P.env + ma * P.hop + lane;
#pragma unroll
            for (int u = 0; u < C::U; ++u) {
                const bool ok = C::FULL || lane + 32 * u < C::NB;
                float g[12], e[12];
#pragma unroll
                for (int m1 = 0; m1 < 12; ++m1) {
                    g[m1] = ok ? pg[32 * u + C::NB * m1] : 0.f;
                    e[m1] = ok ? __ldg(pe + 32 * u + C::NB * m1) : 1.f;
                }
#pragma unroll
                for (int m1 = 0; m1 < 12; ++m1) g[m1] = __fdividef(g[m1], e[m1]);
#pragma unroll
                for (int m1 = 0; m1 < 8; ++m1) {
                    fa[u][m1] = g[m1];
                    fb[u][m1] = g[m1 + 4];
                }
            }
        } else if (MODE == R2C_ISTFT_BWD) {
            // adjoint of "slice [N/2, N/2 + T) of the overlap-add buffer, divided by the envelope"
            const int Ltot = N + P.hop * (P.Tf - 1);
#pragma unroll
            for (int u = 0; u < C::U; ++u)
#pragma unroll
                for (int m1 = 0; m1 < 8; ++m1) {
                    const int n = lane + 32 * u + C::NB * m1;
                    fa[u][m1] = fb[u][m1] = 0.f;
                    if (C::FULL || lane + 32 * u < C::NB) {
                        const int tpa = ma * P.hop + n, ta = tpa - HALF;
                        if (ta >= 0 && ta < P.T && tpa < Ltot) fa[u][m1] = __fdividef(xa[ta], __ldg(&P.env[tpa]));
                        const int tpb = mb * P.hop + n, tb = tpb - HALF;
                        if (vb && tb >= 0 && tb < P.T && tpb < Ltot) fb[u][m1] = __fdividef(xa[tb], __ldg(&P.env[tpb]));
                    }
                }
        } else if (fast) {
            const float* pa = xa + sa + lane;
            const float* pb = xb + sb + lane;
            if (!kPairSig && vb && P.hop == HALF) {
                // consecutive frames of one signal overlap by half: frame b = [second half of frame a | HALF new samples]
#pragma unroll
                for (int u = 0; u < C::U; ++u) {
                    const bool ok = C::FULL || lane + 32 * u < C::NB;
#pragma unroll
                    for (int m1 = 0; m1 < 8; ++m1) fa[u][m1] = ok ? pa[32 * u + C::NB * m1] : 0.f;
#pragma unroll
                    for (int m1 = 0; m1 < 4; ++m1) {
                        fb[u][m1] = fa[u][m1 + 4];
                        fb[u][m1 + 4] = ok ? pa[32 * u + C::NB * (m1 + 8)] : 0.f;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < C::U; ++u) {
                    const bool ok = C::FULL || lane + 32 * u < C::NB;
#pragma unroll
                    for (int m1 = 0; m1 < 8; ++m1) {
                        fa[u][m1] = ok ? pa[32 * u + C::NB * m1] : 0.f;
                        fb[u][m1] = (ok && vb) ? pb[32 * u + C::NB * m1] : 0.f;
                    }
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < C::U; ++u)
#pragma unroll
                for (int m1 = 0; m1 < 8; ++m1) {
                    const int n = lane + 32 * u + C::NB * m1;
                    const bool ok = C::FULL || lane + 32 * u < C::NB;
                    fa[u][m1] = ok ? reflect_load(xa, P.T, sa + n) : 0.f;
                    fb[u][m1] = (ok && vb) ? reflect_load(xb, P.T, sb + n) : 0.f;
                }
        }
        const float* pw = P.window + lane;
#pragma unroll
        for (int u = 0; u < C::U; ++u)
#pragma unroll
            for (int m1 = 0; m1 < 8; ++m1) {
                const float w = (C::FULL || lane + 32 * u < C::NB) ? 0.5f * __ldg(pw + 32 * u + C::NB * m1) : 0.f;
                v[u][m1] = make_float2(fa[u][m1] * w, fb[u][m1] * w);
            }
    }
    float2 XA[R], XB[R];
    fft_dit<R>(v, XA, XB, buf, P.tw, lane);

    // ---- split the two real spectra (thread local) and feed the fused consumers
    const int jA = lane, jB = lane ? 64 - lane : 32;
    const size_t row_a = ((size_t)b * P.Tf + ma) * F;
    const size_t dfb = kPairSig ? 0 : F;          // frame b sits one row further (mb = ma + 1)
    float acc_mag = 0.f, acc_cplx = 0.f;
    const bool l0 = (lane == 0);
    const float2* sx = nullptr;
    const float* smk = nullptr;
    if (MODE == R2C_ISTFT_BWD) {
        cp_async_wait_w();
        __syncwarp();
        sx = reinterpret_cast<const float2*>(st + stage_phase(P.xspec + row_a));
        smk = st + SM::XS + stage_phase(P.mask + row_a);
    }
    // o: element offset of the bin in frame a's row;  sc: iSTFT-adjoint scale of the bin
    auto emit = [&](const size_t o, float2 zk, float2 zn, float sc) {
        float2 A = make_float2(zk.x + zn.x, zk.y - zn.y);
        float2 Bv = make_float2(zk.y + zn.y, zn.x - zk.x);
        if (MODE == R2C_STFT) {
            P.spec_a[o] = A;
            if (P.o0) P.o0[o] = fmaxf(cabs_w(A.x, A.y), P.eps);
            if (vb) {
                P.spec_a[o + dfb] = Bv;
                if (P.o0) P.o0[o + dfb] = fmaxf(cabs_w(Bv.x, Bv.y), P.eps);
            }
        } else if (MODE == R2C_TFF) {
            const float nm = fmaxf(cabs_w(A.x, A.y), P.eps);
            const float cm = fmaxf(cabs_w(Bv.x, Bv.y), P.eps);
            const float nmc = pow_w(nm, P.c);
            P.o0[o] = nm;
            P.o1[o] = __fdividef(pow_w(cm, P.c), nmc + P.gamma);
            P.o2[o] = nmc;
            if (P.spec_a) P.spec_a[o] = A;
            if (P.spec_b) P.spec_b[o] = Bv;
        } else if (MODE == R2C_MRLOSS) {
            const float ma_ = fmaxf(cabs_w(A.x, A.y), P.eps);
            const float mb_ = fmaxf(cabs_w(Bv.x, Bv.y), P.eps);
            const float dm = ma_ - mb_;
            const float dx = A.x - Bv.x, dy = A.y - Bv.y;
            acc_mag = fmaf(dm, dm, acc_mag);
            acc_cplx = fmaf(dx, dx, fmaf(dy, dy, acc_cplx));
        } else {   // R2C_ISTFT_BWD: grad of irfft = (c_k / N) * FFT, c_k = 1 at DC / Nyquist else 2
            A.x *= sc; A.y *= sc; Bv.x *= sc; Bv.y *= sc;
#pragma unroll
            for (int qf = 0; qf < 2; ++qf) {
                if (qf && !vb) continue;
                const float2 G = qf ? Bv : A;
                const size_t oo = qf ? o + dfb : o;
                if (P.mask) {
                    // E = X * max(mask, eps)^(1/c)  (apply_mask(compressed=True), stft.py:282-289)
                    const float2 X = staged ? sx[oo - row_a] : P.xspec[oo];
                    const float mk = staged ? smk[oo - row_a] : P.mask[oo];
                    const float mc = fmaxf(mk, P.eps);
                    const float inv_c = 1.f / P.c;
                    const float lg = lg2_w(mc);
                    const float dlin = (mk >= P.eps) ? inv_c * ex2_w((inv_c - 1.f) * lg) : 0.f;
                    P.gmask[oo] = (X.x * G.x + X.y * G.y) * dlin;
                    if (P.spec_a) {
                        const float lin = ex2_w(inv_c * lg);
                        P.spec_a[oo] = make_float2(G.x * lin, G.y * lin);
                    }
                } else {
                    P.spec_a[oo] = G;
                }
            }
        }
    };
    const size_t oA = row_a + (size_t)jA, oB = row_a + (size_t)jB;
#pragma unroll
    for (int r = 0; r < C::NA; ++r) {
        // partner of bin jA + 64 r: butterfly 64 - lane, slot R - 1 - r  (lane 0: own butterfly, slot (R - r) % R)
        const float2 p0 = XA[(R - r) % R], p1 = XB[R - 1 - r];
        emit(oA + 64 * r, XA[r], make_float2(l0 ? p0.x : p1.x, l0 ? p0.y : p1.y),
             (r == 0 && l0) ? P.sc_edge : P.sc_int);
    }
#pragma unroll
    for (int r = 0; r < C::NBN; ++r) {
        const float2 p0 = XB[R - 1 - r], p1 = XA[R - 1 - r];
        emit(oB + 64 * r, XB[r], make_float2(l0 ? p0.x : p1.x, l0 ? p0.y : p1.y), P.sc_int);
    }
    if (l0) {   // the Nyquist bin (self paired)
        if (C::XTRA_A) emit(oA + 64 * C::NA, XA[C::NA], XA[(R - C::NA) % R], P.sc_edge);
        else emit(oB + 64 * C::NBN, XB[C::NBN], XB[R - 1 - C::NBN], P.sc_edge);
    }
    if (MODE == R2C_MRLOSS) {
        const float s0 = warp_sum(acc_mag), s1 = warp_sum(acc_cplx);
        if (l0) {   // 64 accumulator slots per sum: thousands of warps on one address would serialise in L2
            const int slot_id = (unit + b * 7) & 63;
            atomicAdd(&P.acc[slot_id], s0);
            atomicAdd(&P.acc[64 + slot_id], s1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// one-sided spectra -> windowed real frames -> overlap-add   (hop = N / 2)
// ------------------------------------------------------------------------------------------------
// scale, mask and combine the raw bins p (frame a) and q (frame b) into conj Z[k] and conj Z[N - k]
__device__ __forceinline__ void c2r_combine(float2 p, float2 q, float mka, float mkb, bool has_mask, float eps, float inv_c,
                                            float sca, float scb, bool edge, float2& d, float2& pn) {
    if (has_mask) {
        sca *= pow_w(fmaxf(mka, eps), inv_c);
        scb *= pow_w(fmaxf(mkb, eps), inv_c);
    }
    p.x *= sca; p.y *= sca;
    q.x *= scb; q.y *= scb;
    if (edge) { p.y = 0.f; q.y = 0.f; }      // c2r ignores the imaginary part of DC / Nyquist
    // Z = P + iQ (Hermitian extensions); we transform conj(Z) forward and conjugate the result
    d = make_float2(p.x - q.y, -(p.y + q.x));     // conj Z[k]
    pn = make_float2(p.x + q.y, p.y - q.x);       // conj Z[N - k]
}

template <int R, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) c2r_warp_kernel(const C2RParams P) {
    using C = Cfg<R>;
    constexpr int N = C::N, HALF = C::HALF, F = HALF + 1;
    extern __shared__ float2 smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = smem + (size_t)warp * C::BUF;
    float* tails = reinterpret_cast<float*>(smem + (size_t)WARPS * C::BUF);     // [WARPS][HALF]
    const int b = blockIdx.y;
    const int s0 = blockIdx.x * (2 * WARPS - 1);      // first overlap-add segment (of HALF samples) this CTA writes
    const int fa = s0 - 1 + 2 * warp, fb = fa + 1;    // this warp's two frames
    const bool va = fa >= 0 && fa < P.Tf, vb = fb >= 0 && fb < P.Tf;
    const size_t row_a = ((size_t)b * P.Tf + (va ? fa : 0)) * F;
    const size_t row_b = ((size_t)b * P.Tf + (vb ? fb : 0)) * F;

    float2 v[C::U][8];
    if (va || vb) {
        float2 UA[R], UB[R];
        const int jA = lane, jB = lane ? 64 - lane : 32;
        const bool l0 = (lane == 0);
        float2 DA[C::NA + 1], PA[C::NA + 1], DB[C::NBN + 1], PB[C::NBN + 1];
        const float inv_c = 1.f / P.c;
        const bool has_mask = P.mask != nullptr;
        // all loads first (unconditional: a missing frame reads row 0 of the batch and is scaled by 0), then the math:
        // the kernel is bound by HBM latency, so every lane keeps its 2 x (NA + NBN + 1) bins in flight at once
        constexpr int NBIN = C::NA + C::NBN + 1;       // A side | B side | Nyquist (read by every lane, used by lane 0)
        const float2* sa_ = P.spec + row_a;
        const float2* sb_ = P.spec + row_b;
        const float* ma_ = has_mask ? P.mask + row_a : nullptr;
        const float* mb_ = has_mask ? P.mask + row_b : nullptr;
        float2 rp[NBIN], rq[NBIN];
        float rma[NBIN], rmb[NBIN];
#pragma unroll
        for (int i = 0; i < NBIN; ++i) {
            const int k = (i < C::NA) ? jA + 64 * i : (i < C::NA + C::NBN ? jB + 64 * (i - C::NA) : HALF);
            rp[i] = sa_[k];
            rq[i] = sb_[k];
            rma[i] = has_mask ? ma_[k] : 1.f;
            rmb[i] = has_mask ? mb_[k] : 1.f;
        }
        const float za = va ? 1.f : 0.f, zb = vb ? 1.f : 0.f;
        const float si_a = P.sc_int * za, si_b = P.sc_int * zb, se_a = P.sc_edge * za, se_b = P.sc_edge * zb;
#pragma unroll
        for (int r = 0; r < C::NA; ++r) {
            const bool edge = (r == 0) && l0;
            c2r_combine(rp[r], rq[r], rma[r], rmb[r], has_mask, P.eps, inv_c, edge ? se_a : si_a, edge ? se_b : si_b, edge,
                        DA[r], PA[r]);
        }
#pragma unroll
        for (int r = 0; r < C::NBN; ++r)
            c2r_combine(rp[C::NA + r], rq[C::NA + r], rma[C::NA + r], rmb[C::NA + r], has_mask, P.eps, inv_c, si_a, si_b,
                        false, DB[r], PB[r]);
        {   // Nyquist (lane 0 only uses it)
            float2 d, pn;
            c2r_combine(rp[NBIN - 1], rq[NBIN - 1], rma[NBIN - 1], rmb[NBIN - 1], has_mask, P.eps, inv_c, se_a, se_b, true, d, pn);
            DA[C::NA] = PA[C::NA] = DB[C::NBN] = PB[C::NBN] = make_float2(0.f, 0.f);
            if (C::XTRA_A) { DA[C::NA] = d; PA[C::NA] = pn; }
            else { DB[C::NBN] = d; PB[C::NBN] = pn; }
        }
        // lanes >= 1: UA[r] = DA[r], UB[R-1-r] = PA[r];  UB[r] = DB[r], UA[R-1-r] = PB[r]
        // lane 0:     UA[r] = DA[r], UA[R-r]   = PA[r];  UB[r] = DB[r], UB[R-1-r] = PB[r]   (butterflies 0 and 32)
#pragma unroll
        for (int i = 0; i < R; ++i) {
            // UA[i]
            if (i < C::NA) {
                UA[i] = DA[i];
            } else {
                // lanes >= 1: PB[R-1-i];  lane 0: i == NA (even R only): DA[NA] (Nyquist), else PA[R-i]
                const float2 g = PB[R - 1 - i];
                const float2 z = (C::XTRA_A && i == C::NA) ? DA[C::NA] : PA[R - i];
                UA[i] = make_float2(l0 ? z.x : g.x, l0 ? z.y : g.y);
            }
            // UB[i]
            if (i < C::NBN) {
                UB[i] = DB[i];
            } else {
                // lanes >= 1: PA[R-1-i];  lane 0: i == NBN (odd R only): DB[NBN] (Nyquist), else PB[R-1-i]
                const float2 g = PA[R - 1 - i];
                const float2 z = (!C::XTRA_A && i == C::NBN) ? DB[C::NBN] : PB[R - 1 - i];
                UB[i] = make_float2(l0 ? z.x : g.x, l0 ? z.y : g.y);
            }
        }
        fft_dif<R>(UA, UB, v, buf, P.tw, lane);
    } else {
#pragma unroll
        for (int u = 0; u < C::U; ++u)
#pragma unroll
            for (int m1 = 0; m1 < 8; ++m1) v[u][m1] = make_float2(0.f, 0.f);
    }
    // windowed frames: frame a = Re(conj y) w, frame b = Im(conj y) w = -y.y w.  A missing frame contributes exactly
    // zero (not the rounding noise of the other frame's transform: the envelope division amplifies it at the edges)
    const float wsa = va ? 1.f : 0.f, wsb = vb ? -1.f : 0.f;
#pragma unroll
    for (int u = 0; u < C::U; ++u)
#pragma unroll
        for (int m1 = 0; m1 < 8; ++m1) {
            if (C::FULL || lane + 32 * u < C::NB) {
                const float w = __ldg(&P.window[lane + 32 * u + C::NB * m1]);
                v[u][m1].x *= w * wsa;
                v[u][m1].y *= w * wsb;
            }
        }
    // second half of frame b is the tail the next warp's first segment needs
    float* my_tail = tails + (size_t)warp * HALF;
#pragma unroll
    for (int u = 0; u < C::U; ++u)
#pragma unroll
        for (int m1 = 4; m1 < 8; ++m1)
            if (C::FULL || lane + 32 * u < C::NB) my_tail[lane + 32 * u + C::NB * (m1 - 4)] = v[u][m1].y;
    __syncthreads();
    const float* prev_tail = tails + (size_t)(warp > 0 ? warp - 1 : 0) * HALF;
    // segment s covers padded positions [s HALF, (s + 1) HALF); sum = first half of frame s + second half of frame s - 1
#pragma unroll
    for (int sg = 0; sg < 2; ++sg) {
        const int s = sg ? fb : fa;
        if (!sg && warp == 0) continue;                 // belongs to the previous CTA
        if (s < 0 || s > P.Tf) continue;
        // (iSTFT: output sample t = tp - HALF, so segment 0 is never written and segment s starts at t = (s - 1) HALF)
        const int t0 = (MODE == C2R_ISTFT) ? (s - 1) * HALF : s * HALF;
        const int lim = (MODE == C2R_ISTFT) ? P.length : P.Ltot;
        if (t0 < 0 || t0 >= lim) continue;
        float* po = P.out + (size_t)b * lim + t0 + lane;
        const float* pe = P.env + s * HALF + lane;
        const bool whole = t0 + HALF <= lim;
#pragma unroll
        for (int u = 0; u < C::U; ++u)
#pragma unroll
            for (int m1 = 0; m1 < 4; ++m1) {
                if (!(C::FULL || lane + 32 * u < C::NB)) continue;
                const int i = 32 * u + C::NB * m1;            // (+ lane, folded into the pointers)
                float tot = sg ? (v[u][m1].y + v[u][m1 + 4].x) : (v[u][m1].x + prev_tail[lane + i]);
                if (MODE == C2R_ISTFT) tot = __fdividef(tot, __ldg(pe + i));
                if (whole || t0 + lane + i < lim) po[i] = tot;
            }
    }
    if (MODE == C2R_ISTFT && s0 + 2 * WARPS - 1 > P.Tf) {
        // torch.istft zero-fills when `length` runs past the overlap-add buffer
        for (int t = P.Ltot - HALF + (int)threadIdx.x; t < P.length; t += WARPS * 32)
            if (t >= 0) P.out[(size_t)b * P.length + t] = 0.f;
    }
    (void)N;
}

}  // namespace wf
