// Generator bottleneck operators on channels-last rows [M = B*T'*F', C = 64].
//
// Replaces (reference file:line):
//   nn.LayerNorm(64)                           models/generator.py:126, :132, :238, :244, :577
//   4 x nn.GRU(16,16) per block (bi/uni)       models/generator.py:52-75, :94-110, :169-192, :211-222
//   nn.MultiheadAttention(64, 4 heads)         models/generator.py:78-82, :133, :194-198, :245
// The reference permutes [B,C,T,F] into [B*T,F,C] / [B*F,T,C] copies for every block; here the
// activations stay in one channels-last buffer and a sequence is described by strides
// (outer/inner/step), so the frequency and time blocks read the same memory without a transpose.
//
// The GRU recurrences are latency bound (33 or 129 dependent steps of a 16-wide cell): one
// 16-lane group per (sequence, GRU, direction) keeps W_hh rows in registers and exchanges the
// hidden state with warp shuffles; BPTT recomputes the gates from the saved hidden states and
// accumulates dW_hh in registers.  Attention is one CTA per (sequence, head) with K/V in shared
// memory and an online softmax; the backward recomputes probabilities from the saved
// log-sum-exp (no L x L matrix is ever written).
#include "common.cuh"

namespace {

struct SeqGeom {
    int nseq, L, inner;
    int64_t outer_stride, inner_stride, step_stride;   // in rows
};

__device__ __forceinline__ int64_t seq_row0(const SeqGeom& g, int seq) {
    return (int64_t)(seq / g.inner) * g.outer_stride + (int64_t)(seq % g.inner) * g.inner_stride;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// Gate non-linearities of the GRU recurrence: they sit on the serial per-step dependency chain (33 / 129 steps), so they
// are written on MUFU.EX2 / MUFU.RCP (4 dependent instructions) instead of expf / tanhf / IEEE division (~60):
// absolute error ~1e-7, inside the 2e-5 parity bar of the recurrence tests.
__device__ __forceinline__ float gate_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float gate_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));              // in (0, 1]: no overflow
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over the last dim (C <= 256), one warp per row
// ------------------------------------------------------------------------------------------
constexpr int kLnMaxPerLane = 8;

__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean,
                              float* __restrict__ rstd, int M, int C, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= M) return;
    const float* xr = x + (int64_t)warp * C;
    float v[kLnMaxPerLane];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        int c = lane + 32 * i;
        v[i] = c < C ? xr[c] : 0.f;
        s += v[i];
    }
    const float mu = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        int c = lane + 32 * i;
        float d = c < C ? v[i] - mu : 0.f;
        q += d * d;
    }
    const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        int c = lane + 32 * i;
        if (c < C) y[(int64_t)warp * C + c] = (v[i] - mu) * rs * gamma[c] + beta[c];
    }
    if (lane == 0) {
        mean[warp] = mu;
        rstd[warp] = rs;
    }
}

__global__ void ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                              const float* __restrict__ gamma, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ dres,
                              float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                              int M, int C) {
    const int lane = threadIdx.x & 31;
    const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float ag[kLnMaxPerLane], ab[kLnMaxPerLane], gm[kLnMaxPerLane];
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        ag[i] = ab[i] = 0.f;
        int c = lane + 32 * i;
        gm[i] = c < C ? gamma[c] : 0.f;
    }
    for (int row = warp0; row < M; row += nwarps) {
        const float mu = mean[row], rs = rstd[row];
        float xh[kLnMaxPerLane], g[kLnMaxPerLane];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) {
            int c = lane + 32 * i;
            if (c < C) {
                float d = dy[(int64_t)row * C + c];
                xh[i] = (x[(int64_t)row * C + c] - mu) * rs;
                ag[i] += d * xh[i];
                ab[i] += d;
                g[i] = d * gm[i];
                s1 += g[i];
                s2 += g[i] * xh[i];
            } else {
                xh[i] = g[i] = 0.f;
            }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int i = 0; i < kLnMaxPerLane; ++i) {
            int c = lane + 32 * i;
            if (c < C) {
                float v = rs * (g[i] - s1 - xh[i] * s2);
                if (dres) v += dres[(int64_t)row * C + c];
                dx[(int64_t)row * C + c] = v;
            }
        }
    }
    // block-level reduction of the parameter gradients, then one atomic per channel per CTA
    __shared__ float sg[256], sb[256];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) sg[c] = sb[c] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
        int c = lane + 32 * i;
        if (c < C) {
            atomicAdd(&sg[c], ag[i]);
            atomicAdd(&sb[c], ab[i]);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(&dgamma[c], sg[c]);
        atomicAdd(&dbeta[c], sb[c]);
    }
}

// C = 64 (the generator's only width): half a warp per row with 16-byte accesses, the next row pair prefetched while
// the current one is reduced, parameter gradients kept in registers across rows and reduced once per CTA.
struct LnRow { float4 d, x, r; float mu, rs; };

__global__ void __launch_bounds__(256) ln_bwd64_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                       const float* __restrict__ gamma, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ dres,
                                                       float* __restrict__ dx, float* __restrict__ dgamma,
                                                       float* __restrict__ dbeta, int M) {
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4, warp = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int step = 2 * ((gridDim.x * blockDim.x) >> 5);
    const float4 gm = reinterpret_cast<const float4*>(gamma)[sub];
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
    auto load = [&](int row) {
        LnRow v;
        v.d = v.x = v.r = make_float4(0.f, 0.f, 0.f, 0.f);
        v.mu = 0.f; v.rs = 0.f;
        if (row < M) {
            const size_t o = (size_t)row * 16 + sub;
            v.d = reinterpret_cast<const float4*>(dy)[o];
            v.x = reinterpret_cast<const float4*>(x)[o];
            if (dres) v.r = reinterpret_cast<const float4*>(dres)[o];
            v.mu = mean[row];
            v.rs = rstd[row];
        }
        return v;
    };
    int base = 2 * gw;
    LnRow cur = load(base + half);
    for (; base < M; base += step) {
        const LnRow nxt = load(base + step + half);
        const int row = base + half;
        float4 xh, g;
        xh.x = (cur.x.x - cur.mu) * cur.rs; xh.y = (cur.x.y - cur.mu) * cur.rs;
        xh.z = (cur.x.z - cur.mu) * cur.rs; xh.w = (cur.x.w - cur.mu) * cur.rs;
        g.x = cur.d.x * gm.x; g.y = cur.d.y * gm.y; g.z = cur.d.z * gm.z; g.w = cur.d.w * gm.w;
        ag.x = fmaf(cur.d.x, xh.x, ag.x); ag.y = fmaf(cur.d.y, xh.y, ag.y);
        ag.z = fmaf(cur.d.z, xh.z, ag.z); ag.w = fmaf(cur.d.w, xh.w, ag.w);
        ab.x += cur.d.x; ab.y += cur.d.y; ab.z += cur.d.z; ab.w += cur.d.w;
        float s1 = (g.x + g.y) + (g.z + g.w);
        float s2 = fmaf(g.x, xh.x, g.y * xh.y) + fmaf(g.z, xh.z, g.w * xh.w);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 *= (1.f / 64.f);
        s2 *= (1.f / 64.f);
        if (row < M) {
            float4 v;
            v.x = cur.rs * (g.x - s1 - xh.x * s2) + cur.r.x;
            v.y = cur.rs * (g.y - s1 - xh.y * s2) + cur.r.y;
            v.z = cur.rs * (g.z - s1 - xh.z * s2) + cur.r.z;
            v.w = cur.rs * (g.w - s1 - xh.w * s2) + cur.r.w;
            reinterpret_cast<float4*>(dx)[(size_t)row * 16 + sub] = v;
        }
        cur = nxt;
    }
    // the two rows of the warp, then the 8 warps of the CTA, then one atomic per channel per CTA
    __shared__ float red[8][128];
    float a[4] = {ag.x, ag.y, ag.z, ag.w}, bsum[4] = {ab.x, ab.y, ab.z, ab.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] += __shfl_xor_sync(0xffffffffu, a[i], 16);
        bsum[i] += __shfl_xor_sync(0xffffffffu, bsum[i], 16);
    }
    if (half == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            red[warp][4 * sub + i] = a[i];
            red[warp][64 + 4 * sub + i] = bsum[i];
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        if (threadIdx.x < 64) atomicAdd(&dgamma[threadIdx.x], t);
        else atomicAdd(&dbeta[threadIdx.x - 64], t);
    }
}

// ------------------------------------------------------------------------------------------
// GRU recurrence.  gi: [rows, GD, 48] = W_ih x + b_ih (gates r,z,n), whh: [GD,48,16], bhh: [GD,48],
// hs: [rows, GD, 16].  GD = groups * dirs, dir = gd % D, dir 1 runs the sequence backwards.
// ------------------------------------------------------------------------------------------
constexpr int kH = 16;
constexpr int kGruThreads = 256;
constexpr int kUnitsPerCta = kGruThreads / kH;
constexpr int kGruPre = 6;     // recurrence steps whose inputs are in flight
// backward: 232 registers per thread.  256-thread CTAs meant ONE resident CTA (8 warps) per SM and 3.5 waves = 4 rounds of
// a latency-bound kernel for the frequency blocks (520 CTAs); 64-thread CTAs capped at 204 registers give 5 per SM (10
// warps): the same units finish in 3 rounds, and the closing shared-memory reduction has 4 contenders instead of 16.
constexpr int kGruBwdThreads = 64;
constexpr int kGruBwdUnits = kGruBwdThreads / kH;

__global__ void __launch_bounds__(kGruThreads) gru_fwd_kernel(const float* __restrict__ gi,
                                                              const float* __restrict__ whh,
                                                              const float* __restrict__ bhh, float* __restrict__ hs,
                                                              float* __restrict__ gsave, float* __restrict__ hprev,
                                                              SeqGeom geo, int GD, int D) {
    const int j = threadIdx.x % kH;
    const int seq = blockIdx.x * kUnitsPerCta + threadIdx.x / kH;
    const int gd = blockIdx.y;
    const bool valid = seq < geo.nseq;
    const bool rev = (gd % D) == 1;
    float wr[kH], wz[kH], wn[kH];
    const float* w = whh + (size_t)gd * 3 * kH * kH;
#pragma unroll
    for (int i = 0; i < kH; ++i) {
        wr[i] = w[(0 * kH + j) * kH + i];
        wz[i] = w[(1 * kH + j) * kH + i];
        wn[i] = w[(2 * kH + j) * kH + i];
    }
    const float br = bhh[gd * 3 * kH + j], bz = bhh[gd * 3 * kH + kH + j], bn = bhh[gd * 3 * kH + 2 * kH + j];
    const int64_t row0 = valid ? seq_row0(geo, seq) : 0;
    const int64_t gstride = (int64_t)GD * 3 * kH;
    float h = 0.f;
    __shared__ __align__(16) float hbuf[2][kGruThreads];
    hbuf[0][threadIdx.x] = 0.f;
    __syncwarp();
    const int step0 = rev ? geo.L - 1 : 0;
    const int dstep = rev ? -1 : 1;
    // The input projections of the next kGruPre steps travel while the current step is computed: a step is ~250 cycles
    // of dependent math, an L2 hit costs 600+, so a one-step look-ahead left every step waiting on its load.
    float pr[kGruPre], pz[kGruPre], pn[kGruPre];
    const float* gbase = gi + (row0 + (int64_t)step0 * geo.step_stride) * gstride + gd * 3 * kH + j;
    const int64_t gstep = (int64_t)dstep * geo.step_stride * gstride;
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) {
        pr[u] = pz[u] = pn[u] = 0.f;
        if (valid && u < geo.L) {
            const float* g0 = gbase + u * gstep;
            pr[u] = g0[0]; pz[u] = g0[kH]; pn[u] = g0[2 * kH];
        }
    }
    float* hrow = hs + ((row0 + (int64_t)step0 * geo.step_stride) * GD + gd) * kH + j;
    const int64_t hstep = (int64_t)dstep * geo.step_stride * GD * kH;
    // training: the gates (r, z, n) and the recurrent part of the candidate (hn = W_hn h + b_hn) of every step, and the
    // state each step started from, go out for the backward (which then neither recomputes a 16 x 48 mat-vec per step nor
    // accumulates dW_hh in registers: profiles round 1, gru_bwd_kernel 129-156 us at 200-232 registers)
    float* grow = gsave ? gsave + ((row0 + (int64_t)step0 * geo.step_stride) * GD + gd) * 4 * kH + j : nullptr;
    float* prow = hprev ? hprev + ((row0 + (int64_t)step0 * geo.step_stride) * GD + gd) * kH + j : nullptr;
    for (int s0 = 0; s0 < geo.L; s0 += kGruPre) {
#pragma unroll
        for (int u = 0; u < kGruPre; ++u) {
            const int s = s0 + u;
            if (s >= geo.L) break;
            const float gr = pr[u], gz = pz[u], gn = pn[u];
            if (valid && s + kGruPre < geo.L) {
                const float* g1 = gbase + (int64_t)(s + kGruPre) * gstep;
                pr[u] = g1[0]; pz[u] = g1[kH]; pn[u] = g1[2 * kH];
            }
            float hr = br, hz = bz, hn = bn, hr2 = 0.f, hz2 = 0.f, hn2 = 0.f;      // two partial sums per gate: half the chain
            // the unit's 16 state values come back as four broadcast 16-byte shared-memory loads (written at the end of
            // the previous step).  32 shuffles per step and lane - one warp-shuffle per clock and SM, 16 warps resident -
            // were ~500 of the step's ~900 cycles
            const float4* h4 = reinterpret_cast<const float4*>(hbuf[s & 1] + (threadIdx.x & ~(kH - 1)));
            const float4 q0 = h4[0], q1 = h4[1], q2 = h4[2], q3 = h4[3];
            const float hv[kH] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int i = 0; i < kH; i += 2) {
                hr = fmaf(wr[i], hv[i], hr);
                hz = fmaf(wz[i], hv[i], hz);
                hn = fmaf(wn[i], hv[i], hn);
                hr2 = fmaf(wr[i + 1], hv[i + 1], hr2);
                hz2 = fmaf(wz[i + 1], hv[i + 1], hz2);
                hn2 = fmaf(wn[i + 1], hv[i + 1], hn2);
            }
            hr += hr2; hz += hz2; hn += hn2;
            const float r = gate_sigmoid(gr + hr);
            const float z = gate_sigmoid(gz + hz);
            const float n = gate_tanh(gn + r * hn);
            if (valid && grow) {
                float* g4 = grow + (int64_t)s * hstep * 4;
                g4[0] = r; g4[kH] = z; g4[2 * kH] = n; g4[3 * kH] = hn;
                prow[(int64_t)s * hstep] = h;
            }
            h = (1.f - z) * n + z * h;
            hbuf[(s + 1) & 1][threadIdx.x] = h;        // (the other buffer: this step's readers may still be loading)
            __syncwarp();
            if (valid) hrow[(int64_t)s * hstep] = h;
        }
    }
}

// BPTT from the saved gates.  dh_in: [rows, ldd] gradient of the GRU output (column (gd/D)*16 + j, shared by both
// directions).  Writes dgi [rows, GD, 48] = gradient of the input projections (dr, dz, dn pre-activation) and
// dgh [rows, GD, 48] = gradient of the recurrent projections (dr, dz, dn * r); the parameter gradients are GEMMs /
// column sums over those two arrays (dW_hh = dgh^T hprev, db_ih = colsum dgi, db_hh = colsum dgh: lctgan/gen_impl.py).
// Per step and lane: 6 loads, the gate derivatives, a 16 x 48 transposed mat-vec (W_hh^T dgate) reduce-scattered over
// the 16 lanes of the unit - no gate recomputation, no dW_hh accumulation (round 1: ~420 instructions per step).
__global__ void __launch_bounds__(kGruBwdThreads) gru_bwd_kernel(
    const float* __restrict__ gsave, const float* __restrict__ hprev, const float* __restrict__ whh,
    const float* __restrict__ dh_in, int ldd, float* __restrict__ dgi, float* __restrict__ dgh, SeqGeom geo, int GD,
    int D) {
    const int j = threadIdx.x % kH;
    const int seq = blockIdx.x * kGruBwdUnits + threadIdx.x / kH;
    const int gd = blockIdx.y;
    const bool valid = seq < geo.nseq;
    const bool rev = (gd % D) == 1;
    float wr[kH], wz[kH], wn[kH];
    const float* w = whh + (size_t)gd * 3 * kH * kH;
#pragma unroll
    for (int i = 0; i < kH; ++i) {
        wr[i] = w[(0 * kH + j) * kH + i];
        wz[i] = w[(1 * kH + j) * kH + i];
        wn[i] = w[(2 * kH + j) * kH + i];
    }
    const int64_t row0 = valid ? seq_row0(geo, seq) : 0;
    const int64_t gstride = (int64_t)GD * 3 * kH;
    const int col = (gd / D) * kH + j;
    // walk the recurrence backwards: forward order was step = rev ? L-1..0 : 0..L-1
    const int step = rev ? 0 : geo.L - 1;
    const int dstep = rev ? 1 : -1;   // direction of "previous forward step"
    float dh = 0.f;
    // inputs of the next kGruPre steps in flight (see gru_fwd_kernel); `it` counts steps walked
    float qr[kGruPre], qz[kGruPre], qn[kGruPre], qm[kGruPre], qh[kGruPre], qd[kGruPre];
    auto fetch = [&](int it, float& r, float& z, float& n, float& hn, float& hp, float& dq) {
        r = z = n = hn = hp = dq = 0.f;
        if (valid && it < geo.L) {
            const int64_t rw = row0 + (int64_t)(step + it * dstep) * geo.step_stride;
            const float* g4 = gsave + (rw * GD + gd) * 4 * kH + j;
            r = g4[0]; z = g4[kH]; n = g4[2 * kH]; hn = g4[3 * kH];
            hp = hprev[(rw * GD + gd) * kH + j];
            dq = dh_in[rw * ldd + col];
        }
    };
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) fetch(u, qr[u], qz[u], qn[u], qm[u], qh[u], qd[u]);
    for (int it0 = 0; it0 < geo.L; it0 += kGruPre) {
#pragma unroll
    for (int u = 0; u < kGruPre; ++u) {
        const int it = it0 + u;
        if (it >= geo.L) break;
        const int64_t row = row0 + (int64_t)(step + it * dstep) * geo.step_stride;
        const float r = qr[u], z = qz[u], n = qn[u], hn = qm[u], hp = qh[u], dout = qd[u];
        fetch(it + kGruPre, qr[u], qz[u], qn[u], qm[u], qh[u], qd[u]);
        const float dht = dh + dout;
        const float dn_pre = dht * (1.f - z) * (1.f - n * n);
        const float dz_pre = dht * (hp - n) * z * (1.f - z);
        const float dr_pre = dn_pre * hn * r * (1.f - r);
        const float dhn = dn_pre * r;
        if (valid) {
            float* d0 = dgi + row * gstride + gd * 3 * kH;
            d0[j] = dr_pre; d0[kH + j] = dz_pre; d0[2 * kH + j] = dn_pre;
            float* d1 = dgh + row * gstride + gd * 3 * kH;
            d1[j] = dr_pre; d1[kH + j] = dz_pre; d1[2 * kH + j] = dhn;
        }
        float c[kH];
#pragma unroll
        for (int i = 0; i < kH; ++i) c[i] = wr[i] * dr_pre + wz[i] * dz_pre + wn[i] * dhn;
        // reduce-scatter c[] over the 16 lanes of the unit: lane i ends with sum_j c_j[i]
#pragma unroll
        for (int off = kH / 2, len = kH / 2; off >= 1; off >>= 1, len >>= 1) {
            const bool upper = (j & off) != 0;
#pragma unroll
            for (int t = 0; t < kH / 2; ++t) {
                if (t < len) {
                    const float send = upper ? c[t] : c[t + len];
                    const float keep = upper ? c[t + len] : c[t];
                    c[t] = keep + __shfl_xor_sync(0xffffffffu, send, off, kH);
                }
            }
        }
        dh = dht * z + c[0];
    }
    }
}

// seq[row,c] = x[row,c] + sum_d hs[row,(g,d),j];  gsum (optional, row stride ldg) = the GRU sum alone
__global__ void gru_combine_kernel(const float* __restrict__ x, const float* __restrict__ hs,
                                   float* __restrict__ seq, float* __restrict__ gsum, int ldg, int64_t M, int G,
                                   int D) {
    const int C = G * kH;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * C) return;
    int64_t row = i / C;
    int c = (int)(i - row * C);
    int g = c / kH, j = c - g * kH;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += hs[((row * G + g) * D + d) * kH + j];
    seq[i] = x[i] + s;
    if (gsum) gsum[row * ldg + c] = s;
}

// ------------------------------------------------------------------------------------------
// Multi-head self-attention core: qkv [rows, 3E] (q | k | v, head h at columns h*16), E = H*16
// ------------------------------------------------------------------------------------------
// (exponentials are __expf = MUFU.EX2: two instructions instead of ~10 in loops that run L^2 times per head; relative
// error 2^-21, far inside the 2e-5 parity bar)
constexpr int kHd = 16;
constexpr int kAttnMaxThreads = 256;

// 16-wide head rows as 8 float2 and packed fp32x2 arithmetic (FFMA2, sm_100): the attention loops are issue bound (ncu:
// 43-57 % issue active at ~ 26 % of the FMA rate), and every multiply-add here comes in pairs over the head dimension,
// so the packed form halves the FMA instruction count without changing any product or any addition's operands' values
// (only the order of the 16-term dot-product sums: even / odd lanes first).
__device__ __forceinline__ void load_row2(const float* p, float2 (&r)[kHd / 2]) {      // p: 16-byte aligned
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int d4 = 0; d4 < kHd / 4; ++d4) {
        const float4 v = p4[d4];
        r[2 * d4] = make_float2(v.x, v.y);
        r[2 * d4 + 1] = make_float2(v.z, v.w);
    }
}
__device__ __forceinline__ float dot16(const float2 (&a)[kHd / 2], const float2 (&b)[kHd / 2]) {
    float2 s0 = __fmul2_rn(a[0], b[0]), s1 = __fmul2_rn(a[1], b[1]);
#pragma unroll
    for (int i = 2; i < kHd / 2; i += 2) {
        s0 = __ffma2_rn(a[i], b[i], s0);
        s1 = __ffma2_rn(a[i + 1], b[i + 1], s1);
    }
    return (s0.x + s1.x) + (s0.y + s1.y);
}
__device__ __forceinline__ void axpy16(float a, const float2 (&x)[kHd / 2], float2 (&y)[kHd / 2]) {
    const float2 a2 = make_float2(a, a);
#pragma unroll
    for (int i = 0; i < kHd / 2; ++i) y[i] = __ffma2_rn(a2, x[i], y[i]);
}   // block = sequence length rounded up to a warp (one query/key per thread)

// One query per thread (blockIdx.z selects the block of queries); keys / values stream through shared memory in chunks
// of `chunk` rows - the online softmax does not care where a chunk ends - so the sequence length is unbounded
// (infer.py feeds whole utterances: L = frames + 3 in the time block).
__global__ void __launch_bounds__(kAttnMaxThreads) attn_fwd_kernel(const float* __restrict__ qkv,
                                                                float* __restrict__ out, float* __restrict__ lse,
                                                                SeqGeom geo, int H, float scale, int chunk) {
    extern __shared__ __align__(16) float sm[];
    const int L = geo.L, E = H * kHd;
    float* Ks = sm;                 // [chunk][16]
    float* Vs = sm + (size_t)chunk * kHd;
    const int seq = blockIdx.x, h = blockIdx.y;
    const int64_t row0 = seq_row0(geo, seq);
    const int i = blockIdx.z * (int)blockDim.x + (int)threadIdx.x;
    const bool valid = i < L;
    const int64_t row = row0 + (int64_t)(valid ? i : 0) * geo.step_stride;
    float2 q[kHd / 2], acc[kHd / 2];
#pragma unroll
    for (int d = 0; d < kHd / 2; ++d) {
        q[d] = make_float2(qkv[row * 3 * E + h * kHd + 2 * d] * scale, qkv[row * 3 * E + h * kHd + 2 * d + 1] * scale);
        acc[d] = make_float2(0.f, 0.f);
    }
    float m = -INFINITY, l = 0.f;
    for (int c0 = 0; c0 < L; c0 += chunk) {
        const int n = min(chunk, L - c0);
        if (c0) __syncthreads();                       // everyone is done with the previous chunk
        for (int idx = threadIdx.x; idx < n * kHd; idx += (int)blockDim.x) {
            int t = idx / kHd, d = idx - t * kHd;
            const float* r = qkv + (row0 + (int64_t)(c0 + t) * geo.step_stride) * 3 * E;
            Ks[idx] = r[E + h * kHd + d];
            Vs[idx] = r[2 * E + h * kHd + d];
        }
        __syncthreads();
        if (valid) {
            for (int t = 0; t < n; ++t) {
                float2 k2[kHd / 2], v2[kHd / 2];
                load_row2(Ks + t * kHd, k2);
                const float s = dot16(q, k2);
                const float mn = fmaxf(m, s);
                const float corr = __expf(m - mn);
                const float pr = __expf(s - mn);
                l = l * corr + pr;
                load_row2(Vs + t * kHd, v2);
                const float2 c2 = make_float2(corr, corr), p2 = make_float2(pr, pr);
#pragma unroll
                for (int d = 0; d < kHd / 2; ++d) acc[d] = __ffma2_rn(p2, v2[d], __fmul2_rn(acc[d], c2));
                m = mn;
            }
        }
    }
    if (valid) {
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < kHd / 2; ++d) {
            out[row * E + h * kHd + 2 * d] = acc[d].x * inv;
            out[row * E + h * kHd + 2 * d + 1] = acc[d].y * inv;
        }
        lse[row * H + h] = m + logf(l);
    }
}

__global__ void __launch_bounds__(kAttnMaxThreads) attn_bwd_kernel(const float* __restrict__ qkv,
                                                                const float* __restrict__ out,
                                                                const float* __restrict__ lse,
                                                                const float* __restrict__ dout,
                                                                float* __restrict__ dqkv, SeqGeom geo, int H,
                                                                float scale, int parts) {
    extern __shared__ __align__(16) float sm[];
    const int L = geo.L, E = H * kHd;
    float* Qs = sm;                          // [L][16] (pre-scaled)
    float* Ks = Qs + (size_t)L * kHd;
    float* Vs = Ks + (size_t)L * kHd;
    float* dOs = Vs + (size_t)L * kHd;
    float* Ls = dOs + (size_t)L * kHd;       // [L] log-sum-exp
    float* Ds = Ls + L;                      // [L] rowsum(dO * O)
    float* red = Ds + L;                     // parts > 1: [parts][L][33] partial sums
    const int seq = blockIdx.x, h = blockIdx.y;
    const int64_t row0 = seq_row0(geo, seq);
    for (int idx = threadIdx.x; idx < L * kHd; idx += (int)blockDim.x) {
        int t = idx / kHd, d = idx - t * kHd;
        const int64_t row = row0 + (int64_t)t * geo.step_stride;
        const float* r = qkv + row * 3 * E;
        Qs[idx] = r[h * kHd + d] * scale;
        Ks[idx] = r[E + h * kHd + d];
        Vs[idx] = r[2 * E + h * kHd + d];
        dOs[idx] = dout[row * E + h * kHd + d];
    }
    for (int t = threadIdx.x; t < L; t += (int)blockDim.x) {
        const int64_t row = row0 + (int64_t)t * geo.step_stride;
        float dsum = 0.f;
#pragma unroll
        for (int d = 0; d < kHd; ++d) dsum += dout[row * E + h * kHd + d] * out[row * E + h * kHd + d];
        Ds[t] = dsum;
        Ls[t] = lse[row * H + h];
    }
    __syncthreads();
    if (parts > 1) {
        // short sequences (the frequency blocks, L = 33): one query per thread would leave half of a 64-thread CTA idle
        // and every thread with a serial loop over all keys.  `parts` threads share a query (pass A) / key (pass B),
        // each walking 1/parts of the other axis; the partial sums meet in shared memory.
        const int part = (int)threadIdx.x / L, i = (int)threadIdx.x - part * L;
        const bool act = part < parts;
        const int chunk = (L + parts - 1) / parts;
        const int t0 = part * chunk, t1 = min(L, t0 + chunk);
        if (act) {
            float2 q[kHd / 2], go[kHd / 2], dq[kHd / 2];
            load_row2(Qs + i * kHd, q);
            load_row2(dOs + i * kHd, go);
#pragma unroll
            for (int d = 0; d < kHd / 2; ++d) dq[d] = make_float2(0.f, 0.f);
            const float li = Ls[i], di = Ds[i];
            for (int t = t0; t < t1; ++t) {
                float2 k2[kHd / 2], v2[kHd / 2];
                load_row2(Ks + t * kHd, k2);
                load_row2(Vs + t * kHd, v2);
                const float ds = __expf(dot16(q, k2) - li) * (dot16(go, v2) - di);
                axpy16(ds, k2, dq);
            }
#pragma unroll
            for (int d = 0; d < kHd / 2; ++d) {                                 // odd row stride: no bank conflicts
                red[(part * L + i) * (kHd + 1) + 2 * d] = dq[d].x;
                red[(part * L + i) * (kHd + 1) + 2 * d + 1] = dq[d].y;
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < L * kHd; idx += (int)blockDim.x) {
            const int i2 = idx / kHd, d = idx - i2 * kHd;
            float v = 0.f;
            for (int pp = 0; pp < parts; ++pp) v += red[(pp * L + i2) * (kHd + 1) + d];
            dqkv[(row0 + (int64_t)i2 * geo.step_stride) * 3 * E + h * kHd + d] = v * scale;
        }
        __syncthreads();
        if (act) {
            const int t = i;
            float2 k[kHd / 2], v[kHd / 2], dk[kHd / 2], dv[kHd / 2];
            load_row2(Ks + t * kHd, k);
            load_row2(Vs + t * kHd, v);
#pragma unroll
            for (int d = 0; d < kHd / 2; ++d) dk[d] = dv[d] = make_float2(0.f, 0.f);
            for (int i2 = t0; i2 < t1; ++i2) {
                float2 q2[kHd / 2], g2[kHd / 2];
                load_row2(Qs + i2 * kHd, q2);
                load_row2(dOs + i2 * kHd, g2);
                const float pr = __expf(dot16(q2, k) - Ls[i2]);
                const float ds = pr * (dot16(g2, v) - Ds[i2]);
                axpy16(pr, g2, dv);
                axpy16(ds, q2, dk);
            }
#pragma unroll
            for (int d = 0; d < kHd / 2; ++d) {
                float* r = red + (part * L + t) * (2 * kHd + 1);
                r[2 * d] = dk[d].x; r[2 * d + 1] = dk[d].y;
                r[kHd + 2 * d] = dv[d].x; r[kHd + 2 * d + 1] = dv[d].y;
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < L * 2 * kHd; idx += (int)blockDim.x) {
            const int t2 = idx / (2 * kHd), d = idx - t2 * 2 * kHd;      // d < 16: dK, else dV
            float v = 0.f;
            for (int pp = 0; pp < parts; ++pp) v += red[(pp * L + t2) * (2 * kHd + 1) + d];
            dqkv[(row0 + (int64_t)t2 * geo.step_stride) * 3 * E + E + (d / kHd) * E + h * kHd + (d % kHd)] = v;
        }
        return;
    }
    // pass A: one query per thread -> dQ
    for (int i = threadIdx.x; i < L; i += (int)blockDim.x) {
        float2 q[kHd / 2], go[kHd / 2], dq[kHd / 2];
        load_row2(Qs + i * kHd, q);
        load_row2(dOs + i * kHd, go);
#pragma unroll
        for (int d = 0; d < kHd / 2; ++d) dq[d] = make_float2(0.f, 0.f);
        const float li = Ls[i], di = Ds[i];
        for (int t = 0; t < L; ++t) {
            float2 k2[kHd / 2], v2[kHd / 2];
            load_row2(Ks + t * kHd, k2);
            load_row2(Vs + t * kHd, v2);
            const float ds = __expf(dot16(q, k2) - li) * (dot16(go, v2) - di);
            axpy16(ds, k2, dq);
        }
        const int64_t row = row0 + (int64_t)i * geo.step_stride;
#pragma unroll
        for (int d = 0; d < kHd / 2; ++d) {
            dqkv[row * 3 * E + h * kHd + 2 * d] = dq[d].x * scale;
            dqkv[row * 3 * E + h * kHd + 2 * d + 1] = dq[d].y * scale;
        }
    }
    // pass B: one key per thread -> dK, dV
    for (int t = threadIdx.x; t < L; t += (int)blockDim.x) {
        float2 k[kHd / 2], v[kHd / 2], dk[kHd / 2], dv[kHd / 2];
        load_row2(Ks + t * kHd, k);
        load_row2(Vs + t * kHd, v);
#pragma unroll
        for (int d = 0; d < kHd / 2; ++d) dk[d] = dv[d] = make_float2(0.f, 0.f);
        for (int i = 0; i < L; ++i) {
            float2 q2[kHd / 2], g2[kHd / 2];
            load_row2(Qs + i * kHd, q2);
            load_row2(dOs + i * kHd, g2);
            const float pr = __expf(dot16(q2, k) - Ls[i]);
            const float ds = pr * (dot16(g2, v) - Ds[i]);
            axpy16(pr, g2, dv);
            axpy16(ds, q2, dk);          // Qs is pre-scaled, so this already carries `scale`
        }
        const int64_t row = row0 + (int64_t)t * geo.step_stride;
#pragma unroll
        for (int d = 0; d < kHd / 2; ++d) {
            dqkv[row * 3 * E + E + h * kHd + 2 * d] = dk[d].x;
            dqkv[row * 3 * E + E + h * kHd + 2 * d + 1] = dk[d].y;
            dqkv[row * 3 * E + 2 * E + h * kHd + 2 * d] = dv[d].x;
            dqkv[row * 3 * E + 2 * E + h * kHd + 2 * d + 1] = dv[d].y;
        }
    }
}

__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dpre,
                               int64_t n, int act, float slope) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dpre[i] = dy[i] * act_grad_from_out(y[i], act, slope);
}

int attn_threads(int64_t L) {
    int64_t t = (L + 31) / 32 * 32;
    return (int)(t < 32 ? 32 : (t > kAttnMaxThreads ? kAttnMaxThreads : t));
}

bool geom_ok(SeqGeom& g, int64_t nseq, int64_t L, int64_t inner, int64_t outer_stride, int64_t inner_stride,
             int64_t step_stride) {
    if (nseq <= 0 || L <= 0 || inner <= 0 || nseq >= (1LL << 31) || L >= (1 << 20)) return false;
    g.nseq = (int)nseq; g.L = (int)L; g.inner = (int)inner;
    g.outer_stride = outer_stride; g.inner_stride = inner_stride; g.step_stride = step_stride;
    return true;
}

}  // namespace

LCT_API int lct_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                              float* rstd, int64_t M, int64_t C, float eps, cudaStream_t st) {
    if (!x || !gamma || !beta || !y || !mean || !rstd || M <= 0 || C <= 0 || C > 32 * kLnMaxPerLane) return LCT_EINVAL;
    ln_fwd_kernel<<<(unsigned)ceil_div64(M * 32, 256), 256, 0, st>>>(x, gamma, beta, y, mean, rstd, (int)M, (int)C, eps);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dx = dres (optional) + LN backward; dgamma/dbeta are accumulated (caller zeroes)
LCT_API int lct_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean,
                              const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta,
                              int64_t M, int64_t C, cudaStream_t st) {
    if (!dy || !x || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta || M <= 0 || C <= 0 ||
        C > 32 * kLnMaxPerLane)
        return LCT_EINVAL;
    const auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
    if (C == 64 && al16(dy) && al16(x) && al16(dx) && al16(gamma) && (!dres || al16(dres))) {
        int64_t ctas = ceil_div64(M, 16);
        if (ctas > 148 * 3) ctas = 148 * 3;
        ln_bwd64_kernel<<<(unsigned)ctas, 256, 0, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, (int)M);
        LCT_RETURN_IF_LAUNCH_FAILED();
        return 0;
    }
    int64_t blocks = ceil_div64(M * 32, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ln_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, (int)M, (int)C);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// gsave [rows, GD, 4, 16] (r, z, n, W_hn h + b_hn) and hprev [rows, GD, 16] (state each step started from) are optional
// outputs for the backward: pass both or neither
LCT_API int lct_gru_fwd(const float* gi, const float* whh, const float* bhh, float* hs, float* gsave, float* hprev,
                        int64_t nseq, int64_t L, int64_t GD, int64_t D, int64_t inner, int64_t outer_stride,
                        int64_t inner_stride, int64_t step_stride, cudaStream_t st) {
    SeqGeom g;
    if (!gi || !whh || !bhh || !hs || (gsave == nullptr) != (hprev == nullptr) || GD <= 0 || D <= 0 || GD % D ||
        GD >= 65536 || !geom_ok(g, nseq, L, inner, outer_stride, inner_stride, step_stride))
        return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(nseq, kUnitsPerCta), (unsigned)GD);
    gru_fwd_kernel<<<grid, kGruThreads, 0, st>>>(gi, whh, bhh, hs, gsave, hprev, g, (int)GD, (int)D);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dgi / dgh [rows, GD, 48]: gradients of the input / recurrent gate projections (see gru_bwd_kernel)
LCT_API int lct_gru_bwd(const float* gsave, const float* hprev, const float* whh, const float* dh_in, int64_t ldd,
                        float* dgi, float* dgh, int64_t nseq, int64_t L, int64_t GD, int64_t D, int64_t inner,
                        int64_t outer_stride, int64_t inner_stride, int64_t step_stride, cudaStream_t st) {
    SeqGeom g;
    if (!gsave || !hprev || !whh || !dh_in || !dgi || !dgh || GD <= 0 || D <= 0 || GD % D || GD >= 65536 ||
        !geom_ok(g, nseq, L, inner, outer_stride, inner_stride, step_stride))
        return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(nseq, kGruBwdUnits), (unsigned)GD);
    gru_bwd_kernel<<<grid, kGruBwdThreads, 0, st>>>(gsave, hprev, whh, dh_in, (int)ldd, dgi, dgh, g, (int)GD, (int)D);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_gru_combine(const float* x, const float* hs, float* seq, float* gsum, int64_t ldg, int64_t M,
                            int64_t G, int64_t D, cudaStream_t st) {
    if (!x || !hs || !seq || M <= 0 || G <= 0 || D <= 0) return LCT_EINVAL;
    int64_t n = M * G * kH;
    gru_combine_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(x, hs, seq, gsum, (int)ldg, M, (int)G, (int)D);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_attn_fwd(const float* qkv, float* out, float* lse, int64_t heads, int64_t nseq, int64_t L,
                         int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t step_stride,
                         cudaStream_t st) {
    SeqGeom g;
    if (!qkv || !out || !lse || heads <= 0 || heads >= 65536 ||
        !geom_ok(g, nseq, L, inner, outer_stride, inner_stride, step_stride))
        return LCT_EINVAL;
    // keys / values are streamed through shared memory 1024 rows (128 KB) at a time: any L
    const int chunk = (int)(L < 1024 ? L : 1024);
    size_t smem = (size_t)2 * chunk * kHd * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int threads = attn_threads(L);
    dim3 grid((unsigned)nseq, (unsigned)heads, (unsigned)ceil_div64(L, threads));
    if (grid.z >= 65536) return LCT_EUNSUPPORTED;
    attn_fwd_kernel<<<grid, threads, smem, st>>>(qkv, out, lse, g, (int)heads, 1.f / sqrtf((float)kHd), chunk);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_attn_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv,
                         int64_t heads, int64_t nseq, int64_t L, int64_t inner, int64_t outer_stride,
                         int64_t inner_stride, int64_t step_stride, cudaStream_t st) {
    SeqGeom g;
    if (!qkv || !out || !lse || !dout || !dqkv || heads <= 0 || heads >= 65536 ||
        !geom_ok(g, nseq, L, inner, outer_stride, inner_stride, step_stride))
        return LCT_EINVAL;
    // short sequences: several threads per query / key (see the kernel)
    int parts = L <= 64 ? (int)(128 / L) : 1;
    if (parts > 4) parts = 4;
    if (parts < 1) parts = 1;
    size_t smem = ((size_t)4 * L * kHd + 2 * L + (parts > 1 ? (size_t)parts * L * (2 * kHd + 1) : 0)) * sizeof(float);
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)nseq, (unsigned)heads);
    attn_bwd_kernel<<<grid, attn_threads(parts * L), smem, st>>>(qkv, out, lse, dout, dqkv, g, (int)heads,
                                                              1.f / sqrtf((float)kHd), parts);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dpre = dy * act'(y) with y the post-activation value
LCT_API int lct_act_bwd(const float* y, const float* dy, float* dpre, int64_t n, int act, float slope,
                        cudaStream_t st) {
    if (!y || !dy || !dpre || n <= 0) return LCT_EINVAL;
    act_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(y, dy, dpre, n, act, slope);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
