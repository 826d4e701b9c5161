// "Row GEMM" on the tensor cores with fp32-level accuracy (3xTF32 error-compensated mma.sync):
//
//     C[m, n] = epilogue( sum_k A(m, k) * W(k, n) ),   m < M (34 056 positions at B = 8), n <= 64 per CTA, K = 16 .. 384
//
// for the generator's row operators (reference models/generator.py):
//   * nn.Linear / GRU input projections / attention in- and out-projections and their data gradients
//     (GRUblockf :104, :133, :138; GRUblockt :219, :245, :248): A(m, :) is a contiguous row;
//   * encoder Conv2d(k=(2,3), s=(1,2), p=(1,1)) (:461-481) and decoder ConvTranspose2d (:506-529) on channels-last
//     activations, and each as the data gradient of the other: A(m, :) is gathered on the fly from 2 contiguous runs
//     (one per time tap: 3 C_s floats for the convolution; C_s / 2 C_s floats for the even / odd output columns of
//     the transposed convolution) - an implicit GEMM, no im2col buffer.
//
// Why this shape: the SIMT kernels these replace were bound by one shared-memory load per FMA (gconv: 190 us for
// 0.84 GFLOP) or by scalar staging (gemm: 19-73 us for 17-35 MB).  The problems are HBM-sized (3-6 us), so the goal
// is simply "few issue slots per byte": 128-row tiles staged with 16-byte cp.async (double buffered over K chunks of
// 32), fragments read conflict free, and every product computed as a_hi b_hi + a_hi b_lo + a_lo b_hi in TF32 with fp32
// accumulation, which keeps the fp32 parity tolerances (error ~2^-21 per product) at a third of the tensor rate -
// still far below the memory time.
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 32, kThreads = 128;   // 2 x 2 warps, warp tile 32 x 32
constexpr int AS = BK + 4;        // A tile row stride: (4 g + t) mod 32 distinct for g < 8, t < 4
constexpr int WS_T = BN + 8;      // W tile [k][n] row stride: (8 t + g) mod 32 distinct
constexpr int WS_N = BK + 4;      // W tile [n][k] row stride
constexpr int kStages = 2;

enum { RG_LINEAR = 0, RG_CONV = 1, RG_DECONV = 2 };

struct RowGemmParams {
    const float* A; const float* W; float* C;
    const float* bias; const float* res; float* out2; const float* gmul;
    int M, N, K;
    int lda, ldb, ldc, ldr, ldo, ldg;
    int tb;                         // W(k, n) = W[k * ldb + n] if tb else W[n * ldb + k]
    int act; float slope; float alpha; int gact; float gslope;
    int nbatch, a_div, b_div;
    int64_t sA, sB, sC, sBias, sRes, sOut2;
    // gather geometry (RG_CONV / RG_DECONV): in [B, Ti, Fi, Cs] -> out [B, To, Fo, N]
    int B, Ti, Fi, Cs, To, Fo;
    int M1, K1;                     // RG_DECONV: rows / K of the odd-column problem (blockIdx.z = 1); W1 its weights
    const float* W1;
};

__device__ __forceinline__ void cp16(float* dst, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// v = hi + lo with hi = v truncated to tf32 and lo = the exact remainder (< 2^-10 |v|) rounded to tf32 (add half an ulp;
// the tensor core ignores the 13 low mantissa bits).  3 instructions; the cvt.rna.tf32.f32 pair this replaces is emulated
// on sm_100a as FSETP + predicated IADD + LOP3 each (7 with the subtraction) and was ~70 % of the main loop's issue slots.
// Error per product: the dropped lo*lo term and lo's rounding, ~2^-21 relative - unchanged in order of magnitude.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct RowMeta {       // per tile row: where its A runs start (element offsets), which are valid, where its output row is
    int off0, off1;
    int c0;            // first input column of the runs (may be -1)
    int flags;         // bit 0: run 0 row valid, bit 1: run 1 row valid, bit 2: row < M
    int orow;          // output row
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) rowgemm_kernel(const RowGemmParams p) {
    extern __shared__ __align__(16) float sm[];
    float* As = sm;                                          // [kStages][BM][AS]
    float* Wsm = As + kStages * BM * AS;                     // [kStages][max(BK * WS_T, BN * WS_N)]
    constexpr int WTILE = (BK * WS_T > BN * WS_N) ? BK * WS_T : BN * WS_N;
    RowMeta* meta = reinterpret_cast<RowMeta*>(Wsm + kStages * WTILE);   // [BM]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int z = blockIdx.z;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    int M = p.M, K = p.K;
    const float* A = p.A;
    const float* W = p.W;
    float* C = p.C;
    const float* bias = p.bias;
    const float* res = p.res;
    float* out2 = p.out2;
    int par = 0;
    if (MODE == RG_LINEAR) {
        A += (int64_t)(z / p.a_div) * p.sA;
        W += (int64_t)(z / p.b_div) * p.sB;
        C += (int64_t)z * p.sC;
        if (bias) bias += (int64_t)z * p.sBias;
        if (res) res += (int64_t)z * p.sRes;
        if (out2) out2 += (int64_t)z * p.sOut2;
    } else if (MODE == RG_DECONV) {
        par = z;
        if (par) { M = p.M1; K = p.K1; W = p.W1; }
    }
    if (m0 >= M) return;

    // ---- row metadata
    const int Cs = p.Cs;
    const int RL = (MODE == RG_CONV) ? 3 * Cs : (MODE == RG_DECONV ? (par ? 2 * Cs : Cs) : K);   // run length (floats)
    for (int r = tid; r < BM; r += kThreads) {
        RowMeta mt;
        const int m = m0 + r;
        mt.off0 = mt.off1 = 0; mt.c0 = 0; mt.flags = 0; mt.orow = 0;
        if (m < M) {
            if (MODE == RG_LINEAR) {
                mt.off0 = m * p.lda;
                mt.flags = 1 | 4;
                mt.orow = m;
            } else if (MODE == RG_CONV) {
                const int f = m % p.Fo, bt = m / p.Fo, t = bt % p.To, b = bt / p.To;
                mt.c0 = 2 * f - 1;
                const int t0 = t - 1, t1 = t;
                mt.off0 = ((b * p.Ti + t0) * p.Fi + mt.c0) * Cs;
                mt.off1 = ((b * p.Ti + t1) * p.Fi + mt.c0) * Cs;
                mt.flags = ((t0 >= 0 && t0 < p.Ti) ? 1 : 0) | ((t1 >= 0 && t1 < p.Ti) ? 2 : 0) | 4;
                mt.orow = m;
            } else {
                const int Fp = par ? p.Fo / 2 : (p.Fo + 1) / 2;      // output columns of this parity
                const int fh = m % Fp, bt = m / Fp, t = bt % p.To, b = bt / p.To;
                mt.c0 = fh;
                const int t0 = t + 1, t1 = t;                        // kt = 0, 1  ->  input row t + 1 - kt
                mt.off0 = ((b * p.Ti + t0) * p.Fi + fh) * Cs;
                mt.off1 = ((b * p.Ti + t1) * p.Fi + fh) * Cs;
                mt.flags = ((t0 >= 0 && t0 < p.Ti) ? 1 : 0) | ((t1 >= 0 && t1 < p.Ti) ? 2 : 0) | 4;
                mt.orow = (b * p.To + t) * p.Fo + 2 * fh + par;
            }
        }
        meta[r] = mt;
    }
    __syncthreads();

    const int nk = (K + BK - 1) / BK;
    auto stage = [&](int kc, int sbuf) {
        const int k0 = kc * BK;
        float* as = As + sbuf * BM * AS;
        // A: 128 rows x 8 chunks of 16 bytes; 8 consecutive threads cover 128 contiguous bytes of one row
#pragma unroll
        for (int i = 0; i < (BM * BK / 4) / kThreads; ++i) {
            const int e = tid + i * kThreads;
            const int r = e >> 3, c4 = e & 7;
            const int k = k0 + 4 * c4;
            const RowMeta mt = meta[r];
            bool ok = (k < K) && (mt.flags & 4);
            const float* src = A;
            if (MODE == RG_LINEAR) {
                src = A + mt.off0 + k;
            } else {
                const int run = (k >= RL) ? 1 : 0;
                const int j = k - run * RL;
                const int tap = j / Cs;
                ok = ok && ((mt.flags >> run) & 1) && ((unsigned)(mt.c0 + tap) < (unsigned)p.Fi);
                src = A + (run ? mt.off1 : mt.off0) + j;
            }
            cp16(as + r * AS + 4 * c4, ok ? src : A, ok);
        }
        float* ws = Wsm + sbuf * WTILE;
        if (p.tb) {      // [k][n]: 32 rows x 16 chunks
#pragma unroll
            for (int i = 0; i < (BK * BN / 4) / kThreads; ++i) {
                const int e = tid + i * kThreads;
                const int kr = e >> 4, c4 = e & 15;
                const int k = k0 + kr, n = n0 + 4 * c4;
                const bool ok = (k < K) && (n < p.N);        // N is a multiple of 4 on this path
                cp16(ws + kr * WS_T + 4 * c4, ok ? W + (int64_t)k * p.ldb + n : W, ok);
            }
        } else {         // [n][k]: 64 rows x 8 chunks
#pragma unroll
            for (int i = 0; i < (BN * BK / 4) / kThreads; ++i) {
                const int e = tid + i * kThreads;
                const int nr = e >> 3, c4 = e & 7;
                const int k = k0 + 4 * c4, n = n0 + nr;
                const bool ok = (k < K) && (n < p.N);
                cp16(ws + nr * WS_N + 4 * c4, ok ? W + (int64_t)n * p.ldb + k : W, ok);
            }
        }
    };

    const int wm = warp >> 1, wn = warp & 1;
    constexpr int WNT = BN / 16;                             // n-tiles per warp (4)
    const int ntiles = min(WNT, max(0, (p.N - n0 - wn * 32 + 7) / 8));   // n-tiles of this warp holding real columns
    float acc[2][WNT][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    stage(0, 0);
    cp_commit();
    for (int kc = 0; kc < nk; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nk) stage(kc + 1, cur ^ 1);
        cp_commit();
        cp_wait<1>();
        __syncthreads();
        const float* as = As + cur * BM * AS + (wm * 32) * AS;
        const float* ws = Wsm + cur * WTILE;
        const int ksteps = min(BK, K - kc * BK) >> 3;
        for (int ks = 0; ks < ksteps; ++ks) {
            const int kk = ks * 8;
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* ar = as + (mt * 16 + gq) * AS + kk + tq;
                split_tf32(ar[0], ah[mt][0], al[mt][0]);
                split_tf32(ar[8 * AS], ah[mt][1], al[mt][1]);
                split_tf32(ar[4], ah[mt][2], al[mt][2]);
                split_tf32(ar[8 * AS + 4], ah[mt][3], al[mt][3]);
            }
            uint32_t bh[WNT][2], bl[WNT][2];
#pragma unroll
            for (int nt = 0; nt < WNT; ++nt) {
                const int nc = wn * 32 + nt * 8 + gq;
                float w0, w1;
                if (p.tb) {
                    w0 = ws[(kk + tq) * WS_T + nc];
                    w1 = ws[(kk + tq + 4) * WS_T + nc];
                } else {
                    w0 = ws[nc * WS_N + kk + tq];
                    w1 = ws[nc * WS_N + kk + tq + 4];
                }
                split_tf32(w0, bh[nt][0], bl[nt][0]);
                split_tf32(w1, bh[nt][1], bl[nt][1]);
            }
            // the three partial products of every (m-tile, n-tile) as three sweeps: dependent mmas sit 8 apart
#pragma unroll
            for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    if (nt < ntiles) mma8(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);      // small terms first
#pragma unroll
            for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    if (nt < ntiles) mma8(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
            for (int nt = 0; nt < WNT; ++nt)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    if (nt < ntiles) mma8(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
        }
        __syncthreads();      // everyone done with stage `cur` before it is refilled
    }

    // ---- epilogue: thread holds rows (gq, gq + 8) of two m-tiles, columns nt * 8 + 2 tq + {0, 1}
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const RowMeta rm = meta[wm * 32 + mt * 16 + gq + 8 * h];
            if (!(rm.flags & 4)) continue;
            const int64_t orow = rm.orow;
#pragma unroll
            for (int nt = 0; nt < WNT; ++nt) {
                if (nt >= ntiles) continue;
                const int gn = n0 + wn * 32 + nt * 8 + 2 * tq;
                if (gn >= p.N) continue;          // N is even on this path
                float v0 = acc[mt][nt][2 * h] * p.alpha, v1 = acc[mt][nt][2 * h + 1] * p.alpha;
                if (bias) { v0 += bias[gn]; v1 += bias[gn + 1]; }
                v0 = apply_act(v0, p.act, p.slope);
                v1 = apply_act(v1, p.act, p.slope);
                if (p.gmul) {
                    const float2 g = *reinterpret_cast<const float2*>(p.gmul + orow * p.ldg + gn);
                    v0 *= act_grad_from_out(g.x, p.gact, p.gslope);
                    v1 *= act_grad_from_out(g.y, p.gact, p.gslope);
                }
                *reinterpret_cast<float2*>(C + orow * p.ldc + gn) = make_float2(v0, v1);
                if (out2) {
                    float2 r2 = make_float2(0.f, 0.f);
                    if (res) r2 = *reinterpret_cast<const float2*>(res + orow * p.ldr + gn);
                    *reinterpret_cast<float2*>(out2 + orow * p.ldo + gn) = make_float2(v0 + r2.x, v1 + r2.y);
                }
            }
        }
}

constexpr size_t smem_bytes() {
    constexpr int WTILE = (BK * WS_T > BN * WS_N) ? BK * WS_T : BN * WS_N;
    return (size_t)(kStages * BM * AS + kStages * WTILE) * sizeof(float) + BM * sizeof(RowMeta);
}

template <int MODE>
int launch(const RowGemmParams& p, int64_t m_max, int gz, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes());
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    dim3 grid((unsigned)ceil_div64(m_max, BM), (unsigned)ceil_div64(p.N, BN), (unsigned)gz);
    rowgemm_kernel<MODE><<<grid, kThreads, smem_bytes(), st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradients of the row operators:  C[m, n] += sum_r A[r, m] * B[r, n]   (both operands row-major over the
// r = 34 056 positions, C is 48..192 x 16..128 and zero-filled by the caller).  The SIMT split-K kernel these calls used
// took 30-96 us each (12 per backward, ~0.6 ms on the parameter-gradient stream).  Here every CTA walks 32-row chunks
// (16-byte cp.async, double buffered) of a slab of rows, multiplies them on the tensor cores (3xTF32, as above) with the
// output tile held in registers across its 4 warps - WM x WN warps side by side over the output, WK warps interleaved
// over the rows for the tiny 48 x 16 GRU blocks - and finishes with one atomic per (element, warp).
// ---------------------------------------------------------------------------------------------------------------
struct TnParams {
    const float* A; const float* B; float* C;
    int M, N, R;                    // output rows / columns, contraction length (positions)
    int lda, ldb, ldc;
    int a_div, b_div;
    int64_t sA, sB, sC;
    // GATHER (weight gradient of the encoder / decoder convolutions, gen_conv.cu lct_gconv_wgrad): row r = (b, t, f) of
    // the small grid [B, Ts, Fs]; B(r, (kt, kf, c)) = Lg[b, t + kt - 1, 2 f + kf - 1, c] - two contiguous runs of
    // 3 Cc floats, zero outside the large grid [B, Tl, Fl, Cc]; C[m][(kt*3 + kf) * Cc + c] lands in dW[m][c][kt][kf]
    int Ts, Fs, Tl, Fl, Cc;
    uint32_t mulFs, mulTs;          // floor(2^32 / d) + 1: r / Fs and (r / Fs) / Ts by multiply-high (r < 2^22)
};

template <int WM, int WN, int WK, int MT, int NT, bool GATHER>
__global__ void __launch_bounds__(kThreads) rowgemm_tn_kernel(const TnParams p) {
    static_assert(WM * WN * WK == kThreads / 32, "4 warps");
    extern __shared__ __align__(16) float sm[];
    constexpr int RK = 32;                               // rows per chunk
    constexpr int Mc = WM * MT * 16, Nc = WN * NT * 8;
    constexpr int SA = Mc + 8, SB = Nc + 8;              // (8 t + g) mod 32 or (24 t + g) mod 32: conflict-free fragments
    constexpr int STAGE = RK * (SA + SB);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int wk = warp % WK, wn = (warp / WK) % WN, wm = warp / (WK * WN);
    const int z = blockIdx.y;
    const float* A = p.A + (int64_t)(z / p.a_div) * p.sA;
    const float* B = p.B + (int64_t)(z / p.b_div) * p.sB;
    float* C = p.C + (int64_t)z * p.sC;
    const int nchunks = (p.R + RK - 1) / RK;

    auto stage = [&](int chunk, float* buf) {
        const int r0 = chunk * RK;
        float* as = buf;
        float* bs = buf + RK * SA;
        for (int idx = tid; idx < RK * (Mc / 4); idx += kThreads) {
            const int r = idx / (Mc / 4), c4 = idx - r * (Mc / 4);
            const bool ok = (r0 + r) < p.R;
            cp16(as + r * SA + 4 * c4, ok ? A + (int64_t)(r0 + r) * p.lda + 4 * c4 : A, ok);
        }
        for (int idx = tid; idx < RK * (Nc / 4); idx += kThreads) {
            const int r = idx / (Nc / 4), c4 = idx - r * (Nc / 4);
            bool ok = (r0 + r) < p.R;
            const float* src = B;
            if (GATHER) {
                const int rg = r0 + r;
                const int bt = p.Fs == 1 ? rg : (int)__umulhi((uint32_t)rg, p.mulFs);      // b * Ts + t
                const int f = rg - bt * p.Fs;
                const int b = p.Ts == 1 ? bt : (int)__umulhi((uint32_t)bt, p.mulTs);
                const int t = bt - b * p.Ts;
                const int run = 3 * p.Cc / 4;                  // 16-byte pieces per time tap
                const int kt = c4 >= run ? 1 : 0;
                const int o = (c4 - kt * run) * 4;             // float offset inside the run (kf * Cc + c)
                const int tl = t + kt - 1;
                const int fl = 2 * f - 1 + o / p.Cc;
                ok = ok && tl >= 0 && tl < p.Tl && fl >= 0 && fl < p.Fl;
                if (ok) src = B + (((int64_t)b * p.Tl + tl) * p.Fl + (2 * f - 1)) * p.Cc + o;
            } else if (ok) {
                src = B + (int64_t)(r0 + r) * p.ldb + 4 * c4;
            }
            cp16(bs + r * SB + 4 * c4, src, ok);
        }
    };

    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    int chunk = blockIdx.x;
    if (chunk < nchunks) stage(chunk, sm);
    cp_commit();
    int cur = 0;
    for (; chunk < nchunks; chunk += gridDim.x, cur ^= 1) {
        const float* as = sm + cur * STAGE;
        const float* bs = as + RK * SA;
        cp_wait<0>();
        __syncthreads();
        if (chunk + (int)gridDim.x < nchunks) stage(chunk + gridDim.x, sm + (cur ^ 1) * STAGE);
        cp_commit();
#pragma unroll
        for (int ks = wk; ks < RK / 8; ks += WK) {
            const float* a0 = as + (ks * 8 + tq) * SA + wm * MT * 16 + gq;
            const float* b0 = bs + (ks * 8 + tq) * SB + wn * NT * 8 + gq;
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                split_tf32(b0[nt * 8], bh[nt][0], bl[nt][0]);
                split_tf32(b0[4 * SB + nt * 8], bh[nt][1], bl[nt][1]);
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                uint32_t ah[4], al[4];
                split_tf32(a0[mt * 16], ah[0], al[0]);
                split_tf32(a0[mt * 16 + 8], ah[1], al[1]);
                split_tf32(a0[4 * SA + mt * 16], ah[2], al[2]);
                split_tf32(a0[4 * SA + mt * 16 + 8], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma8(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
                    mma8(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
                    mma8(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
                }
            }
        }
    }
    cp_wait<0>();
    if (WK > 1) {
        // the WK row-interleaved warps hold partial sums of the same tile: add them in shared memory first
        __syncthreads();
        float* red = sm;                                 // [WK][Mc * Nc]  (<= the stage buffers)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m = (wm * MT + mt) * 16 + gq + 8 * (i >> 1);
                    const int n = (wn * NT + nt) * 8 + 2 * tq + (i & 1);
                    red[wk * Mc * Nc + m * Nc + n] = acc[mt][nt][i];
                }
        __syncthreads();
        for (int idx = tid; idx < Mc * Nc; idx += kThreads) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < WK; ++w) v += red[w * Mc * Nc + idx];
            const int m = idx / Nc, n = idx - m * Nc;
            if (m < p.M && n < p.N) atomicAdd(&C[(int64_t)m * p.ldc + n], v);      // (WK > 1 is never GATHER)
        }
        return;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = (wm * MT + mt) * 16 + gq + 8 * (i >> 1);
                const int n = (wn * NT + nt) * 8 + 2 * tq + (i & 1);
                if (m < p.M && n < p.N) {
                    if (GATHER) {
                        const int k = n / p.Cc, c = n - k * p.Cc;
                        atomicAdd(&C[((int64_t)m * p.Cc + c) * 6 + k], acc[mt][nt][i]);
                    } else {
                        atomicAdd(&C[(int64_t)m * p.ldc + n], acc[mt][nt][i]);
                    }
                }
            }
}

template <int WM, int WN, int WK, int MT, int NT, bool GATHER = false>
int launch_tn(const TnParams& p, int nbatch, cudaStream_t st) {
    static_assert(!GATHER || WK == 1, "the gather epilogue is the direct one");
    constexpr int Mc = WM * MT * 16, Nc = WN * NT * 8;
    constexpr size_t smem = (size_t)2 * 32 * (Mc + 8 + Nc + 8) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rowgemm_tn_kernel<WM, WN, WK, MT, NT, GATHER>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int nchunks = (p.R + 31) / 32;
    static_assert(WK == 1 || (size_t)WK * Mc * Nc <= (size_t)2 * 32 * (Mc + 8 + Nc + 8), "reduction fits the stages");
    // every CTA ends with M x N atomics: one CTA per SM for the large tiles; the 48 x 16 GRU blocks are latency bound
    // (one k-step per warp and chunk) and want many resident CTAs instead.
    const int target = Mc * Nc >= 2048 ? 148 : 148 * 8;
    int gx = target / nbatch;
    if (gx > nchunks) gx = nchunks;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)nbatch);
    rowgemm_tn_kernel<WM, WN, WK, MT, NT, GATHER><<<grid, kThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

bool al16(const void* q) { return ((uintptr_t)q & 15) == 0; }
bool al8(const void* q) { return ((uintptr_t)q & 7) == 0; }

// weight images for the gather forms: conv  img[(kt, kf, s)][d] = w[d][s][kt][kf];
// deconv even columns img0[(kt, s)][d] = w[s][d][kt][1];  odd columns img1[(kt, e, s)][d] = w[s][d][kt][e ? 0 : 2]
__global__ void gconv_image_kernel(const float* __restrict__ w, float* __restrict__ img, int Cs, int Cd, int transposed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 6 * Cs * Cd) return;
    const int d = idx % Cd;
    int r = idx / Cd;
    if (!transposed) {
        const int s = r % Cs; r /= Cs;
        const int kf = r % 3, kt = r / 3;
        img[idx] = w[((size_t)(d * Cs + s) * 2 + kt) * 3 + kf];
    } else if (r < 2 * Cs) {
        const int s = r % Cs, kt = r / Cs;
        img[idx] = w[((size_t)(s * Cd + d) * 2 + kt) * 3 + 1];
    } else {
        r -= 2 * Cs;
        const int s = r % Cs; r /= Cs;
        const int e = r % 2, kt = r / 2;
        img[idx] = w[((size_t)(s * Cd + d) * 2 + kt) * 3 + (e ? 0 : 2)];
    }
}

}  // namespace

// 1 if lct_rowgemm covers this GEMM (otherwise use lct_gemm's SIMT kernel).  Called by lct_gemm itself.
// Tensor-core form of lct_gconv_wgrad (gen_conv.cu) for the shapes the generator has: (Ca, Cc) = (64, 32), (32, 16).
// Returns 1 if it handled the call (result code in *rc), 0 if the caller should use its SIMT kernels.
int lct_gconv_wgrad_try(const float* S, const float* Lg, float* dW, int64_t B, int64_t Ts, int64_t Fs, int64_t Ca,
                        int64_t Tl, int64_t Fl, int64_t Cc, cudaStream_t st, int* rc) {
    const int64_t R = B * Ts * Fs;
    if (!al16(S) || !al16(Lg) || R >= (1LL << 22) || B * Tl * Fl * Cc >= (1LL << 31)) return 0;
    TnParams t = {};
    t.A = S; t.B = Lg; t.C = dW; t.M = (int)Ca; t.N = (int)(6 * Cc); t.R = (int)R;
    t.lda = (int)Ca; t.ldb = 0; t.ldc = 0; t.a_div = 1; t.b_div = 1;
    t.Ts = (int)Ts; t.Fs = (int)Fs; t.Tl = (int)Tl; t.Fl = (int)Fl; t.Cc = (int)Cc;
    t.mulFs = (uint32_t)((1ULL << 32) / (uint64_t)Fs) + 1u;
    t.mulTs = (uint32_t)((1ULL << 32) / (uint64_t)Ts) + 1u;
    if (Ca == 64 && Cc == 32) *rc = launch_tn<1, 4, 1, 4, 6, true>(t, 1, st);
    else if (Ca == 32 && Cc == 16) *rc = launch_tn<1, 4, 1, 2, 3, true>(t, 1, st);
    else return 0;
    return 1;
}

int lct_rowgemm_try(const float* A, const float* B, float* C, const float* bias, const float* res, float* out2, int64_t M,
                    int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int64_t ldo, int ta, int tb,
                    int act, float slope, float alpha, int accumulate, int64_t ksplit, int64_t nbatch, int64_t a_div,
                    int64_t b_div, int64_t sA, int64_t sB, int64_t sC, int64_t sBias, int64_t sRes, int64_t sOut2,
                    cudaStream_t st, int* rc) {
    if (ta) {
        // weight-gradient form C[M x N] += A^T B over K positions (caller zero-fills C: that is the ksplit > 1 contract)
        if (!tb || ksplit <= 1 || accumulate || bias || res || out2 || act != LCT_ACT_NONE || alpha != 1.f) return 0;
        if ((lda & 3) || (ldb & 3) || (sA & 3) || (sB & 3) || !al16(A) || !al16(B) || nbatch >= 65536 || K >= (1LL << 30))
            return 0;
        TnParams t = {};
        t.A = A; t.B = B; t.C = C; t.M = (int)M; t.N = (int)N; t.R = (int)K;
        t.lda = (int)lda; t.ldb = (int)ldb; t.ldc = (int)ldc; t.a_div = (int)a_div; t.b_div = (int)b_div;
        t.sA = sA; t.sB = sB; t.sC = sC;
        if (lda < M || ldb < N) return 0;
        if (M == 64 && N == 128) *rc = launch_tn<1, 4, 1, 4, 4>(t, (int)nbatch, st);
        else if (M == 64 && N == 64) *rc = launch_tn<2, 2, 1, 2, 4>(t, (int)nbatch, st);
        else if (M == 192 && N == 64) *rc = launch_tn<4, 1, 1, 3, 8>(t, (int)nbatch, st);
        else if (M == 48 && N == 16) *rc = launch_tn<1, 1, 4, 3, 2>(t, (int)nbatch, st);
        else return 0;
        return 1;
    }
    if (accumulate || ksplit != 1) return 0;
    if ((K & 7) || (N & 3) || (lda & 3) || (ldb & 3) || (ldc & 1) || (sA & 3) || (sB & 3) || (sC & 1)) return 0;
    if (!al16(A) || !al16(B) || !al8(C) || M * lda >= (1LL << 31) || nbatch >= 65536) return 0;
    if (out2 && (!al8(out2) || (ldo & 1) || (sOut2 & 1))) return 0;
    if (res && (!al8(res) || (ldr & 1) || (sRes & 1))) return 0;
    RowGemmParams p = {};
    p.A = A; p.W = B; p.C = C; p.bias = bias; p.res = res; p.out2 = out2;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.lda = (int)lda; p.ldb = (int)ldb; p.ldc = (int)ldc; p.ldr = (int)ldr; p.ldo = (int)ldo;
    p.tb = tb; p.act = act; p.slope = slope; p.alpha = alpha;
    p.nbatch = (int)nbatch; p.a_div = (int)a_div; p.b_div = (int)b_div;
    p.sA = sA; p.sB = sB; p.sC = sC; p.sBias = sBias; p.sRes = sRes; p.sOut2 = sOut2;
    *rc = launch<RG_LINEAR>(p, M, (int)nbatch, st);
    return 1;
}

// Number of floats of the weight image lct_gconv_weight_image writes (6 * Cs * Cd).
LCT_API int lct_gconv_image_len(int64_t Cs, int64_t Cd) { return (int)(6 * Cs * Cd); }

// transposed = 0: w [Cd][Cs][2][3] (Conv2d weight, or a ConvTranspose2d weight used for its data gradient);
// transposed = 1: w [Cs][Cd][2][3].  img: 6 * Cs * Cd floats.
LCT_API int lct_gconv_weight_image(const float* w, float* img, int transposed, int64_t Cs, int64_t Cd, cudaStream_t st) {
    if (!w || !img || Cs <= 0 || Cd <= 0) return LCT_EINVAL;
    const int n = (int)(6 * Cs * Cd);
    gconv_image_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, img, (int)Cs, (int)Cd, transposed);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// 1 if lct_gconv_mma covers the layer: Cs a multiple of 4 (16-byte taps), Cd a multiple of 4 up to 64
LCT_API int lct_gconv_mma_supported(int64_t Cs, int64_t Cd) { return (Cs >= 4 && (Cs & 3) == 0 && Cd >= 8 && (Cd & 3) == 0 && Cd <= 64) ? 1 : 0; }

// Same operator as lct_gconv (gen_conv.cu) through the implicit row GEMM; img from lct_gconv_weight_image.
LCT_API int lct_gconv_mma(const float* in, const float* img, const float* bias, float* out, const float* gmul,
                          int transposed, int64_t B, int64_t Ti, int64_t Fi, int64_t Cs, int64_t To, int64_t Fo, int64_t Cd,
                          int act, float slope, int gact, float gslope, cudaStream_t st) {
    if (!in || !img || !out || B <= 0 || Ti <= 0 || Fi <= 0 || To <= 0 || Fo <= 0 || !lct_gconv_mma_supported(Cs, Cd))
        return LCT_EINVAL;
    if (!al16(in) || !al16(img) || !al8(out) || (gmul && !al8(gmul))) return LCT_EUNSUPPORTED;
    if (B * Ti * Fi * Cs >= (1LL << 31) || B * To * Fo * Cd >= (1LL << 31)) return LCT_EUNSUPPORTED;
    RowGemmParams p = {};
    p.A = in; p.C = out; p.bias = bias; p.gmul = gmul;
    p.N = (int)Cd; p.ldb = (int)Cd; p.ldc = (int)Cd; p.ldg = (int)Cd; p.tb = 1;
    p.act = act; p.slope = slope; p.alpha = 1.f; p.gact = gact; p.gslope = gslope;
    p.B = (int)B; p.Ti = (int)Ti; p.Fi = (int)Fi; p.Cs = (int)Cs; p.To = (int)To; p.Fo = (int)Fo;
    if (!transposed) {
        p.W = img; p.M = (int)(B * To * Fo); p.K = (int)(6 * Cs);
        return launch<RG_CONV>(p, p.M, 1, st);
    }
    p.W = img; p.M = (int)(B * To * ((Fo + 1) / 2)); p.K = (int)(2 * Cs);
    p.W1 = img + 2 * Cs * Cd; p.M1 = (int)(B * To * (Fo / 2)); p.K1 = (int)(4 * Cs);
    return launch<RG_DECONV>(p, p.M > p.M1 ? p.M : p.M1, 2, st);
}
