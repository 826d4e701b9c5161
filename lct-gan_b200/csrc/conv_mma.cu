// Grouped strided convolutions of the waveform discriminators on the tensor cores (TF32 mma.sync,
// fp32 accumulate), forward + data gradient + weight gradient.
//
// Same operator as conv_disc.cu (reference models/discriminators.py:37-67, :93-98, :166-196, :215-220):
// a convolution along L of [B, C, L, P] with groups, stride S, "same" padding.  conv_disc.cu is the
// fp32 SIMT version (kept for exact-fp32 mode); profiling showed it latency/issue bound (fma pipe 18-41 %,
// top stall long_scoreboard, < 1 wave per launch), so this file restructures the work as an implicit GEMM:
//
//   * per group:  D[position, n] = sum_kk A[position, kk] * W[kk, n]   with A gathered on the fly from a
//     shared-memory window of the input (no im2col buffer).  Forward: kk = (ci, tap), n = out channel.
//     Data gradient in polyphase form: kk = (co, t), n = (in channel, stride phase r) with r fastest
//     (weight tap k = r + S t), so one GEMM produces all S phases.
//   * persistent CTAs (one group each, looping over (batch, position tile); grid = the CTAs that are really resident,
//     rounded down to one wave) with the next tile's window prefetched by 16-byte cp.async (per-channel alignment
//     shift, zero-fill for the padding) while the current one is multiplied.
//   * m16n8k8 TF32 fragments are read straight from the window through a LUT of byte offsets (rows = 8 consecutive
//     positions, columns = 4 consecutive taps -> 32 consecutive words for S <= 4: conflict free; the data gradient's
//     rows sit at immediate offsets) and rounded to tf32 by one integer add; weights come pre-arranged and pre-rounded
//     from the weight-norm launch.
//   * data-gradient epilogue: the accumulator tile is transposed through shared memory so that every in-channel owns
//     one contiguous run of outputs, then a coalesced pass applies (+ FM gradient) x LeakyReLU'(saved activation)
//     with 16 unconditional loads in flight per thread.
//   * weight gradient: M = out channels, N = (ci, tap), K = positions; the dW tile of a group lives in registers,
//     split by columns over two warp groups when it is wide (8 warps per CTA), and is flushed with one atomic per
//     (weight, CTA).
//   (round-1 ncu source pages behind these choices: profiles/README.md)
//
// Precision: TF32 operands (10-bit mantissa), fp32 accumulation - the "bf16 training" configuration of
// BASELINE.json (configs[2]) at higher operand precision than bf16.
#include <cstdlib>
#include <type_traits>
#include "common.cuh"

namespace {

constexpr int kThreads = 128;   // 4 warps
enum { MODE_FWD = 0, MODE_DGRAD = 1 };

// n / d by multiplication (d small and n * d < 2^32: rows of P, batch index of a tile); d == 1 and the unsafe range
// fall back to the identity / a real unsigned division.  The signed `j / p.P` sequences were ~20 instructions each,
// several per thread and tile.
struct FastDiv { uint32_t mul, d; };
__device__ __forceinline__ int fdiv(int n, FastDiv f) {
    return f.d == 1 ? n : (f.mul ? (int)__umulhi((uint32_t)n, f.mul) : (int)((uint32_t)n / f.d));
}
FastDiv make_fdiv(int64_t d, int64_t max_n) {
    FastDiv f;
    f.d = (uint32_t)d;
    f.mul = (d > 1 && max_n * d < (1LL << 32)) ? (uint32_t)((1ULL << 32) / (uint64_t)d) + 1u : 0u;
    return f;
}

struct MmaParams {
    const float* x;        // gathered tensor  [B][Cx][Lx][P]   (fwd: input, dgrad: dY)
    const float* w;        // conv weight      [Cout][Cin_g][K]
    const float* wimg;     // optional: the staged image [G][KKpad][NS] (tf32-rounded, zero padded) built by the
                           // weight-norm kernel (lct_mt_weight_norm_fwd); skips the in-kernel arrangement
    float* out;            // fwd: y [B][Cout][Lout][P];  dgrad: dx [B][Cin][Lin][P]
    const float* bias;     // fwd
    const float* gextra;   // dgrad (optional)
    const float* xact;     // dgrad (optional)
    int B, Cx, Cxg, Lx, P, Co, N, Lo;      // Co = channels of `out`, N = GEMM columns per group (fwd: Cout/G; dgrad: Cin/G * S)
    int Cig;                               // dgrad: in-channels per group
    int K, S, pad, Tmax;                   // conv geometry (Tmax = ceil(K / S))
    int Sg, pad_eff, Tspan, opad, Q;       // gather stride, window origin shift, taps spanned, output shift, rows
    int KK, KKpad, Npad, NS, WS;
    int act; float slope;
    int tiles_per_b, ntiles;
    int TPe;                               // positions per tile (dgrad: whole rows of P only, <= the CTA's 64*MTW rows)
    int ES;                                // dgrad: plane stride of the staged output tile
    int ot_alias;                          // dgrad: the staged tile fits into (and reuses) a window buffer
    FastDiv fP, fT;                        // division by P / by tiles_per_b
};

// fp32 -> tf32, round to nearest / ties away (= cvt.rna.tf32.f32 for every finite input): add half a tf32 ulp to the
// magnitude; the tensor core ignores the 13 low mantissa bits.  cvt.rna itself is emulated on sm_100a as FSETP(|v|<inf)
// + predicated integer add, twice the instructions (ncu source page of round 1), in the innermost loop.
__device__ __forceinline__ uint32_t f2tf32(float v) { return __float_as_uint(v) + 0x1000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(src) : "memory");
}
// Copy n_e floats src[f0 .. f0+n_e) (zero outside [0, lim)) to dst[0 .. n_e) with one warp.  dst and src + f0 must
// have the same alignment phase modulo 16 bytes (the caller shifts dst by ((src + f0) & 3) floats), so the body
// moves 16 bytes per cp.async: the per-element 4-byte version was the bottleneck of these kernels.
__device__ __forceinline__ void warp_copy_row(float* dst, const float* src, int f0, int n_e, int lim, int phase,
                                              const float* safe, int lane) {
    const int h0 = min((4 - phase) & 3, n_e);
    if (lane < h0) {
        const int f = f0 + lane;
        const bool ok = f >= 0 && f < lim;
        cp_async4(dst + lane, ok ? src + f : safe, ok);
    }
    const int nvec = (n_e - h0) >> 2;
    for (int v = lane; v < nvec; v += 32) {
        const int e = h0 + 4 * v, f = f0 + e;
        if (f >= 0 && f + 3 < lim) {
            cp_async16(dst + e, src + f);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = (f + i) >= 0 && (f + i) < lim;
                cp_async4(dst + e + i, ok ? src + f + i : safe, ok);
            }
        }
    }
    for (int e = h0 + 4 * nvec + lane; e < n_e; e += 32) {
        const int f = f0 + e;
        const bool ok = f >= 0 && f < lim;
        cp_async4(dst + e, ok ? src + f : safe, ok);
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage the input window of `tile` (all Cxg channels of group g) into `buf` with cp.async and write the tile's
// gather LUT: lutd[kk] = lut0[kk] + (alignment shift of kk's channel)
__device__ __forceinline__ void stage_window(const MmaParams& p, int g, int tile, float* buf, int* lutd,
                                             const int* lut0, const int* chk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = fdiv(tile, p.fT), jt = tile - b * p.tiles_per_b;
    const int j0 = jt * p.TPe;
    const int jend = min(j0 + p.TPe, p.Q * p.P);
    const int row_first = fdiv(j0, p.fP), row_last = fdiv(jend - 1, p.fP);
    const int n_e = ((row_last - row_first) * p.Sg + p.Tspan) * p.P;
    const int e0 = (row_first * p.Sg - p.pad_eff) * p.P;   // flat start inside a channel (may be negative)
    const int lim = p.Lx * p.P;
    const int64_t cb0 = ((int64_t)b * p.Cx + (int64_t)g * p.Cxg) * lim;
    for (int ch = warp; ch < p.Cxg; ch += kThreads / 32) {
        const int64_t cb = cb0 + (int64_t)ch * lim;
        const int phase = (int)((cb + e0) & 3);
        warp_copy_row(buf + ch * p.WS + phase, p.x + cb, e0, n_e, lim, phase, p.x, lane);
    }
    for (int kk = threadIdx.x; kk < p.KKpad; kk += kThreads) {
        const int64_t cb = cb0 + (int64_t)chk[kk] * lim;
        lutd[kk] = (lut0[kk] + (int)((cb + e0) & 3)) * 4;   // byte offset
    }
}

template <int MODE, int NT, int MTW>
__global__ void __launch_bounds__(kThreads) conv_mma_kernel(const MmaParams p) {
    extern __shared__ __align__(16) float sm[];
    // MTW m-tiles of 16 positions per warp: 64 * MTW positions per tile (p.TPe of them used)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int g = blockIdx.x;
    float* wsm = sm;                                              // [KKpad][NS]
    int* lut0 = reinterpret_cast<int*>(wsm + p.KKpad * p.NS);       // [KKpad] static part of the gather LUT
    int* chk = lut0 + p.KKpad;                                    // [KKpad] window channel of kk
    int* lutd0 = chk + p.KKpad;                                   // [2][KKpad] per-buffer LUT (with alignment shift)
    int* lutd1 = lutd0 + p.KKpad;
    float* win0 = reinterpret_cast<float*>(lutd1 + p.KKpad);
    const int winsz = p.Cxg * p.WS;
    float* win1 = win0 + winsz;
    float* otile = win1 + winsz;                                  // dgrad: [Cig][ES] staged output tile

    // ---- prologue: this group's weights (tf32-rounded, [kk][n]) and the gather LUT
    if (p.wimg) {
        const float* src = p.wimg + (size_t)g * p.KKpad * p.NS;     // 16-byte aligned: KKpad * NS is a multiple of 64
        for (int idx = tid * 4; idx < p.KKpad * p.NS; idx += kThreads * 4) cp_async16(wsm + idx, src + idx);
    } else
    for (int idx = tid; idx < p.KKpad * p.Npad; idx += kThreads) {
        const int n = idx % p.Npad;
        const int kk = idx / p.Npad;
        float v = 0.f;
        if (kk < p.KK && n < p.N) {
            if (MODE == MODE_FWD) {
                // kk = ci*K + tap; w[(g*N + n)][ci][tap] is contiguous in kk
                v = p.w[((size_t)g * p.N + n) * p.KK + kk];
            } else {
                // kk = co*Tmax + t;  n = ci*S + r;  tap k = r + S*t
                const int co = kk / p.Tmax, t = kk - co * p.Tmax;
                const int ci = n / p.S, r = n - ci * p.S;
                const int k = r + p.S * t;
                if (k < p.K) v = p.w[(((size_t)g * p.Cxg + co) * p.Cig + ci) * p.K + k];
            }
        }
        wsm[kk * p.NS + n] = __uint_as_float(f2tf32(v));
    }
    for (int kk = tid; kk < p.KKpad; kk += kThreads) {
        int off = 0, chn = 0;
        if (kk < p.KK) {
            if (MODE == MODE_FWD) {
                const int ci = kk / p.K, tap = kk - ci * p.K;
                off = ci * p.WS + tap * p.P;
                chn = ci;
            } else {
                const int co = kk / p.Tmax, t = kk - co * p.Tmax;
                off = co * p.WS + (p.Tmax - 1 - t) * p.P;
                chn = co;
            }
        }
        lut0[kk] = off;
        chk[kk] = chn;
    }
    __syncthreads();
    // this thread's accumulator columns: n = nt*8 + 2*tq + c  ->  output offset of the column (channel plane + row
    // shift), bias;  dgrad: offset inside the staged output tile
    int col_off[NT][2];
    bool col_ok[NT][2];
    float col_bias[NT][2];
    const int lo_p = p.Lo * p.P;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int n = nt * 8 + 2 * tq + c;
            col_ok[nt][c] = n < p.N;
            col_off[nt][c] = 0;
            col_bias[nt][c] = 0.f;
            if (n < p.N) {
                if (MODE == MODE_FWD) {
                    col_off[nt][c] = (g * p.N + n) * lo_p;
                    if (p.bias) col_bias[nt][c] = p.bias[g * p.N + n];
                } else {
                    // offset inside the staged output tile: plane of the in-channel, row shift of the stride phase
                    const int ci = n / p.S;
                    col_off[nt][c] = ci * p.ES + (n - ci * p.S) * p.P;
                }
            }
        }
    const bool has_g = p.gextra != nullptr, has_x = p.xact != nullptr;
    // act'(saved output) = 1 where it is positive, else: LeakyReLU slope / 0 for ReLU / 1 without activation
    const float dslope = !has_x ? 1.f : (p.act == LCT_ACT_LRELU ? p.slope : (p.act == LCT_ACT_RELU ? 0.f : 1.f));
    int tile = blockIdx.y;
    if (tile < p.ntiles) stage_window(p, g, tile, win0, lutd0, lut0, chk);
    cp_async_commit();

    int cur = 0;
    for (; tile < p.ntiles; tile += gridDim.y, cur ^= 1) {
        float* win = cur ? win1 : win0;
        const int* lut = cur ? lutd1 : lutd0;
        cp_async_wait_all();
        __syncthreads();                       // window + LUT `cur` (and, first time, the weights) visible to all
        const int nxt = tile + gridDim.y;
        if (nxt < p.ntiles) stage_window(p, g, nxt, cur ? win0 : win1, cur ? lutd0 : lutd1, lut0, chk);
        cp_async_commit();

        const int b = fdiv(tile, p.fT), jt = tile - b * p.tiles_per_b;
        const int j0 = jt * p.TPe;
        const int jtot = min(p.Q * p.P, j0 + p.TPe);
        const int row_first = fdiv(j0, p.fP);
        if (MODE == MODE_DGRAD && (has_g || has_x)) {
            // The epilogue below reads the FM gradient / saved activation of this tile's output run in 2-4 dependent
            // rounds (16 loads per thread each): ask the L2 for those lines now, so that the rounds cost L2 hits instead
            // of DRAM round trips with nothing else of this CTA in flight (ncu: long_scoreboard was the top stall).
            const int nrows_p = fdiv(jtot - j0, p.fP);
            const int E_p = nrows_p * p.S * p.P;
            const int gpos_p = (p.S * row_first - p.opad) * p.P;
            const int lo_c = max(gpos_p, 0), hi_c = min(gpos_p + E_p, lo_p);          // the run, clipped to the map
            const int cb_p = b * p.Co * lo_p + g * p.Cig * lo_p;
            const int lines = (hi_c - lo_c + 31) / 32 + 1;     // 128-byte lines per channel (+ 1: the run is not aligned)
            for (int i = tid; i < p.Cig * lines; i += kThreads) {
                const int ci = i / lines, ln = i - ci * lines;
                const int off = cb_p + ci * lo_p + min(lo_c + 32 * ln, hi_c - 1);
                if (has_g) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gextra + off));
                if (has_x) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.xact + off));
            }
        }
        // per-thread rows: m-tile mt, half h -> position j0 + warp*16*MTW + mt*16 + gq + 8h
        // abase = window address of the row (LUT entries are byte offsets);  orow = output offset of the row
        const char* abase[MTW][2];
        int orow[MTW][2];
        bool rok[MTW][2];
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + warp * 16 * MTW + mt * 16 + gq + 8 * h;
                rok[mt][h] = j < jtot;
                const int rq = fdiv(j, p.fP), pp = j - rq * p.P;
                // dgrad gathers with row stride 1: the window offset is linear in j, rows past the end of the map
                // read (and discard) in-bounds garbage; forward rows past the end are pointed at the window start
                const int base = MODE == MODE_FWD ? (rok[mt][h] ? (rq - row_first) * p.Sg * p.P + pp : 0)
                                                  : j - row_first * p.P;
                abase[mt][h] = reinterpret_cast<const char*>(win + base);
                // forward: offset of the row in an output plane;  dgrad: in a plane of the staged tile
                orow[mt][h] = MODE == MODE_FWD ? rq * p.P + pp : (rq - row_first) * p.S * p.P + pp;
            }
        float acc[MTW][NT][4];
#pragma unroll
        for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

        const float* wr0 = wsm + tq * p.NS + gq;
        const float* wr1 = wr0 + 4 * p.NS;
        const int wstep = 8 * p.NS;
        for (int k0 = 0; k0 < p.KKpad; k0 += 8, wr0 += wstep, wr1 += wstep) {
            const int l0 = lut[k0 + tq], l1 = lut[k0 + tq + 4];
            uint32_t a[MTW][4];
            if (MODE == MODE_FWD) {
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) {
                    a[mt][0] = f2tf32(*reinterpret_cast<const float*>(abase[mt][0] + l0));
                    a[mt][1] = f2tf32(*reinterpret_cast<const float*>(abase[mt][1] + l0));
                    a[mt][2] = f2tf32(*reinterpret_cast<const float*>(abase[mt][0] + l1));
                    a[mt][3] = f2tf32(*reinterpret_cast<const float*>(abase[mt][1] + l1));
                }
            } else {
                // rows are 8 / 16 / 24 ... floats apart: one address per LUT entry, immediates for the rest
                const float* a0 = reinterpret_cast<const float*>(abase[0][0] + l0);
                const float* a1 = reinterpret_cast<const float*>(abase[0][0] + l1);
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) {
                    a[mt][0] = f2tf32(a0[16 * mt]);
                    a[mt][1] = f2tf32(a0[16 * mt + 8]);
                    a[mt][2] = f2tf32(a1[16 * mt]);
                    a[mt][3] = f2tf32(a1[16 * mt + 8]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint32_t b0 = __float_as_uint(wr0[nt * 8]);
                const uint32_t b1 = __float_as_uint(wr1[nt * 8]);
#pragma unroll
                for (int mt = 0; mt < MTW; ++mt) mma_tf32(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
            }
        }
        const int bbase = b * p.Co * lo_p;
        if (MODE == MODE_FWD) {
            // ---- forward epilogue: bias + activation; 8 consecutive positions of a channel per 8 lanes (full sectors)
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const float v = apply_act(acc[mt][nt][2 * h + c] + col_bias[nt][c], p.act, p.slope);
                            if (rok[mt][h] && col_ok[nt][c]) p.out[bbase + col_off[nt][c] + orow[mt][h]] = v;
                        }
        } else {
            // ---- data-gradient epilogue.  The accumulator layout scatters a channel's run over lanes (stride-S rows,
            // 2 of 4 lanes per phase pair): written straight to global memory it cost 3-4x the sectors in loads of the
            // FM gradient / saved activation and in partial-sector stores, one element at a time.  Instead the tile is
            // transposed through shared memory (a tile covers whole rows of P, so every in-channel owns ONE contiguous
            // run of nrows*S*P outputs) and finished by a coalesced pass:  dx = (tile + gextra) * act'(xact), with the
            // 16 loads of a thread issued before their first use.
            // the staged tile lives in the window that was just consumed when it fits there (one more barrier), else in
            // its own buffer
            float* ot = p.ot_alias ? win : otile;
            if (p.ot_alias) __syncthreads();
#pragma unroll
            for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            if (rok[mt][h] && col_ok[nt][c]) ot[col_off[nt][c] + orow[mt][h]] = acc[mt][nt][2 * h + c];
            __syncthreads();
            const int nrows = fdiv(jtot - j0, p.fP);                // whole rows (TPe is a multiple of P)
            const int E = nrows * p.S * p.P;
            const int gpos0 = (p.S * row_first - p.opad) * p.P;   // position of the run's first element in its plane
            const int cbase = bbase + g * p.Cig * lo_p + gpos0;
            const float* __restrict__ gep = p.gextra;
            const float* __restrict__ xap = p.xact;
            float* __restrict__ outp = p.out;
            // 4 channels x 2 elements (128 apart) per thread and step.  The loads are unconditional (elements outside the
            // tile / the map read a valid dummy address) and specialised on which tensors exist, so that the compiler
            // issues all 16 of them before the first use; only the stores are predicated.  (A predicated version was
            // compiled into load -> compare -> next load chains.)
            const int safe = cbase + min(max(-gpos0, 0), lo_p - 1 - gpos0);   // (32-bit offsets: B*C*L*P < 2^31)
            auto pass = [&](auto HG, auto HX) {
                constexpr int CU = 4, EU = 2;
                for (int ci0 = 0; ci0 < p.Cig; ci0 += CU)
                    for (int e0 = tid; e0 < E; e0 += EU * kThreads) {
                        float ge[CU][EU], xa[CU][EU];
                        int gi[CU][EU];
                        bool ok[CU][EU];
#pragma unroll
                        for (int q = 0; q < EU; ++q) {
                            const int e = e0 + q * kThreads;
                            const bool eok = e < E && (unsigned)(gpos0 + e) < (unsigned)lo_p;
#pragma unroll
                            for (int u = 0; u < CU; ++u) {
                                ok[u][q] = eok && (ci0 + u) < p.Cig;
                                gi[u][q] = ok[u][q] ? cbase + (ci0 + u) * lo_p + e : safe;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < EU; ++q)
#pragma unroll
                            for (int u = 0; u < CU; ++u) {
                                ge[u][q] = decltype(HG)::value ? __ldg(gep + gi[u][q]) : 0.f;
                                xa[u][q] = decltype(HX)::value ? __ldg(xap + gi[u][q]) : 1.f;
                            }
#pragma unroll
                        for (int q = 0; q < EU; ++q)
#pragma unroll
                            for (int u = 0; u < CU; ++u)
                                if (ok[u][q])
                                    outp[gi[u][q]] = (ot[(ci0 + u) * p.ES + e0 + q * kThreads] + ge[u][q]) *
                                                     (xa[u][q] > 0.f ? 1.f : dslope);
                    }
            };
            if (has_g && has_x) pass(std::true_type{}, std::true_type{});
            else if (has_x) pass(std::false_type{}, std::true_type{});
            else if (has_g) pass(std::true_type{}, std::false_type{});
            else pass(std::false_type{}, std::false_type{});
            // (the next tile's scatter comes after the barrier at the top of the loop: no second barrier needed)
        }
    }
    cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------
// weight gradient:  dW[g*N + n][ci][tap] += sum_{b, pos} dY[b][g*N + n][pos] * X[b][g*Cig + ci][row(pos)*S + tap - pad]
// mma roles: M = out channel (MT m-tiles of 16), N = kk = (ci, tap) (NT n-tiles of 8), K = positions
// ---------------------------------------------------------------------------------------------
struct MmaWgradParams {
    const float* x;     // [B][Cin][Lin][P]
    const float* dy;    // [B][Cout][Lout][P]
    float* dw;          // [Cout][Cin_g][K]  (accumulated)
    float* db;          // [Cout] (accumulated, optional)
    int B, Cin, Cig, Lin, P, Cout, N, Lout;
    int K, S, pad;
    int KK, WS, DS;
    int tiles_per_b, ntiles;
    FastDiv fP, fT;                        // division by P / by tiles_per_b
};

// NT n-tiles per warp, NSPLIT warp groups side by side over the (ci, tap) columns: 4 * NSPLIT warps per CTA.  (With all
// 21 n-tiles of the MSD layers in one warp the kernel needed 195 registers -> 8 resident warps per SM, and it is
// bound by the latency of its shared-memory gathers; two groups of 11 n-tiles halve the accumulators.)
// TP = positions per tile (128 or 256): the longer tile halves the barriers and window halos per position and won
// 15-20 % on the k = 41 scale layers; it lost as much on the short maps and on the two-m-tile period layers, and on the
// 512 -> 1024 period layer its shared memory (84 KB) costs more resident CTAs than it gains (36 us against 29 us with
// four 128-position CTAs per SM; tools/bench_disc_layers.py) - chosen per layer in launch_wgrad
template <int MT, int NT, int NSPLIT, int TP>
__global__ void __launch_bounds__(kThreads * NSPLIT) conv_mma_wgrad_kernel(const MmaWgradParams p) {
    extern __shared__ __align__(16) float sm[];
    // TP positions per tile; each position warp takes a quarter of them
    constexpr int NW = 4 * NSPLIT;
    constexpr int NTH = kThreads * NSPLIT;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wp = warp & 3, wn = warp >> 2;
    const int gq = lane >> 2, tq = lane & 3;
    const int g = blockIdx.x;
    const int winsz = p.Cig * p.WS;
    const int dysz = 16 * MT * p.DS;
    float* win0 = sm;
    float* win1 = win0 + winsz;
    float* dy0 = win1 + winsz;
    float* dy1 = dy0 + dysz;
    int* shw0 = reinterpret_cast<int*>(dy1 + dysz);   // [2][Cig]   alignment shift of every window channel
    int* shw1 = shw0 + p.Cig;
    int* shd0 = shw1 + p.Cig;                         // [2][16*MT] alignment shift of every dY row
    int* shd1 = shd0 + 16 * MT;

    // LUT of this thread's B columns (kk = (wn*NT + nt)*8 + gq), constant over tiles (the per-tile shift is added below)
    int lutn[NT], cin[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int kk = (wn * NT + nt) * 8 + gq;
        int off = 0, ci = 0;
        if (kk < p.KK) {
            ci = kk / p.K;
            off = ci * p.WS + (kk - ci * p.K) * p.P;
        }
        lutn[nt] = off;
        cin[nt] = ci;
    }
    float acc[MT][NT][4];
    float sdy[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        sdy[mt][0] = sdy[mt][1] = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    }
    const int jtot = p.Lout * p.P;

    auto stage = [&](int tile, float* wbuf, float* dbuf, int* shw, int* shd) {
        const int b = fdiv(tile, p.fT), jt = tile - b * p.tiles_per_b;
        const int j0 = jt * TP;
        const int jend = min(j0 + TP, jtot);
        const int row_first = fdiv(j0, p.fP), row_last = fdiv(jend - 1, p.fP);
        const int n_e = ((row_last - row_first) * p.S + p.K) * p.P;
        const int e0 = (row_first * p.S - p.pad) * p.P;
        const int lim = p.Lin * p.P;
        for (int ch = warp; ch < p.Cig; ch += NW) {
            const int64_t cb = ((int64_t)b * p.Cin + (int64_t)g * p.Cig + ch) * lim;
            const int phase = (int)((cb + e0) & 3);
            warp_copy_row(wbuf + ch * p.WS + phase, p.x + cb, e0, n_e, lim, phase, p.x, lane);
            if (lane == 0) shw[ch] = phase;
        }
        for (int oc = warp; oc < 16 * MT; oc += NW) {
            const bool chok = oc < p.N;
            const int64_t cb = ((int64_t)b * p.Cout + (int64_t)g * p.N + (chok ? oc : 0)) * jtot;
            const int phase = (int)((cb + j0) & 3);
            // rows of padded out-channels are zero: give them an empty valid range
            warp_copy_row(dbuf + oc * p.DS + phase, p.dy + cb, j0, TP, chok ? jtot : 0, phase, p.dy, lane);
            if (lane == 0) shd[oc] = phase;
        }
    };

    int tile = blockIdx.y;
    if (tile < p.ntiles) stage(tile, win0, dy0, shw0, shd0);
    cp_async_commit();
    int cur = 0;
    for (; tile < p.ntiles; tile += gridDim.y, cur ^= 1) {
        const float* win = cur ? win1 : win0;
        const float* dyt = cur ? dy1 : dy0;
        const int* shw = cur ? shw1 : shw0;
        const int* shd = cur ? shd1 : shd0;
        cp_async_wait_all();
        __syncthreads();
        const int nxt = tile + gridDim.y;
        if (nxt < p.ntiles) stage(nxt, cur ? win0 : win1, cur ? dy0 : dy1, cur ? shw0 : shw1, cur ? shd0 : shd1);
        cp_async_commit();
        int lc[NT], sd[MT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) lc[nt] = (lutn[nt] + shw[cin[nt]]) * 4;      // byte offsets
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            sd[mt][0] = shd[mt * 16 + gq];
            sd[mt][1] = shd[mt * 16 + gq + 8];
        }

        const int jt = tile - fdiv(tile, p.fT) * p.tiles_per_b;
        const int j0 = jt * TP;
        const int row_first = fdiv(j0, p.fP);
        // this warp's TP / 4 positions in k-steps of 8
#pragma unroll
        for (int ks = 0; ks < TP / 32; ++ks) {
            const int pl = wp * (TP / 4) + ks * 8;           // tile-local position of this k-step
            int bs[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = j0 + pl + tq + 4 * h;
                if (j < jtot) {
                    const int rq = fdiv(j, p.fP), pp = j - rq * p.P;
                    bs[h] = (rq - row_first) * p.S * p.P + pp;
                } else {
                    bs[h] = 0;                               // dY is zero there
                }
            }
            uint32_t a[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const float* d0 = dyt + (mt * 16 + gq) * p.DS + sd[mt][0] + pl;
                const float* d1 = dyt + (mt * 16 + gq + 8) * p.DS + sd[mt][1] + pl;
                const float v0 = d0[tq], v1 = d1[tq], v2 = d0[tq + 4], v3 = d1[tq + 4];
                sdy[mt][0] += v0 + v2;
                sdy[mt][1] += v1 + v3;
                a[mt][0] = f2tf32(v0); a[mt][1] = f2tf32(v1); a[mt][2] = f2tf32(v2); a[mt][3] = f2tf32(v3);
            }
            const char* wb0 = reinterpret_cast<const char*>(win + bs[0]);
            const char* wb1 = reinterpret_cast<const char*>(win + bs[1]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint32_t b0 = f2tf32(*reinterpret_cast<const float*>(wb0 + lc[nt]));
                const uint32_t b1 = f2tf32(*reinterpret_cast<const float*>(wb1 + lc[nt]));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
            }
        }
    }
    cp_async_wait_all();
    // ---- flush: reduce the 4 position warps' partial tiles in shared memory (the window buffers are free now),
    // then one atomic per (weight, CTA)
    __syncthreads();
    float* red = sm;                                   // [16*MT][NT*NSPLIT*8]
    constexpr int RS = NT * NSPLIT * 8;
    for (int wv = 0; wv < 4; ++wv) {
        if (wp == wv) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int oc = mt * 16 + gq + 8 * h;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int kk = (wn * NT + nt) * 8 + 2 * tq + c;
                            if (wv == 0) red[oc * RS + kk] = acc[mt][nt][2 * h + c];
                            else red[oc * RS + kk] += acc[mt][nt][2 * h + c];
                        }
                }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < p.N * p.KK; idx += NTH) {
        const int oc = idx / p.KK, kk = idx - oc * p.KK;
        atomicAdd(&p.dw[((size_t)g * p.N + oc) * p.KK + kk], red[oc * RS + kk]);
    }
    if (p.db && wn == 0) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float s = sdy[mt][h];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const int oc = mt * 16 + gq + 8 * h;
                if (tq == 0 && oc < p.N) atomicAdd(&p.db[g * p.N + oc], s);
            }
    }
}

// row stride of the staged weights: (t * NS + g) mod 32 must be distinct for t < 4, g < 8  ->  NS = 8 or 24 (mod 32)
int pick_ns(int npad) {
    int ns = npad;
    while (ns % 32 != 8 && ns % 32 != 24) ns += 8;
    return ns;
}

// resident CTAs per SM targeted by the forward / data-gradient grids.  Whole-step A/B on one box (with the wgrad target
// below at 2): 4 -> 9.975 ms, 5 -> 9.99 ms, 6 (the old default, wgrad 3) -> 10.15 ms; 5 also keeps the standalone
// kernels at their occupancy limit.
int g_ctas_per_sm = 5;
int g_force_mtw = 2;

int grid_y(int G, int ntiles, int ctas_per_sm) {
    int per = 148 * ctas_per_sm / G;                 // rounded down: one CTA too many per group would be a second wave
    if (per < 1) per = 1;
    if (per > ntiles) per = ntiles;
    return per;
}

template <int MODE, int NT, int MTW>
int launch_mma(MmaParams& p, cudaStream_t st) {
    constexpr int TP = 32 * MTW * (kThreads / 32) / 2;
    const int rows_max = TP / p.P + 2;
    const int nr_max = (rows_max - 1) * p.Sg + p.Tspan;
    p.WS = (nr_max * p.P + 3 + 3) & ~3;      // + up to 3 floats of alignment shift
    p.TPe = MODE == MODE_FWD ? TP : (TP / p.P) * p.P;      // dgrad tiles hold whole rows of P (P <= 16 << TP)
    p.ES = MODE == MODE_FWD ? 0 : ((TP / p.P) * p.S * p.P) | 1;
    p.tiles_per_b = (int)ceil_div64((int64_t)p.Q * p.P, p.TPe);
    p.ntiles = p.B * p.tiles_per_b;
    p.fP = make_fdiv(p.P, (int64_t)p.Q * p.P + TP);
    p.fT = make_fdiv(p.tiles_per_b, p.ntiles);
    p.ot_alias = MODE != MODE_FWD && p.Cig * p.ES <= p.Cxg * p.WS;
    size_t smem = ((size_t)p.KKpad * p.NS + (size_t)4 * p.KKpad + (size_t)2 * p.Cxg * p.WS +
                   ((MODE == MODE_FWD || p.ot_alias) ? 0 : (size_t)p.Cig * p.ES)) * sizeof(float);
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv_mma_kernel<MODE, NT, MTW>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int G = p.Cx / p.Cxg;
    // persistent grid = what is really resident (registers / shared memory may allow fewer CTAs than the target:
    // a grid sized for more would run a second, mostly idle wave)
    static int regs = 0;                    // per instantiation
    const int occ = lct_resident_ctas(conv_mma_kernel<MODE, NT, MTW>, regs, kThreads, smem, 0);
    dim3 grid((unsigned)G, (unsigned)grid_y(G, p.ntiles, occ < g_ctas_per_sm ? occ : g_ctas_per_sm));
    conv_mma_kernel<MODE, NT, MTW><<<grid, kThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int MODE, int NT>
int launch_mma_mtw(MmaParams& p, cudaStream_t st) {
    // short layers: 128-position tiles keep more CTAs busy
    if (g_force_mtw == 2) return launch_mma<MODE, NT, 2>(p, st);
    if (g_force_mtw == 4) return launch_mma<MODE, NT, 4>(p, st);
    if ((int64_t)p.Q * p.P * p.B * (p.Cx / p.Cxg) < 148 * 256) return launch_mma<MODE, NT, 2>(p, st);
    return launch_mma<MODE, NT, 4>(p, st);
}

bool shape_ok(int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin,
              int64_t P) {
    return B > 0 && B < 65536 && Cin > 0 && Cout > 0 && G > 0 && G < 65536 && Cin % G == 0 && Cout % G == 0 &&
           K > 0 && K <= 64 && S >= 1 && S <= 4 && pad >= 0 && Lin > 0 && P > 0 && P <= 16 &&
           Lin + 2 * pad >= K && Lin * P < (1LL << 28) && B * Cin * Lin * P < (1LL << 31) &&
           B * Cout * ((Lin + 2 * pad - K) / S + 1) * P < (1LL << 31);
}

}  // namespace

// geometry of the staged weight images: mode 0 forward ([Cin/G*K -> pad 8] x NS), mode 1 data gradient
// ([Cout/G*ceil(K/S) -> pad 8] x NS); out[0] = KKpad, out[1] = NS  (HOST pointer)
LCT_API int lct_conv_mma_image_geometry(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int mode,
                                        int64_t* out) {
    if (!out || G <= 0 || Cin % G || Cout % G || S < 1) return LCT_EINVAL;
    const int64_t cig = Cin / G, cog = Cout / G, tmax = (K + S - 1) / S;
    const int64_t kk = mode ? cog * tmax : cig * K;
    const int64_t n = mode ? cig * S : cog;
    out[0] = (kk + 7) & ~(int64_t)7;
    out[1] = pick_ns((int)((n + 7) & ~(int64_t)7));
    return 0;
}

// 1 if the tensor-core kernels cover this layer (otherwise use lct_conv1d_*)
LCT_API int lct_conv_mma_supported(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P) {
    if (G <= 0 || Cin % G || Cout % G) return 0;
    const int64_t cig = Cin / G, cog = Cout / G;
    if (K > 64 || S < 1 || S > 4 || S == 2 || P < 1 || P > 16) return 0;
    if (cog > 32 || cig > 16) return 0;                 // dense layers have their own tcgen05 kernel
    if (cig * K > 168) return 0;                        // wgrad keeps <= 21 n-tiles of dW in registers
    return 1;
}

LCT_API int lct_conv_mma_fwd(const float* x, const float* w, const float* wimg, const float* bias, float* y, int64_t B,
                             int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin,
                             int64_t P, int act, float slope, cudaStream_t st) {
    if (!x || !w || !y || !shape_ok(B, Cin, Cout, G, K, S, pad, Lin, P) ||
        !lct_conv_mma_supported(Cin, Cout, G, K, S, P))
        return LCT_EINVAL;
    MmaParams p = {};
    p.x = x; p.w = w; p.wimg = wimg; p.out = y; p.bias = bias; p.act = act; p.slope = slope;
    p.B = (int)B; p.Cx = (int)Cin; p.Cxg = (int)(Cin / G); p.Lx = (int)Lin; p.P = (int)P;
    p.Co = (int)Cout; p.N = (int)(Cout / G); p.Lo = (int)((Lin + 2 * pad - K) / S + 1);
    p.K = (int)K; p.S = (int)S; p.pad = (int)pad; p.Tmax = (int)((K + S - 1) / S);
    p.Sg = (int)S; p.pad_eff = (int)pad; p.Tspan = (int)K; p.opad = 0; p.Q = p.Lo;
    p.KK = p.Cxg * p.K; p.KKpad = (p.KK + 7) & ~7;
    p.Npad = (p.N + 7) & ~7; p.NS = pick_ns(p.Npad);
    switch (p.Npad) {
        case 8: return launch_mma_mtw<MODE_FWD, 1>(p, st);
        case 16: return launch_mma_mtw<MODE_FWD, 2>(p, st);
        case 32: return launch_mma_mtw<MODE_FWD, 4>(p, st);
        default: return LCT_EUNSUPPORTED;
    }
}

LCT_API int lct_conv_mma_dgrad(const float* dy, const float* w, const float* wimg, float* dx, const float* gextra,
                               const float* xact, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S,
                               int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t st) {
    if (!dy || !w || !dx || !shape_ok(B, Cin, Cout, G, K, S, pad, Lin, P) ||
        !lct_conv_mma_supported(Cin, Cout, G, K, S, P))
        return LCT_EINVAL;
    MmaParams p = {};
    p.x = dy; p.w = w; p.wimg = wimg; p.out = dx; p.gextra = gextra; p.xact = xact; p.act = act; p.slope = slope;
    const int Lout = (int)((Lin + 2 * pad - K) / S + 1);
    p.B = (int)B; p.Cx = (int)Cout; p.Cxg = (int)(Cout / G); p.Lx = Lout; p.P = (int)P;
    p.Co = (int)Cin; p.Cig = (int)(Cin / G); p.N = p.Cig * (int)S; p.Lo = (int)Lin;
    p.K = (int)K; p.S = (int)S; p.pad = (int)pad; p.Tmax = (int)((K + S - 1) / S);
    p.Sg = 1; p.pad_eff = p.Tmax - 1; p.Tspan = p.Tmax; p.opad = (int)pad; p.Q = (int)((Lin + pad + S - 1) / S);
    p.KK = p.Cxg * p.Tmax; p.KKpad = (p.KK + 7) & ~7;
    p.Npad = (p.N + 7) & ~7; p.NS = pick_ns(p.Npad);
    switch (p.Npad) {
        case 8: return launch_mma_mtw<MODE_DGRAD, 1>(p, st);
        case 16: return launch_mma_mtw<MODE_DGRAD, 2>(p, st);
        case 24: return launch_mma_mtw<MODE_DGRAD, 3>(p, st);
        case 32: return launch_mma_mtw<MODE_DGRAD, 4>(p, st);
        case 48: return launch_mma<MODE_DGRAD, 6, 2>(p, st);
        case 64: return launch_mma<MODE_DGRAD, 8, 2>(p, st);
        default: return LCT_EUNSUPPORTED;
    }
}

// resident CTAs per SM targeted by the weight-gradient grid: 2 measured best in the whole step (3: +0.4 .. 1 %);
// lct_set_wgrad_ctas(1) while the G step's dead gradients run beside the generator's backward (lctgan/config.py)
int g_wgrad_ctas_per_sm = 2;

template <int MT, int NT, int NSPLIT, int TP>
int launch_wgrad_tp(MmaWgradParams& p, cudaStream_t st) {
    const int rows_max = TP / p.P + 2;
    const int nr_max = (rows_max - 1) * p.S + p.K;
    p.WS = (nr_max * p.P + 3 + 3) & ~3;      // + up to 3 floats of alignment shift
    p.DS = TP + 4;                           // 132 / 260: holds the shift, and DS mod 32 = 4 spreads the A rows over banks
    p.tiles_per_b = (int)ceil_div64((int64_t)p.Lout * p.P, TP);
    p.ntiles = p.B * p.tiles_per_b;
    p.fP = make_fdiv(p.P, (int64_t)p.Lout * p.P + TP);
    p.fT = make_fdiv(p.tiles_per_b, p.ntiles);
    size_t smem = ((size_t)2 * p.Cig * p.WS + (size_t)2 * 16 * MT * p.DS + 2 * p.Cig + 2 * 16 * MT) * sizeof(float);
    const size_t red = (size_t)16 * MT * NT * NSPLIT * 8 * sizeof(float);
    if (smem < red) smem = red;
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    auto kern = conv_mma_wgrad_kernel<MT, NT, NSPLIT, TP>;
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int G = p.Cin / p.Cig;
    // fewer, longer-lived CTAs than fwd/dgrad (every CTA ends with N x KK atomics), and never more than are resident
    static int regs = 0;                    // per instantiation
    const int occ = lct_resident_ctas(kern, regs, kThreads * NSPLIT, smem, 0);
    // the target counts 8-warp CTAs; the 4-warp instantiations (NSPLIT == 1) get twice as many for the same warps per SM
    // (ncu: 11 % of the warp slots occupied with 2 x 4 warps)
    const int want = g_wgrad_ctas_per_sm * (NSPLIT == 1 ? 2 : 1);
    const int per_sm = occ < want ? occ : want;
    int gy = 148 * per_sm / G;                // rounded down: no second wave
    if (gy > p.ntiles) gy = p.ntiles;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)G, (unsigned)gy);
    kern<<<grid, kThreads * NSPLIT, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

template <int MT, int NT, int NSPLIT>
int launch_wgrad(MmaWgradParams& p, cudaStream_t st) {
    // 256-position tiles where they were measured faster (see the kernel), 128 elsewhere
    constexpr bool kLongOk = (MT == 1 && NT == 11);
    if (kLongOk && (int64_t)p.Lout * p.P >= 384) return launch_wgrad_tp<MT, NT, NSPLIT, kLongOk ? 256 : 128>(p, st);
    return launch_wgrad_tp<MT, NT, NSPLIT, 128>(p, st);
}

// dw [Cout][Cin/G][K] and db [Cout] (optional) are accumulated (caller zeroes)
LCT_API int lct_conv_mma_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t Cin,
                               int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P,
                               cudaStream_t st) {
    if (!x || !dy || !dw || !shape_ok(B, Cin, Cout, G, K, S, pad, Lin, P) ||
        !lct_conv_mma_supported(Cin, Cout, G, K, S, P))
        return LCT_EINVAL;
    MmaWgradParams p = {};
    p.x = x; p.dy = dy; p.dw = dw; p.db = db;
    p.B = (int)B; p.Cin = (int)Cin; p.Cig = (int)(Cin / G); p.Lin = (int)Lin; p.P = (int)P;
    p.Cout = (int)Cout; p.N = (int)(Cout / G); p.Lout = (int)((Lin + 2 * pad - K) / S + 1);
    p.K = (int)K; p.S = (int)S; p.pad = (int)pad;
    p.KK = p.Cig * p.K;
    const int nt = (p.KK + 7) / 8;
    const int mt = (p.N + 15) / 16;
#define LCT_WG(MTV, NTV, NSV) return launch_wgrad<MTV, NTV, NSV>(p, st)
    if (mt == 1) {
        if (nt <= 1) LCT_WG(1, 1, 1);
        if (nt <= 2) LCT_WG(1, 2, 1);
        if (nt <= 5) LCT_WG(1, 5, 1);
        if (nt <= 10) LCT_WG(1, 5, 2);
        if (nt <= 21) LCT_WG(1, 11, 2);
    } else if (mt == 2) {
        if (nt <= 1) LCT_WG(2, 1, 1);
        if (nt <= 5) LCT_WG(2, 5, 1);
        if (nt <= 10) LCT_WG(2, 5, 2);
    }
#undef LCT_WG
    return LCT_EUNSUPPORTED;
}

// n >= 1: weight-gradient grids of lct_conv_mma_wgrad launched from now on target n resident CTAs per SM; 0: the default.
LCT_API int lct_set_wgrad_ctas(int n) {
    if (n < 0 || n > 8) return LCT_EINVAL;
    g_wgrad_ctas_per_sm = n > 0 ? n : 2;
    return 0;
}
