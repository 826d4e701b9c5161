// Grouped strided convolutions of the waveform discriminators on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM): forward and data gradient.
//
// Operator (reference models/discriminators.py:37-67, :93-98 PeriodDiscriminator; :166-196, :215-220
// ScaleDiscriminator): convolution along L of [B, C, L, P] (P = period, 1 for the scale discriminators) with
// groups, stride S and "same" padding K/2 - Cin/G = 1..16 input and Cout/G = 4..32 output channels per group.
// These layers are HBM bound (2-40 flop/B); round 1 ran them as mma.sync implicit GEMMs whose A fragments were
// gathered by the threads and measured them instruction-issue bound (22 warp instructions per m16n8k8 mma, 0.29
// of the HBM roofline).  tcgen05.mma reads BOTH operands from shared memory through descriptors, so here the
// threads only move data once (global -> registers -> shared memory, with the tf32 rounding on the way) and one
// thread issues ~20 MMAs per 128-position tile.
//
// The implicit GEMM without im2col.  A tile is 128 consecutive flat output positions m = l * P + p of one
// (batch, group).  The input window is stored in shared memory as float4 "slots" of 4 channels (a channel QUAD),
// one PLANE per (stride phase rho, quad):   plane[rho][q][j] = x[4q .. 4q+3][S * u + rho - pad][p],  j = u * P + p.
// For tap k = S * a + rho the 128 x 4 operand chunk "tap k, quad q" is then the 128 consecutive slots
// j = m + a * P of plane (rho, q): rows 16 bytes apart - exactly the K-major no-swizzle core-matrix layout of a
// UMMA shared-memory descriptor (8 rows x 16 bytes contiguous, SBO = 128).  So every (tap, quad) chunk is just a
// descriptor start address, a K = 8 MMA pairs two chunks through the descriptor's leading byte offset (next quad:
// LBO = plane pitch; or next tap of the same phase: LBO = 16 P), and overlapping windows cost nothing.
//   forward        D[m, co]       = sum_{k, ci} X[ci][S l + k - pad][p] W[co][ci][k]           M = positions, N = Cout/G
//   data gradient  D[(q,p),(ci,r)] = sum_{o, co} dY[co][q + o][p] W[co][ci][r + pad - S o]     N = (Cin/G) x S phases
//                  dX[ci][S q + r][p] = D;  one GEMM produces all S output phases, fused (+ FM gradient) x LeakyReLU'
//   weight gradient: stays on the TF32 mma.sync kernel of conv_mma.cu.  Positions are its contraction dimension, so the
//                  window operand would have to be MN-major (taps x channels contiguous per position); tools/umma_probe.cu
//                  shows that tcgen05.mma kind::tf32 returns exact zeros for ANY MN-major operand in the no-swizzle
//                  layout (K-major is exact for M = 64 and 128), and the K-major alternative needs four shifted copies
//                  of every phase plane and reads 3 x the forward's operand bytes per tile - slower than mma.sync.
// Weights arrive pre-arranged and tf32-rounded (round to nearest) as per-group images in exactly the shared-memory
// layout (lct_conv_tc_images, one launch per layer stack).  Precision contract: TF32 operands rounded to nearest,
// fp32 accumulation (tests/test_gpu_conv_mma.py: 2e-4 against fp64 on tf32-rounded operands).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kThreads = 128;     // 4 warps: warp w owns TMEM lanes [32 w, 32 w + 32)
constexpr int kTileM = 128;
constexpr int kMaxMma = 48;
enum { MODE_FWD = 0, MODE_DGRAD = 1 };

struct FastDiv { uint32_t mul, d; };
__host__ __device__ __forceinline__ int fdiv(int n, FastDiv f) {       // n >= 0
#ifdef __CUDA_ARCH__
    return f.d == 1 ? n : (int)__umulhi((uint32_t)n, f.mul);
#else
    return (int)((uint32_t)n / f.d);
#endif
}
FastDiv make_fdiv(int d) {          // exact for 0 <= n < 2^20 and d <= 64
    FastDiv f;
    f.d = (uint32_t)d;
    f.mul = d > 1 ? (uint32_t)((1ULL << 32) / (uint64_t)d) + 1u : 0u;
    return f;
}

// ---------------------------------------------------------------------------------------------------------
// Geometry of one layer in one mode; the MMA enumeration is a pure function of it (host and device agree).
// ---------------------------------------------------------------------------------------------------------
struct Geom {
    int mode;            // MODE_FWD / MODE_DGRAD
    int S, K, pad, P;
    int cig, cog;        // conv input / output channels per group
    int ca, nqa;         // channels of the A operand per group (fwd: cig, dgrad: cog) and its quads
    int ncol, Npad;      // real / padded (multiple of 16) GEMM columns (fwd: cog, dgrad: cig * S)
    int nphase;          // planes per quad (fwd: S stride phases, dgrad: 1)
    int o_min;           // dgrad: tap offsets o = o_min .. o_min + ntap[0] - 1 (l = q + o)
    int ntap[4];         // taps per phase
    int by_tap;          // 1: a K = 8 MMA pairs taps (t, t + 1) of one quad (nqa == 1); 0: quads (2 i, 2 i + 1) of one tap
    int tapmax;          // slots of look-ahead per plane, in taps (by_tap: rounded up to even)
    int nmma;
};

__host__ __device__ inline int phase_mmas(const Geom& g, int ph) {
    return g.by_tap ? (g.ntap[ph] + 1) / 2 : g.ntap[ph] * (g.nqa / 2);
}
// MMA j -> (phase, first tap, first quad); chunk 1 is (tap + 1, quad) if by_tap (valid iff tap + 1 < ntap) else (tap, quad + 1)
__host__ __device__ inline void decode_mma(const Geom& g, int j, int& ph, int& tap, int& quad) {
    ph = 0;
    while (ph < g.nphase - 1 && j >= phase_mmas(g, ph)) { j -= phase_mmas(g, ph); ++ph; }
    if (g.by_tap) { tap = 2 * j; quad = 0; }
    else { const int h = g.nqa / 2; tap = j / h; quad = 2 * (j - tap * h); }
}

bool make_geom(Geom& g, int mode, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t P) {
    if (G <= 0 || Cin % G || Cout % G || S < 1 || S > 4 || K < 1 || K > 64 || P < 1 || P > 16 || pad != K / 2) return false;
    g = Geom{};
    g.mode = mode; g.S = (int)S; g.K = (int)K; g.pad = (int)pad; g.P = (int)P;
    g.cig = (int)(Cin / G); g.cog = (int)(Cout / G);
    if (g.cig > 16 || g.cog > 32) return false;
    if (mode == MODE_FWD) {
        g.ca = g.cig; g.ncol = g.cog; g.nphase = g.S; g.o_min = 0;
        for (int r = 0; r < g.S; ++r) g.ntap[r] = r < g.K ? (g.K - 1 - r) / g.S + 1 : 0;
    } else {
        g.ca = g.cog; g.ncol = g.cig * g.S; g.nphase = 1;
        // k = r + pad - S o in [0, K) for some phase r in [0, S): o from ceil((pad - K + 1) / S) to floor((S - 1 + pad) / S)
        const int lo = g.pad - g.K + 1;
        g.o_min = lo >= 0 ? (lo + g.S - 1) / g.S : -((-lo) / g.S);
        const int o_max = (g.S - 1 + g.pad) / g.S;
        g.ntap[0] = o_max - g.o_min + 1;
    }
    g.nqa = (g.ca + 3) / 4;
    if (g.nqa != 1 && (g.nqa & 1)) return false;
    g.by_tap = g.nqa == 1;
    g.Npad = (g.ncol + 15) & ~15;
    if (g.Npad > 32) return false;
    g.tapmax = 0;
    g.nmma = 0;
    for (int ph = 0; ph < g.nphase; ++ph) {
        if (g.ntap[ph] > g.tapmax) g.tapmax = g.ntap[ph];
        g.nmma += phase_mmas(g, ph);
    }
    if (g.by_tap) g.tapmax = (g.tapmax + 1) & ~1;
    return g.nmma >= 1 && g.nmma <= kMaxMma;
}

// ---------------------------------------------------------------------------------------------------------
// Weight images: per group, per MMA j a [2 chunks][Npad rows][4] block (K-major, SBO = 128, LBO = 16 Npad bytes)
// ---------------------------------------------------------------------------------------------------------
struct ImgJob {
    Geom g;
    const float* w;      // [Cout][cig][K]
    float* img;          // [G][nmma][2][Npad][4]
    int G;
};
constexpr int kMaxImgJobs = 16;
struct ImgJobs { ImgJob job[kMaxImgJobs]; int n; };

__global__ void __launch_bounds__(256) tc_image_kernel(const ImgJobs J) {
    const ImgJob& jb = J.job[blockIdx.y];
    const Geom g = jb.g;
    const int per_group = g.nmma * 2 * g.Npad;             // float4 rows per group: [mma j][chunk c][row n]
    const int64_t total = (int64_t)jb.G * per_group;
    float4* img4 = reinterpret_cast<float4*>(jb.img);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(idx / per_group);
        int r = (int)(idx - (int64_t)grp * per_group);
        const int n = r % g.Npad; r /= g.Npad;
        const int c = r & 1;
        const int j = r >> 1;
        int ph, tap, quad;
        decode_mma(g, j, ph, tap, quad);
        bool valid = true;
        if (c) {
            if (g.by_tap) { ++tap; valid = tap < g.ntap[ph]; }
            else ++quad;
        }
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid && n < g.ncol) {
            if (g.mode == MODE_FWD) {
                const int k = g.S * tap + ph;
                const float* w = jb.w + ((int64_t)(grp * g.cog + n) * g.cig) * g.K + k;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * quad + e < g.ca) v[e] = w[(int64_t)(4 * quad + e) * g.K];
            } else {
                const int ci = n / g.S, rr = n - ci * g.S;
                const int k = rr + g.pad - g.S * (g.o_min + tap);
                if (k >= 0 && k < g.K) {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * quad + e < g.ca) v[e] = jb.w[((int64_t)(grp * g.cog + 4 * quad + e) * g.cig + ci) * g.K + k];
                }
            }
        }
        img4[idx] = make_float4(tc::tf32_rna(v[0]), tc::tf32_rna(v[1]), tc::tf32_rna(v[2]), tc::tf32_rna(v[3]));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Operand staging: global [B][C][Ls][P] fp32 -> registers -> tf32-rounded float4 slots in shared memory
// ---------------------------------------------------------------------------------------------------------
struct SrcMap {
    const float* src;
    int C, Ls, P;        // channels of the tensor, its length along L, period
    int Sg, ishift;      // source position of (row u, phase rho): i = Sg * u + rho + ishift
    int nch;             // real channels per group (the rest of the last quad is zero)
    int nslots;          // slots to fill per plane
    int PP;              // plane pitch in bytes
    FastDiv fP, fS;
};

template <int NQ, int NIT>
struct Staged { float v[NIT][NQ * 4]; };

__device__ __forceinline__ float ld_nc(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_global(float* p, float v) { asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// Per-thread staging plan.  Element e = tid + 128 it of a tile is (source row ir = e / P, p = e % P), i.e. stride phase
// rho = ir % Sg of plane row du = ir / Sg: none of this depends on the tile, only the tile's first slot p0 and its first
// source position do.  So the divisions happen once per CTA; per tile and element remain two range checks.
template <int NQ, int NIT>
struct StagePlan {
    int ir[NIT];        // source row of element it, relative to the tile's first source position
    int jrel[NIT];      // du * P + p: slot of the element relative to slot -p0
    int dst[NIT];       // byte offset of that slot in its plane: rho * NQ * PP + 16 jrel
    __device__ __forceinline__ void init(const SrcMap& s) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int e = (int)threadIdx.x + kThreads * it;
            const int r = fdiv(e, s.fP), pp = e - r * s.P;
            const int du = fdiv(r, s.fS), rho = r - du * s.Sg;
            ir[it] = r;
            jrel[it] = du * s.P + pp;
            dst[it] = rho * NQ * s.PP + 16 * jrel[it];
        }
    }
};

// issue the loads of one tile (coalesced: consecutive threads read consecutive floats of every channel row)
template <int NQ, int NIT>
__device__ __forceinline__ void stage_load(Staged<NQ, NIT>& R, const StagePlan<NQ, NIT>& pl, const SrcMap& s, int b,
                                           int cbase, int m0) {
    const int u0 = fdiv(m0, s.fP);
    const int ibase = s.Sg * u0 + s.ishift;
    const int64_t chs = (int64_t)s.Ls * s.P;
    const float* row0 = s.src + ((int64_t)b * s.C + cbase) * chs + (int64_t)ibase * s.P + threadIdx.x;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        const bool ok = (unsigned)(ibase + pl.ir[it]) < (unsigned)s.Ls;
#pragma unroll
        for (int c = 0; c < NQ * 4; ++c)
            R.v[it][c] = (ok && c < s.nch) ? ld_nc(row0 + c * chs + kThreads * it) : 0.f;
    }
}

// round to tf32 (one integer add: the tensor core ignores the 13 low mantissa bits) and store as float4 slots
template <int NQ, int NIT>
__device__ __forceinline__ void stage_store(const Staged<NQ, NIT>& R, const StagePlan<NQ, NIT>& pl, const SrcMap& s,
                                            uint8_t* base, int m0) {
    const int p0 = m0 - fdiv(m0, s.fP) * s.P;
    uint8_t* b0 = base - 16 * p0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
        if ((unsigned)(pl.jrel[it] - p0) < (unsigned)s.nslots) {
            uint8_t* dst = b0 + pl.dst[it];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                *reinterpret_cast<uint4*>(dst + q * s.PP) =
                    make_uint4(__float_as_uint(R.v[it][4 * q]) + 0x1000u, __float_as_uint(R.v[it][4 * q + 1]) + 0x1000u,
                               __float_as_uint(R.v[it][4 * q + 2]) + 0x1000u, __float_as_uint(R.v[it][4 * q + 3]) + 0x1000u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward / data gradient
// ---------------------------------------------------------------------------------------------------------
// data-gradient epilogue of 16 accumulator columns [C0, C0 + 16) of one tile row (q, p): column n = ci * S + r is
// dX[ci][S q + r][p]; fused (+ FM gradient) x LeakyReLU'(saved activation).  S and C0 are compile-time so that the
// column -> (ci, r) split costs nothing; VEC: S == 4, P == 1, 16-byte aligned rows -> one float4 per (row, channel).
// o / ge / xa point at element (ci = 0, position S q, p) of dX / FM gradient / saved activation of this row.
// The fused operands of the vector epilogue (S == 4, P == 1): FM gradient and saved activation of the <= 4 input channels
// a tile row touches, one float4 (the 4 output phases) each.  They are requested BEFORE the wait on the tile's MMAs, all
// at once, as 16-byte cp.async copies into the thread's own shared-memory slots (no registers held while they travel):
// issued inside the epilogue - one channel at a time, behind the previous channel's store - they cost a DRAM round trip
// per channel and tile with nothing else in flight.  Slot (array k, channel c, thread t) = fs[(k * 4 + c) * 128 + t].
constexpr int kFusedBytes = 2 * 4 * kThreads * 16;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void dgrad_fused_request(float4* fs, int nci, int64_t chs, const float* ge, const float* xa) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (ge && c < nci) cp_async16(fs + c * kThreads + threadIdx.x, ge + (int64_t)c * chs);
        if (xa && c < nci) cp_async16(fs + (4 + c) * kThreads + threadIdx.x, xa + (int64_t)c * chs);
    }
    cp_async_commit();
}

// data-gradient epilogue of 16 accumulator columns [C0, C0 + 16) of one tile row (q, p): column n = ci * S + r is
// dX[ci][S q + r][p]; fused (+ FM gradient) x LeakyReLU'(saved activation).  Vector form: S == 4, P == 1, 16-byte aligned
// rows -> one float4 per (row, channel); o points at element (ci = 0, position 4 q) of dX.
__device__ __forceinline__ void dgrad_store16_vec(const uint32_t (&v)[16], int ncol, int64_t chs, float* o, const float4* fs,
                                                  bool has_g, bool has_x, float neg) {
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
        if (4 * ci < ncol) {
            float4 a = make_float4(__uint_as_float(v[4 * ci]), __uint_as_float(v[4 * ci + 1]), __uint_as_float(v[4 * ci + 2]),
                                   __uint_as_float(v[4 * ci + 3]));
            if (has_g) {
                const float4 g4 = fs[ci * kThreads + threadIdx.x];
                a.x += g4.x; a.y += g4.y; a.z += g4.z; a.w += g4.w;
            }
            if (has_x) {
                const float4 x4 = fs[(4 + ci) * kThreads + threadIdx.x];
                a.x *= x4.x > 0.f ? 1.f : neg; a.y *= x4.y > 0.f ? 1.f : neg;
                a.z *= x4.z > 0.f ? 1.f : neg; a.w *= x4.w > 0.f ? 1.f : neg;
            }
            *reinterpret_cast<float4*>(o + (int64_t)ci * chs) = a;
        }
    }
}

// the same 16 columns written to the shared-memory staging tile instead: T[ci][(S (q - q0) + r) P + p] (t points at
// row element r = 0 of channel 0; TS = floats per channel) - the coalesced pass over the tile follows
template <int S, int C0>
__device__ __forceinline__ void dgrad_stage16(const uint32_t (&v)[16], int ncol, float* t, int TS, int P) {
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const int col = C0 + n;
        const int ci = col / S, r = col - ci * S;
        if (col < ncol) t[ci * TS + r * P] = __uint_as_float(v[n]);
    }
}

struct ConvParams {
    Geom g;
    SrcMap a;                // the gathered tensor (fwd: x, dgrad: dY)
    const float* wimg;       // [G][nmma][2][Npad][4]
    float* out;              // fwd: y [B][Cout][Lout][P]; dgrad: dx [B][Cin][Lin][P]
    const float* bias;       // fwd (optional)
    const float* gextra;     // dgrad (optional)
    const float* xact;       // dgrad (optional)
    int B, Cin, Cout, Lin, Lout;
    int Mtot;                // flat rows per batch (fwd: Lout * P; dgrad: ceil(Lin / S) * P)
    int mtile;               // flat rows per tile: 128 (fwd); dgrad: whole rows of P only, (128 / P) * P
    int TS;                  // dgrad: floats per channel of the staged output tile
    int tiles_per_b, ntiles;
    int a_bytes, b_bytes;    // shared-memory sizes of the A planes and of the weight image
    int fused_bytes;         // dgrad: vector epilogue: the two prefetched operands; staged epilogue: output tile + the two
    float neg;               // slope of the activation for negative inputs: LeakyReLU slope, ReLU 0, none 1
    int vec_ok;              // dgrad: dx / gextra / xact are 16-byte aligned
    FastDiv fT;
    uint16_t aoff[kMaxMma];  // MMA j: offset of its first A chunk from the start of the planes, in 16-byte units
};

template <int MODE, int NQ, int NIT>
__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* A = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* Bw = A + p.a_bytes;
    float* bias_s = reinterpret_cast<float*>(Bw + p.b_bytes);          // [32]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(bias_s + 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    // data gradient only (p.fused_bytes): vector epilogue [2][4][128] float4 of prefetched operands; staged epilogue
    // [3][cig][TS] floats = output tile, prefetched FM gradient, prefetched saved activation
    float4* fused = reinterpret_cast<float4*>(tmem_slot + 2);

    const Geom& g = p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp_u = tc::warp_uniform_id();
    const int grp = blockIdx.x;
    const int cbase_a = grp * g.ca;

    // ---- the first tile's global loads go out before anything else: they fly during the set-up below
    StagePlan<NQ, NIT> plan;
    plan.init(p.a);
    Staged<NQ, NIT> R0;
    const int gstep = (int)gridDim.y;
    int t0 = blockIdx.y;
    if (t0 < p.ntiles) {
        const int b = fdiv(t0, p.fT);
        stage_load<NQ, NIT>(R0, plan, p.a, b, cbase_a, (t0 - b * p.tiles_per_b) * p.mtile);
    }

    // ---- one-time setup: barrier, TMEM, weight image, descriptors, bias
    if (threadIdx.x == 0) {
        tc::mbar_init(mbar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<64>(tmem_slot);       // two accumulator buffers of <= 32 columns
    {
        const float4* src = reinterpret_cast<const float4*>(p.wimg + (size_t)grp * (p.b_bytes / 4));
        float4* dst = reinterpret_cast<float4*>(Bw);
        for (int i = threadIdx.x; i < p.b_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
    }
    if (threadIdx.x < 32) {
        const int n = threadIdx.x;
        bias_s[n] = (MODE == MODE_FWD && p.bias && n < g.cog) ? p.bias[grp * g.cog + n] : 0.f;
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = tc::idesc_tf32(kTileM, g.Npad, 0, 0);
    // MMA j reads A chunks at planes + 16 aoff[j] (second chunk LBO further: next quad's plane, or next tap = 16 P bytes)
    // and its [2][Npad][4] block of the weight image; only the start-address field of the descriptors changes
    const uint64_t da0 = tc::smem_desc(tc::smem_u32(A), g.by_tap ? (uint32_t)(g.P * 16) : (uint32_t)p.a.PP, 128);
    const uint64_t db0 = tc::smem_desc(tc::smem_u32(Bw), (uint32_t)(g.Npad * 16), 128);
    const uint32_t bstep = (uint32_t)(g.Npad * 2);                      // 32 Npad bytes per MMA, in 16-byte units

    const int64_t chs = (int64_t)p.Lin * g.P;
    const bool vec = MODE == MODE_DGRAD && g.S == 4 && g.P == 1 && g.ncol <= 16 && (p.Lin & 3) == 0 && p.vec_ok;
    // Two tiles deep: accumulators alternate between two TMEM buffers, and tile i + 1 is staged and its MMAs issued
    // BEFORE the epilogue of tile i - the tensor core works through the epilogue's loads and stores (nothing of tile i
    // lives in the operand planes after its MMAs: the staged data-gradient epilogue has its own output tile).
    const bool has_fused = MODE == MODE_DGRAD && (p.gextra != nullptr || p.xact != nullptr);
    float* Tst = reinterpret_cast<float*>(fused);                 // staged epilogue: [cig][TS] output tile,
    float* Gst = Tst + g.cig * p.TS;                              //   prefetched FM gradient,
    float* Xst = Gst + g.cig * p.TS;                              //   prefetched saved activation

    // registers -> shared memory, MMAs into TMEM buffer `buf`, then the next tile's loads into the same registers (they
    // fly during this tile's MMAs and epilogue).  (A second register set = loads two tiles ahead was measured SLOWER:
    // the extra 30-60 registers cost one or two of the 4-6 resident CTAs per SM.)
    auto stage = [&](Staged<NQ, NIT>& R, const int tile, const uint32_t buf) {
        const int b = fdiv(tile, p.fT);
        const int m0 = (tile - b * p.tiles_per_b) * p.mtile;
        stage_store<NQ, NIT>(R, plan, p.a, A, m0);
        tc::fence_proxy_async_smem();            // st.shared (generic proxy) -> tcgen05.mma operand reads (async proxy)
        tc::fence_before_sync();                 // (deep: this thread's tcgen05.ld of the buffer being re-used are done)
        __syncthreads();
        // one elected lane of warp 0 (a warp-uniform branch, so the descriptors stay in uniform registers and the MMAs
        // go out back to back: see tc::elect_one) issues the tile's MMAs; the tensor core works from here on
        if (warp_u == 0) {
            if (tc::elect_one()) {
                tc::fence_after_sync();
                const uint32_t td = tmem_base + buf * 32u;
#pragma unroll 4
                for (int j = 0; j < g.nmma; ++j)
                    tc::umma_tf32(td, da0 + (uint64_t)p.aoff[j], db0 + (uint64_t)((uint32_t)j * bstep), idesc,
                                  (uint32_t)(j != 0));
                tc::umma_commit(mbar);
            }
            __syncwarp();
        }
        const int next = tile + gstep;
        if (next < p.ntiles) {
            const int nb = fdiv(next, p.fT);
            stage_load<NQ, NIT>(R, plan, p.a, nb, cbase_a, (next - nb * p.tiles_per_b) * p.mtile);
        }
    };
    // data gradient, vector epilogue: tile row ml = q - q0 (P == 1); the fused operands of the row travel as cp.async
    // copies while the tile's MMAs run
    auto request_fused = [&](const int tile) {
        const int b = fdiv(tile, p.fT);
        const int m0 = (tile - b * p.tiles_per_b) * p.mtile;
        if (vec) {
            const int ml = warp * 32 + lane;
            if (m0 + ml < p.Mtot) {
                const int64_t vbase = ((int64_t)b * p.Cin + (int64_t)grp * g.cig) * chs + 4 * (int64_t)(m0 + ml);
                dgrad_fused_request(fused, g.cig, chs, p.gextra ? p.gextra + vbase : nullptr,
                                    p.xact ? p.xact + vbase : nullptr);
            }
            return;
        }
        // staged epilogue: the tile's contiguous output run of every channel, nval = S * rows * P elements starting at
        // any alignment.  The shared-memory copy is shifted by the run's phase modulo 16 bytes (element t lives at
        // [ci * TS + phase + t], TS % 4 == 0), so the body moves 16 bytes per cp.async and only head and tail go by 4
        const int j0 = g.S * fdiv(m0, p.a.fP) * g.P;
        int nval = g.S * p.mtile;
        if (nval > (int)chs - j0) nval = (int)chs - j0;
        const int64_t cb = ((int64_t)b * p.Cin + (int64_t)grp * g.cig) * chs + j0;
        for (int ci = 0; ci < g.cig; ++ci) {
            const int64_t e0 = cb + (int64_t)ci * chs;
            const int phase = p.vec_ok ? (int)(e0 & 3) : 0;
            const int h0 = p.vec_ok ? min((4 - phase) & 3, nval) : nval;         // 4-byte head (everything if unaligned)
            const int nvec = (nval - h0) >> 2;
            float* gd = Gst + ci * p.TS + phase;
            float* xd = Xst + ci * p.TS + phase;
            for (int t = (int)threadIdx.x; t < h0; t += kThreads) {
                if (p.gextra) cp_async4(gd + t, p.gextra + e0 + t);
                if (p.xact) cp_async4(xd + t, p.xact + e0 + t);
            }
            for (int v = (int)threadIdx.x; v < nvec; v += kThreads) {
                const int t = h0 + 4 * v;
                if (p.gextra) cp_async16(gd + t, p.gextra + e0 + t);
                if (p.xact) cp_async16(xd + t, p.xact + e0 + t);
            }
            for (int t = h0 + 4 * nvec + (int)threadIdx.x; t < nval; t += kThreads) {
                if (p.gextra) cp_async4(gd + t, p.gextra + e0 + t);
                if (p.xact) cp_async4(xd + t, p.xact + e0 + t);
            }
        }
        cp_async_commit();
    };
    auto epilogue = [&](const int tile, const uint32_t buf) {
        const int b = fdiv(tile, p.fT);
        const int m0 = (tile - b * p.tiles_per_b) * p.mtile;
        const uint32_t trow = tmem_base + buf * 32u + ((uint32_t)(warp * 32) << 16);
        const int64_t vbase = ((int64_t)b * p.Cin + (int64_t)grp * g.cig) * chs + 4 * (int64_t)(m0 + warp * 32 + lane);
        const bool vrow = m0 + warp * 32 + lane < p.Mtot;
        // ---- epilogue: TMEM lane = tile row
        const int m = m0 + warp * 32 + lane;
        if (MODE == MODE_FWD) {
            const int64_t LoP = (int64_t)p.Mtot;
            float* yp = p.out + ((int64_t)b * p.Cout + (int64_t)grp * g.cog) * LoP + m;
            for (int c0 = 0; c0 < g.Npad; c0 += 16) {
                uint32_t v[16];
                tc::tmem_ld16(trow + (uint32_t)c0, v);
                tc::tmem_ld_wait();
                if (m < p.Mtot) {
                    const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0);
                    const int nval = g.cog - c0;
#pragma unroll
                    for (int n4 = 0; n4 < 4; ++n4) {
                        const float4 bb = b4[n4];
                        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int n = 4 * n4 + k;
                            const float t = __uint_as_float(v[n]) + bv[k];
                            if (n < nval) st_global(yp, t > 0.f ? t : t * p.neg);
                            yp += LoP;
                        }
                    }
                }
            }
        } else {
            if (vec) {
                // one float4 (the 4 output phases) per row and channel: consecutive lanes, consecutive 16 bytes
                uint32_t v[16];
                tc::tmem_ld16(trow, v);
                tc::tmem_ld_wait();
                cp_async_wait_all();
                if (vrow) dgrad_store16_vec(v, g.ncol, chs, p.out + vbase, fused, p.gextra != nullptr, p.xact != nullptr, p.neg);
            } else {
                // tile row ml = (q - q0, p) (dgrad tiles hold whole rows of P): dX[ci][S q + r][p] for the S phases r.
                // Stage the tile in shared memory, then one coalesced pass per channel over the contiguous run of
                // S * rows * P outputs this tile owns; the fused operands of that run were prefetched by cp.async
                const int ml = warp * 32 + lane;
                const int qrel = fdiv(ml, p.a.fP), pp = ml - qrel * g.P;
                const int q0 = fdiv(m0, p.a.fP);
                const bool rowok = ml < p.mtile && m0 + ml < p.Mtot;
                const int64_t cb = ((int64_t)b * p.Cin + (int64_t)grp * g.cig) * chs;
                float* T = Tst;
                float* t = T + g.S * ml - (g.S - 1) * pp;
#define LCT_DS(SV, C0V) dgrad_stage16<SV, C0V>(v, g.ncol, t, p.TS, g.P)
                {
                    uint32_t v[16];
                    tc::tmem_ld16(trow, v);
                    tc::tmem_ld_wait();
                    if (rowok) {
                        if (g.S == 4) LCT_DS(4, 0);
                        else if (g.S == 3) LCT_DS(3, 0);
                        else if (g.S == 2) LCT_DS(2, 0);
                        else LCT_DS(1, 0);
                    }
                }
                if (g.Npad > 16) {
                    uint32_t v[16];
                    tc::tmem_ld16(trow + 16u, v);
                    tc::tmem_ld_wait();
                    if (rowok) {
                        if (g.S == 4) LCT_DS(4, 16);
                        else if (g.S == 3) LCT_DS(3, 16);
                        else if (g.S == 2) LCT_DS(2, 16);
                        else LCT_DS(1, 16);
                    }
                }
#undef LCT_DS
                cp_async_wait_all();                                            // (this thread's own requests)
                __syncthreads();
                const int j0 = g.S * q0 * g.P;                                  // first output (flat) of the tile
                int nval = g.S * p.mtile;                                       // = S * rows * P
                if (nval > (int)chs - j0) nval = (int)chs - j0;
                for (int ci = 0; ci < g.cig; ++ci) {                            // <= S <= 4 elements per thread and channel
                    const int64_t ib = cb + (int64_t)ci * chs + j0 + (int)threadIdx.x;
                    const int so = ci * p.TS + (int)threadIdx.x;
                    const int sp = so + (p.vec_ok ? (int)((cb + (int64_t)ci * chs + j0) & 3) : 0);   // (see request_fused)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (u * kThreads + (int)threadIdx.x < nval) {
                            float a = T[so + u * kThreads];
                            if (p.gextra) a += Gst[sp + u * kThreads];
                            if (p.xact) a *= Xst[sp + u * kThreads] > 0.f ? 1.f : p.neg;
                            st_global(p.out + ib + u * kThreads, a);
                        }
                    }
                }
            }
        }
    };

    uint32_t phase = 0, buf = 0;
    int tile = t0;
    if (tile < p.ntiles) {
        stage(R0, tile, 0);
        if (has_fused) request_fused(tile);
    }
    while (tile < p.ntiles) {
        const int next = tile + gstep;
        tc::mbar_wait(mbar, phase);
        phase ^= 1;
        tc::fence_after_sync();
        if (next < p.ntiles) stage(R0, next, buf ^ 1u);
        epilogue(tile, buf);
        if (has_fused && next < p.ntiles) request_fused(next);        // (after the epilogue has read this tile's slots)
        buf ^= 1u;
        tile = next;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(tmem_base);
}

int g_tc_ctas_per_sm = 6;

template <int MODE, int NQ, int NIT>
int launch_conv(ConvParams& p, int G, cudaStream_t st) {
    const bool vec = MODE == MODE_DGRAD && p.g.S == 4 && p.g.P == 1 && p.g.ncol <= 16 && (p.Lin & 3) == 0 && p.vec_ok;
    p.fused_bytes = MODE != MODE_DGRAD ? 0 : (vec ? kFusedBytes : 3 * p.g.cig * p.TS * 4);
    const size_t smem = 128 + (size_t)p.a_bytes + p.b_bytes + 32 * 4 + 32 + (size_t)p.fused_bytes;
    if (smem > 200 * 1024) return LCT_EUNSUPPORTED;
    auto kern = conv_tc_kernel<MODE, NQ, NIT>;
    static bool attr_set = false;           // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    static int regs = 0;
    int occ = lct_resident_ctas(kern, regs, kThreads, smem, 64);
    if (occ > g_tc_ctas_per_sm) occ = g_tc_ctas_per_sm;
    int gy = 148 * occ / G;                 // persistent: at most one wave of resident CTAs
    if (gy > p.ntiles) gy = p.ntiles;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int rounds = (p.ntiles + gy - 1) / gy;            // every CTA of a group gets the same number of tiles (+- 1):
    gy = (p.ntiles + rounds - 1) / rounds;                  // 504 tiles on 222 CTAs would be 3 rounds at 76 % fill
    kern<<<dim3((unsigned)G, (unsigned)gy), kThreads, smem, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// fill the staging map + shared-memory sizes; returns the NIT needed (0 if unsupported)
int setup_conv(ConvParams& p, const Geom& g, const float* src, int Csrc, int Ls) {
    p.g = g;
    SrcMap& a = p.a;
    a.src = src; a.C = Csrc; a.Ls = Ls; a.P = g.P;
    a.Sg = g.mode == MODE_FWD ? g.S : 1;
    a.ishift = g.mode == MODE_FWD ? -g.pad : g.o_min;
    a.nch = g.ca;
    a.nslots = kTileM + g.tapmax * g.P;
    const int padmod = a.Sg == 4 ? 32 : (a.Sg == 3 ? 48 : (a.Sg == 2 ? 64 : 0));   // spreads the planes over the banks
    a.PP = ((a.nslots * 16 + 127) & ~127) + padmod;
    a.fP = make_fdiv(g.P); a.fS = make_fdiv(a.Sg);
    p.a_bytes = (g.nphase * g.nqa * a.PP + 127) & ~127;
    p.b_bytes = g.nmma * g.Npad * 32;
    for (int j = 0; j < g.nmma; ++j) {
        int ph, tap, quad;
        decode_mma(g, j, ph, tap, quad);
        p.aoff[j] = (uint16_t)(((ph * g.nqa + quad) * a.PP + tap * g.P * 16) >> 4);
    }
    // elements per tile and channel: Sg * rows * P with rows <= ceil((P - 1 + nslots) / P)
    const int rows = (g.P - 1 + a.nslots + g.P - 1) / g.P;
    const int nE = a.Sg * rows * g.P;
    return (nE + kThreads - 1) / kThreads;
}

template <int MODE>
int dispatch_conv(ConvParams& p, int G, int nit, cudaStream_t st) {
    const int nq = p.g.nqa;
    if (nit <= 2) {
        if (nq == 1) return launch_conv<MODE, 1, 2>(p, G, st);
        if (nq == 2) return launch_conv<MODE, 2, 2>(p, G, st);
        if (nq == 4) return launch_conv<MODE, 4, 2>(p, G, st);
        if (nq == 8) return launch_conv<MODE, 8, 2>(p, G, st);
    } else if (nit <= 3) {
        if (nq == 1) return launch_conv<MODE, 1, 3>(p, G, st);
        if (nq == 2) return launch_conv<MODE, 2, 3>(p, G, st);
        if (nq == 4) return launch_conv<MODE, 4, 3>(p, G, st);
        if (nq == 8) return launch_conv<MODE, 8, 3>(p, G, st);
    } else if (nit <= 5) {
        if (nq == 1) return launch_conv<MODE, 1, 5>(p, G, st);
        if (nq == 2) return launch_conv<MODE, 2, 5>(p, G, st);
        if (nq == 4) return launch_conv<MODE, 4, 5>(p, G, st);
    }
    return LCT_EUNSUPPORTED;
}


float act_neg(int act, float slope) { return act == LCT_ACT_LRELU ? slope : (act == LCT_ACT_RELU ? 0.f : 1.f); }

bool dims_ok(int64_t B, int64_t Cin, int64_t Cout, int64_t Lin, int64_t Lout, int64_t P) {
    return B > 0 && B < 65536 && Lin > 0 && Lout > 0 && Lin * P < (1LL << 19) && B * Cin * Lin * P < (1LL << 31) &&
           B * Cout * Lout * P < (1LL << 31);
}

}  // namespace

// 1 if the tcgen05 kernels cover this layer (grouped / first layers of the MPD and MSD stacks)
LCT_API int lct_conv_tc_supported(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P) {
    Geom gf, gd;
    if (!make_geom(gf, MODE_FWD, Cin, Cout, G, K, S, K / 2, P) || !make_geom(gd, MODE_DGRAD, Cin, Cout, G, K, S, K / 2, P))
        return 0;
    ConvParams p = {};
    if (setup_conv(p, gf, nullptr, 0, 1) > 5 || setup_conv(p, gd, nullptr, 0, 1) > 5) return 0;
    if (gd.nqa == 8 && setup_conv(p, gd, nullptr, 0, 1) > 3) return 0;
    return 1;
}

// floats of the per-layer weight image: out[0] forward, out[1] data gradient (HOST pointer)
LCT_API int lct_conv_tc_image_len(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P, int64_t* out) {
    Geom gf, gd;
    if (!out || !make_geom(gf, MODE_FWD, Cin, Cout, G, K, S, K / 2, P) || !make_geom(gd, MODE_DGRAD, Cin, Cout, G, K, S, K / 2, P))
        return LCT_EINVAL;
    out[0] = G * (int64_t)gf.nmma * gf.Npad * 8;
    out[1] = G * (int64_t)gd.nmma * gd.Npad * 8;
    return 0;
}

// Build the forward (img_f[i]) and / or data-gradient (img_d[i]) images of n <= 8 layers in one launch.
// w[i]: fp32 [Cout][Cin/G][K]; geo: HOST array of 6 int64 per layer (Cin, Cout, G, K, S, P); NULL image = skip.
LCT_API int lct_conv_tc_images(const void* const* w, void* const* img_f, void* const* img_d, const int64_t* geo,
                               int64_t n, cudaStream_t st) {
    if (!w || !geo || n <= 0 || 2 * n > kMaxImgJobs) return LCT_EINVAL;
    ImgJobs J;
    J.n = 0;
    int64_t maxtot = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t* q = geo + 6 * i;
        for (int mode = 0; mode < 2; ++mode) {
            void* img = mode == MODE_FWD ? (img_f ? img_f[i] : nullptr) : (img_d ? img_d[i] : nullptr);
            if (!img) continue;
            if (!w[i]) return LCT_EINVAL;
            ImgJob& jb = J.job[J.n];
            if (!make_geom(jb.g, mode, q[0], q[1], q[2], q[3], q[4], q[3] / 2, q[5])) return LCT_EINVAL;
            jb.w = (const float*)w[i]; jb.img = (float*)img; jb.G = (int)q[2];
            const int64_t tot = q[2] * (int64_t)jb.g.nmma * jb.g.Npad * 8;
            if (tot > maxtot) maxtot = tot;
            ++J.n;
        }
    }
    if (J.n == 0) return 0;
    int gx = (int)ceil_div64(maxtot / 4, 256);
    if (gx > 592) gx = 592;
    if (gx < 1) gx = 1;
    tc_image_kernel<<<dim3((unsigned)gx, (unsigned)J.n), 256, 0, st>>>(J);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_conv_tc_fwd(const float* x, const float* wimg, const float* bias, float* y, int64_t B, int64_t Cin,
                            int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act,
                            float slope, cudaStream_t st) {
    Geom g;
    if (!x || !wimg || !y || !make_geom(g, MODE_FWD, Cin, Cout, G, K, S, pad, P)) return LCT_EINVAL;
    const int64_t Lout = (Lin + 2 * pad - K) / S + 1;
    if (Lin + 2 * pad < K || !dims_ok(B, Cin, Cout, Lin, Lout, P)) return LCT_EINVAL;
    ConvParams p = {};
    const int nit = setup_conv(p, g, x, (int)Cin, (int)Lin);
    p.wimg = wimg; p.out = y; p.bias = bias; p.neg = act_neg(act, slope);
    p.B = (int)B; p.Cin = (int)Cin; p.Cout = (int)Cout; p.Lin = (int)Lin; p.Lout = (int)Lout;
    p.Mtot = (int)(Lout * P);
    p.mtile = kTileM;
    p.tiles_per_b = (p.Mtot + kTileM - 1) / kTileM;
    p.ntiles = p.B * p.tiles_per_b;
    p.fT = make_fdiv(p.tiles_per_b);
    if (p.ntiles >= (1 << 20)) return LCT_EUNSUPPORTED;
    return dispatch_conv<MODE_FWD>(p, (int)G, nit, st);
}

LCT_API int lct_conv_tc_dgrad(const float* dy, const float* wimg, float* dx, const float* gextra, const float* xact,
                              int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad,
                              int64_t Lin, int64_t P, int act, float slope, cudaStream_t st) {
    Geom g;
    if (!dy || !wimg || !dx || !make_geom(g, MODE_DGRAD, Cin, Cout, G, K, S, pad, P)) return LCT_EINVAL;
    const int64_t Lout = (Lin + 2 * pad - K) / S + 1;
    if (Lin + 2 * pad < K || !dims_ok(B, Cin, Cout, Lin, Lout, P)) return LCT_EINVAL;
    ConvParams p = {};
    const int nit = setup_conv(p, g, dy, (int)Cout, (int)Lout);
    p.wimg = wimg; p.out = dx; p.gextra = gextra; p.xact = xact; p.neg = act_neg(act, slope);
    p.B = (int)B; p.Cin = (int)Cin; p.Cout = (int)Cout; p.Lin = (int)Lin; p.Lout = (int)Lout;
    p.Mtot = (int)(((Lin + S - 1) / S) * P);
    p.vec_ok = (((uintptr_t)dx | (uintptr_t)gextra | (uintptr_t)xact) & 15) == 0;
    p.mtile = (kTileM / (int)P) * (int)P;             // whole rows of P: the tile's outputs are one contiguous run
    p.TS = ((int)S * p.mtile + 4 + 3) & ~3;      // + 3 floats of alignment shift for the prefetched operands, multiple of 4
    p.tiles_per_b = (p.Mtot + p.mtile - 1) / p.mtile;
    p.ntiles = p.B * p.tiles_per_b;
    p.fT = make_fdiv(p.tiles_per_b);
    if (p.ntiles >= (1 << 20)) return LCT_EUNSUPPORTED;
    return dispatch_conv<MODE_DGRAD>(p, (int)G, nit, st);
}
