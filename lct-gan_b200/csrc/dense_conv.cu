// The one true dense contraction of the training step on the 5th-generation tensor cores:
// MSD convs.5 = Conv1d(1024 -> 1024, k = 5, stride 1, pad 2)  (reference models/discriminators.py:166-196,
// layer 6 of ScaleDiscriminator; 10.5 GFLOP per call at B = 8, AI 359-719 flop/B: SURVEY.md section 8a D3).
//
// tcgen05 implicit GEMM, no im2col:
//   forward  Y[r, co]  = sum_{tap, ci} Xp[r + tap, ci] * Wt[tap][co][ci]       r = padded (batch, position) row
//   dgrad    dX[r, ci] = sum_{tap, co} dYp[r + tap, co] * Wd[tap][ci][co]      Wd[tap] = W[:, :, K-1-tap]^T
//   wgrad    dW[co, ci, tap] = sum_r dYq[co][r] * Xq[tap][ci][r]                positions are the contraction dim;
//            Xq[tap] is X shifted by `tap` positions (one staged copy per tap: a TMA box must start on a
//            16-byte boundary of the innermost dimension, so the shift cannot ride in the box coordinate)
// Operands are bf16 copies staged once per call (channels-last rows with K/2 zero rows between batches for
// fwd/dgrad, position-major rows for wgrad); accumulation is fp32 in TMEM.  Because a convolution tap is a
// pure row (or column) offset in these layouts, every A/B tile is a single TMA box load with the tap folded
// into the box coordinate; padding and ragged edges are the TMA's out-of-bounds zero fill.
//
// Kernel shape: one 128 x BN output tile per CTA (BN = 64 conv, 128 wgrad), 256 threads:
//   warp 0 lane 0 : TMA producer   (cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier expect_tx)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16; tcgen05.commit frees slots)
//   warp 2        : TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4..7    : epilogue       (tcgen05.ld 32x32b -> bias / activation / dgrad fusion -> coalesced fp32 stores)
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 "version 1"):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major, 1), [32,46) SBO >> 4 = 1024 B
//   (8 rows x 128 B per swizzle atom), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

enum { EPI_CONV = 0, EPI_WGRAD = 1 };

struct DenseParams {
    // k-block schedule: kb -> A box (a0 + (kb % kdiv) * BK, m0 + (kb / kdiv) * a_step1)
    //                        B box (b0 + (kb % kdiv) * BK, n0 + (kb / kdiv) * b_step1)
    int nkb, kdiv, a0, a_step1, b0, b_step1, b_tap_rows;
    // EPI_CONV: rows are padded (batch, position) pairs; out is fp32 [B, Cn, L]
    int R, Lp, L, Cn;
    const float* bias;      // forward
    const float* gextra;    // dgrad (optional)
    const float* xact;      // dgrad (optional)
    int act; float slope;
    float* out;
    // EPI_WGRAD: out is dW [M_total, N_total, K]; tap = blockIdx.z
    int Ntot, Ktaps;
};

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
dense_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp_u = tc::warp_uniform_id();
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tap = (EPI == EPI_WGRAD) ? blockIdx.z : 0;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_u == 0 && tc::elect_one()) {
        // ===== TMA producer =====
        for (int kb = 0; kb < p.nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], STAGE_BYTES);
            const int kin = kb % p.kdiv, kout = kb / p.kdiv;
            uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
            tma_load_2d(sa, &tmA, &full[s], p.a0 + kin * BK, m0 + kout * p.a_step1);
            tma_load_2d(sa + A_BYTES, &tmB, &full[s], p.b0 + kin * BK, n0 + kout * p.b_step1 + tap * p.b_tap_rows);
        }
    } else if (warp_u == 1 && tc::elect_one()) {
        // ===== MMA issuer (one elected lane of a warp-uniform branch: operands stay in uniform registers) =====
        constexpr uint32_t idesc = make_idesc(BM, BN);
        for (int kb = 0; kb < p.nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
            const uint32_t sb = sa + A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
                // advancing K inside the 128-byte swizzle row = +32 bytes on the start address
                umma(tmem_base, make_desc(sa + k * UMMA_K * 2), make_desc(sb + k * UMMA_K * 2), idesc,
                     (uint32_t)((kb | k) != 0));
            }
            umma_commit(&empty[s]);          // arrives once these MMAs have consumed the smem slot
        }
        umma_commit(tmem_full);              // accumulator complete
    } else if (warp_u >= 4) {
        // ===== epilogue: TMEM lane = tile row; warp (w % 4) owns lanes [32 (w%4), +32) =====
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (EPI == EPI_CONV) {
                const int b = row / p.Lp, l = row - b * p.Lp;
                if (row < p.R && l < p.L) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int ch = n0 + c0 + j;
                        const size_t idx = ((size_t)b * p.Cn + ch) * p.L + l;
                        float a = __uint_as_float(v[j]);
                        if (p.bias) a += __ldg(&p.bias[ch]);
                        if (p.gextra) a += p.gextra[idx];
                        if (p.xact) a *= act_grad_from_out(p.xact[idx], p.act, p.slope);
                        else a = apply_act(a, p.act, p.slope);
                        p.out[idx] = a;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c0 + j;
                    p.out[((size_t)row * p.Ntot + col) * p.Ktaps + tap] = __uint_as_float(v[j]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient, all K taps of a 128 x 64 (co x ci) block in ONE CTA.  The generic form above gives every tap its own
// CTA, which then owns every K-th float of dW [Co][Ci][K]: 16 384 scattered 4-byte stores per tile - 95 % of its
// 78 us (ncu: issue active 2.8 %, stalls barrier / lg_throttle; the 17-33 k-blocks of MMA take ~3 us).  Here the dY
// tile is loaded once per k-block and multiplied with the K shifted X tiles into K accumulators side by side in TMEM
// (K * 64 <= 512 columns); an epilogue thread then owns, for its co row, runs of 8 ci x K taps = 40 consecutive floats
// of dW and writes them as 16-byte stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

template <int KT, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
dense_wgrad_taps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int BN = 64;
    constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + KT * B_BYTES;
    constexpr int TMEM_COLS = 512;
    static_assert(KT * BN <= TMEM_COLS, "accumulators fit TMEM");
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp_u = tc::warp_uniform_id();
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_u == 0 && tc::elect_one()) {
        // ===== TMA producer: dY tile + the KT shifted X tiles of this k-block =====
        for (int kb = 0; kb < p.nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], STAGE_BYTES);
            uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
            tma_load_2d(sa, &tmA, &full[s], kb * BK, m0);
#pragma unroll
            for (int tap = 0; tap < KT; ++tap)
                tma_load_2d(sa + A_BYTES + tap * B_BYTES, &tmB, &full[s], kb * BK, n0 + tap * p.b_tap_rows);
        }
    } else if (warp_u == 1 && tc::elect_one()) {
        // ===== MMA issuer (one elected lane of a warp-uniform branch: operands stay in uniform registers) =====
        constexpr uint32_t idesc = make_idesc(BM, BN);
        for (int kb = 0; kb < p.nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE_BYTES);
#pragma unroll
            for (int tap = 0; tap < KT; ++tap) {
                const uint32_t sb = sa + A_BYTES + tap * B_BYTES;
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma(tmem_base + (uint32_t)(tap * BN), make_desc(sa + k * UMMA_K * 2), make_desc(sb + k * UMMA_K * 2),
                         idesc, (uint32_t)((kb | k) != 0));
            }
            umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
    } else if (warp_u >= 4) {
        // ===== epilogue: TMEM lane = co row; 8 ci x KT taps = 8 KT consecutive floats of dW per step =====
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        float* orow = p.out + ((size_t)row * p.Ntot + n0) * KT;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 8) {
            uint32_t v[KT][8];
#pragma unroll
            for (int tap = 0; tap < KT; ++tap)
                tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * BN + c0), v[tap]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float w[8 * KT];
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int tap = 0; tap < KT; ++tap) w[j * KT + tap] = __uint_as_float(v[tap][j]);
            float4* o4 = reinterpret_cast<float4*>(orow + (size_t)c0 * KT);      // 16-byte aligned: 8 KT floats per step
#pragma unroll
            for (int i = 0; i < 2 * KT; ++i) o4[i] = make_float4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// bf16 matrix [rows, cols] (cols contiguous, row pitch `pitch` elements) -> map with box {BK cols, box_rows}
int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return LCT_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch * 2) % 16) return LCT_EINVAL;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : LCT_EINVAL;
}


template <int BN, int STAGES, int EPI>
int launch_dense(const CUtensorMap& a, const CUtensorMap& b, DenseParams& p, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256;
    cudaError_t e = cudaFuncSetAttribute(dense_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    dense_kernel<BN, STAGES, EPI><<<grid, kThreads, smem, st>>>(a, b, p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// bf16 operand staging
// ---------------------------------------------------------------------------------------------
// x fp32 [B, C, L] -> out bf16 [B, L + 2*pad, C] with `pad` zero rows on both sides of every batch
__global__ void stage_nlc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int L, int pad) {
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const int l0 = blockIdx.x * 32 - pad;       // tile of output rows [l0, l0+32) in unpadded coordinates
    const int c0 = blockIdx.y * 32;
    const int Lp = L + 2 * pad;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int c = c0 + i, l = l0 + threadIdx.x;
        t[i][threadIdx.x] = (c < C && l >= 0 && l < L) ? x[((size_t)b * C + c) * L + l] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int l = l0 + i, c = c0 + threadIdx.x;
        int lp = l + pad;
        if (c < C && lp >= 0 && lp < Lp) out[((size_t)b * Lp + lp) * C + c] = __float2bfloat16(t[threadIdx.x][i]);
    }
}

// x fp32 [B, C, L] -> out bf16 [copies, C, pitch]: out[k][c][b * Lp + (shift - k) + l] = x[b, c, l], zeros elsewhere
// (copy k read at position r equals copy 0 read at r + k); rowsum (optional) [C] += sum_{b,l} x (bias gradient)
__global__ void stage_ncl_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int C, int L,
                                 int Lp, int shift, int pitch, float* __restrict__ rowsum) {
    __shared__ float red[32];
    const int c = blockIdx.x, k = blockIdx.y;
    const int sh = shift - k;
    __nv_bfloat16* o = out + ((size_t)k * C + c) * pitch;
    float s = 0.f;
    for (int i = threadIdx.x; i < pitch; i += blockDim.x) {
        int b = i / Lp, j = i - b * Lp, l = j - sh;
        float v = 0.f;
        if (b < B && l >= 0 && l < L) v = x[((size_t)b * C + c) * L + l];
        s += v;
        o[i] = __float2bfloat16(v);
    }
    if (rowsum && k == 0) {
        float tot = block_sum(s, red);
        if (threadIdx.x == 0) atomicAdd(&rowsum[c], tot);
    }
}

// w fp32 [Co, Ci, K] -> wt bf16 [K, Co, Ci] (forward) and wd bf16 [K, Ci, Co] with taps flipped (dgrad)
__global__ void stage_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt,
                                     __nv_bfloat16* __restrict__ wd, int Co, int Ci, int K) {
    __shared__ float t[32][33];
    const int k = blockIdx.z;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int co = co0 + i, ci = ci0 + threadIdx.x;
        float v = (co < Co && ci < Ci) ? w[((size_t)co * Ci + ci) * K + k] : 0.f;
        t[i][threadIdx.x] = v;
        if (wt && co < Co && ci < Ci) wt[((size_t)k * Co + co) * Ci + ci] = __float2bfloat16(v);
    }
    __syncthreads();
    if (wd) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            int ci = ci0 + i, co = co0 + threadIdx.x;
            if (co < Co && ci < Ci) wd[((size_t)(K - 1 - k) * Ci + ci) * Co + co] = __float2bfloat16(t[threadIdx.x][i]);
        }
    }
}

}  // namespace

LCT_API int lct_dense_supported(int64_t Cin, int64_t Cout, int64_t K) {
    return (Cin % 128 == 0 && Cout % 128 == 0 && K >= 1 && K <= 8 && (K & 1) && get_encode() != nullptr) ? 1 : 0;
}

// x fp32 [B,C,L] -> bf16 [B, L+2*pad, C]
LCT_API int lct_stage_nlc_bf16(const float* x, void* out, int64_t B, int64_t C, int64_t L, int64_t pad, cudaStream_t st) {
    if (!x || !out || B <= 0 || B >= 65536 || C <= 0 || L <= 0 || pad < 0) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(L + 2 * pad, 32), (unsigned)ceil_div64(C, 32), (unsigned)B);
    stage_nlc_kernel<<<grid, dim3(32, 8), 0, st>>>(x, (__nv_bfloat16*)out, (int)C, (int)L, (int)pad);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// x fp32 [B,C,L] -> bf16 [C, pitch] position-major (see stage_ncl_kernel); rowsum optional
LCT_API int lct_stage_ncl_bf16(const float* x, void* out, float* rowsum, int64_t B, int64_t C, int64_t L, int64_t Lp,
                               int64_t shift, int64_t pitch, int64_t copies, cudaStream_t st) {
    if (!x || !out || B <= 0 || C <= 0 || C >= (1LL << 31) || L <= 0 || Lp < L + shift || pitch < B * Lp || pitch % 8 ||
        copies < 1 || copies > 64)
        return LCT_EINVAL;
    stage_ncl_kernel<<<dim3((unsigned)C, (unsigned)copies), 256, 0, st>>>(x, (__nv_bfloat16*)out, (int)B, (int)C, (int)L, (int)Lp, (int)shift,
                                                  (int)pitch, rowsum);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// w fp32 [Co,Ci,K] -> wt bf16 [K,Co,Ci] and/or wd bf16 [K,Ci,Co] (taps flipped)
LCT_API int lct_stage_dense_weights(const float* w, void* wt, void* wd, int64_t Co, int64_t Ci, int64_t K,
                                    cudaStream_t st) {
    if (!w || (!wt && !wd) || Co <= 0 || Ci <= 0 || K <= 0 || K >= 65536) return LCT_EINVAL;
    dim3 grid((unsigned)ceil_div64(Ci, 32), (unsigned)ceil_div64(Co, 32), (unsigned)K);
    stage_weights_kernel<<<grid, dim3(32, 8), 0, st>>>(w, (__nv_bfloat16*)wt, (__nv_bfloat16*)wd, (int)Co, (int)Ci,
                                                       (int)K);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// out fp32 [B, Cn, L] = epilogue( sum_{tap, ca} a[b*Lp + l + tap, ca] * w[tap][cn][ca] ),  Lp = L + K - 1
//   forward: a = staged input, w = wt, bias/act set, gextra = xact = NULL
//   dgrad:   a = staged dY,    w = wd, bias NULL, out = (acc + gextra) * act'(xact)
LCT_API int lct_dense_conv(const void* a, const void* w, const float* bias, const float* gextra, const float* xact,
                           float* out, int64_t B, int64_t L, int64_t Ca, int64_t Cn, int64_t K, int act, float slope,
                           cudaStream_t st) {
    if (!a || !w || !out || B <= 0 || L <= 0 || !lct_dense_supported(Ca, Cn, K)) return LCT_EINVAL;
    const int64_t Lp = L + K - 1, R = B * Lp;
    CUtensorMap ma, mb;
    int rc = make_map(&ma, a, (uint64_t)R, (uint64_t)Ca, (uint64_t)Ca, BM);
    if (rc) return rc;
    // 128 x 128 tiles move 16 KB of operands per 128 x 64 x 64 MACs instead of 24 KB (the layer is bound by the
    // aggregate L2 -> SM throughput, not by the tensor pipe); 128 x 64 tiles when that would leave SMs without a tile
    const int64_t mt = ceil_div64(R, BM);
    const int bn = (mt * (Cn / 128) >= 100) ? 128 : 64;
    rc = make_map(&mb, w, (uint64_t)(K * Cn), (uint64_t)Ca, (uint64_t)Ca, (uint32_t)bn);
    if (rc) return rc;
    DenseParams p = {};
    p.kdiv = (int)(Ca / BK); p.nkb = (int)(K * p.kdiv);
    p.a0 = 0; p.a_step1 = 1; p.b0 = 0; p.b_step1 = (int)Cn;
    p.R = (int)R; p.Lp = (int)Lp; p.L = (int)L; p.Cn = (int)Cn;
    p.bias = bias; p.gextra = gextra; p.xact = xact; p.act = act; p.slope = slope; p.out = out;
    dim3 grid((unsigned)mt, (unsigned)(Cn / bn), 1);
    if (bn == 128) return launch_dense<128, 5, EPI_CONV>(ma, mb, p, grid, st);
    return launch_dense<64, 6, EPI_CONV>(ma, mb, p, grid, st);
}

// dw fp32 [Co, Ci, K] = sum_r dyq[co][r] * xq[tap][ci][r]   (dyq: 1 copy, shift 0; xq: K copies, shift K/2; both from
// lct_stage_ncl_bf16 with the same Lp = L + K - 1 and row pitch `pitch`)
LCT_API int lct_dense_wgrad(const void* dyq, const void* xq, float* dw, int64_t Co, int64_t Ci, int64_t K,
                            int64_t pitch, cudaStream_t st) {
    if (!dyq || !xq || !dw || pitch <= 0 || pitch % 8 || !lct_dense_supported(Ci, Co, K)) return LCT_EINVAL;
    CUtensorMap ma, mb;
    int rc = make_map(&ma, dyq, (uint64_t)Co, (uint64_t)pitch, (uint64_t)pitch, BM);
    if (rc) return rc;
    DenseParams p = {};
    p.nkb = (int)ceil_div64(pitch, BK); p.kdiv = p.nkb;
    p.a0 = 0; p.a_step1 = 0; p.b0 = 0; p.b_step1 = 0; p.b_tap_rows = (int)Ci;
    p.out = dw; p.Ntot = (int)Ci; p.Ktaps = (int)K;
    if (K == 5 && ((uintptr_t)dw & 15) == 0) {
        // all 5 taps per CTA (see dense_wgrad_taps_kernel)
        rc = make_map(&mb, xq, (uint64_t)(K * Ci), (uint64_t)pitch, (uint64_t)pitch, 64);
        if (rc) return rc;
        constexpr int KT = 5, ST = 3;
        constexpr size_t smem = (size_t)ST * (BM * BK * 2 + KT * 64 * BK * 2) + 1024 + 256;
        cudaError_t e = cudaFuncSetAttribute(dense_wgrad_taps_kernel<KT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid5((unsigned)(Co / BM), (unsigned)(Ci / 64));
        dense_wgrad_taps_kernel<KT, ST><<<grid5, kThreads, smem, st>>>(ma, mb, p);
        LCT_RETURN_IF_LAUNCH_FAILED();
        return 0;
    }
    rc = make_map(&mb, xq, (uint64_t)(K * Ci), (uint64_t)pitch, (uint64_t)pitch, 128);
    if (rc) return rc;
    dim3 grid((unsigned)(Co / BM), (unsigned)(Ci / 128), (unsigned)K);
    return launch_dense<128, 5, EPI_WGRAD>(ma, mb, p, grid, st);
}
