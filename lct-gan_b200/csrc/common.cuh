// Shared helpers for the lctgan sm_100a kernels.  Everything in csrc/ is compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
// and exported through the C ABI declared in include/lctgan.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LCT_API extern "C" __attribute__((visibility("default")))

// Error convention: 0 ok, <0 invalid argument, >0 cudaError_t.
#define LCT_EINVAL (-1)
#define LCT_EUNSUPPORTED (-2)

// every kernel launch in this library is followed by this macro; it also feeds the launch
// counter behind lct_kernel_launches() (bench.py reports it as gpu_launches).
// (launches come from the main thread and from autograd's backward threads: relaxed atomic)
#include <atomic>
extern std::atomic<int> g_lct_kernel_launches;
#define LCT_RETURN_IF_LAUNCH_FAILED()                                        \
    do {                                                                     \
        cudaError_t _e = cudaGetLastError();                                 \
        if (_e != cudaSuccess) return (int)_e;                               \
        g_lct_kernel_launches.fetch_add(1, std::memory_order_relaxed);       \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Resident CTAs per SM of a kernel, from its register count (queried once per instantiation: pass a static int as
// cache), its threads, its dynamic shared memory and its TMEM columns (0 = none).  Computed by hand for the persistent
// grids: cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 1 on the first call of an instantiation and while a
// stream was being captured (ncu, round 2: grids of 148 CTAs instead of 888).
template <typename Kern>
static inline int lct_resident_ctas(Kern kern, int& regs_cache, int threads, size_t smem, int tmem_cols) {
    if (regs_cache == 0) {
        cudaFuncAttributes fa;
        regs_cache = (cudaFuncGetAttributes(&fa, kern) == cudaSuccess && fa.numRegs > 0) ? fa.numRegs : 128;
        (void)cudaGetLastError();
    }
    const int regs = (regs_cache + 7) & ~7;                        // allocation granularity: 8 registers per thread
    const int warps = (threads + 31) / 32;
    int occ = 65536 / (regs * 32 * warps);
    const int by_smem = (int)((227 * 1024) / (smem + 1024));      // + 1 KB reserved per CTA
    if (by_smem < occ) occ = by_smem;
    if (64 / warps < occ) occ = 64 / warps;
    if (tmem_cols > 0 && 512 / tmem_cols < occ) occ = 512 / tmem_cols;
    if (occ > 32) occ = 32;
    return occ < 1 ? 1 : occ;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum for blockDim.x <= 1024; result valid in thread 0 (and broadcast to all).
__device__ __forceinline__ float block_sum(float v, float* red /* >= 32 floats */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
        r = warp_sum(r);
        if (lane == 0) red[0] = r;
    }
    __syncthreads();
    return red[0];
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// activation codes shared by host and device
enum { LCT_ACT_NONE = 0, LCT_ACT_LRELU = 1, LCT_ACT_RELU = 2 };

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == LCT_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == LCT_ACT_RELU) return v > 0.f ? v : 0.f;
    return v;
}
// derivative expressed through the *post*-activation value y (sign(y) == sign(pre) for both)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
    if (act == LCT_ACT_LRELU) return y > 0.f ? 1.f : slope;
    if (act == LCT_ACT_RELU) return y > 0.f ? 1.f : 0.f;
    return 1.f;
}
