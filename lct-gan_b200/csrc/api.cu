// Library-level entry points of liblctgan_sm100.so (see include/lctgan.h).
#include "common.cuh"

std::atomic<int> g_lct_kernel_launches{0};

LCT_API int lct_version(void) { return 2; }
LCT_API int lct_kernel_launches(void) { return g_lct_kernel_launches.load(std::memory_order_relaxed); }
LCT_API int lct_reset_kernel_launches(void) {
    g_lct_kernel_launches.store(0, std::memory_order_relaxed);
    return 0;
}

namespace {
// 16-byte stores over the aligned middle, bytes at the two ends
__global__ void zero_fill_kernel(uint8_t* p, int64_t bytes) {
    const int64_t head = (16 - (int64_t)(reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    const int64_t h = head < bytes ? head : bytes;
    const int64_t n16 = (bytes - h) / 16;
    uint4* q = reinterpret_cast<uint4*>(p + h);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = t; i < n16; i += stride) q[i] = make_uint4(0u, 0u, 0u, 0u);
    if (t < h) p[t] = 0;
    const int64_t tail0 = h + n16 * 16;
    if (tail0 + t < bytes && t < 16) p[tail0 + t] = 0;
}
}  // namespace

// Zero `bytes` bytes at `p` on `st`.  Gradient accumulators of a whole layer stack / of the whole generator are carved
// from ONE buffer cleared by one such call.  A kernel, not cudaMemsetAsync: captured into the step's CUDA graph a memset
// node does not carry its stream's priority, and the one at the head of a sub-discriminator chain was measured waiting
// 570 us behind the other chains' kernels (CUPTI timeline, profiles/timeline_r2_*.txt).
LCT_API int lct_memset_zero(void* p, int64_t bytes, cudaStream_t st) {
    if (!p || bytes < 0) return LCT_EINVAL;
    if (bytes == 0) return 0;
    int64_t blocks = (bytes / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    zero_fill_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<uint8_t*>(p), bytes);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
