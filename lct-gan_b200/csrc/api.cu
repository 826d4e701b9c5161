// Library-level entry points of liblctgan_sm100.so (see include/lctgan.h).
#include "common.cuh"

std::atomic<int> g_lct_kernel_launches{0};

LCT_API int lct_version(void) { return 2; }
LCT_API int lct_kernel_launches(void) { return g_lct_kernel_launches.load(std::memory_order_relaxed); }
LCT_API int lct_reset_kernel_launches(void) {
    g_lct_kernel_launches.store(0, std::memory_order_relaxed);
    return 0;
}

// Zero `bytes` bytes at `p` on `st` (a memset node when captured into a CUDA graph: no kernel is launched).  Gradient
// accumulators of a whole layer stack / of the whole generator are carved from ONE buffer cleared by one such call.
LCT_API int lct_memset_zero(void* p, int64_t bytes, cudaStream_t st) {
    if (!p || bytes < 0) return LCT_EINVAL;
    if (bytes == 0) return 0;
    cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, st);
    return e == cudaSuccess ? 0 : (int)e;
}
