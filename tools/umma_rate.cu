// Bring-up probe: what does ONE tcgen05.mma cost as a function of its shape, operand type and shared-memory layout?
// Every CTA (R per SM, 128 threads) lets one thread issue a chain of `iters` MMAs that read the same operand bytes
// (values are irrelevant: the buffers are zero) and reports clock64 cycles per MMA; the slowest CTA and the kernel's
// wall time (CUDA events) are printed.  This is the measurement behind DESIGN.md section 6's "operand fetch" bound of
// the grouped-convolution kernels (M = 128 positions, N = 16 / 32 columns, K = 8 tf32).  Build (no GPU needed):
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O2 -I lct-gan_b200/csrc -o tools/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

struct Cfg {
    int kind;      // 0 = tf32 (K = 8), 1 = bf16 (kind::f16, K = 16): 32 bytes of K per row either way
    int M, N;
    int layout;    // 0 = no swizzle, dense (SBO 128, LBO 16 * rows); 1 = no swizzle, overlapping chunks (LBO 16: the
                   // conv kernel's "next tap" pairing); 2 = 128-byte swizzle (rows 128 B apart, SBO 1024)
    int nacc;      // accumulators used round-robin (1 = every MMA depends on the previous one's D)
    int walk;      // 1: every MMA reads a different A start address (as the conv kernel's taps do); 0: the same
    int issue;     // 0: `if (threadIdx.x == 0)` + descriptors built in vector registers (the compiler wraps every MMA in an
                   //    ELECT / R2UR / BRA.U.ANY loop); 1: warp-uniform branch + elect.sync, descriptors from uniform values
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void umma_any(int kind, uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    if (kind == 0)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc)
                     : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
                     "l"(da), "l"(db), "r"(idesc)
                     : "memory");
}

__device__ __forceinline__ uint64_t desc_of(uint32_t addr, int layout, int rows) {
    if (layout == 2) {
        uint64_t d = (uint64_t)((addr & 0x3FFFF) >> 4);
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)(1024 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)2 << 61;
        return d;
    }
    return tc::smem_desc(addr, layout == 1 ? 16u : (uint32_t)(16 * rows), 128);
}

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, int iters, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm) + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 56 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.f;
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<512>(&tslot);
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tslot;
    const uint32_t a0 = tc::smem_u32(base), b0 = a0 + 24 * 1024;
    const uint32_t idesc = (1u << 4) | ((c.kind == 0 ? 2u : 1u) << 7) | ((c.kind == 0 ? 2u : 1u) << 10) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    if (c.issue == 0) {
        if (threadIdx.x == 0) {
            const uint64_t db = desc_of(b0, c.layout, c.N);
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t aoff = c.walk ? (uint32_t)((i & 7) * (c.layout == 2 ? 32 : 16)) : 0u;
                umma_any(c.kind, tb + (uint32_t)((i % c.nacc) * c.N), desc_of(a0 + aoff, c.layout, c.M), db, idesc);
            }
            tc::umma_commit(&mbar);
            tc::mbar_wait(&mbar, 0);
            cycles[blockIdx.x] = clock64() - t0;
        }
    } else {
        const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
        if (warp_u == 0) {
            if (elect_one()) {
                const uint64_t db = desc_of(b0, c.layout, c.N);
                const uint64_t da0 = desc_of(a0, c.layout, c.M);
                const uint32_t step = c.walk ? (c.layout == 2 ? 2u : 1u) : 0u;       // in 16-byte units
                const uint32_t dstep = c.nacc == 2 ? (uint32_t)c.N : 0u;
                const long long t0 = clock64();
#pragma unroll 4
                for (int i = 0; i < iters; ++i)
                    umma_any(c.kind, tb + (uint32_t)(i & 1) * dstep, da0 + (uint64_t)((uint32_t)(i & 7) * step), db, idesc);
                tc::umma_commit(&mbar);
                tc::mbar_wait(&mbar, 0);
                cycles[blockIdx.x] = clock64() - t0;
            }
            __syncwarp();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<512>(tb);
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long* d_cyc;
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
    const int iters = 4096;
    std::vector<Cfg> cfgs;
    for (int kind : {0, 1})
        for (int M : {128, 64})
            for (int N : {16, 32, 64, 128, 256})
                for (int layout : {0, 1, 2}) {
                    if (layout == 1 && N > 32) continue;
                    if (M == 64 && N != 16 && N != 256) continue;
                    cfgs.push_back(Cfg{kind, M, N, layout, 1, 1, 1});
                    if (N <= 32 && layout != 2) {
                        cfgs.push_back(Cfg{kind, M, N, layout, 2, 1, 1});
                        cfgs.push_back(Cfg{kind, M, N, layout, 1, 0, 1});
                        if (kind == 0 && M == 128) cfgs.push_back(Cfg{kind, M, N, layout, 1, 1, 0});
                    }
                }
    printf("one CTA per SM (the whole TMEM each), %d MMAs per CTA, clock64 cycles per MMA (slowest CTA) and kernel us\n", iters);
    printf("%-5s %4s %4s %-22s %4s %4s %5s %10s %10s %12s\n", "kind", "M", "N", "layout", "nacc", "walk", "issue", "cyc/mma", "us", "A+B B/clk");
    for (const Cfg& c : cfgs) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        rate_kernel<<<sms, 128, 58 * 1024>>>(c, 64, d_cyc);                 // warm-up
        cudaEventRecord(e0);
        rate_kernel<<<sms, 128, 58 * 1024>>>(c, iters, d_cyc);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> cyc(sms);
        cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        const double worst = (double)*std::max_element(cyc.begin(), cyc.end()) / iters;
        const double bytes = 32.0 * (c.M + c.N);
        printf("%-5s %4d %4d %-22s %4d %4d %5s %10.1f %10.1f %12.1f\n", c.kind ? "bf16" : "tf32", c.M, c.N,
               c.layout == 0 ? "noswz dense" : c.layout == 1 ? "noswz overlap(LBO16)" : "swizzle128", c.nacc, c.walk,
               c.issue ? "unif" : "vec", worst,
               ms * 1e3, bytes / worst);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    return 0;
}
