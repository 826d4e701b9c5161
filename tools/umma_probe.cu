// Bring-up probe: which no-swizzle shared-memory layouts / descriptor settings does tcgen05.mma kind::tf32 accept for
// K-major and MN-major operands?  One CTA, one MMA (M x N x 8), operands written to shared memory according to the
// hypothesis under test, D read back from TMEM and compared with the exact product.  Build (no GPU needed):
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O2 -I lct-gan_b200/csrc -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

struct Variant {
    int M, N;
    int a_mn, b_mn;          // operand majors (0 = K-major, 1 = MN-major)
    int a_lbo, a_sbo, b_lbo, b_sbo;   // descriptor byte offsets
    // how the host lays the operands out in shared memory (byte offset of element (row, k)):
    //   K-major : (row / 8) * sbo_l + (row % 8) * 16 + (k / 4) * lbo_l + (k % 4) * 4
    //   MN-major: (row / 4) * blk_l + (row % 4) * 4 + k * 16              (blk_l = stride between 4-row blocks)
    int a_l0, a_l1, b_l0, b_l1;       // K-major: (sbo_l, lbo_l); MN-major: (blk_l, unused)
};

__global__ void probe_kernel(const float* a_img, const float* b_img, int a_bytes, int b_bytes, Variant v, float* d_out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* A = sm;
    uint8_t* B = sm + ((a_bytes + 127) & ~127);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < a_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(A)[i] = a_img[i];
    for (int i = threadIdx.x; i < b_bytes / 4; i += blockDim.x) reinterpret_cast<float*>(B)[i] = b_img[i];
    if (threadIdx.x == 0) { tc::mbar_init(&mbar, 1); tc::mbar_fence_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc<32>(&tslot);
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tb = tslot;
    if (threadIdx.x == 0) {
        const uint64_t da = tc::smem_desc(tc::smem_u32(A), v.a_lbo, v.a_sbo);
        const uint64_t db = tc::smem_desc(tc::smem_u32(B), v.b_lbo, v.b_sbo);
        tc::umma_tf32(tb, da, db, tc::idesc_tf32(v.M, v.N, v.a_mn, v.b_mn), 0);
        tc::umma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::fence_after_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < v.N; c0 += 8) {
        uint32_t r[8];
        tc::tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        for (int n = 0; n < 8; ++n) d_out[(warp * 32 + lane) * 32 + c0 + n] = __uint_as_float(r[n]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc<32>(tb);
}

static int off_of(int mn, int l0, int l1, int row, int k) {
    if (!mn) return (row / 8) * l0 + (row % 8) * 16 + (k / 4) * l1 + (k % 4) * 4;
    return (row / 4) * l0 + (row % 4) * 4 + k * 16;
}

int main() {
    std::vector<Variant> vs;
    auto add = [&](int M, int N, int amn, int bmn, int albo, int asbo, int blbo, int bsbo, int al0, int al1, int bl0, int bl1) {
        vs.push_back(Variant{M, N, amn, bmn, albo, asbo, blbo, bsbo, al0, al1, bl0, bl1});
    };
    // K-major operands: 8-row groups 128 B apart, K chunks after all rows (dense): A chunk stride = M*16, B = N*16
    for (int M : {128, 64}) {
        const int N = 16;
        add(M, N, 0, 0, M * 16, 128, N * 16, 128, 128, M * 16, 128, N * 16);                       // both K-major (sanity)
        // MN-major operand: blocks of 4 rows x 8 k-rows (128 B) dense: block stride 128 B
        for (int swap = 0; swap < 2; ++swap) {
            const int lbo = swap ? 128 : 0, sbo = swap ? 0 : 128;                                   // block stride in SBO or in LBO
            add(M, N, 1, 0, lbo, sbo, N * 16, 128, 128, 0, 128, N * 16);                            // A MN-major, B K-major
            add(M, N, 0, 1, M * 16, 128, lbo, sbo, 128, M * 16, 128, 0);                            // A K-major, B MN-major
            add(M, N, 1, 1, lbo, sbo, lbo, sbo, 128, 0, 128, 0);                                    // both MN-major
        }
        add(M, N, 1, 1, 128, 128, 128, 128, 128, 0, 128, 0);                                        // both fields = block stride
        // Toeplitz-style MN-major A: blocks only 16 B apart (overlapping), data = ramp; B K-major
        add(M, N, 1, 0, 16, 16, N * 16, 128, 16, 0, 128, N * 16);
    }
    float *d_a, *d_b, *d_d;
    cudaMalloc(&d_a, 1 << 16); cudaMalloc(&d_b, 1 << 16); cudaMalloc(&d_d, 128 * 32 * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    // ---- address decoding: the MN-major operand is a ramp (value = float index in shared memory), the other operand a
    // K-major one-hot that selects k = column (resp. row) % 8: D shows WHICH shared-memory float the tensor core read for
    // every (row, k) of the MN-major operand, for distinct LBO / SBO values
    for (int which = 0; which < 2; ++which) {
        for (int M : {128, 64}) {
            const int N = 8, K = 8;
            Variant v{M, N, which == 0, which == 1, 0, 0, 0, 0, 0, 0, 0, 0};
            std::vector<float> a_img(1 << 13, 0.f), b_img(1 << 12, 0.f);
            if (which == 0) {   // decode A (MN-major): LBO = 1024 B, SBO = 128 B; B K-major one-hot
                v.a_lbo = 1024; v.a_sbo = 128; v.b_lbo = N * 16; v.b_sbo = 128;
                for (int j = 0; j < 2048; ++j) a_img[j] = (float)j;
                for (int n = 0; n < N; ++n) b_img[off_of(0, 128, N * 16, n, n % 8) / 4] = 1.f;
            } else {            // decode B (MN-major): LBO = 1024 B, SBO = 128 B; A K-major one-hot
                v.b_lbo = 1024; v.b_sbo = 128; v.a_lbo = M * 16; v.a_sbo = 128;
                for (int j = 0; j < 1024; ++j) b_img[j] = (float)j;
                for (int r = 0; r < M; ++r) a_img[off_of(0, 128, M * 16, r, r % 8) / 4] = 1.f;
            }
            cudaMemcpy(d_a, a_img.data(), 1 << 15, cudaMemcpyHostToDevice);
            cudaMemcpy(d_b, b_img.data(), 1 << 14, cudaMemcpyHostToDevice);
            cudaMemset(d_d, 0xff, 128 * 32 * 4);
            probe_kernel<<<1, 128, 60 * 1024>>>(d_a, d_b, 1 << 15, 1 << 14, v, d_d);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * 32);
            cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost);
            printf("decode %s MN-major (lbo 1024 = 256 floats, sbo 128 = 32 floats), M=%d, err=%d\n", which ? "B" : "A", M, (int)e);
            if (which == 0) {
                for (int r = 0; r < (M < 24 ? M : 24); ++r) {
                    const int lane = M == 128 ? r : (r % 16) + 32 * (r / 16);
                    printf("  A row %2d: float index read for k=0..7:", r);
                    for (int n = 0; n < 8; ++n) printf(" %5.0f", D[lane * 32 + n]);
                    printf("\n");
                }
            } else {
                for (int n = 0; n < 8; ++n) {
                    printf("  B row %2d: float index read for k=0..7:", n);
                    for (int r = 0; r < 8; ++r) printf(" %5.0f", D[r * 32 + n]);      // rows 0..7 select k = 0..7
                    printf("\n");
                }
            }
        }
    }
    int vi = 0;
    for (const Variant& v : vs) {
        const int K = 8;
        std::vector<float> Am(v.M * K), Bm(v.N * K);
        const bool toeplitz = v.a_mn && v.a_l0 == 16;
        std::vector<float> a_img(1 << 14, 0.f), b_img(1 << 14, 0.f);
        if (toeplitz) {
            // shared memory holds a ramp s[j] (j = float index); element (row, k) reads s[(row/4)*4 + row%4 + 4k] = s[row + 4 k]
            for (int j = 0; j < 4096; ++j) a_img[j] = (float)((j * 7) % 13 - 6);
            for (int r = 0; r < v.M; ++r) for (int k = 0; k < K; ++k) Am[r * K + k] = a_img[r + 4 * k];
        } else {
            for (int r = 0; r < v.M; ++r) for (int k = 0; k < K; ++k) {
                Am[r * K + k] = (float)(((r * 5 + k * 3) % 11) - 5);
                a_img[off_of(v.a_mn, v.a_l0, v.a_l1, r, k) / 4] = Am[r * K + k];
            }
        }
        for (int n = 0; n < v.N; ++n) for (int k = 0; k < K; ++k) {
            Bm[n * K + k] = (float)(((n * 3 + k * 7) % 9) - 4);
            b_img[off_of(v.b_mn, v.b_l0, v.b_l1, n, k) / 4] = Bm[n * K + k];
        }
        const int a_bytes = 1 << 15, b_bytes = 1 << 14;
        cudaMemcpy(d_a, a_img.data(), a_bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(d_b, b_img.data(), b_bytes, cudaMemcpyHostToDevice);
        cudaMemset(d_d, 0xff, 128 * 32 * 4);
        probe_kernel<<<1, 128, 60 * 1024>>>(d_a, d_b, a_bytes, b_bytes, v, d_d);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> D(128 * 32);
        cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, zero = 0, tot = 0;
        for (int r = 0; r < v.M; ++r) {
            const int lane = v.M == 128 ? r : (r % 16) + 32 * (r / 16);
            for (int n = 0; n < v.N; ++n) {
                float ref = 0.f;
                for (int k = 0; k < K; ++k) ref += Am[r * K + k] * Bm[n * K + k];
                const float got = D[lane * 32 + n];
                ++tot;
                if (got != ref) ++bad;
                if (got == 0.f && ref != 0.f) ++zero;
            }
        }
        printf("variant %2d: M=%3d N=%2d A %s (lbo %4d sbo %4d) B %s (lbo %4d sbo %4d)%s -> %s (%d / %d wrong, %d unexpectedly zero) err=%d\n",
               vi++, v.M, v.N, v.a_mn ? "MN" : "K ", v.a_lbo, v.a_sbo, v.b_mn ? "MN" : "K ", v.b_lbo, v.b_sbo,
               toeplitz ? " toeplitz" : "", bad ? "MISMATCH" : "exact", bad, tot, zero, (int)e);
        if (e != cudaSuccess) break;
    }
    return 0;
}
