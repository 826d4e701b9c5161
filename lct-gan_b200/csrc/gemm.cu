// Skinny fp32 GEMMs of the generator bottleneck (nn.Linear, GRU input projections, attention
// in/out projections) and their gradients.
//
// Replaces (reference file:line): GRUblockf/GRUblockt linear algebra, models/generator.py:104,
// :133, :138, :219, :245, :248 (aten::linear / addmm inside nn.GRU, nn.MultiheadAttention,
// nn.Linear -> cuBLAS).  M = B*T'*F' = 34 056 rows at B=8, N and K are 16..192: AI < 50 flop/B,
// HBM/L2 bound, so a register-tiled SIMT kernel is the right tool (one pass over A, weights
// resident in L1/L2).  Three operand layouts cover forward, dgrad and wgrad; wgrad splits the
// long row dimension across CTAs and finishes with atomics.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);   // 256

struct GemmParams {
    const float* A; const float* B; float* C;
    const float* bias; const float* res; float* out2;
    int M, N, K;
    int lda, ldb, ldc, ldr, ldo;
    int ta, tb;          // ta: A(m,k) = A[k*lda+m];  tb: B(k,n) = B[k*ldb+n], else B[n*ldb+k]
    int act; float slope; float alpha;
    int accumulate;      // C += result (no atomics)
    int ksplit;          // >1: split K across blockIdx.z, atomicAdd epilogue (C pre-zeroed / accumulated)
    int nbatch, a_div, b_div;
    int64_t sA, sB, sC, sBias, sRes, sOut2;
};

__global__ void __launch_bounds__(kThreads) gemm_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int batch = z / p.ksplit, ks = z - batch * p.ksplit;
    const float* A = p.A + (int64_t)(batch / p.a_div) * p.sA;
    const float* B = p.B + (int64_t)(batch / p.b_div) * p.sB;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kchunk = ((p.K + p.ksplit - 1) / p.ksplit + BK - 1) / BK * BK;
    const int k_begin = ks * kchunk;
    const int k_end = min(p.K, k_begin + kchunk);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        // A tile: BM x BK
#pragma unroll
        for (int i = 0; i < (BM * BK) / kThreads; ++i) {
            int e = tid + i * kThreads;
            int m, k;
            if (p.ta) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < p.M && gk < k_end) v = p.ta ? A[(int64_t)gk * p.lda + gm] : A[(int64_t)gm * p.lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / kThreads; ++i) {
            int e = tid + i * kThreads;
            int n, k;
            if (p.tb) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < p.N && gk < k_end) v = p.tb ? B[(int64_t)gk * p.ldb + gn] : B[(int64_t)gn * p.ldb + gk];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * TN]);
            float a[TM] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    float* C = p.C + (int64_t)batch * p.sC;
    const float* bias = p.bias ? p.bias + (int64_t)batch * p.sBias : nullptr;
    const float* res = p.res ? p.res + (int64_t)batch * p.sRes : nullptr;
    float* out2 = p.out2 ? p.out2 + (int64_t)batch * p.sOut2 : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * TM + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * TN + j;
            if (gn >= p.N) continue;
            float v = acc[i][j] * p.alpha;
            float* c = C + (int64_t)gm * p.ldc + gn;
            if (p.ksplit > 1) {
                atomicAdd(c, v);
            } else {
                if (bias) v += bias[gn];
                v = apply_act(v, p.act, p.slope);
                if (p.accumulate) v += *c;
                *c = v;
                if (out2) out2[(int64_t)gm * p.ldo + gn] = v + (res ? res[(int64_t)gm * p.ldr + gn] : 0.f);
            }
        }
    }
}

// out[n] += sum_m X[m, n]: a 32-column x 8-row-group CTA walks 256 rows, reduces the row groups in shared memory
// and issues one atomic per column
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N,
                                                     int ld, int rows_per_cta) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.y * 32 + tx;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(M, r0 + rows_per_cta);
    float a = 0.f;
    if (n < N)
        for (int r = r0 + ty; r < r1; r += 8) a += X[(int64_t)r * ld + n];
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && n < N) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][tx];
        atomicAdd(&out[n], s);
    }
}


// ------------------------------------------------------------------------------------------------
// TF32 tensor-core version (mma.sync m16n8k8, fp32 accumulate) of the same GEMM: 128 x 64 x 16 CTA tile, four warps
// of 32 x 64, operands rounded to TF32 while they are staged into shared memory.  Used in tensor-core mode
// (lct_set_tensor_core_gemm); the SIMT kernel above stays for exact-fp32 mode.
// ------------------------------------------------------------------------------------------------
constexpr int MB_M = 128, MB_N = 64, MB_K = 16, MB_THREADS = 128;
constexpr int MB_AS = MB_K + 4;     // A tile row stride:  (g * 20 + t) mod 32 distinct for g < 8, t < 4
constexpr int MB_BS = MB_N + 8;     // B tile row stride:  (t * 72 + g) mod 32 distinct

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return u;
}

__global__ void __launch_bounds__(MB_THREADS, 4) gemm_mma_kernel(const GemmParams p) {
    __shared__ __align__(16) uint32_t As[MB_M * MB_AS];
    __shared__ __align__(16) uint32_t Bs[MB_K * MB_BS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int z = blockIdx.z;
    const int batch = z / p.ksplit, ks = z - batch * p.ksplit;
    const float* A = p.A + (int64_t)(batch / p.a_div) * p.sA;
    const float* B = p.B + (int64_t)(batch / p.b_div) * p.sB;
    const int m0 = blockIdx.x * MB_M, n0 = blockIdx.y * MB_N;
    const int kchunk = ((p.K + p.ksplit - 1) / p.ksplit + MB_K - 1) / MB_K * MB_K;
    const int k_begin = ks * kchunk;
    const int k_end = min(p.K, k_begin + kchunk);
    const int ntiles = min(MB_N / 8, (p.N - n0 + 7) / 8);      // n-tiles that hold real columns (CTA uniform)

    float acc[2][MB_N / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < MB_N / 8; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    for (int k0 = k_begin; k0 < k_end; k0 += MB_K) {
#pragma unroll
        for (int i = 0; i < (MB_M * MB_K) / MB_THREADS; ++i) {
            const int e = tid + i * MB_THREADS;
            int m, k;
            if (p.ta) { m = e % MB_M; k = e / MB_M; } else { k = e % MB_K; m = e / MB_K; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < p.M && gk < k_end) v = p.ta ? A[(int64_t)gk * p.lda + gm] : A[(int64_t)gm * p.lda + gk];
            As[m * MB_AS + k] = to_tf32(v);
        }
#pragma unroll
        for (int i = 0; i < (MB_N * MB_K) / MB_THREADS; ++i) {
            const int e = tid + i * MB_THREADS;
            int n, k;
            if (p.tb) { n = e % MB_N; k = e / MB_N; } else { k = e % MB_K; n = e / MB_K; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < p.N && gk < k_end) v = p.tb ? B[(int64_t)gk * p.ldb + gn] : B[(int64_t)gn * p.ldb + gk];
            Bs[k * MB_BS + n] = to_tf32(v);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < MB_K; kk += 8) {
            uint32_t a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t* ar = As + (warp * 32 + mt * 16 + gq) * MB_AS + kk + tq;
                a[mt][0] = ar[0];
                a[mt][1] = ar[8 * MB_AS];
                a[mt][2] = ar[4];
                a[mt][3] = ar[8 * MB_AS + 4];
            }
#pragma unroll
            for (int nt = 0; nt < MB_N / 8; ++nt) {
                if (nt < ntiles) {
                    const uint32_t b0 = Bs[(kk + tq) * MB_BS + nt * 8 + gq];
                    const uint32_t b1 = Bs[(kk + tq + 4) * MB_BS + nt * 8 + gq];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        asm volatile(
                            "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                            "{%0,%1,%2,%3};"
                            : "+f"(acc[mt][nt][0]), "+f"(acc[mt][nt][1]), "+f"(acc[mt][nt][2]), "+f"(acc[mt][nt][3])
                            : "r"(a[mt][0]), "r"(a[mt][1]), "r"(a[mt][2]), "r"(a[mt][3]), "r"(b0), "r"(b1));
                    }
                }
            }
        }
        __syncthreads();
    }

    float* C = p.C + (int64_t)batch * p.sC;
    const float* bias = p.bias ? p.bias + (int64_t)batch * p.sBias : nullptr;
    const float* res = p.res ? p.res + (int64_t)batch * p.sRes : nullptr;
    float* out2 = p.out2 ? p.out2 + (int64_t)batch * p.sOut2 : nullptr;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int gm = m0 + warp * 32 + mt * 16 + gq + 8 * h;
            if (gm >= p.M) continue;
#pragma unroll
            for (int nt = 0; nt < MB_N / 8; ++nt)
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    const int gn = n0 + nt * 8 + 2 * tq + c2;
                    if (gn >= p.N) continue;
                    float v = acc[mt][nt][2 * h + c2] * p.alpha;
                    float* c = C + (int64_t)gm * p.ldc + gn;
                    if (p.ksplit > 1) {
                        atomicAdd(c, v);
                    } else {
                        if (bias) v += bias[gn];
                        v = apply_act(v, p.act, p.slope);
                        if (p.accumulate) v += *c;
                        *c = v;
                        if (out2) out2[(int64_t)gm * p.ldo + gn] = v + (res ? res[(int64_t)gm * p.ldr + gn] : 0.f);
                    }
                }
        }
}

int g_gemm_tensor_cores = 0;   // measured slower than the SIMT kernel on the generator's shapes (tools/bench_gemm.py): opt-in
int g_rowgemm = 1;             // route eligible NT / NN GEMMs to the 3xTF32 row GEMM (rowgemm.cu)

}  // namespace

// rowgemm.cu: launches the 3xTF32 tensor-core row GEMM if the operands qualify (returns 1 and sets *rc), else 0
int lct_rowgemm_try(const float* A, const float* B, float* C, const float* bias, const float* res, float* out2, int64_t M,
                    int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int64_t ldo, int ta, int tb,
                    int act, float slope, float alpha, int accumulate, int64_t ksplit, int64_t nbatch, int64_t a_div,
                    int64_t b_div, int64_t sA, int64_t sB, int64_t sC, int64_t sBias, int64_t sRes, int64_t sOut2,
                    cudaStream_t st, int* rc);

// (other translation units: is the 3xTF32 tensor-core path selected?)
int lct_rowgemm_enabled() { return g_rowgemm && !g_gemm_tensor_cores; }

// 1 (default): NT / NN GEMMs run on the fp32-accurate 3xTF32 row GEMM; 0: fp32 SIMT kernel for everything
LCT_API int lct_set_rowgemm(int on) {
    g_rowgemm = on ? 1 : 0;
    return 0;
}

// 0 (default): fp32 SIMT GEMM; 1: TF32 mma.sync GEMM (experimental: slower on the generator's skinny shapes)
LCT_API int lct_set_tensor_core_gemm(int on) {
    g_gemm_tensor_cores = on ? 1 : 0;
    return 0;
}

// C[M,N] = act(alpha * op(A) op(B) + bias)  (+ C if accumulate);  out2 = C + res  (optional).
// Batched over blockIdx.z: A += (z / a_div) * sA, B += (z / b_div) * sB, C/bias/res/out2 += z * s*.
// ksplit > 1 splits K over CTAs and atomically adds into C (bias/act/res/out2 must then be unset).
LCT_API int lct_gemm(const float* A, const float* B, float* C, const float* bias, const float* res, float* out2,
                     int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr,
                     int64_t ldo, int ta, int tb, int act, float slope, float alpha, int accumulate, int64_t ksplit,
                     int64_t nbatch, int64_t a_div, int64_t b_div, int64_t sA, int64_t sB, int64_t sC, int64_t sBias,
                     int64_t sRes, int64_t sOut2, cudaStream_t st) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || nbatch <= 0 || ksplit <= 0 || a_div <= 0 || b_div <= 0)
        return LCT_EINVAL;
    if (ksplit > 1 && (bias || res || out2 || act != LCT_ACT_NONE || accumulate)) return LCT_EINVAL;
    if (nbatch * ksplit >= 65536) return LCT_EINVAL;
    if (g_rowgemm && !g_gemm_tensor_cores) {
        int rc = 0;
        if (lct_rowgemm_try(A, B, C, bias, res, out2, M, N, K, lda, ldb, ldc, ldr, ldo, ta, tb, act, slope, alpha, accumulate,
                            ksplit, nbatch, a_div, b_div, sA, sB, sC, sBias, sRes, sOut2, st, &rc))
            return rc;
    }
    GemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.res = res; p.out2 = out2;
    p.M = (int)M; p.N = (int)N; p.K = (int)K;
    p.lda = (int)lda; p.ldb = (int)ldb; p.ldc = (int)ldc; p.ldr = (int)ldr; p.ldo = (int)ldo;
    p.ta = ta; p.tb = tb; p.act = act; p.slope = slope; p.alpha = alpha; p.accumulate = accumulate;
    p.ksplit = (int)ksplit; p.nbatch = (int)nbatch; p.a_div = (int)a_div; p.b_div = (int)b_div;
    p.sA = sA; p.sB = sB; p.sC = sC; p.sBias = sBias; p.sRes = sRes; p.sOut2 = sOut2;
    if (g_gemm_tensor_cores) {
        dim3 grid((unsigned)ceil_div64(M, MB_M), (unsigned)ceil_div64(N, MB_N), (unsigned)(nbatch * ksplit));
        gemm_mma_kernel<<<grid, MB_THREADS, 0, st>>>(p);
        LCT_RETURN_IF_LAUNCH_FAILED();
        return 0;
    }
    dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)ceil_div64(N, BN), (unsigned)(nbatch * ksplit));
    gemm_kernel<<<grid, kThreads, 0, st>>>(p);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// out[N] += column sums of X[M,N] (row stride ld); `out` zeroed / accumulated by the caller
LCT_API int lct_colsum(const float* X, float* out, int64_t M, int64_t N, int64_t ld, cudaStream_t st) {
    if (!X || !out || M <= 0 || N <= 0) return LCT_EINVAL;
    const int rows = 256;
    dim3 grid((unsigned)ceil_div64(M, rows), (unsigned)ceil_div64(N, 32));
    colsum_kernel<<<grid, 256, 0, st>>>(X, out, (int)M, (int)N, (int)ld, rows);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
