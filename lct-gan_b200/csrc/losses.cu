// Multi-tensor loss reductions.
//
// Replaces (reference file:line):
//   discriminator_loss      losses.py:110-135   (8 + 8 logit tensors, ls | hinge)
//   generator_adv_loss      losses.py:138-151
//   feature_matching_loss   losses.py:154-173   (51 feature-map pairs, 121 M elements per side at B=8)
//   mask_mse_loss           losses.py:176-181
// One launch reduces up to kMaxSeg tensors (the reference issues one mse/l1 kernel chain per
// tensor).  HBM bound: one read per operand, vectorised float4, warp-shuffle + one atomic per CTA.
#include "common.cuh"

namespace {

constexpr int kMaxSeg = 64;
constexpr int kThreads = 256;
constexpr int kChunk = kThreads * 4 * 8;   // elements per CTA

enum { OP_SQ_CONST = 0, OP_SQ_DIFF = 1, OP_ABS_DIFF = 2, OP_RELU_AFFINE = 3, OP_SUM = 4 };

struct Segs {
    const float* a[kMaxSeg];
    const float* b[kMaxSeg];
    float* g[kMaxSeg];
    int64_t n[kMaxSeg];
    int chunk0[kMaxSeg + 1];
    float scale[kMaxSeg];
    int nseg;
    int op;
    float k0, k1;
};

__device__ __forceinline__ float op_val(int op, float a, float b, float k0, float k1) {
    switch (op) {
        case OP_SQ_CONST: { float d = a - k0; return d * d; }
        case OP_SQ_DIFF: { float d = a - b; return d * d; }
        case OP_ABS_DIFF: return fabsf(a - b);
        case OP_RELU_AFFINE: return fmaxf(k0 + k1 * a, 0.f);
        default: return a;
    }
}
__device__ __forceinline__ float op_grad(int op, float a, float b, float k0, float k1) {
    switch (op) {
        case OP_SQ_CONST: return 2.f * (a - k0);
        case OP_SQ_DIFF: return 2.f * (a - b);
        case OP_ABS_DIFF: { float d = a - b; return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
        case OP_RELU_AFFINE: return (k0 + k1 * a) > 0.f ? k1 : 0.f;
        default: return 1.f;
    }
}

__device__ __forceinline__ int find_seg(const Segs& S, int chunk) {
    int lo = 0, hi = S.nseg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (S.chunk0[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <bool GRAD>
__global__ void __launch_bounds__(kThreads) mt_kernel(const Segs S, float* __restrict__ out,
                                                      const float* __restrict__ upstream) {
    __shared__ float red[32];
    const int seg = find_seg(S, blockIdx.x);
    const int64_t start = (int64_t)(blockIdx.x - S.chunk0[seg]) * kChunk;
    const int64_t n = S.n[seg];
    const int64_t end = min(start + (int64_t)kChunk, n);
    const float* __restrict__ a = S.a[seg];
    const float* __restrict__ b = S.b[seg];
    float* __restrict__ g = S.g[seg];
    const bool has_b = (S.op == OP_SQ_DIFF || S.op == OP_ABS_DIFF);
    const float gs = GRAD ? S.scale[seg] * (upstream ? upstream[0] : 1.f) : 0.f;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) & 15) == 0) &&
                     (!has_b || (reinterpret_cast<uintptr_t>(b) & 15) == 0) &&
                     (!GRAD || (reinterpret_cast<uintptr_t>(g) & 15) == 0);
    float acc = 0.f;
    if (vec) {
        const int64_t end4 = start + ((end - start) & ~(int64_t)3);
        // 4 float4 per operand in flight per thread (8 independent 16-byte loads before the first use); streaming loads:
        // every element is read exactly once
        constexpr int U = 4;
        for (int64_t i0 = start + (int64_t)threadIdx.x * 4; i0 < end4; i0 += (int64_t)U * kThreads * 4) {
            float4 va[U], vb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * kThreads * 4;
                const bool ok = i < end4;
                va[u] = ok ? __ldcs(reinterpret_cast<const float4*>(a + i)) : make_float4(0, 0, 0, 0);
                vb[u] = (ok && has_b) ? __ldcs(reinterpret_cast<const float4*>(b + i)) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = i0 + (int64_t)u * kThreads * 4;
                if (i >= end4) break;
                if (GRAD) {
                    float4 r;
                    r.x = gs * op_grad(S.op, va[u].x, vb[u].x, S.k0, S.k1);
                    r.y = gs * op_grad(S.op, va[u].y, vb[u].y, S.k0, S.k1);
                    r.z = gs * op_grad(S.op, va[u].z, vb[u].z, S.k0, S.k1);
                    r.w = gs * op_grad(S.op, va[u].w, vb[u].w, S.k0, S.k1);
                    *reinterpret_cast<float4*>(g + i) = r;
                } else {
                    acc += op_val(S.op, va[u].x, vb[u].x, S.k0, S.k1) + op_val(S.op, va[u].y, vb[u].y, S.k0, S.k1) +
                           op_val(S.op, va[u].z, vb[u].z, S.k0, S.k1) + op_val(S.op, va[u].w, vb[u].w, S.k0, S.k1);
                }
            }
        }
        for (int64_t i = end4 + threadIdx.x; i < end; i += kThreads) {
            float va = a[i], vb = has_b ? b[i] : 0.f;
            if (GRAD) g[i] = gs * op_grad(S.op, va, vb, S.k0, S.k1);
            else acc += op_val(S.op, va, vb, S.k0, S.k1);
        }
    } else {
        for (int64_t i = start + threadIdx.x; i < end; i += kThreads) {
            float va = a[i], vb = has_b ? b[i] : 0.f;
            if (GRAD) g[i] = gs * op_grad(S.op, va, vb, S.k0, S.k1);
            else acc += op_val(S.op, va, vb, S.k0, S.k1);
        }
    }
    if (!GRAD) {
        float s = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(out, s * S.scale[seg]);
    }
}

__global__ void __launch_bounds__(kThreads) mt_copy_kernel(const Segs S) {
    const int seg = find_seg(S, blockIdx.x);
    const int64_t start = (int64_t)(blockIdx.x - S.chunk0[seg]) * kChunk;
    const int64_t end = min(start + (int64_t)kChunk, S.n[seg]);
    const float* __restrict__ a = S.a[seg];
    float* __restrict__ g = S.g[seg];
    for (int64_t i = start + threadIdx.x; i < end; i += kThreads) g[i] = a[i];
}

int build(Segs& S, const void* const* a, const void* const* b, void* const* g, const int64_t* n,
          const float* scale, int64_t nseg, int op, float k0, float k1, bool grad) {
    if (!a || !n || !scale || nseg <= 0 || nseg > kMaxSeg || op < 0 || op > OP_SUM) return LCT_EINVAL;
    const bool has_b = (op == OP_SQ_DIFF || op == OP_ABS_DIFF);
    if (has_b && !b) return LCT_EINVAL;
    if (grad && !g) return LCT_EINVAL;
    int chunks = 0;
    for (int i = 0; i < nseg; ++i) {
        if (!a[i] || n[i] <= 0 || (has_b && !b[i]) || (grad && !g[i])) return LCT_EINVAL;
        S.a[i] = (const float*)a[i];
        S.b[i] = has_b ? (const float*)b[i] : nullptr;
        S.g[i] = grad ? (float*)g[i] : nullptr;
        S.n[i] = n[i];
        S.scale[i] = scale[i];
        S.chunk0[i] = chunks;
        chunks += (int)ceil_div64(n[i], kChunk);
    }
    S.chunk0[nseg] = chunks;
    S.nseg = (int)nseg;
    S.op = op;
    S.k0 = k0;
    S.k1 = k1;
    return chunks;
}

// ---- fused multi-tensor AdamW (torch.optim.AdamW semantics, train.py:601-610) ----
constexpr int kMaxAdam = 48;
struct AdamSegs {
    float* p[kMaxAdam];
    const float* g[kMaxAdam];
    float* m[kMaxAdam];
    float* v[kMaxAdam];
    int64_t n[kMaxAdam];
    int chunk0[kMaxAdam + 1];
    int nseg;
};

__global__ void __launch_bounds__(kThreads) mt_adamw_kernel(const AdamSegs S, const float* __restrict__ step_ptr,
                                                            float lr, float b1, float b2, float eps, float wd,
                                                            float gmul) {
    int lo = 0, hi = S.nseg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (S.chunk0[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const int seg = lo;
    const int64_t start = (int64_t)(blockIdx.x - S.chunk0[seg]) * kChunk;
    const int64_t end = min(start + (int64_t)kChunk, S.n[seg]);
    // bias corrections 1 - beta^t = -expm1(t log beta) in fp32 (relative error ~2e-7 against torch.optim.AdamW's double
    // arithmetic on the same float betas), by ONE thread per CTA.  Double-precision pow() here was ~10 us of dependent
    // latency per launch on this GPU's fp64 pipe - in every thread at first - and the optimiser's 7 launches per step all
    // sit on the step's critical path.
    __shared__ float corr[2];
    if (threadIdx.x == 0) {
        const float t = step_ptr[0];                            // already incremented for this step
        corr[0] = -expm1f(t * logf(b1));
        corr[1] = sqrtf(-expm1f(t * logf(b2)));
    }
    __syncthreads();
    const float bc1 = corr[0], bc2_sqrt = corr[1];
    const float step_size = lr / bc1;
    const float decay = 1.f - lr * wd;
    float* __restrict__ pp = S.p[seg];
    const float* __restrict__ gg = S.g[seg];
    float* __restrict__ mm = S.m[seg];
    float* __restrict__ vv = S.v[seg];
    auto upd = [&](float& p, float gr, float& m, float& v) {
        const float g = gr * gmul;
        p *= decay;
        m = m + (g - m) * (1.f - b1);
        v = b2 * v + (1.f - b2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        p -= step_size * (m / denom);
    };
    const bool vec = (((uintptr_t)pp | (uintptr_t)gg | (uintptr_t)mm | (uintptr_t)vv) & 15) == 0;   // (kChunk % 4 == 0)
    int64_t i = start + 4 * (int64_t)threadIdx.x;
    if (vec) {
        for (; i + 3 < end; i += 4 * kThreads) {
            float4 p4 = *reinterpret_cast<const float4*>(pp + i);
            const float4 g4 = *reinterpret_cast<const float4*>(gg + i);
            float4 m4 = *reinterpret_cast<const float4*>(mm + i);
            float4 v4 = *reinterpret_cast<const float4*>(vv + i);
            upd(p4.x, g4.x, m4.x, v4.x); upd(p4.y, g4.y, m4.y, v4.y);
            upd(p4.z, g4.z, m4.z, v4.z); upd(p4.w, g4.w, m4.w, v4.w);
            *reinterpret_cast<float4*>(pp + i) = p4;
            *reinterpret_cast<float4*>(mm + i) = m4;
            *reinterpret_cast<float4*>(vv + i) = v4;
        }
    }
    // scalar form: unaligned tensors, and the < 4 elements at the end of a tensor (at most one thread gets there)
    for (; i < end; i += vec ? end : 4 * kThreads) {
        for (int64_t e = i; e < min(i + 4, end); ++e) {
            float p = pp[e], m = mm[e], v = vv[e];
            upd(p, gg[e], m, v);
            pp[e] = p; mm[e] = m; vv[e] = v;
        }
    }
}

__global__ void add_scalar_kernel(float* x, float v) { x[0] += v; }

// torch.nn.utils.clip_grad_norm_ (train.py:246-248) over up to kMaxSeg gradient tensors, in place:
//   total = pre * sqrt(sumsq),  g <- g * pre * min(1, max_norm / (total + 1e-6))
// `sumsq` = sum of squares of all gradients (lct_mt_reduce, op 0, k0 = 0); `pre` folds the 1 / world_size of a
// data-parallel SUM all-reduce into the same pass.  norm_out (optional) receives `total`.
__global__ void __launch_bounds__(kThreads) mt_clip_kernel(const Segs S, const float* __restrict__ sumsq,
                                                           float max_norm, float pre, float* __restrict__ norm_out) {
    const int seg = find_seg(S, blockIdx.x);
    const int64_t start = (int64_t)(blockIdx.x - S.chunk0[seg]) * kChunk;
    const int64_t end = min(start + (int64_t)kChunk, S.n[seg]);
    const float total = sqrtf(sumsq[0]) * pre;
    const float coef = fminf(1.f, max_norm / (total + 1e-6f)) * pre;
    float* __restrict__ g = S.g[seg];
    for (int64_t i = start + threadIdx.x; i < end; i += kThreads) g[i] *= coef;
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = total;
}

}  // namespace

// x[0] += v  (device-side step counter of the fused optimiser, CUDA-graph friendly)
LCT_API int lct_add_scalar(float* x, float v, cudaStream_t st) {
    if (!x) return LCT_EINVAL;
    add_scalar_kernel<<<1, 1, 0, st>>>(x, v);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// One AdamW update (decoupled weight decay, bias correction, torch.optim.AdamW defaults' formulas) of up to 48 tensors
// per launch.  p/g/m/v: HOST arrays of device pointers, n: HOST array of element counts; step: device float holding
// the (already incremented) step number; grad_scale multiplies every gradient as it is read (1 / world size when the
// gradients hold the SUM of a data-parallel all-reduce; 1 otherwise).
LCT_API int lct_mt_adamw(void* const* p, const void* const* g, void* const* m, void* const* v, const int64_t* n,
                         int64_t nseg, const float* step, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, cudaStream_t st) {
    if (!p || !g || !m || !v || !n || !step || nseg <= 0 || nseg > kMaxAdam) return LCT_EINVAL;
    AdamSegs S;
    int chunks = 0;
    for (int i = 0; i < nseg; ++i) {
        if (!p[i] || !g[i] || !m[i] || !v[i] || n[i] <= 0) return LCT_EINVAL;
        S.p[i] = (float*)p[i]; S.g[i] = (const float*)g[i]; S.m[i] = (float*)m[i]; S.v[i] = (float*)v[i];
        S.n[i] = n[i];
        S.chunk0[i] = chunks;
        chunks += (int)ceil_div64(n[i], kChunk);
    }
    S.chunk0[nseg] = chunks;
    S.nseg = (int)nseg;
    mt_adamw_kernel<<<chunks, kThreads, 0, st>>>(S, step, lr, beta1, beta2, eps, weight_decay, grad_scale);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_mt_adamw_max_segments(void) { return kMaxAdam; }

// out[0] += sum_i scale[i] * sum_j op(a_i[j], b_i[j]);  `out` must be zeroed by the caller.
//   op 0: (a-k0)^2   1: (a-b)^2   2: |a-b|   3: relu(k0 + k1*a)   4: a
LCT_API int lct_mt_reduce(const void* const* a, const void* const* b, const int64_t* n, const float* scale,
                          int64_t nseg, int op, float k0, float k1, float* out, cudaStream_t st) {
    Segs S;
    if (!out) return LCT_EINVAL;
    int chunks = build(S, a, b, nullptr, n, scale, nseg, op, k0, k1, false);
    if (chunks < 0) return chunks;
    mt_kernel<false><<<chunks, kThreads, 0, st>>>(S, out, nullptr);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// g_i[j] = upstream[0] * scale[i] * d op / d a   (upstream may be null = 1)
LCT_API int lct_mt_grad(const void* const* a, const void* const* b, void* const* g, const int64_t* n,
                        const float* scale, int64_t nseg, int op, float k0, float k1, const float* upstream,
                        cudaStream_t st) {
    Segs S;
    int chunks = build(S, a, b, g, n, scale, nseg, op, k0, k1, true);
    if (chunks < 0) return chunks;
    mt_kernel<true><<<chunks, kThreads, 0, st>>>(S, nullptr, upstream);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

// dst_i[j] = src_i[j] for nseg tensors in one launch (parameter packing for the GRU banks)
LCT_API int lct_mt_copy(const void* const* src, void* const* dst, const int64_t* n, int64_t nseg, cudaStream_t st) {
    Segs S;
    float ones[kMaxSeg];
    for (int i = 0; i < kMaxSeg; ++i) ones[i] = 1.f;
    int chunks = build(S, src, nullptr, dst, n, ones, nseg, OP_SUM, 0.f, 0.f, true);
    if (chunks < 0) return chunks;
    mt_copy_kernel<<<chunks, kThreads, 0, st>>>(S);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}

LCT_API int lct_mt_max_segments(void) { return kMaxSeg; }

// In-place gradient clipping by global norm (see mt_clip_kernel).  g: HOST array of device pointers; sumsq: device
// float holding the sum of squares over ALL gradients being clipped (several launches may share it); norm_out optional.
LCT_API int lct_mt_clip(void* const* g, const int64_t* n, int64_t nseg, const float* sumsq, float max_norm, float pre,
                        float* norm_out, cudaStream_t st) {
    if (!g || !n || !sumsq || nseg <= 0 || nseg > kMaxSeg || !(max_norm > 0.f) || !(pre > 0.f)) return LCT_EINVAL;
    Segs S;
    int chunks = 0;
    for (int i = 0; i < nseg; ++i) {
        if (!g[i] || n[i] <= 0) return LCT_EINVAL;
        S.g[i] = (float*)g[i];
        S.n[i] = n[i];
        S.chunk0[i] = chunks;
        chunks += (int)ceil_div64(n[i], kChunk);
    }
    S.chunk0[nseg] = chunks;
    S.nseg = (int)nseg;
    mt_clip_kernel<<<chunks, kThreads, 0, st>>>(S, sumsq, max_norm, pre, norm_out);
    LCT_RETURN_IF_LAUNCH_FAILED();
    return 0;
}
