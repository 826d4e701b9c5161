/* lctgan.h -- C ABI of liblctgan_sm100.so: hand-written sm_100a kernels for the LCT-GAN
 * adversarial training step (reference: jqshang/LCT-GAN).
 *
 * The reference has no FFI of its own: its hot path is Python calling PyTorch operators
 * (SURVEY.md section 8b).  This library is what sits under the reference-compatible Python modules
 * (lct-gan_b200/{datasets,models,losses.py}); every entry point names the reference call it
 * replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - Every function returns int: 0 = ok, <0 = invalid / unsupported argument (-1 EINVAL,
 *     -2 unsupported shape), >0 = cudaError_t of the failed launch.  No exceptions, no allocation:
 *     all buffers (inputs, outputs, workspaces) are device pointers owned by the caller, work is
 *     enqueued on `stream` and nothing synchronises.
 *   - float* = fp32 device memory, densely packed in the layout written next to it.
 *   - Spectrograms are [B, Tf, F] (frequency innermost) complex-interleaved, i.e. the memory
 *     layout behind the [B, F, Tf] view torch.stft returns.
 *   - Generator activations are channels-last [B, T, F, C]; discriminator activations are the
 *     reference's NCHW [B, C, L, P] (P = period, 1 for the scale discriminators).
 *   - act codes: 0 none, 1 LeakyReLU(slope), 2 ReLU.
 *   - Accumulating outputs (documented per function) must be zeroed by the caller.
 *
 * The parser in lct-gan_b200/lctgan/_lib.py reads this file to build the ctypes signatures, so
 * keep one declaration per statement: `LCT_API int name(type arg, ...);`.
 */
#ifndef LCTGAN_H_
#define LCTGAN_H_

#include <stdint.h>

#ifdef __cplusplus
#define LCT_API extern "C"
#else
#define LCT_API
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ---------------------------------------------------------------- library info */
LCT_API int lct_version(void);                 /* ABI version, bumped on any signature change */
LCT_API int lct_kernel_launches(void);         /* number of kernels this library has launched in this process */
LCT_API int lct_reset_kernel_launches(void);

/* ---------------------------------------------------------------- STFT / iSTFT  (datasets/stft.py) */
/* 1 if n_fft is supported (even, 8..2048, prime factors 2/3/5). */
LCT_API int lct_fft_supported(int64_t n_fft);
/* Number of complex entries of the twiddle buffer for n_fft (2 n_fft + 64). */
LCT_API int lct_fft_twiddle_len(int64_t n_fft);
/* tw[lct_fft_twiddle_len(n_fft)][2], generated in double precision on the device:
 * [0, N) exp(-2 pi i m / N) | [N, N+64) W_64^(m0 k1) at m0*8+k1 | [N+64, 2N+64) W_N^(q j) at q*64+j
 * (the last two feed the register-resident warp FFT used for n_fft = 320 / 512 / 768). */
LCT_API int lct_fft_twiddles(float* tw, int64_t n_fft, cudaStream_t stream);
/* 1: force the generic shared-memory Stockham kernels for every n_fft (A/B measurements); 0: default dispatch. */
LCT_API int lct_fft_force_generic(int on);
/* env[n_fft + hop*(n_frames-1)] = overlap-added squared window (torch.istft's window_envelop). */
LCT_API int lct_ola_envelope(const float* window, float* env, int64_t n_fft, int64_t hop, int64_t n_frames, cudaStream_t stream);
/* ComplexSTFT.forward, stft.py:59-88 (torch.stft: center, reflect, onesided, unnormalised).
 * x [B,T] -> spec [B,Tf,F,2], Tf = 1 + T/hop, F = n_fft/2+1; optional mag [B,Tf,F] = max(|X|, eps)
 * (magnitude, stft.py:138-160). */
LCT_API int lct_stft_fwd(const float* x, const float* window, const float* tw, float* spec, float* mag, int64_t B, int64_t T, int64_t n_fft, int64_t hop, float eps, cudaStream_t stream);
/* Adjoint of lct_stft_fwd: gspec [B,Tf,F,2] -> gx [B,T]; work holds B*(n_fft + hop*(Tf-1)) floats. */
LCT_API int lct_stft_bwd(const float* gspec, const float* window, const float* tw, float* work, float* gx, int64_t B, int64_t T, int64_t n_fft, int64_t hop, cudaStream_t stream);
/* ComplexSTFT.istft, stft.py:90-132 (torch.istft with length).  If mask_c != NULL the spectrum is first
 * multiplied by max(mask_c, eps)^(1/c) (apply_mask(compressed=True), stft.py:243-290). */
LCT_API int lct_istft_fwd(const float* spec, const float* mask_c, const float* window, const float* tw, const float* env, float* y, int64_t B, int64_t n_frames, int64_t n_fft, int64_t hop, int64_t length, float c, float eps, cudaStream_t stream);
/* Adjoint of lct_istft_fwd: gy [B,length] -> gspec.  With mask_c/xspec: gmask = dL/dmask_c and gspec (optional) = dL/dxspec. */
LCT_API int lct_istft_bwd(const float* gy, const float* window, const float* tw, const float* env, float* gspec, const float* xspec, const float* mask_c, float* gmask, int64_t B, int64_t n_frames, int64_t n_fft, int64_t hop, int64_t length, float c, float eps, cudaStream_t stream);
/* TFFeatures.forward, tf_features.py:85-146: both STFTs, |X|, IRM^c and |X|^c in one kernel. */
LCT_API int lct_tf_features_fwd(const float* noisy, const float* clean, const float* window, const float* tw, float* noisy_mag, float* irm_c, float* noisy_mag_c, float* noisy_spec, float* clean_spec, int64_t B, int64_t T, int64_t n_fft, int64_t hop, float c, float gamma, float eps, cudaStream_t stream);
/* One resolution of MultiResolutionSTFTLoss, losses.py:66-80, without materialising spectrograms:
 * acc is [2][64] (zeroed by the caller): acc[0][s] += partial sums of (|Yh|_eps - |Y|_eps)^2, acc[1][s] += partial sums of
 * |Yh - Y|^2 (64 slots each so that thousands of CTAs do not serialise on one address; the caller adds the slots). */
LCT_API int lct_mrstft_sums(const float* y_hat, const float* y, const float* window, const float* tw, float* acc, int64_t B, int64_t T, int64_t n_fft, int64_t hop, float eps, cudaStream_t stream);
/* dL/dYh for that resolution: k_mag * d/dYh (|Yh|_eps-|Y|_eps)^2 + k_cplx * d/dYh |Yh-Y|^2, times upstream[0]. */
LCT_API int lct_mrstft_grad_spec(const float* spec_hat, const float* spec_ref, float* gspec, int64_t n_bins, float eps, float k_mag, float k_cplx, const float* upstream, cudaStream_t stream);

/* ---------------------------------------------------------------- spectral elementwise (datasets/stft.py) */
LCT_API int lct_magnitude_fwd(const float* spec, float* mag, int64_t n, float power, float eps, cudaStream_t stream);      /* stft.py:138-160 */
LCT_API int lct_magnitude_bwd(const float* spec, const float* gmag, float* gspec, int64_t n, float power, float eps, cudaStream_t stream);
LCT_API int lct_powclamp_fwd(const float* x, float* y, int64_t n, float e, float eps, cudaStream_t stream);                /* compress/decompress, stft.py:163-178 */
LCT_API int lct_powclamp_bwd(const float* x, const float* gy, float* gx, int64_t n, float e, float eps, cudaStream_t stream);
LCT_API int lct_irm_fwd(const float* clean_spec, const float* noisy_spec, float* irm_c, int64_t n, float c, float gamma, float eps, cudaStream_t stream);   /* stft.py:184-218 */
LCT_API int lct_apply_mask_fwd(const float* spec, const float* mask, float* out, int64_t n, int compressed, float c, float eps, cudaStream_t stream);      /* stft.py:243-290 */
LCT_API int lct_apply_mask_bwd(const float* spec, const float* mask, const float* gout, float* gmask, float* gspec, int64_t n, int compressed, float c, float eps, cudaStream_t stream);
LCT_API int lct_add2d(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo, int64_t M, int64_t N, cudaStream_t stream);   /* out[m,n] = a[m,n] + b[m,n], row-strided */
LCT_API int lct_axpby(const float* a, const float* b, float* out, int64_t n, float ka, float kb, cudaStream_t stream);

/* ---------------------------------------------------------------- discriminator framing (models/discriminators.py) */
/* F.pad(x, (0,pad), "reflect") of PeriodDiscriminator.forward, discriminators.py:84-88: integer indexing, bit exact. */
LCT_API int lct_reflect_pad_right_fwd(const float* x, float* y, int64_t B, int64_t T, int64_t pad, cudaStream_t stream);
LCT_API int lct_reflect_pad_right_bwd(const float* gy, float* gx, int64_t B, int64_t T, int64_t pad, cudaStream_t stream);
/* AvgPool1d(4, 2, padding=2, count_include_pad=False), discriminators.py:252-255: [B,L] -> [B,L/2+1]. */
LCT_API int lct_avgpool4_fwd(const float* x, float* y, int64_t B, int64_t L, cudaStream_t stream);
LCT_API int lct_avgpool4_bwd(const float* gy, float* gx, int64_t B, int64_t L, cudaStream_t stream);

/* ---------------------------------------------------------------- discriminator convolutions */
/* weight_norm (torch.nn.utils.weight_norm, dim 0; discriminators.py:46-66, :176-196): w = g * v / ||v||, rows of length `row`. */
LCT_API int lct_weight_norm_fwd(const float* g, const float* v, float* w, float* norm, int64_t Cout, int64_t row, cudaStream_t stream);
LCT_API int lct_weight_norm_bwd(const float* g, const float* v, const float* dw, float* dg, float* dv, int64_t Cout, int64_t row, cudaStream_t stream);
/* The same for up to 16 layers in one launch (g/v/w/dw/dg/dv: HOST arrays of device pointers; rows = Cout, rowlen per layer). */
/* img_f/img_d (optional HOST arrays, NULL entries allowed): also write the zero-padded TF32 weight images the tensor-core conv
 * kernels stage (caller zero-fills them); geo: HOST int64[nseg][8] = {Cout/G, K, S, ceil(K/S), KKpad_f, NS_f, KKpad_d, NS_d}. */
LCT_API int lct_mt_weight_norm_fwd(const void* const* g, const void* const* v, void* const* w, const int64_t* rows, const int64_t* rowlen, void* const* img_f, void* const* img_d, const int64_t* geo, int64_t nseg, cudaStream_t stream);
LCT_API int lct_mt_weight_norm_bwd(const void* const* g, const void* const* v, const void* const* dw, void* const* dg, void* const* dv, const int64_t* rows, const int64_t* rowlen, int64_t nseg, cudaStream_t stream);
/* Conv2d(k=(K,1), stride=(S,1), pad=(pad,0), groups=G) / Conv1d(K,S,pad,G) + bias + activation
 * (discriminators.py:93-98, :215-220).  x [B,Cin,Lin,P] -> y [B,Cout,Lout,P], w [Cout,Cin/G,K]. */
LCT_API int lct_conv1d_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);
/* dx = (conv^T(dy, w) + gextra) * act'(xact); gextra (feature-matching gradient of the same map) and xact (the
 * post-activation input of the layer) are optional. */
LCT_API int lct_conv1d_dgrad(const float* dy, const float* w, float* dx, const float* gextra, const float* xact, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);
/* dw [Cout,Cin/G,K] and db [Cout] (optional) are accumulated. */
LCT_API int lct_conv1d_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, cudaStream_t stream);

/* ---- grouped strided convolutions on tcgen05 (csrc/conv_tc.cu): UMMA kind::tf32 with both operands read from shared
 * memory through descriptors, accumulators in TMEM.  Same operator and tensor layouts as lct_conv1d_* (reference
 * models/discriminators.py:37-67, :93-98, :166-196, :215-220): x [B,Cin,Lin,P], w [Cout,Cin/G,K], pad = K/2.
 * Weights are passed as per-layer IMAGES (tf32-rounded, arranged per group in the kernels' shared-memory layout) built
 * by lct_conv_tc_images for up to 8 layers per launch; geo = HOST array of 6 int64 per layer (Cin, Cout, G, K, S, P). */
LCT_API int lct_conv_tc_supported(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P);
LCT_API int lct_conv_tc_image_len(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P, int64_t* out);   /* out[0] fwd, out[1] dgrad floats (HOST) */
LCT_API int lct_conv_tc_images(const void* const* w, void* const* img_f, void* const* img_d, const int64_t* geo, int64_t n, cudaStream_t stream);
LCT_API int lct_conv_tc_fwd(const float* x, const float* wimg, const float* bias, float* y, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);
LCT_API int lct_conv_tc_dgrad(const float* dy, const float* wimg, float* dx, const float* gextra, const float* xact, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);

/* The same grouped convolutions on the tensor cores: TF32 mma.sync implicit GEMM, fp32 accumulation, persistent CTAs with
 * cp.async double-buffered input windows (conv_mma.cu).  Same arguments as lct_conv1d_*; lct_conv_mma_supported says
 * whether a layer shape is covered (groups with <= 16 in / <= 32 out channels, stride 1/3/4, Cin/G * K <= 168). */
LCT_API int lct_conv_mma_supported(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t P);
LCT_API int lct_conv_mma_image_geometry(int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int mode, int64_t* out);   /* out (HOST) = {KKpad, NS} of the staged weight image; mode 0 fwd, 1 dgrad */
LCT_API int lct_conv_mma_fwd(const float* x, const float* w, const float* wimg, const float* bias, float* y, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);
LCT_API int lct_conv_mma_dgrad(const float* dy, const float* w, const float* wimg, float* dx, const float* gextra, const float* xact, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, int act, float slope, cudaStream_t stream);
/* weight-gradient grids launched from now on target n resident CTAs per SM (0: default 2): the G step's dead gradients run with 1 beside the generator's backward */
LCT_API int lct_set_wgrad_ctas(int n);
LCT_API int lct_conv_mma_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t Cin, int64_t Cout, int64_t G, int64_t K, int64_t S, int64_t pad, int64_t Lin, int64_t P, cudaStream_t stream);

/* conv_post (C -> 1 channel, odd K <= 8, stride 1, pad K/2; discriminators.py:59-66, :188-196): channel-reduction kernels.
 * y [B,1,L,P] is overwritten (each CTA reduces all channels of its positions: no atomics); dw/db accumulate. */
LCT_API int lct_conv_post_fwd(const float* x, const float* w, const float* bias, float* y, int64_t B, int64_t C, int64_t L, int64_t P, int64_t K, cudaStream_t stream);
LCT_API int lct_conv_post_wgrad(const float* x, const float* dy, float* dw, float* db, int64_t B, int64_t C, int64_t L, int64_t P, int64_t K, cudaStream_t stream);
LCT_API int lct_conv_post_dgrad(const float* dy, const float* w, float* dx, const float* gextra, const float* xact, int64_t B, int64_t C, int64_t L, int64_t P, int64_t K, int act, float slope, cudaStream_t stream);

/* The dense layer MSD convs.5 (Conv1d 1024->1024, k=5, s=1; discriminators.py:166-196) on tcgen05 tensor cores:
 * bf16 operands staged once per call, fp32 accumulation in TMEM, TMA-fed implicit GEMM (no im2col).
 * lct_dense_supported: 1 if (Cin, Cout multiples of 128, odd K <= 8) and the driver exposes cuTensorMapEncodeTiled. */
LCT_API int lct_dense_supported(int64_t Cin, int64_t Cout, int64_t K);
LCT_API int lct_stage_nlc_bf16(const float* x, void* out, int64_t B, int64_t C, int64_t L, int64_t pad, cudaStream_t stream);   /* [B,C,L] f32 -> [B,L+2pad,C] bf16, zero rows between batches */
LCT_API int lct_stage_ncl_bf16(const float* x, void* out, float* rowsum, int64_t B, int64_t C, int64_t L, int64_t Lp, int64_t shift, int64_t pitch, int64_t copies, cudaStream_t stream);   /* [B,C,L] f32 -> [copies,C,pitch] bf16, out[k][c][b*Lp+shift-k+l]; rowsum[C] += sums (bias grad) */
LCT_API int lct_stage_dense_weights(const float* w, void* wt, void* wd, int64_t Co, int64_t Ci, int64_t K, cudaStream_t stream);   /* [Co,Ci,K] f32 -> wt [K,Co,Ci], wd [K,Ci,Co] taps flipped (bf16) */
/* out f32 [B,Cn,L] = epilogue(sum_{tap,ca} a[b*(L+K-1)+l+tap, ca] * w[tap][cn][ca]); forward: bias+act; dgrad: (acc+gextra)*act'(xact). */
LCT_API int lct_dense_conv(const void* a, const void* w, const float* bias, const float* gextra, const float* xact, float* out, int64_t B, int64_t L, int64_t Ca, int64_t Cn, int64_t K, int act, float slope, cudaStream_t stream);
/* dw f32 [Co,Ci,K] = sum_r dyq[co][r] * xq[tap][ci][r]  (dyq: 1 copy, shift 0; xq: K copies, shift K/2, from lct_stage_ncl_bf16). */
LCT_API int lct_dense_wgrad(const void* dyq, const void* xq, float* dw, int64_t Co, int64_t Ci, int64_t K, int64_t pitch, cudaStream_t stream);

/* ---------------------------------------------------------------- generator (models/generator.py) */
/* nn.Linear / GRU input projections / attention projections and their gradients (generator.py:104, :133, :138, :219, :245, :248).
 * C[M,N] = act(alpha * op(A) op(B) + bias) (+C); out2 = C + res.  ta: A(m,k)=A[k*lda+m]; tb: B(k,n)=B[k*ldb+n] else B[n*ldb+k].
 * Batched over z: A += (z/a_div)*sA, B += (z/b_div)*sB, C/bias/res/out2 += z*s.  ksplit>1: split-K with atomic adds. */
LCT_API int lct_gemm(const float* A, const float* B, float* C, const float* bias, const float* res, float* out2, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int64_t ldo, int ta, int tb, int act, float slope, float alpha, int accumulate, int64_t ksplit, int64_t nbatch, int64_t a_div, int64_t b_div, int64_t sA, int64_t sB, int64_t sC, int64_t sBias, int64_t sRes, int64_t sOut2, cudaStream_t stream);
LCT_API int lct_set_rowgemm(int on);            /* 1 (default): NT/NN GEMMs with aligned operands, the TN weight-gradient GEMMs (split-K form) and lct_gconv_wgrad run on the fp32-accurate 3xTF32 tensor-core row GEMMs (rowgemm.cu); 0: fp32 SIMT */
LCT_API int lct_set_tensor_core_gemm(int on);   /* 1: lct_gemm runs on TF32 mma.sync (experimental, slower on the skinny generator shapes); 0 (default): fp32 SIMT */
LCT_API int lct_colsum(const float* X, float* out, int64_t M, int64_t N, int64_t ld, cudaStream_t stream);   /* out[N] += column sums (bias gradients) */
/* nn.LayerNorm(C) over rows [M,C] (generator.py:126, :132, :238, :244, :577). */
LCT_API int lct_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int64_t M, int64_t C, float eps, cudaStream_t stream);
LCT_API int lct_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, const float* dres, float* dx, float* dgamma, float* dbeta, int64_t M, int64_t C, cudaStream_t stream);
/* The recurrent part of the 4 (x2 directions) nn.GRU(16,16) of a block (generator.py:94-110, :211-222).
 * gi [rows,GD,48] = W_ih x + b_ih; whh [GD,48,16]; bhh [GD,48]; hs [rows,GD,16].  gd = group*D + dir, dir 1 = reverse.
 * Sequence s starts at row (s/inner)*outer_stride + (s%inner)*inner_stride and steps by step_stride rows. */
LCT_API int lct_gru_fwd(const float* gi, const float* whh, const float* bhh, float* hs, float* gsave, float* hprev, int64_t nseq, int64_t L, int64_t GD, int64_t D, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t step_stride, cudaStream_t stream);
/* BPTT: dh_in [rows,ldd] (column (gd/D)*16+j) -> dgi [rows,GD,48]; dwhh/dbih/dbhh accumulated. */
LCT_API int lct_gru_bwd(const float* gsave, const float* hprev, const float* whh, const float* dh_in, int64_t ldd, float* dgi, float* dgh, int64_t nseq, int64_t L, int64_t GD, int64_t D, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t step_stride, cudaStream_t stream);
/* seq = x + sum_dir hs (generator.py:105-107, :128);  gsum (optional, row stride ldg) = the GRU output alone. */
LCT_API int lct_gru_combine(const float* x, const float* hs, float* seq, float* gsum, int64_t ldg, int64_t M, int64_t G, int64_t D, cudaStream_t stream);
/* softmax(q k^T / 4) v per head of nn.MultiheadAttention(64,4) (generator.py:133, :245); qkv [rows,3*16*heads]. */
LCT_API int lct_attn_fwd(const float* qkv, float* out, float* lse, int64_t heads, int64_t nseq, int64_t L, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t step_stride, cudaStream_t stream);
LCT_API int lct_attn_bwd(const float* qkv, const float* out, const float* lse, const float* dout, float* dqkv, int64_t heads, int64_t nseq, int64_t L, int64_t inner, int64_t outer_stride, int64_t inner_stride, int64_t step_stride, cudaStream_t stream);
LCT_API int lct_act_bwd(const float* y, const float* dy, float* dpre, int64_t n, int act, float slope, cudaStream_t stream);
/* Conv2d(k=(2,3), s=(1,2), p=(1,1)) (transposed=0, w [Cd,Cs,2,3]; generator.py:461-481) and
 * ConvTranspose2d(k=(2,3), s=(1,2), p=(1,1), op=(0,1)) (transposed=1, w [Cs,Cd,2,3]; generator.py:506-529) on
 * channels-last [B,T,F,C]; each is also the other's data gradient (gmul: multiply by act'(gmul)). */
LCT_API int lct_gconv(const float* in, const float* w, const float* bias, float* out, const float* gmul, int transposed, int64_t B, int64_t Ti, int64_t Fi, int64_t Cs, int64_t To, int64_t Fo, int64_t Cd, int act, float slope, int gact, float gslope, cudaStream_t stream);
/* dW[Ca][Cc][2][3] += sum S[b,t,f,a] * Lg[b,t+kt-1,2f+kf-1,c]  (conv: S=dOut, Lg=in; deconv: S=in, Lg=dOut). */
/* Implicit row-GEMM form of lct_gconv on the tensor cores (3xTF32, fp32-level accuracy; rowgemm.cu).
 * lct_gconv_weight_image re-arranges a [Cd][Cs][2][3] (transposed = 0) or [Cs][Cd][2][3] (transposed = 1) weight into
 * the [K][Cd] image(s) the kernel reads (lct_gconv_image_len floats); lct_gconv_mma then equals lct_gconv. */
LCT_API int lct_gconv_mma_supported(int64_t Cs, int64_t Cd);
LCT_API int lct_gconv_image_len(int64_t Cs, int64_t Cd);
LCT_API int lct_gconv_weight_image(const float* w, float* img, int transposed, int64_t Cs, int64_t Cd, cudaStream_t stream);
LCT_API int lct_gconv_mma(const float* in, const float* img, const float* bias, float* out, const float* gmul, int transposed, int64_t B, int64_t Ti, int64_t Fi, int64_t Cs, int64_t To, int64_t Fo, int64_t Cd, int act, float slope, int gact, float gslope, cudaStream_t stream);
LCT_API int lct_gconv_wgrad(const float* S, const float* Lg, float* dW, int64_t B, int64_t Ts, int64_t Fs, int64_t Ca, int64_t Tl, int64_t Fl, int64_t Cc, cudaStream_t stream);
/* h[:, :To, :Fo] + skipN(mag)[:, :To, :Fo] with skipN the 1x1 Conv2d(1->C) (generator.py:484-498, :587-598). */
LCT_API int lct_skip_add_fwd(const float* h, const float* mag, const float* w, const float* bias, float* out, int64_t B, int64_t Th, int64_t Fh, int64_t Tm, int64_t Fm, int64_t C, cudaStream_t stream);
LCT_API int lct_skip_add_bwd(const float* g, const float* mag, float* dh, float* dw, float* db, int64_t B, int64_t Th, int64_t Fh, int64_t Tm, int64_t Fm, int64_t C, cudaStream_t stream);
LCT_API int lct_crop_pad(const float* src, float* dst, int64_t B, int64_t Ts, int64_t Fs, int64_t Td, int64_t Fd, int64_t C, cudaStream_t stream);
/* crop / zero-pad to [T,F] then sigmoid (generator.py:601-630). */
LCT_API int lct_final_mask_fwd(const float* y, float* mask, int64_t B, int64_t Ty, int64_t Fy, int64_t T, int64_t F, int use_sigmoid, cudaStream_t stream);
LCT_API int lct_final_mask_bwd(const float* y, const float* mask, const float* gmask, float* dpre, int64_t B, int64_t Ty, int64_t Fy, int64_t T, int64_t F, int use_sigmoid, int act, float slope, cudaStream_t stream);

/* ---------------------------------------------------------------- losses (losses.py) */
/* a, b, g are HOST arrays of device pointers; n and scale are HOST arrays.
 * out[0] += sum_i scale[i] * sum_j op(a_i[j], b_i[j]) over nseg <= lct_mt_max_segments() tensors.
 * op 0: (a-k0)^2  1: (a-b)^2  2: |a-b|  3: relu(k0+k1*a)  4: a.  Covers discriminator_loss (losses.py:110-135),
 * generator_adv_loss (:138-151), feature_matching_loss (:154-173), mask_mse_loss (:176-181). */
LCT_API int lct_mt_reduce(const void* const* a, const void* const* b, const int64_t* n, const float* scale, int64_t nseg, int op, float k0, float k1, float* out, cudaStream_t stream);
/* g_i[j] = upstream[0] * scale[i] * d op / d a. */
LCT_API int lct_mt_grad(const void* const* a, const void* const* b, void* const* g, const int64_t* n, const float* scale, int64_t nseg, int op, float k0, float k1, const float* upstream, cudaStream_t stream);
/* dst_i[j] = src_i[j] for nseg tensors in one launch (packs the 16/32 GRU parameter tensors of a block). */
LCT_API int lct_mt_copy(const void* const* src, void* const* dst, const int64_t* n, int64_t nseg, cudaStream_t stream);
LCT_API int lct_mt_max_segments(void);
/* Fused multi-tensor AdamW (the optimiser train.py:601-610 builds): p/g/m/v are HOST arrays of device
 * pointers (<= lct_mt_adamw_max_segments() tensors per launch), n HOST element counts, step a device float that already
 * holds this update's step number (lct_add_scalar increments it on the stream: CUDA-graph friendly). */
LCT_API int lct_mt_adamw(void* const* p, const void* const* g, void* const* m, void* const* v, const int64_t* n, int64_t nseg, const float* step, float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, cudaStream_t stream);
LCT_API int lct_mt_adamw_max_segments(void);
LCT_API int lct_add_scalar(float* x, float v, cudaStream_t stream);
/* torch.nn.utils.clip_grad_norm_ (reference train.py:246-248) in place over <= lct_mt_max_segments() gradient tensors:
 * total = pre * sqrt(sumsq[0]); g *= pre * min(1, max_norm / (total + 1e-6)).  sumsq: device float = sum of squares of
 * ALL clipped gradients (lct_mt_reduce op 0); pre folds the 1/world_size of a data-parallel SUM; norm_out optional. */
LCT_API int lct_mt_clip(void* const* g, const int64_t* n, int64_t nseg, const float* sumsq, float max_norm, float pre, float* norm_out, cudaStream_t stream);
/* cudaMemsetAsync(p, 0, bytes) on `stream` (a memset node in a captured graph; no kernel).  Replaces the per-tensor
 * zero-fill kernels of gradient accumulators: one buffer per layer stack, cleared once. */
LCT_API int lct_memset_zero(void* p, int64_t bytes, cudaStream_t stream);

/* ---- callers on either side of the hot path (csrc/pipeline.cu) ----
 * lct_si_sdr: out[b] = SI-SDR (dB) of est[b, :L] against ref[b, :L], L = min(lengths[b], T_ref, T_est), both zero-mean over
 * that span (reference train.py:261-282 _si_sdr_torch, there one call + host sync per utterance); lengths (device int64)
 * optional.  lct_crop_segments: out[b, t] = src[offset[b] + start[b] + t] while start[b] + t < end[b], else 0 - the loader's
 * segment cropping (datasets/datasets.py:131-156 _crop_pair) and collate_fn zero padding (:187-230) on utterances packed in
 * one device buffer; offset / end / start: device int64 [B]; clean side optional. */
LCT_API int lct_si_sdr(const float* ref, const float* est, const int64_t* lengths, float* out, int64_t B, int64_t T_ref, int64_t T_est, float eps, cudaStream_t stream);
LCT_API int lct_crop_segments(const float* noisy, const float* clean, const int64_t* offset_n, const int64_t* offset_c, const int64_t* end_n, const int64_t* end_c, const int64_t* start, float* out_n, float* out_c, int64_t B, int64_t T, cudaStream_t stream);

#endif /* LCTGAN_H_ */
