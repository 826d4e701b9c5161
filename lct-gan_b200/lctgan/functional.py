"""Autograd bindings of the front-end and loss kernels (torch.autograd.Function wrappers).

Every forward/backward here enqueues lctgan kernels only; PyTorch is used to allocate buffers
and to carry the autograd graph.  Spectrograms cross this boundary as the reference's
[B, F, Tf] complex64 tensors (a transposed view of the physical [B, Tf, F] buffer, exactly what
torch.stft returns).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("lctgan is CUDA only (sm_100a); got a CPU tensor and there is no CPU fallback")


# --------------------------------------------------------------------------------------- STFT
class STFTFn(torch.autograd.Function):
    """ComplexSTFT.forward (reference datasets/stft.py:59-88)."""

    @staticmethod
    def forward(ctx, x, window, n_fft, hop):
        _require_cuda(x, window)
        spec, _ = ops.stft_fwd(x, window, n_fft, hop)
        ctx.save_for_backward(window)
        ctx.geom = (x.shape[1], n_fft, hop)
        return ops.spec_view(spec)

    @staticmethod
    def backward(ctx, g):
        (window,) = ctx.saved_tensors
        T, n_fft, hop = ctx.geom
        return ops.stft_bwd(ops.spec_phys(g), window, T, n_fft, hop), None, None, None


class ISTFTFn(torch.autograd.Function):
    """ComplexSTFT.istft (reference datasets/stft.py:90-132)."""

    @staticmethod
    def forward(ctx, spec, window, n_fft, hop, length):
        _require_cuda(spec, window)
        phys = ops.spec_phys(spec)
        ctx.save_for_backward(window)
        ctx.geom = (n_fft, hop, phys.shape[1])
        return ops.istft_fwd(phys, window, n_fft, hop, length)

    @staticmethod
    def backward(ctx, gy):
        (window,) = ctx.saved_tensors
        n_fft, hop, Tf = ctx.geom
        gspec, _ = ops.istft_bwd(gy, window, n_fft, hop, Tf)
        return ops.spec_view(gspec), None, None, None, None


class MaskedISTFTFn(torch.autograd.Function):
    """apply_mask(compressed=True) + istft fused (reference models/generator.py:684-695).
    The noisy spectrum is data (no gradient); the gradient goes to the compressed mask."""

    @staticmethod
    def forward(ctx, spec_phys, mask_phys, window, n_fft, hop, length, c, eps):
        ctx.save_for_backward(spec_phys, mask_phys, window)
        ctx.cfg = (n_fft, hop, length, c, eps)
        return ops.istft_fwd(spec_phys, window, n_fft, hop, length, mask_c=mask_phys, c=c, eps=eps)

    @staticmethod
    def backward(ctx, gy):
        spec_phys, mask_phys, window = ctx.saved_tensors
        n_fft, hop, length, c, eps = ctx.cfg
        _, gmask = ops.istft_bwd(gy, window, n_fft, hop, spec_phys.shape[1], xspec=spec_phys, mask_c=mask_phys,
                                 want_gspec=False, c=c, eps=eps)
        return None, gmask, None, None, None, None, None, None


class MagnitudeFn(torch.autograd.Function):
    """magnitude (reference datasets/stft.py:138-160)."""

    @staticmethod
    def forward(ctx, spec, power, eps):
        _require_cuda(spec)
        ctx.transposed = spec.dim() >= 2 and not spec.is_contiguous() and spec.transpose(-1, -2).is_contiguous()
        phys = spec.transpose(-1, -2) if ctx.transposed else spec.contiguous()
        ctx.save_for_backward(phys)
        ctx.cfg = (power, eps)
        mag = ops.magnitude_fwd(phys, power, eps)
        return mag.transpose(-1, -2) if ctx.transposed else mag

    @staticmethod
    def backward(ctx, g):
        (phys,) = ctx.saved_tensors
        power, eps = ctx.cfg
        gp = g.transpose(-1, -2) if ctx.transposed else g
        gs = ops.magnitude_bwd(phys, gp.contiguous(), power, eps)
        return (gs.transpose(-1, -2) if ctx.transposed else gs), None, None


def _same_layout(x):
    """Return (dense tensor sharing x's memory order, restore fn) for elementwise kernels."""
    if x.is_contiguous():
        return x, (lambda y: y)
    if x.dim() >= 2 and x.transpose(-1, -2).is_contiguous():
        return x.transpose(-1, -2), (lambda y: y.transpose(-1, -2))
    return x.contiguous(), (lambda y: y)


class PowClampFn(torch.autograd.Function):
    """compress / decompress: max(x, eps)^e (reference datasets/stft.py:163-178)."""

    @staticmethod
    def forward(ctx, x, e, eps):
        _require_cuda(x)
        dense, restore = _same_layout(x)
        ctx.save_for_backward(dense)
        ctx.cfg = (e, eps)
        ctx.transposed = dense is not x and dense.data_ptr() == x.data_ptr()
        return restore(ops.powclamp_fwd(dense, e, eps))

    @staticmethod
    def backward(ctx, g):
        (dense,) = ctx.saved_tensors
        e, eps = ctx.cfg
        gd = g.transpose(-1, -2) if ctx.transposed else g
        gx = ops.powclamp_bwd(dense, gd.contiguous(), e, eps)
        return (gx.transpose(-1, -2) if ctx.transposed else gx), None, None


class ApplyMaskFn(torch.autograd.Function):
    """apply_mask (reference datasets/stft.py:243-290); spec, mask are [B, F, T]."""

    @staticmethod
    def forward(ctx, spec, mask, compressed, c, eps):
        _require_cuda(spec, mask)
        sp = ops.spec_phys(spec)
        mp = ops.spec_phys(mask)
        ctx.save_for_backward(sp, mp)
        ctx.cfg = (compressed, c, eps)
        return ops.spec_view(ops.apply_mask_fwd(sp, mp, compressed, c, eps))

    @staticmethod
    def backward(ctx, g):
        sp, mp = ctx.saved_tensors
        compressed, c, eps = ctx.cfg
        need_s, need_m = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gmask, gspec = ops.apply_mask_bwd(sp, mp, ops.spec_phys(g), compressed, want_gmask=need_m or not need_s,
                                          want_gspec=need_s, c=c, eps=eps)
        return (ops.spec_view(gspec) if need_s else None, ops.spec_view(gmask) if need_m else None, None, None, None)


# --------------------------------------------------------------------------------------- waveform framing
class ReflectPadRightFn(torch.autograd.Function):
    """F.pad(x, (0, pad), mode='reflect') (reference models/discriminators.py:84-88)."""

    @staticmethod
    def forward(ctx, x, pad):
        ctx.geom = (x.shape[1], pad)
        return ops.reflect_pad_right_fwd(x.contiguous(), pad)

    @staticmethod
    def backward(ctx, g):
        T, pad = ctx.geom
        return ops.reflect_pad_right_bwd(g.contiguous(), T, pad), None


class AvgPool4Fn(torch.autograd.Function):
    """AvgPool1d(4, 2, padding=2, count_include_pad=False) (reference models/discriminators.py:252-255)."""

    @staticmethod
    def forward(ctx, x):
        ctx.L = x.shape[1]
        return ops.avgpool4_fwd(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return ops.avgpool4_bwd(g.contiguous(), ctx.L)


# --------------------------------------------------------------------------------------- losses
class MTLossFn(torch.autograd.Function):
    """sum_i scale_i * sum op(a_i, b_i) over a list of tensors in one launch (losses.py:110-181)."""

    @staticmethod
    def forward(ctx, op, k0, k1, scales, n_a, *tensors):
        a = list(tensors[:n_a])
        b = list(tensors[n_a:]) if len(tensors) > n_a else None
        _require_cuda(*a)
        a = [t.contiguous() for t in a]
        if b is not None:
            _require_cuda(*b)
            b = [t.contiguous() for t in b]
        out = ops.mt_reduce(a, b, scales, op, k0, k1)
        ctx.cfg = (op, k0, k1, list(scales), n_a)
        ctx.save_for_backward(*(a + (b or [])))
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        op, k0, k1, scales, n_a = ctx.cfg
        saved = ctx.saved_tensors
        a = list(saved[:n_a])
        b = list(saved[n_a:]) if len(saved) > n_a else None
        need_a = any(ctx.needs_input_grad[5:5 + n_a])
        need_b = b is not None and any(ctx.needs_input_grad[5 + n_a:])
        ga: List[Optional[torch.Tensor]] = [None] * n_a
        gb: List[Optional[torch.Tensor]] = [None] * (len(b) if b is not None else 0)
        if need_a or need_b:
            grads = ops.mt_grad(a, b, scales, op, k0, k1, upstream=g.reshape(1).contiguous())
            if need_a:
                ga = [gr if ctx.needs_input_grad[5 + i] else None for i, gr in enumerate(grads)]
            if need_b:   # difference ops only: d/db = -d/da
                gb = [(-gr) if ctx.needs_input_grad[5 + n_a + i] else None for i, gr in enumerate(grads)]
        return (None, None, None, None, None, *ga, *gb)


class DLossBatchedFn(torch.autograd.Function):
    """discriminator_loss (losses.py:68-85) over logits that hold [real; fake] in ONE batch (rows [0, nb) real): the two
    terms are reduced from the two batch halves inside the node and the backward writes both halves of one gradient
    tensor.  Slicing the logits in autograd instead (t[:nb], t[nb:]) costs a zero fill, a copy and an add per logits
    tensor and half in backward: ~55 tiny torch kernels between the loss and the start of the D backward."""

    @staticmethod
    def forward(ctx, loss_type, nb, *logits):
        _require_cuda(*logits)
        ts = [t.contiguous() for t in logits]
        n = max(len(ts), 1)
        real, fake = [t[:nb] for t in ts], [t[nb:] for t in ts]
        sr = [1.0 / (t.numel() * n) for t in real]
        sf = [1.0 / (t.numel() * n) for t in fake]
        if loss_type == "ls":
            cfg = ((ops.OP_SQ_CONST, 1.0, 0.0), (ops.OP_SQ_CONST, 0.0, 0.0))
        elif loss_type == "hinge":
            cfg = ((ops.OP_RELU_AFFINE, 1.0, -1.0), (ops.OP_RELU_AFFINE, 1.0, 1.0))
        else:
            raise ValueError(f"Unknown loss_type: {loss_type}")
        out = ops.mt_reduce(real, None, sr, cfg[0][0], cfg[0][1], cfg[0][2])
        ops.mt_reduce(fake, None, sf, cfg[1][0], cfg[1][1], cfg[1][2], out=out)
        ctx.cfg, ctx.nb, ctx.scales = cfg, nb, (sr, sf)
        ctx.save_for_backward(*ts)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        ts = list(ctx.saved_tensors)
        nb, (sr, sf), cfg = ctx.nb, ctx.scales, ctx.cfg
        grads = [torch.empty_like(t) for t in ts]
        up = g.reshape(1).contiguous()
        ops.mt_grad([t[:nb] for t in ts], None, sr, cfg[0][0], cfg[0][1], cfg[0][2], upstream=up, out=[x[:nb] for x in grads])
        ops.mt_grad([t[nb:] for t in ts], None, sf, cfg[1][0], cfg[1][1], cfg[1][2], upstream=up, out=[x[nb:] for x in grads])
        return (None, None, *[gr if ctx.needs_input_grad[2 + i] else None for i, gr in enumerate(grads)])


def d_loss_batched(logits: Sequence[torch.Tensor], nb: int, loss_type: str = "ls") -> torch.Tensor:
    return DLossBatchedFn.apply(loss_type, nb, *logits)


class WeightedSumFn(torch.autograd.Function):
    """sum_i w_i * s_i over scalar tensors in one launch (and one launch for all the gradients): the loss combinations
    ``lr + lf`` (losses.py:135) and ``mr + l_mask * m + l_adv * (adv + l_fm * fm)`` (train.py:240-243) without torch's
    one-kernel-per-scalar-operator chain."""

    @staticmethod
    def forward(ctx, weights, *scalars):
        _require_cuda(*scalars)
        a = [t.reshape(1).contiguous() for t in scalars]
        ctx.weights = [float(w) for w in weights]
        ctx.save_for_backward(*a)
        return ops.mt_reduce(a, None, ctx.weights, ops.OP_SUM).view(())

    @staticmethod
    def backward(ctx, g):
        a = list(ctx.saved_tensors)
        grads = ops.mt_grad(a, None, ctx.weights, ops.OP_SUM, upstream=g.reshape(1).contiguous())
        return (None, *[gr.view(()) if ctx.needs_input_grad[1 + i] else None for i, gr in enumerate(grads)])


def weighted_sum(scalars: Sequence[torch.Tensor], weights: Sequence[float]) -> torch.Tensor:
    return WeightedSumFn.apply(tuple(weights), *scalars)


def mt_loss(op: int, a: Sequence[torch.Tensor], b: Optional[Sequence[torch.Tensor]], scales: Sequence[float],
            k0: float = 0.0, k1: float = 0.0) -> torch.Tensor:
    tensors = list(a) + (list(b) if b is not None else [])
    return MTLossFn.apply(op, k0, k1, list(scales), len(a), *tensors)


class MRSTFTLossFn(torch.autograd.Function):
    """MultiResolutionSTFTLoss.forward (reference losses.py:54-100).  The forward reduces the
    spectral errors inside the STFT kernel (no spectrogram is written); the backward rebuilds the
    two spectra per resolution, forms dL/dY_hat and runs the adjoint STFT."""

    @staticmethod
    def forward(ctx, y_hat, y, res_cfg, mag_weight, complex_weight, eps, *windows):
        _require_cuda(y_hat, y)
        y_hat = y_hat.contiguous()
        y = y.contiguous()
        B, T = y_hat.shape
        nres = len(res_cfg)
        acc = ops.zeros((nres, 2, 64), y_hat.device)   # 64 partial-sum slots per sum
        wsum = sum(w for (_, _, w) in res_cfg)
        k_total, k_mag, k_cplx = [], [], []
        for i, ((n_fft, hop, w), win) in enumerate(zip(res_cfg, windows)):
            ops.mrstft_sums(y_hat, y, win, n_fft, hop, acc[i], eps)
            n = B * (n_fft // 2 + 1) * (1 + T // hop)
            norm = w / (wsum * n) if wsum > 0 else w / n
            k_mag.append(norm)
            k_cplx.append(norm)
        segs = [acc[i, j] for i in range(nres) for j in range(2)]
        sc_total = [k * (mag_weight if j == 0 else complex_weight) for i, k in enumerate(k_mag) for j in range(2)]
        sc_mag = [k if j == 0 else 0.0 for k in k_mag for j in range(2)]
        sc_cplx = [k if j == 1 else 0.0 for k in k_cplx for j in range(2)]
        total = ops.mt_reduce(segs, None, sc_total, ops.OP_SUM)
        mag_t = ops.mt_reduce(segs, None, sc_mag, ops.OP_SUM)
        cplx_t = ops.mt_reduce(segs, None, sc_cplx, ops.OP_SUM)
        ctx.save_for_backward(y_hat, y, *windows)
        ctx.cfg = (res_cfg, mag_weight, complex_weight, eps, k_mag)
        ctx.mark_non_differentiable(mag_t, cplx_t)
        return total.view(()), mag_t.view(()), cplx_t.view(())

    @staticmethod
    def backward(ctx, g, _gm, _gc):
        y_hat, y, *windows = ctx.saved_tensors
        res_cfg, mag_weight, complex_weight, eps, k = ctx.cfg
        if ctx.needs_input_grad[1]:
            raise RuntimeError("MultiResolutionSTFTLoss: gradient w.r.t. the reference waveform is not implemented")
        up = g.reshape(1).contiguous()
        T = y_hat.shape[1]
        gy = None
        for (n_fft, hop, _w), win, ki in zip(res_cfg, windows, k):
            sh, _ = ops.stft_fwd(y_hat, win, n_fft, hop)
            sr, _ = ops.stft_fwd(y, win, n_fft, hop)
            gs = ops.mrstft_grad_spec(sh, sr, ki * mag_weight, ki * complex_weight, upstream=up, eps=eps)
            gr = ops.stft_bwd(gs, win, T, n_fft, hop)
            if gy is None:
                gy = gr
            else:
                ops.call("lct_axpby", gy, gr, gy, gy.numel(), 1.0, 1.0)
        return (gy, None, None, None, None, None) + (None,) * len(windows)
