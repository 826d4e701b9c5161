"""Thin tensor-level wrappers over the C ABI (no autograd).  Each function allocates its outputs
with torch (memory plumbing only) and enqueues the lctgan kernels on the current stream.

Layouts: spectrograms are physically [B, Tf, F] complex64 (``spec_view`` gives the reference's
[B, F, Tf] view); generator activations are channels-last [B, T, F, C]; discriminator
activations are [B, C, L, P].
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import config
from ._lib import call, call_ret

ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2

_TW: Dict[Tuple[int, int], torch.Tensor] = {}
_ENV: Dict[tuple, list] = {}


def zeros(shape, device) -> torch.Tensor:
    """Zero-filled fp32 tensor: torch.empty + lct_memset_zero (one 16-byte-store kernel of this library on the current
    stream, which keeps the stream's priority inside a captured graph) instead of torch.zeros' fill kernel."""
    t = torch.empty(shape, dtype=torch.float32, device=device)
    if t.numel():
        call("lct_memset_zero", t, t.numel() * 4)
    return t


class Arena:
    """Bump allocator over ONE zero-filled buffer (one clearing launch) for the many small gradient accumulators of a backward
    pass; falls back to separate zeroed tensors when exhausted.  Views are 16-byte aligned."""

    def __init__(self, numel: int, device):
        self.buf = zeros((int(numel),), device)
        self.off = 0

    def __call__(self, *shape) -> torch.Tensor:
        n = int(math.prod(shape))
        if self.off + n > self.buf.numel():
            return zeros(shape, self.buf.device)
        v = self.buf[self.off:self.off + n].view(shape)
        self.off += (n + 3) // 4 * 4
        return v


def _dev_index(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def fft_supported(n_fft: int) -> bool:
    return bool(call_ret("lct_fft_supported", n_fft))


def twiddles(n_fft: int, like: torch.Tensor) -> torch.Tensor:
    """Twiddle tables (see lct_fft_twiddles in include/lctgan.h), a pure function of (N, device): cached."""
    key = (n_fft, _dev_index(like))
    tw = _TW.get(key)
    if tw is None:
        if not fft_supported(n_fft):
            raise RuntimeError(f"n_fft={n_fft} unsupported (need even, 8..2048, prime factors 2/3/5)")
        tw = torch.empty(call_ret("lct_fft_twiddle_len", n_fft), 2, device=like.device, dtype=torch.float32)
        call("lct_fft_twiddles", tw, n_fft)
        _TW[key] = tw
    return tw


_ENV_MAX = 64


def ola_envelope(window: torch.Tensor, n_fft: int, hop: int, n_frames: int) -> torch.Tensor:
    """Overlap-added squared window (iSTFT normalisation), cached.

    The key is (device, window storage address, window version, geometry) and the entry keeps a reference to the window
    tensor itself, so the address cannot be recycled for another window while the entry lives: a key match implies the
    same values.  Least-recently-used entries are dropped beyond 64 - except those created or used while a CUDA graph
    was being captured: the captured graph holds their addresses, so they are pinned for the life of the process."""
    key = (_dev_index(window), window.data_ptr(), window._version, n_fft, hop, n_frames)
    capturing = window.is_cuda and torch.cuda.is_current_stream_capturing()
    ent = _ENV.get(key)
    if ent is not None:
        _ENV[key] = _ENV.pop(key)                      # move to the most-recently-used end
        if capturing:
            ent[2] = True
        return ent[1]
    env = torch.empty(n_fft + hop * (n_frames - 1), device=window.device, dtype=torch.float32)
    call("lct_ola_envelope", window, env, n_fft, hop, n_frames)
    _ENV[key] = [window, env, capturing]
    if len(_ENV) > _ENV_MAX:
        for k in [k for k, e in _ENV.items() if not e[2]][:len(_ENV) - _ENV_MAX]:
            del _ENV[k]
    return env


def spec_view(phys: torch.Tensor) -> torch.Tensor:
    """[B, Tf, F] physical -> the reference's [B, F, Tf] view (same memory, like torch.stft)."""
    return phys.transpose(1, 2)


def spec_phys(spec: torch.Tensor) -> torch.Tensor:
    """[B, F, Tf] (any strides) -> contiguous [B, Tf, F]."""
    p = spec.transpose(1, 2)
    return p if p.is_contiguous() else p.contiguous()


def _check_wave(x: torch.Tensor) -> torch.Tensor:
    if x.dtype != torch.float32:
        raise RuntimeError(f"expected float32 waveform, got {x.dtype}")
    return x if x.is_contiguous() else x.contiguous()


# ----------------------------------------------------------------------------- STFT family
def stft_fwd(x, window, n_fft, hop, want_mag=False, eps=1e-12):
    x = _check_wave(x)
    B, T = x.shape
    Tf, F = 1 + T // hop, n_fft // 2 + 1
    spec = torch.empty(B, Tf, F, dtype=torch.complex64, device=x.device)
    mag = torch.empty(B, Tf, F, dtype=torch.float32, device=x.device) if want_mag else None
    call("lct_stft_fwd", x, window, twiddles(n_fft, x), spec, mag, B, T, n_fft, hop, eps)
    return spec, mag


def stft_bwd(gspec, window, T, n_fft, hop):
    B, Tf, F = gspec.shape
    work = torch.empty(B, n_fft + hop * (Tf - 1), dtype=torch.float32, device=gspec.device)
    gx = torch.empty(B, T, dtype=torch.float32, device=gspec.device)
    call("lct_stft_bwd", gspec, window, twiddles(n_fft, gspec), work, gx, B, T, n_fft, hop)
    return gx


def istft_fwd(spec, window, n_fft, hop, length, mask_c=None, c=0.3, eps=1e-12):
    B, Tf, F = spec.shape
    env = ola_envelope(window, n_fft, hop, Tf)
    y = torch.empty(B, length, dtype=torch.float32, device=spec.device)
    call("lct_istft_fwd", spec, mask_c, window, twiddles(n_fft, spec), env, y, B, Tf, n_fft, hop, length, c, eps)
    return y


def istft_bwd(gy, window, n_fft, hop, Tf, xspec=None, mask_c=None, want_gspec=True, c=0.3, eps=1e-12):
    gy = _check_wave(gy)
    B, length = gy.shape
    F = n_fft // 2 + 1
    env = ola_envelope(window, n_fft, hop, Tf)
    gspec = torch.empty(B, Tf, F, dtype=torch.complex64, device=gy.device) if want_gspec else None
    gmask = torch.empty(B, Tf, F, dtype=torch.float32, device=gy.device) if mask_c is not None else None
    call("lct_istft_bwd", gy, window, twiddles(n_fft, gy), env, gspec, xspec, mask_c, gmask, B, Tf, n_fft, hop,
         length, c, eps)
    return gspec, gmask


def tf_features_fwd(noisy, clean, window, n_fft, hop, c=0.3, gamma=1e-12, eps=1e-12, want_specs=False):
    noisy, clean = _check_wave(noisy), _check_wave(clean)
    B, T = noisy.shape
    Tf, F = 1 + T // hop, n_fft // 2 + 1
    mk = lambda dt: torch.empty(B, Tf, F, dtype=dt, device=noisy.device)
    nmag, irm, nmagc = mk(torch.float32), mk(torch.float32), mk(torch.float32)
    ns = mk(torch.complex64) if want_specs else None
    cs = mk(torch.complex64) if want_specs else None
    call("lct_tf_features_fwd", noisy, clean, window, twiddles(n_fft, noisy), nmag, irm, nmagc, ns, cs, B, T, n_fft,
         hop, c, gamma, eps)
    return nmag, irm, nmagc, ns, cs


def mrstft_sums(y_hat, y, window, n_fft, hop, acc, eps=1e-12):
    B, T = y_hat.shape
    call("lct_mrstft_sums", y_hat, y, window, twiddles(n_fft, y_hat), acc, B, T, n_fft, hop, eps)


def mrstft_grad_spec(spec_hat, spec_ref, k_mag, k_cplx, upstream=None, eps=1e-12):
    g = torch.empty_like(spec_hat)
    call("lct_mrstft_grad_spec", spec_hat, spec_ref, g, spec_hat.numel(), eps, k_mag, k_cplx, upstream)
    return g


# ----------------------------------------------------------------------------- spectral elementwise
def magnitude_fwd(spec, power=1.0, eps=1e-12):
    mag = torch.empty(spec.shape, dtype=torch.float32, device=spec.device)
    call("lct_magnitude_fwd", spec, mag, spec.numel(), power, eps)
    return mag


def magnitude_bwd(spec, gmag, power=1.0, eps=1e-12):
    g = torch.empty_like(spec)
    call("lct_magnitude_bwd", spec, gmag, g, spec.numel(), power, eps)
    return g


def powclamp_fwd(x, e, eps=1e-12):
    y = torch.empty_like(x)
    call("lct_powclamp_fwd", x, y, x.numel(), e, eps)
    return y


def powclamp_bwd(x, gy, e, eps=1e-12):
    gx = torch.empty_like(x)
    call("lct_powclamp_bwd", x, gy, gx, x.numel(), e, eps)
    return gx


def irm_fwd(clean_spec, noisy_spec, c=0.3, gamma=1e-12, eps=1e-12):
    out = torch.empty(clean_spec.shape, dtype=torch.float32, device=clean_spec.device)
    call("lct_irm_fwd", clean_spec, noisy_spec, out, clean_spec.numel(), c, gamma, eps)
    return out


def apply_mask_fwd(spec, mask, compressed, c=0.3, eps=1e-12):
    out = torch.empty_like(spec)
    call("lct_apply_mask_fwd", spec, mask, out, spec.numel(), int(compressed), c, eps)
    return out


def apply_mask_bwd(spec, mask, gout, compressed, want_gmask=True, want_gspec=False, c=0.3, eps=1e-12):
    gmask = torch.empty_like(mask) if want_gmask else None
    gspec = torch.empty_like(spec) if want_gspec else None
    call("lct_apply_mask_bwd", spec, mask, gout, gmask, gspec, spec.numel(), int(compressed), c, eps)
    return gmask, gspec


# ----------------------------------------------------------------------------- discriminator pieces
def reflect_pad_right_fwd(x, pad):
    B, T = x.shape
    y = torch.empty(B, T + pad, dtype=torch.float32, device=x.device)
    call("lct_reflect_pad_right_fwd", x, y, B, T, pad)
    return y


def reflect_pad_right_bwd(gy, T, pad):
    B = gy.shape[0]
    gx = torch.empty(B, T, dtype=torch.float32, device=gy.device)
    call("lct_reflect_pad_right_bwd", gy, gx, B, T, pad)
    return gx


def avgpool4_fwd(x):
    B, L = x.shape
    y = torch.empty(B, L // 2 + 1, dtype=torch.float32, device=x.device)
    call("lct_avgpool4_fwd", x, y, B, L)
    return y


def avgpool4_bwd(gy, L):
    B = gy.shape[0]
    gx = torch.empty(B, L, dtype=torch.float32, device=gy.device)
    call("lct_avgpool4_bwd", gy, gx, B, L)
    return gx


def weight_norm_fwd(g, v):
    cout = v.shape[0]
    row = v.numel() // cout
    w = torch.empty_like(v)
    call("lct_weight_norm_fwd", g, v, w, None, cout, row)
    return w


def weight_norm_bwd(g, v, dw):
    cout = v.shape[0]
    row = v.numel() // cout
    dg = torch.empty_like(g)
    dv = torch.empty_like(v)
    call("lct_weight_norm_bwd", g, v, dw, dg, dv, cout, row)
    return dg, dv


_SUPPORTED: Dict[tuple, bool] = {}      # (entry point, layer geometry) -> the library's answer (pure functions of the shape)


def _supported(fn: str, *geom) -> bool:
    key = (fn,) + geom
    r = _SUPPORTED.get(key)
    if r is None:
        r = _SUPPORTED[key] = bool(call_ret(fn, *geom))
    return r


def _tc_ok(cin, cout, k, groups, stride, pad, P):
    return (config.dense_tensor_cores and config.grouped_conv_tcgen05 and pad == k // 2 and
            _supported("lct_conv_tc_supported", cin, cout, groups, k, stride, P))


def _mma_ok(cin, cout, k, groups, stride, pad, P):
    return (config.dense_tensor_cores and pad == k // 2 and
            _supported("lct_conv_mma_supported", cin, cout, groups, k, stride, P))


# Which of the two tensor-core implementations of the grouped / first-layer convolutions runs a pass.  Both compute the
# same TF32 (round-to-nearest operands, fp32 accumulate) result; the choice is the measured one per layer family at the
# D-step batch 2B = 16 (tools/bench_disc_layers.py, profiles/README.md round 2):
#   forward        tcgen05 (conv_tc.cu) everywhere except the 1 -> 16, k = 15, stride-1 first MSD layer
#                  (12.7 vs 15.3 us: eight half-empty K = 8 MMAs per tile for one input channel)
#   data gradient  tcgen05 for the scale discriminators (P = 1: 23-38 vs 28-61 us per layer) and the one-channel first
#                  layers; the TF32 mma.sync kernel (conv_mma.cu) for the period discriminators' grouped layers, whose
#                  128-row tiles carry too little data to hide the single-buffered tile pipeline's latencies
#                  (35-78 vs 25-49 us)
#   weight gradient  mma.sync: positions are the contraction dimension, which needs an MN-major window operand, and
#                  tcgen05.mma kind::tf32 returns zeros for MN-major no-swizzle operands (tools/umma_probe.cu)
def _use_tc(cin, cout, k, groups, stride, pad, P):
    return _tc_ok(cin, cout, k, groups, stride, pad, P) and not (cin // groups == 1 and stride == 1)


def _use_tc_dgrad(cin, cout, k, groups, stride, pad, P):
    return _tc_ok(cin, cout, k, groups, stride, pad, P) and (P == 1 or cin // groups == 1 or config.tc_dgrad_periods)


def _use_mma(cin, cout, k, groups, stride, pad, P):
    return _mma_ok(cin, cout, k, groups, stride, pad, P) and not _use_tc(cin, cout, k, groups, stride, pad, P)


def _use_mma_dgrad(cin, cout, k, groups, stride, pad, P):
    return _mma_ok(cin, cout, k, groups, stride, pad, P) and not _use_tc_dgrad(cin, cout, k, groups, stride, pad, P)


def conv_tc_images(ws, specs, P, want_f=True, want_d=True):
    """Weight images of the tcgen05 grouped convolutions, ONE launch per stack: ws[i] [Cout, Cin/G, K] normalised
    weights, specs[i] = (k, stride, pad, groups); want_f / want_d: bool or per-layer list (forward / data-gradient
    image wanted).  Returns (imgs_f, imgs_d) with None where no image was asked for or the kernels do not cover the
    layer (conv_post, the dense layer)."""
    n = len(ws)
    wf = list(want_f) if isinstance(want_f, (list, tuple)) else [bool(want_f)] * n
    wd = list(want_d) if isinstance(want_d, (list, tuple)) else [bool(want_d)] * n
    imgs_f, imgs_d = [None] * n, [None] * n
    idx, shapes = [], []
    buf = (ctypes.c_int64 * 2)()
    for i, (k, s, pad, g) in enumerate(specs):
        cout, cig = ws[i].shape[0], ws[i].shape[1]
        if not (wf[i] or wd[i]) or _is_post(cout, k, g, s, pad) or not _tc_ok(cig * g, cout, k, g, s, pad, P):
            continue
        call_ret("lct_conv_tc_image_len", cig * g, cout, g, k, s, P, buf)
        idx.append(i)
        shapes += [(int(buf[0]) if wf[i] else 0,), (int(buf[1]) if wd[i] else 0,)]
    if not idx:
        return imgs_f, imgs_d
    views = _flat_views(shapes, ws[0].device)
    maxn = 8
    for c0 in range(0, len(idx), maxn):
        part = idx[c0:c0 + maxn]
        geo = []
        for i in part:
            k, s, pad, g = specs[i]
            geo += [ws[i].shape[1] * g, ws[i].shape[0], g, k, s, P]
        f = [views[2 * (c0 + j)] if wf[i] else None for j, i in enumerate(part)]
        d = [views[2 * (c0 + j) + 1] if wd[i] else None for j, i in enumerate(part)]
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() if t is not None else None for t in ts])
        call("lct_conv_tc_images", _ptr_array([ws[i].contiguous() for i in part]), arr(f), arr(d),
             (ctypes.c_int64 * len(geo))(*geo), len(part))
        for j, i in enumerate(part):
            imgs_f[i], imgs_d[i] = f[j], d[j]
    return imgs_f, imgs_d


def conv_tc_fwd(x, img, bias, Cout, groups, K, stride, pad, act=ACT_NONE, slope=0.2, out=None):
    B, Cin, Lin, P = x.shape
    y = out if out is not None else torch.empty(B, Cout, conv_out_len(Lin, K, stride, pad), P, dtype=torch.float32,
                                                device=x.device)
    call("lct_conv_tc_fwd", x, img, bias, y, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
    return y


def conv_tc_dgrad(dy, img_d, x_shape, Cout, groups, K, stride, pad, gextra=None, xact=None, act=ACT_NONE, slope=0.2):
    B, Cin, Lin, P = x_shape
    dx = torch.empty(B, Cin, Lin, P, dtype=torch.float32, device=dy.device)
    call("lct_conv_tc_dgrad", dy, img_d, dx, gextra, xact, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
    return dx


def _is_post(cout, k, groups, stride, pad):
    """conv_post shape: one output channel, "same" odd kernel -> dedicated channel-reduction kernels."""
    return cout == 1 and groups == 1 and stride == 1 and (k & 1) and k <= 8 and pad == k // 2


def _flat_views(shapes, device, zero=False, return_flat=False):
    """One allocation (optionally zero-filled: one memset) carved into contiguous tensors of `shapes`."""
    sizes = [int(math.prod(s)) for s in shapes]
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) // 4 * 4            # keep every view 16-byte aligned
    flat = zeros((tot,), device) if zero else torch.empty(tot, dtype=torch.float32, device=device)
    views = [flat[o:o + n].view(s) for o, n, s in zip(offs, sizes, shapes)]
    return (views, flat) if return_flat else views


def mt_weight_norm_fwd(gs, vs, specs=None, P=1):
    """w_i = g_i * v_i / ||v_i|| for all layers of a stack in one launch.  With `specs` ((k, stride, pad, groups) per
    layer) also the staged TF32 weight images of the tensor-core conv kernels, per layer and direction for the kernel
    that runs it (see _use_tc / _use_tc_dgrad): the mma.sync images come out of the weight-norm launch itself, the
    tcgen05 images out of one more launch.  Returns (ws, imgs_f, imgs_d)."""
    n = len(vs)
    ws = _flat_views([tuple(v.shape) for v in vs], vs[0].device)
    rows = (ctypes.c_int64 * n)(*[v.shape[0] for v in vs])
    rowlen = (ctypes.c_int64 * n)(*[v.numel() // v.shape[0] for v in vs])
    imgs_f, imgs_d = [None] * n, [None] * n
    pf = pd = geo = None
    tc_f, tc_d = [False] * n, [False] * n
    if specs is not None and config.dense_tensor_cores:
        shapes, idxs, geos = [], [], [0] * (8 * n)
        buf = (ctypes.c_int64 * 2)()
        mf, md = [None] * n, [None] * n
        for i, (k, s, pad, g) in enumerate(specs):
            cout, cig = vs[i].shape[0], vs[i].shape[1]
            cin = cig * g
            if _is_post(cout, k, g, s, pad):
                continue
            tc_f[i], tc_d[i] = _use_tc(cin, cout, k, g, s, pad, P), _use_tc_dgrad(cin, cout, k, g, s, pad, P)
            if (tc_f[i] and tc_d[i]) or not _mma_ok(cin, cout, k, g, s, pad, P):
                continue
            call_ret("lct_conv_mma_image_geometry", cin, cout, g, k, s, 0, buf)
            kkf, nsf = int(buf[0]), int(buf[1])
            call_ret("lct_conv_mma_image_geometry", cin, cout, g, k, s, 1, buf)
            kkd, nsd = int(buf[0]), int(buf[1])
            geos[8 * i:8 * i + 8] = [cout // g, k, s, (k + s - 1) // s, kkf, nsf, kkd, nsd]
            shapes += [(g, kkf, nsf), (g, kkd, nsd)]
            idxs.append(i)
        if idxs:
            views = _flat_views(shapes, vs[0].device, zero=True)
            for j, i in enumerate(idxs):
                mf[i], md[i] = views[2 * j], views[2 * j + 1]
            arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() if t is not None else None for t in ts])
            pf, pd, geo = arr(mf), arr(md), (ctypes.c_int64 * (8 * n))(*geos)
        imgs_f, imgs_d = mf, md
    call("lct_mt_weight_norm_fwd", _ptr_array(gs), _ptr_array(vs), _ptr_array(ws), rows, rowlen, pf, pd, geo, n)
    if any(tc_f) or any(tc_d):
        tf, td = conv_tc_images(ws, specs, P, want_f=tc_f, want_d=tc_d)
        imgs_f = [tf[i] if tc_f[i] else imgs_f[i] for i in range(n)]
        imgs_d = [td[i] if tc_d[i] else imgs_d[i] for i in range(n)]
    return ws, imgs_f, imgs_d


def mt_weight_norm_bwd(gs, vs, dws, out=None):
    """out: optional (dgs, dvs) lists of pre-allocated outputs (views of the stack's gradient arena)."""
    if out is not None:
        dgs, dvs = out
    else:
        dgs = _flat_views([tuple(g.shape) for g in gs], gs[0].device)
        dvs = _flat_views([tuple(v.shape) for v in vs], vs[0].device)
    rows = (ctypes.c_int64 * len(vs))(*[v.shape[0] for v in vs])
    rowlen = (ctypes.c_int64 * len(vs))(*[v.numel() // v.shape[0] for v in vs])
    call("lct_mt_weight_norm_bwd", _ptr_array(gs), _ptr_array(vs), _ptr_array(dws), _ptr_array(dgs), _ptr_array(dvs),
         rows, rowlen, len(vs))
    return dgs, dvs


def conv_out_len(lin, k, s, pad):
    return (lin + 2 * pad - k) // s + 1


def conv1d_fwd(x, w, bias, groups, stride, pad, act=ACT_NONE, slope=0.2, wimg=None, out=None):
    """x [B,Cin,Lin,P], w [Cout,Cin/G,K] (any trailing singleton dims) -> [B,Cout,Lout,P] (written into `out` if given:
    a contiguous batch slice of a larger buffer)."""
    B, Cin, Lin, P = x.shape
    Cout, K = w.shape[0], w.shape[2]
    Lout = conv_out_len(Lin, K, stride, pad)
    if out is not None and (tuple(out.shape) != (B, Cout, Lout, P) or not out.is_contiguous()):
        raise ValueError(f"conv1d_fwd: out has shape {tuple(out.shape)}, expected contiguous {(B, Cout, Lout, P)}")
    if _is_post(Cout, K, groups, stride, pad) and act == ACT_NONE:
        y = out if out is not None else torch.empty(B, 1, Lin, P, dtype=torch.float32, device=x.device)
        call("lct_conv_post_fwd", x, w, bias, y, B, Cin, Lin, P, K)
        return y
    if _use_tc(Cin, Cout, K, groups, stride, pad, P):
        if wimg is None:
            wimg = conv_tc_images([w.reshape(Cout, Cin // groups, K)], [(K, stride, pad, groups)], P, want_d=False)[0][0]
        return conv_tc_fwd(x, wimg, bias, Cout, groups, K, stride, pad, act, slope, out=out)
    y = out if out is not None else torch.empty(B, Cout, Lout, P, dtype=torch.float32, device=x.device)
    if _use_mma(Cin, Cout, K, groups, stride, pad, P):
        call("lct_conv_mma_fwd", x, w, wimg, bias, y, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
        return y
    call("lct_conv1d_fwd", x, w, bias, y, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
    return y


def conv1d_dgrad(dy, w, x_shape, groups, stride, pad, gextra=None, xact=None, act=ACT_NONE, slope=0.2, wimg=None):
    B, Cin, Lin, P = x_shape
    Cout, K = w.shape[0], w.shape[2]
    dx = torch.empty(B, Cin, Lin, P, dtype=torch.float32, device=dy.device)
    if _is_post(Cout, K, groups, stride, pad):
        call("lct_conv_post_dgrad", dy, w, dx, gextra, xact, B, Cin, Lin, P, K, act, slope)
        return dx
    if _use_tc_dgrad(Cin, Cout, K, groups, stride, pad, P):
        if wimg is None:
            wimg = conv_tc_images([w.reshape(Cout, Cin // groups, K)], [(K, stride, pad, groups)], P, want_f=False)[1][0]
        call("lct_conv_tc_dgrad", dy, wimg, dx, gextra, xact, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
        return dx
    if _use_mma_dgrad(Cin, Cout, K, groups, stride, pad, P):
        call("lct_conv_mma_dgrad", dy, w, wimg, dx, gextra, xact, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
        return dx
    call("lct_conv1d_dgrad", dy, w, dx, gextra, xact, B, Cin, Cout, groups, K, stride, pad, Lin, P, act, slope)
    return dx


def conv1d_wgrad(x, dy, w_shape, groups, stride, pad, want_bias=True, dw=None, db=None):
    """dw/db may be passed in pre-zeroed (views of one flat buffer: one fill per stack instead of two per layer)."""
    B, Cin, Lin, P = x.shape
    Cout, K = w_shape[0], w_shape[2]
    if dw is None:
        dw = zeros(tuple(w_shape), x.device)
    if db is None and want_bias:
        db = zeros((Cout,), x.device)
    if _is_post(Cout, K, groups, stride, pad):
        call("lct_conv_post_wgrad", x, dy, dw, db, B, Cin, Lin, P, K)
        return dw, db
    if _mma_ok(Cin, Cout, K, groups, stride, pad, P):
        call("lct_conv_mma_wgrad", x, dy, dw, db, B, Cin, Cout, groups, K, stride, pad, Lin, P)
        return dw, db
    call("lct_conv1d_wgrad", x, dy, dw, db, B, Cin, Cout, groups, K, stride, pad, Lin, P)
    return dw, db


# ----------------------------------------------------------------------------- generator pieces
def gemm(A, B, C, M, N, K, lda, ldb, ldc, ta=False, tb=False, bias=None, act=ACT_NONE, slope=0.2, alpha=1.0,
         accumulate=False, res=None, ldr=0, out2=None, ldo=0, ksplit=1, nbatch=1, a_div=1, b_div=1, sA=0, sB=0, sC=0,
         sBias=0, sRes=0, sOut2=0, a_off=0, b_off=0, c_off=0):
    """Raw strided GEMM on flat fp32 buffers; *_off are element offsets into A/B/C."""
    a = A.view(-1)[a_off:] if a_off else A
    b = B.view(-1)[b_off:] if b_off else B
    c = C.view(-1)[c_off:] if c_off else C
    call("lct_gemm", _raw(a), _raw(b), _raw(c), bias, res, out2, M, N, K, lda, ldb, ldc, ldr, ldo, int(ta), int(tb),
         act, slope, alpha, int(accumulate), ksplit, nbatch, a_div, b_div, sA, sB, sC, sBias, sRes, sOut2)


def _raw(t):
    if t is None:
        return None
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("lctgan gemm operands must be CUDA float32")
    return ctypes.c_void_p(t.data_ptr())


def gemm_free_add(a, b, out, M, N, ldb):
    """out[M,N] = a[M,N] + b[:, :N] where b has row stride ldb."""
    call("lct_add2d", a, N, _raw(b), ldb, out, N, M, N)


def colsum(X, out, M, N, ld):
    call("lct_colsum", _raw(X), out, M, N, ld)


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    mean = torch.empty(M, dtype=torch.float32, device=x.device)
    rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    call("lct_layernorm_fwd", x, gamma, beta, y, mean, rstd, M, C, eps)
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dres=None):
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    call("lct_layernorm_bwd", dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, M, C)
    return dx


def act_bwd(y, dy, act=ACT_LRELU, slope=0.2):
    d = torch.empty_like(y)
    call("lct_act_bwd", y, dy, d, y.numel(), act, slope)
    return d


def gconv(x, w, bias, out_tf, cd, transposed, act=ACT_NONE, slope=0.2, gmul=None, gact=ACT_NONE, gslope=0.2):
    """x [B,Ti,Fi,Cs] -> [B,To,Fo,Cd] with (To,Fo) = out_tf."""
    B, Ti, Fi, Cs = x.shape
    To, Fo = out_tf
    out = torch.empty(B, To, Fo, cd, dtype=torch.float32, device=x.device)
    if config.gconv_tensor_cores and call_ret("lct_gconv_mma_supported", Cs, cd):
        # implicit row GEMM on the tensor cores (3xTF32: fp32-level accuracy); the weight image is a 6*Cs*Cd re-arrangement
        img = torch.empty(call_ret("lct_gconv_image_len", Cs, cd), dtype=torch.float32, device=x.device)
        call("lct_gconv_weight_image", w, img, int(transposed), Cs, cd)
        call("lct_gconv_mma", x, img, bias, out, gmul, int(transposed), B, Ti, Fi, Cs, To, Fo, cd, act, slope, gact,
             gslope)
        return out
    call("lct_gconv", x, w, bias, out, gmul, int(transposed), B, Ti, Fi, Cs, To, Fo, cd, act, slope, gact, gslope)
    return out


def gconv_wgrad(S, Lg, w_shape, out=None):
    """`out`: optional pre-zeroed accumulator of shape w_shape."""
    B, Ts, Fs, Ca = S.shape
    _, Tl, Fl, Cc = Lg.shape
    dW = out if out is not None else zeros(tuple(w_shape), S.device)
    call("lct_gconv_wgrad", S, Lg, dW, B, Ts, Fs, Ca, Tl, Fl, Cc)
    return dW


def skip_add_fwd(h, mag, w, bias):
    B, Th, Fh, C = h.shape
    _, Tm, Fm = mag.shape
    out = torch.empty(B, min(Th, Tm), min(Fh, Fm), C, dtype=torch.float32, device=h.device)
    call("lct_skip_add_fwd", h, mag, w, bias, out, B, Th, Fh, Tm, Fm, C)
    return out


def skip_add_bwd(g, mag, h_shape, dw, db, want_dh=True):
    B, Th, Fh, C = h_shape
    _, Tm, Fm = mag.shape
    same = (g.shape[1] == Th and g.shape[2] == Fh)
    dh = None
    if want_dh and not same:
        dh = torch.empty(h_shape, dtype=torch.float32, device=g.device)
    call("lct_skip_add_bwd", g, mag, dh, dw, db, B, Th, Fh, Tm, Fm, C)
    if want_dh and same:
        dh = g
    return dh


def final_mask_fwd(y, T, F, use_sigmoid=True):
    B, Ty, Fy = y.shape[0], y.shape[1], y.shape[2]
    mask = torch.empty(B, T, F, dtype=torch.float32, device=y.device)
    call("lct_final_mask_fwd", y, mask, B, Ty, Fy, T, F, int(use_sigmoid))
    return mask


def final_mask_bwd(y, mask, gmask, use_sigmoid=True, act=ACT_RELU, slope=0.0):
    B, Ty, Fy = y.shape[0], y.shape[1], y.shape[2]
    _, T, F = mask.shape
    dpre = torch.empty_like(y)
    call("lct_final_mask_bwd", y, mask, gmask, dpre, B, Ty, Fy, T, F, int(use_sigmoid), act, slope)
    return dpre


# ----------------------------------------------------------------------------- multi-tensor losses
OP_SQ_CONST, OP_SQ_DIFF, OP_ABS_DIFF, OP_RELU_AFFINE, OP_SUM = 0, 1, 2, 3, 4


def _flat_ok(t):
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("loss inputs must be CUDA float32 tensors (no CPU fallback)")
    return t if t.is_contiguous() else t.contiguous()


def _ptr_array(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def mt_reduce(a: Sequence[torch.Tensor], b: Optional[Sequence[torch.Tensor]], scale: Sequence[float], op: int,
              k0: float = 0.0, k1: float = 0.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[0] += sum_i scale_i * sum op(a_i, b_i).  Returns the 1-element accumulator."""
    a = [_flat_ok(t) for t in a]
    b = [_flat_ok(t) for t in b] if b is not None else None
    if out is None:
        out = zeros((1,), a[0].device)
    maxseg = call_ret("lct_mt_max_segments")
    for s in range(0, len(a), maxseg):
        aa = a[s:s + maxseg]
        n = (ctypes.c_int64 * len(aa))(*[t.numel() for t in aa])
        sc = (ctypes.c_float * len(aa))(*scale[s:s + maxseg])
        call("lct_mt_reduce", _ptr_array(aa), _ptr_array(b[s:s + maxseg]) if b is not None else None, n, sc, len(aa),
             op, k0, k1, out)
    return out


def mt_grad(a: Sequence[torch.Tensor], b: Optional[Sequence[torch.Tensor]], scale: Sequence[float], op: int,
            k0: float = 0.0, k1: float = 0.0, upstream: Optional[torch.Tensor] = None,
            out: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """g_i = upstream * scale_i * d op(a_i, b_i) / d a_i; `out`: pre-allocated g_i (e.g. slices of larger buffers)."""
    a = [_flat_ok(t) for t in a]
    b = [_flat_ok(t) for t in b] if b is not None else None
    g = [_flat_ok(t) for t in out] if out is not None else [torch.empty_like(t) for t in a]
    maxseg = call_ret("lct_mt_max_segments")
    for s in range(0, len(a), maxseg):
        aa = a[s:s + maxseg]
        n = (ctypes.c_int64 * len(aa))(*[t.numel() for t in aa])
        sc = (ctypes.c_float * len(aa))(*scale[s:s + maxseg])
        call("lct_mt_grad", _ptr_array(aa), _ptr_array(b[s:s + maxseg]) if b is not None else None,
             _ptr_array(g[s:s + maxseg]), n, sc, len(aa), op, k0, k1, upstream)
    return g


def mt_copy(srcs: Sequence[torch.Tensor], dsts: Sequence[torch.Tensor]) -> None:
    """dst_i <- src_i (flat), one launch per 64 tensors."""
    maxseg = call_ret("lct_mt_max_segments")
    for s in range(0, len(srcs), maxseg):
        ss, dd = srcs[s:s + maxseg], dsts[s:s + maxseg]
        n = (ctypes.c_int64 * len(ss))(*[t.numel() for t in ss])
        call("lct_mt_copy", _ptr_array(ss), _ptr_array(dd), n, len(ss))


# ----------------------------------------------------------------------------- dense tcgen05 convolution
def dense_supported(cin: int, cout: int, k: int) -> bool:
    return bool(call_ret("lct_dense_supported", cin, cout, k))


def stage_nlc_bf16(x, pad):
    """[B,C,L,1] / [B,C,L] fp32 -> [B, L+2*pad, C] bf16 (zero rows around every batch)."""
    B, C, L = x.shape[0], x.shape[1], x.shape[2]
    out = torch.empty(B, L + 2 * pad, C, dtype=torch.bfloat16, device=x.device)
    call("lct_stage_nlc_bf16", x, out, B, C, L, pad)
    return out


def stage_ncl_bf16(x, Lp, shift, rowsum=None, copies=1):
    """[B,C,L] fp32 -> [copies, C, pitch] bf16 with out[k][c][b*Lp + shift - k + l] = x[b,c,l] (copy k is copy 0
    advanced by k positions); pitch = B*Lp rounded up to 8."""
    B, C, L = x.shape[0], x.shape[1], x.shape[2]
    pitch = (B * Lp + 7) // 8 * 8
    out = torch.empty(copies, C, pitch, dtype=torch.bfloat16, device=x.device)
    call("lct_stage_ncl_bf16", x, out, rowsum, B, C, L, Lp, shift, pitch, copies)
    return out


def stage_dense_weights(w, want_wt=True, want_wd=True):
    Co, Ci, K = w.shape[0], w.shape[1], w.shape[2]
    wt = torch.empty(K, Co, Ci, dtype=torch.bfloat16, device=w.device) if want_wt else None
    wd = torch.empty(K, Ci, Co, dtype=torch.bfloat16, device=w.device) if want_wd else None
    call("lct_stage_dense_weights", w, wt, wd, Co, Ci, K)
    return wt, wd


def dense_conv(a_staged, w_staged, B, L, Ca, Cn, K, bias=None, gextra=None, xact=None, act=ACT_NONE, slope=0.2, out=None):
    if out is None:
        out = torch.empty(B, Cn, L, 1, dtype=torch.float32, device=a_staged.device)
    elif tuple(out.shape) != (B, Cn, L, 1) or not out.is_contiguous():
        raise ValueError(f"dense_conv: out has shape {tuple(out.shape)}, expected contiguous {(B, Cn, L, 1)}")
    call("lct_dense_conv", a_staged, w_staged, bias, gextra, xact, out, B, L, Ca, Cn, K, act, slope)
    return out


def dense_wgrad(dyq, xq, Co, Ci, K, w_shape, out=None):
    dw = out if out is not None else torch.empty(w_shape, dtype=torch.float32, device=dyq.device)
    call("lct_dense_wgrad", dyq, xq, dw, Co, Ci, K, dyq.shape[-1])
    return dw
