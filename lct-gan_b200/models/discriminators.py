"""Multi-period and multi-scale waveform discriminators with the reference's API
(jqshang/LCT-GAN models/discriminators.py), executing on the lctgan sm_100a kernels.

Parameters live in ``weight_norm(nn.Conv2d/Conv1d)`` containers built in the reference's order, so
``state_dict`` keys (``discriminators.i.convs.j.{bias,weight_g,weight_v}``, ``conv_post.*``), shapes,
iteration order and seeded initial values are identical; the containers' forward (and the
weight-norm pre-hook) is never invoked - each sub-discriminator is one autograd node
(``lctgan.disc_impl.ConvStackFn``) that returns every feature map like the reference.

Reference anchors: PeriodDiscriminator :9-103, MultiPeriodDiscriminator :106-147,
ScaleDiscriminator :150-224, MultiScaleDiscriminator :227-286.
"""
from typing import List, Tuple

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm, weight_norm

from lctgan import config as _cfg
from lctgan import functional as LF
from lctgan import ops
from lctgan.disc_impl import conv_stack, prepare_stack, stack_forward


def _run_concurrently(discs, inputs, no_grad=None, first_stream=0, post=None):
    """Evaluate discs[i](inputs[i]) for all i, each on its own CUDA stream (fork/join around the caller's stream).
    no_grad[i] evaluates that call under torch.no_grad().  A sub-discriminator that appears more than once (the G step
    runs every one on the clean and on the enhanced batch) prepares its weights - weight norm, staged conv images - ONCE,
    on the stream of its first call; the later calls wait for that event and reuse the buffers."""
    x0 = inputs[0]
    flags = no_grad if no_grad is not None else [False] * len(discs)
    shared = {id(d) for d in discs if sum(1 for e in discs if e is d) > 1 and x0.is_cuda}
    preps = {}

    def call(d, x, ng, stream=None):
        if id(d) in shared:
            if id(d) not in preps:
                prep = d.prepare(need_dgrad=not all(f for e, f in zip(discs, flags) if e is d))
                ev = torch.cuda.Event() if stream is not None else None
                if ev is not None:
                    ev.record(stream)
                preps[id(d)] = (prep, ev, stream)
            else:
                prep, ev, s0 = preps[id(d)]
                if ev is not None and stream is not None and stream != s0:
                    stream.wait_event(ev)
            d._prep = preps[id(d)][0]
        try:
            if ng:
                with torch.no_grad():
                    return d(x)
            return d(x)
        finally:
            d._prep = None

    if not (_cfg.concurrent_discriminators and x0.is_cuda and len(discs) > 1):
        out = []
        for i, (d, x, ng) in enumerate(zip(discs, inputs, flags)):
            out.append(call(d, x, ng))
            if post is not None:
                post(i, out, None, None)
        return out
    cur = torch.cuda.current_stream(x0.device)
    streams = _cfg.side_streams(first_stream + len(discs), x0.device)[first_stream:]
    out = [None] * len(discs)
    done = [None] * len(discs)
    for i, (d, x, s, ng) in enumerate(zip(discs, inputs, streams, flags)):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            out[i] = call(d, x, ng, s)
            if post is not None:
                # post(i, outputs so far, this call's stream, done events of the earlier calls): work that follows call i
                # on ITS stream, before the join (e.g. the feature-matching term of one sub-discriminator)
                done[i] = torch.cuda.Event()
                done[i].record(s)
                post(i, out, s, done)
    for s in streams:
        cur.wait_stream(s)
    return out


def _msd_inputs(msd, x):
    if x.dim() == 2:
        x = x.unsqueeze(1)
    B, C, T = x.shape
    x_scale = x.reshape(B, T) if C == 1 else x
    inputs = []
    for i in range(len(msd.discriminators)):
        inputs.append(x_scale if x_scale.dim() == 3 else x_scale.unsqueeze(1))
        if i + 1 < len(msd.discriminators):   # the reference pools once more and discards the result
            x_scale = LF.AvgPool4Fn.apply(x_scale)
    return inputs


def run_discriminators(mpd, msd, waves, no_grad=None, first_stream=0, post=None):
    """Scheduling helper (not part of the reference API): evaluate `mpd(w)` and `msd(w)` for every waveform w in
    `waves` with ALL 8 * len(waves) sub-discriminators forked at once instead of one module call after the other.
    Returns [(mpd_logits, mpd_fmaps, msd_logits, msd_fmaps) per waveform]; values are identical to the module calls."""
    flags = no_grad if no_grad is not None else [False] * len(waves)
    discs, inputs, ng = [], [], []
    for w, f in zip(waves, flags):
        discs += list(mpd.discriminators)
        inputs += [w] * len(mpd.discriminators)
        ng += [f] * len(mpd.discriminators)
        if f:
            with torch.no_grad():
                mi = _msd_inputs(msd, w)
        else:
            mi = _msd_inputs(msd, w)
        discs += list(msd.discriminators)
        inputs += mi
        ng += [f] * len(msd.discriminators)
    res = _run_concurrently(discs, inputs, ng, first_stream, post=post)
    out, per = [], len(mpd.discriminators) + len(msd.discriminators)
    for k in range(len(waves)):
        r = res[k * per:(k + 1) * per]
        np_ = len(mpd.discriminators)
        out.append(([t[0] for t in r[:np_]], [t[1] for t in r[:np_]], [t[0] for t in r[np_:]], [t[1] for t in r[np_:]]))
    return out

def begin_split_forward(mpd, msd, first_half, first_stream=0, after=None):
    """Scheduling helper (not part of the reference API) for a pass over a batch whose second half does not exist yet:
    the D step evaluates every sub-discriminator on [clean; enhanced] (train.py:188-193) and only `clean` is known while
    the generator runs.  On its own side stream (forked from the caller's stream HERE) every sub-discriminator prepares
    its weights and pushes `first_half` through its kernels into feature-map buffers sized for the whole batch.
    Returns the state finish_split_forward() continues from; nothing is joined yet.
    `after` (a CUDA event recorded earlier on the caller's stream): fork from THAT point instead of the stream's current
    one - the caller may then enqueue the generator first (a captured graph launches ready branches in the order they
    were created: enqueued second, the generator's first kernel was measured waiting 0.49 ms behind this pass although
    nothing but creation order made it wait) and still have this pass depend on nothing the generator does."""
    discs = list(mpd.discriminators) + list(msd.discriminators)
    cur = torch.cuda.current_stream(first_half.device)
    streams = _cfg.side_streams(first_stream + len(discs), first_half.device)[first_stream:]
    fork = (lambda s: s.wait_event(after)) if after is not None else (lambda s: s.wait_stream(cur))
    if after is not None:
        # the pooled inputs of the scale discriminators are made on the first scale stream, not on the caller's
        s0 = streams[len(mpd.discriminators)]
        fork(s0)
        with torch.cuda.stream(s0), torch.no_grad():
            pooled = _msd_inputs(msd, first_half)
            pooled_ev = torch.cuda.Event()
            pooled_ev.record(s0)
    else:
        pooled, pooled_ev = _msd_inputs(msd, first_half), None
    inputs = [first_half] * len(mpd.discriminators) + pooled
    state = []
    for k, (d, x, s) in enumerate(zip(discs, inputs, streams)):
        fork(s)
        if pooled_ev is not None and k > len(mpd.discriminators):
            s.wait_event(pooled_ev)
        with torch.cuda.stream(s), torch.no_grad():
            prep = d.prepare(need_dgrad=False)
            xa = d._input4(x).contiguous()
            nb = xa.shape[0]
            x4 = torch.empty(2 * nb, *xa.shape[1:], dtype=xa.dtype, device=xa.device)
            ops.mt_copy([xa], [x4[:nb]])
            fmaps = stack_forward(x4, d._specs, _stack_params(d), prep, None, 0, nb)
        # (x is kept: the pooled scale inputs were allocated on the caller's stream, which is NOT joined with s before
        # finish_split_forward - released here, their memory could be handed out again while s still reads it)
        state.append((prep, x4, fmaps, nb, x))
    return state


def finish_split_forward(mpd, msd, second_half, state, first_stream=0):
    """Second half of begin_split_forward's batch (no gradient flows into the waveform: the D step's detached enhanced
    batch).  Returns (period logits, period feature maps, scale logits, scale feature maps) of the WHOLE batch, each
    sub-discriminator one autograd node over it, exactly like run_discriminators(mpd, msd, [cat(first, second)])."""
    discs = list(mpd.discriminators) + list(msd.discriminators)
    second_half = second_half.detach()
    inputs = [second_half] * len(mpd.discriminators) + _msd_inputs(msd, second_half)
    cur = torch.cuda.current_stream(second_half.device)
    streams = _cfg.side_streams(first_stream + len(discs), second_half.device)[first_stream:]
    res = []
    for d, x, s, (prep, x4, fmaps, nb, _) in zip(discs, inputs, streams, state):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            with torch.no_grad():
                xb = d._input4(x).contiguous()
                ops.mt_copy([xb], [x4[nb:]])
            d._prep = dict(prep, partial=(fmaps, nb))
            try:
                res.append(d._finish(*d._run(x4)))
            finally:
                d._prep = None
    for s in streams:
        cur.wait_stream(s)
    np_ = len(mpd.discriminators)
    return ([t[0] for t in res[:np_]], [t[1] for t in res[:np_]], [t[0] for t in res[np_:]], [t[1] for t in res[np_:]])


# (out_channels, kernel, stride, groups)
_PERIOD_LAYERS = ((32, 5, 3, 1), (128, 5, 3, 4), (512, 5, 3, 16), (1024, 5, 3, 64), (1024, 5, 1, 64))
_SCALE_LAYERS = ((16, 15, 1, 1), (64, 41, 4, 4), (256, 41, 4, 16), (1024, 41, 4, 64), (1024, 41, 4, 256),
                 (1024, 5, 1, 1))


def _stack_params(mod: nn.Module) -> List[torch.Tensor]:
    """(bias, weight_g, weight_v) of every conv then conv_post - the parameter order of the reference."""
    out: List[torch.Tensor] = []
    for conv in list(mod.convs) + [mod.conv_post]:
        out += [conv.bias, conv.weight_g, conv.weight_v]
    return out


class _SubDiscriminator(nn.Module):
    #: when True, the (discarded) discriminator weight gradients of the generator step are not computed
    #: if the input waveform itself requires grad.  Default False = the reference's exact behaviour.
    skip_param_grads_when_input_requires_grad = False
    #: weight preparation shared between several passes of one run_discriminators call (set and cleared there)
    _prep = None

    def prepare(self, need_dgrad: bool = True) -> dict:
        return prepare_stack(self._specs, _stack_params(self), getattr(self, "period", 1), need_dgrad=need_dgrad)

    def _finish(self, logits, fmaps):
        return logits, fmaps

    def _run(self, x4: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        if self._spectral:
            raise RuntimeError("use_spectral_norm=True is not supported by the lctgan kernels "
                               "(the reference's train.py never enables it)")
        fmaps = conv_stack(x4, self._specs, _stack_params(self), self.skip_param_grads_when_input_requires_grad,
                           prep=self._prep)
        return fmaps[-1], fmaps


class PeriodDiscriminator(_SubDiscriminator):
    """[B, T] or [B, 1, T] -> (logits [B, 1, H, P], [6 feature maps]); the waveform is right-reflect-padded
    to a multiple of the period and folded to [B, 1, T/P, P]."""

    def __init__(self, period: int, use_spectral_norm: bool = False):
        super().__init__()
        self.period = period
        self._spectral = bool(use_spectral_norm)
        norm_f = spectral_norm if use_spectral_norm else weight_norm
        convs, in_ch, specs = [], 1, []
        for out_ch, k, s, g in _PERIOD_LAYERS:
            convs.append(norm_f(nn.Conv2d(in_ch, out_ch, kernel_size=(k, 1), stride=(s, 1), padding=(k // 2, 0),
                                          groups=g)))
            specs.append((k, s, k // 2, g))
            in_ch = out_ch
        self.convs = nn.ModuleList(convs)
        self.conv_post = norm_f(nn.Conv2d(in_ch, 1, kernel_size=(3, 1), stride=(1, 1), padding=(1, 0)))
        specs.append((3, 1, 1, 1))
        self._specs = tuple(specs)
        self.activation = nn.LeakyReLU(0.2, inplace=True)

    def _input4(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            x = x.unsqueeze(1)
        B, C, T = x.shape
        assert C == 1, "PeriodDiscriminator expects shape [B, 1, T] or [B, T]."
        x2 = x.reshape(B, T)
        if T % self.period != 0:
            pad = self.period - (T % self.period)
            x2 = LF.ReflectPadRightFn.apply(x2, pad)
            T = T + pad
        return x2.reshape(B, 1, T // self.period, self.period)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        return self._run(self._input4(x))


class MultiPeriodDiscriminator(nn.Module):
    def __init__(self, periods: List[int] = (2, 3, 5, 7, 11), use_spectral_norm: bool = False):
        super().__init__()
        self.discriminators = nn.ModuleList(
            [PeriodDiscriminator(p, use_spectral_norm=use_spectral_norm) for p in periods])

    def forward(self, x: torch.Tensor) -> Tuple[List[torch.Tensor], List[List[torch.Tensor]]]:
        res = _run_concurrently(list(self.discriminators), [x] * len(self.discriminators))
        return [r[0] for r in res], [r[1] for r in res]


class ScaleDiscriminator(_SubDiscriminator):
    """[B, T] or [B, 1, T] -> (logits [B, 1, L], [7 feature maps])."""

    def __init__(self, use_spectral_norm: bool = False):
        super().__init__()
        self._spectral = bool(use_spectral_norm)
        norm_f = spectral_norm if use_spectral_norm else weight_norm
        convs, in_ch, specs = [], 1, []
        for out_ch, k, s, g in _SCALE_LAYERS:
            convs.append(norm_f(nn.Conv1d(in_ch, out_ch, kernel_size=k, stride=s, padding=k // 2, groups=g)))
            specs.append((k, s, k // 2, g))
            in_ch = out_ch
        self.convs = nn.ModuleList(convs)
        self.conv_post = norm_f(nn.Conv1d(in_ch, 1, kernel_size=3, stride=1, padding=1))
        specs.append((3, 1, 1, 1))
        self._specs = tuple(specs)
        self.activation = nn.LeakyReLU(0.2, inplace=True)

    def _input4(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 2:
            x = x.unsqueeze(1)
        B, C, T = x.shape
        assert C == 1, "ScaleDiscriminator expects shape [B, 1, T] or [B, T]."
        return x.reshape(B, 1, T, 1)

    def _finish(self, logits, fmaps):
        # the kernels carry a trailing period axis of 1; hand back the reference's [B, C, L] shapes
        fmaps = [f.squeeze(-1) for f in fmaps]
        return fmaps[-1], fmaps

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        return self._finish(*self._run(self._input4(x)))


class MultiScaleDiscriminator(nn.Module):
    def __init__(self, num_scales: int = 3, use_spectral_norm: bool = False):
        super().__init__()
        assert num_scales >= 1, "num_scales must be >= 1"
        self.num_scales = num_scales
        self.discriminators = nn.ModuleList(
            [ScaleDiscriminator(use_spectral_norm=(use_spectral_norm and i == 0)) for i in range(num_scales)])
        self.avg_pool = nn.AvgPool1d(kernel_size=4, stride=2, padding=2, count_include_pad=False)

    def forward(self, x: torch.Tensor) -> Tuple[List[torch.Tensor], List[List[torch.Tensor]]]:
        res = _run_concurrently(list(self.discriminators), _msd_inputs(self, x))
        return [r[0] for r in res], [r[1] for r in res]
