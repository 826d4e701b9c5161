"""One waveform sub-discriminator (a PeriodDiscriminator or a ScaleDiscriminator conv stack) as a
single autograd node: weight-norm -> grouped strided conv -> bias -> LeakyReLU per layer,
returning every (post-activation) feature map like the reference does
(models/discriminators.py:69-103, :199-224).

Backward, per layer from the top: the data-gradient kernel of layer i+1 adds the
feature-matching gradient that arrived for map i and multiplies by LeakyReLU'(map i) in its
epilogue, so the pre-activation gradient of every layer is produced exactly once and feeds both
the weight-gradient and the next data-gradient kernel.

Chain length matters at batch 8 (every kernel of a sub-discriminator depends on the previous one and
fills at most one wave of the GPU), so everything that does not depend on the activations is batched:
the weight norms of all layers are ONE launch in forward and ONE in backward, and all weight / bias
gradient accumulators of a stack are views of one zero-filled buffer (one fill).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import config, ops

LRELU_SLOPE = 0.2


def _use_dense(x_shape, w_shape, k, s, g) -> bool:
    """MSD convs.5-like layers (dense, stride 1, P == 1) go to the tcgen05 kernel in tensor-core mode."""
    return (config.dense_tensor_cores and g == 1 and s == 1 and x_shape[3] == 1 and w_shape[0] > 1
            and ops.dense_supported(w_shape[1], w_shape[0], k))


# (kernel, stride, padding, groups) per conv; the last entry is conv_post (no activation)
LayerSpec = Tuple[int, int, int, int]


def prepare_stack(specs: Sequence[LayerSpec], params: Sequence[torch.Tensor], P: int, need_dgrad: bool = True) -> dict:
    """Everything a stack's passes derive from its WEIGHTS alone: the normalised weights (one batched weight-norm
    launch), the staged TF32 images of the tensor-core conv kernels and the bf16 images of the dense tcgen05 layer.
    A pure function of the parameters, so two passes over different inputs under the same weights (the G step's real
    and fake passes, train.py:221-227) share one preparation - models.discriminators.run_discriminators does that."""
    n = len(specs)
    with torch.no_grad():
        gs = [params[3 * i + 1].contiguous() for i in range(n)]
        vs = [params[3 * i + 2].contiguous() for i in range(n)]
        weights, imgs_f, imgs_d = ops.mt_weight_norm_fwd(gs, vs, specs, P)
        wt, wd = {}, {}
        for i, (k, s, pad, g) in enumerate(specs):
            w = weights[i]
            if _use_dense((1, w.shape[1] * g, 1, P), w.shape, k, s, g) and pad == k // 2:
                wt[i], wd[i] = ops.stage_dense_weights(w, want_wt=True, want_wd=need_dgrad)
    return dict(weights=weights, imgs_f=imgs_f, imgs_d=imgs_d, wt=wt, wd=wd)


def stack_forward(x4: torch.Tensor, specs: Sequence[LayerSpec], params: Sequence[torch.Tensor], prep: dict,
                  fmaps: List[torch.Tensor] = None, lo: int = 0, hi: int = None) -> List[torch.Tensor]:
    """The stack's kernels for rows [lo, hi) of the batch x4 [B, 1, L, P]; every feature map is a buffer for the WHOLE
    batch (allocated here when `fmaps` is None) of which only those rows are written.  Batch slices of the [B, C, L, P]
    maps are contiguous, so a batch can be pushed through in parts - the D step's clean half while the generator is
    still producing the enhanced half (models.discriminators.begin_split_forward) - and differentiated in one piece."""
    n = len(specs)
    B = x4.shape[0]
    hi = B if hi is None else hi
    weights, imgs_f = prep["weights"], prep["imgs_f"]
    alloc = fmaps is None
    if alloc:
        fmaps = []
    h_full = x4
    for i, (k, s, pad, g) in enumerate(specs):
        bias, w = params[3 * i], weights[i]
        act = ops.ACT_NONE if i == n - 1 else ops.ACT_LRELU
        Lin, P = h_full.shape[2], h_full.shape[3]
        if alloc:
            fmaps.append(torch.empty(B, w.shape[0], ops.conv_out_len(Lin, k, s, pad), P, dtype=torch.float32,
                                     device=x4.device))
        h, out = h_full[lo:hi], fmaps[i][lo:hi]
        if i in prep["wt"]:
            ops.dense_conv(ops.stage_nlc_bf16(h, pad), prep["wt"][i], hi - lo, Lin, w.shape[1], w.shape[0], k, bias=bias,
                           act=act, slope=LRELU_SLOPE, out=out)
        else:
            ops.conv1d_fwd(h, w, bias, g, s, pad, act=act, slope=LRELU_SLOPE, wimg=imgs_f[i], out=out)
        h_full = fmaps[i]
    return fmaps


def _accumulate(olds: List[torch.Tensor], news: List[torch.Tensor]) -> None:
    """olds[j] += news[j].  Both lists are normally views of two gradient arenas with the same layout (the D step's and
    the G step's buffers of one stack): then it is ONE launch over the flat range instead of torch's multi-tensor add."""
    o0, n0 = olds[0], news[0]
    same = all(o.is_contiguous() and n.is_contiguous() and o.numel() == n.numel() and
               o.data_ptr() - o0.data_ptr() == n.data_ptr() - n0.data_ptr() >= 0 for o, n in zip(olds, news))
    if same:
        last = max(range(len(olds)), key=lambda j: olds[j].data_ptr())
        total = (olds[last].data_ptr() - o0.data_ptr()) // 4 + olds[last].numel()
        so, sn = o0.untyped_storage(), n0.untyped_storage()
        room_o = (so.data_ptr() + so.nbytes() - o0.data_ptr()) // 4
        room_n = (sn.data_ptr() + sn.nbytes() - n0.data_ptr()) // 4
        if total <= room_o and total <= room_n:
            fo = torch.as_strided(o0, (total,), (1,))
            fn_ = torch.as_strided(n0, (total,), (1,))
            ops.call("lct_axpby", fo, fn_, fo, total, 1.0, 1.0)     # (the 16-byte alignment gaps between views are zeros)
            return
    torch._foreach_add_(olds, news)


class ConvStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x4, specs: Sequence[LayerSpec], skip_param_grads: bool, prep, *params):
        """x4: [B, 1, L, P]; params: (bias, weight_g, weight_v) per layer; prep: prepare_stack(...) of the same
        parameters or None (prepared here); prep["partial"] = (feature-map buffers, rows already done) continues a
        forward that stack_forward began on the first rows of the batch.  Returns all feature maps."""
        if not x4.is_cuda:
            raise RuntimeError("lctgan discriminators are CUDA only (sm_100a); there is no CPU fallback")
        n = len(specs)
        assert len(params) == 3 * n
        x4 = x4.contiguous()
        if prep is None:
            prep = prepare_stack(specs, params, x4.shape[3], need_dgrad=False)
        weights, imgs_d = prep["weights"], prep["imgs_d"]
        done, lo = prep.get("partial") or (None, 0)
        fmaps = stack_forward(x4, specs, params, prep, done, lo, x4.shape[0])
        dense: List[int] = [i for i in range(n) if i in prep["wt"]]
        # feature maps that receive no gradient (all but the logits in the D step) must arrive as None in backward, not
        # as materialised zero tensors: autograd's default filled a map-sized zero buffer per unused output (64 fill
        # kernels per step in the round-2 launch list) which the data-gradient kernels then read as "FM gradient"
        ctx.set_materialize_grads(False)
        ctx.specs = list(specs)
        ctx.dense = dense
        ctx.skip_param_grads = skip_param_grads
        ctx.param_objs = params   # the Parameter objects themselves (deferred mode accumulates into their .grad)
        ctx.imgs_d = imgs_d       # (internal buffers, not outputs: safe to keep on ctx)
        ctx.dense_wd = prep["wd"]
        ctx.save_for_backward(x4, *fmaps, *weights, *params)
        return tuple(fmaps)

    @staticmethod
    def backward(ctx, *gouts):
        specs = ctx.specs
        n = len(specs)
        saved = ctx.saved_tensors
        x4 = saved[0]
        fmaps = saved[1:1 + n]
        weights = saved[1 + n:1 + 2 * n]
        params = saved[1 + 2 * n:]
        need_x = ctx.needs_input_grad[0]
        need_p = [ctx.needs_input_grad[4 + j] for j in range(3 * n)]
        want_params = any(need_p) and not (ctx.skip_param_grads and need_x)
        gparams: List = [None] * (3 * n)
        gouts = [g.contiguous() if g is not None else None for g in gouts]
        dws = dbs = dgs_o = dvs_o = arena = None
        if want_params:
            # the weight-gradient accumulators (internal: gradients w.r.t. the NORMALISED weights) in one cleared buffer,
            # and the stack's final parameter gradients (bias, weight_g, weight_v of every layer) in another one - the
            # "arena" a data-parallel exchange all-reduces in place (lctgan.parallel)
            dws = ops._flat_views([tuple(w.shape) for w in weights], x4.device, zero=True)
            fin, arena = ops._flat_views([(w.shape[0],) for w in weights] + [tuple(params[3 * i + 1].shape) for i in range(n)] +
                                         [tuple(params[3 * i + 2].shape) for i in range(n)], x4.device, zero=True,
                                         return_flat=True)
            dbs, dgs_o, dvs_o = fin[:n], fin[n:2 * n], fin[2 * n:]

        dpre = gouts[n - 1]      # conv_post has no activation
        gx = None
        # weight gradients run on a helper stream beside the data-gradient chain (joined before returning)
        cur = torch.cuda.current_stream(x4.device)
        aux = config.aux_stream_for(cur) if (want_params and config.concurrent_discriminators) else None
        keep = []

        def on_aux(t):
            keep.append(t)
            if aux is None:
                return torch.cuda.stream(cur)
            aux.wait_stream(cur)
            return torch.cuda.stream(aux)

        # G step: nobody downstream reads the parameter gradients (config.defer_dead_param_grads): collect the work and
        # hand it to config.defer() - it runs later, beside the generator's backward
        late = bool(want_params and need_x and config.defer_dead_param_grads and x4.is_cuda)
        jobs = []

        def param_work(t, fn):
            if late:
                jobs.append(fn)
            else:
                with on_aux(t):
                    fn()

        for i in range(n - 1, -1, -1):
            k, s, pad, g = specs[i]
            inp = x4 if i == 0 else fmaps[i - 1]
            if dpre is not None and i in ctx.dense:
                # tcgen05 path: bf16 staged operands, fp32 accumulation
                B_, Ci_, L_ = inp.shape[0], inp.shape[1], inp.shape[2]
                Co_ = weights[i].shape[0]
                if want_params:
                    def dense_w(dpre=dpre, inp=inp, i=i, k=k, pad=pad, L_=L_, Co_=Co_, Ci_=Ci_):
                        Lp = L_ + k - 1
                        dyq = ops.stage_ncl_bf16(dpre, Lp, 0, rowsum=dbs[i])
                        xq = ops.stage_ncl_bf16(inp, Lp, pad, copies=k)
                        ops.dense_wgrad(dyq, xq, Co_, Ci_, k, weights[i].shape, out=dws[i])
                    param_work(dpre, dense_w)
                wd = ctx.dense_wd.get(i)
                if wd is None:
                    _, wd = ops.stage_dense_weights(weights[i], want_wt=False, want_wd=True)
                dpre = ops.dense_conv(ops.stage_nlc_bf16(dpre, pad), wd, B_, L_, Co_, Ci_, k, gextra=gouts[i - 1],
                                      xact=inp, act=ops.ACT_LRELU, slope=LRELU_SLOPE)
            elif dpre is not None:
                if want_params:
                    def conv_w(dpre=dpre, inp=inp, i=i, g=g, s=s, pad=pad):
                        ops.conv1d_wgrad(inp, dpre, weights[i].shape, g, s, pad, want_bias=True, dw=dws[i], db=dbs[i])
                    param_work(dpre, conv_w)
                if i > 0:
                    dpre = ops.conv1d_dgrad(dpre, weights[i], inp.shape, g, s, pad, gextra=gouts[i - 1], xact=inp,
                                            act=ops.ACT_LRELU, slope=LRELU_SLOPE, wimg=ctx.imgs_d[i])
                elif need_x:
                    gx = ops.conv1d_dgrad(dpre, weights[i], inp.shape, g, s, pad, wimg=ctx.imgs_d[i])
            elif i > 0 and gouts[i - 1] is not None:
                dpre = ops.act_bwd(inp, gouts[i - 1], ops.ACT_LRELU, LRELU_SLOPE)
        if want_params:
            gs = [params[3 * i + 1].contiguous() for i in range(n)]
            vs = [params[3 * i + 2].contiguous() for i in range(n)]
            if late:
                param_objs = ctx.param_objs

                def finish():
                    # weight gradients, weight-norm backward and the accumulation into .grad; autograd got None
                    for fn in jobs:
                        fn()
                    dgs, dvs = ops.mt_weight_norm_bwd(gs, vs, dws, out=(dgs_o, dvs_o))
                    olds, news = [], []
                    for i in range(n):
                        for j, t in ((3 * i, dbs[i]), (3 * i + 1, dgs[i]), (3 * i + 2, dvs[i])):
                            if not need_p[j]:
                                continue
                            pobj = param_objs[j]
                            if pobj.grad is None:
                                pobj.grad = t
                            else:
                                olds.append(pobj.grad)
                                news.append(t)
                    if olds:
                        _accumulate(olds, news)

                config.defer(finish)
                return (gx, None, None, None, *gparams)
            if aux is not None:
                cur.wait_stream(aux)
            keep.clear()
            dgs, dvs = ops.mt_weight_norm_bwd(gs, vs, dws, out=(dgs_o, dvs_o))
            if config.stack_grad_hook is not None:      # data parallel: this stack's gradients are complete - start
                config.stack_grad_hook(arena)           # their all-reduce while the other stacks are still in backward
            for i in range(n):
                gparams[3 * i] = dbs[i] if need_p[3 * i] else None
                gparams[3 * i + 1] = dgs[i] if need_p[3 * i + 1] else None
                gparams[3 * i + 2] = dvs[i] if need_p[3 * i + 2] else None
        return (gx, None, None, None, *gparams)


def conv_stack(x4: torch.Tensor, specs: Sequence[LayerSpec], params: Sequence[torch.Tensor],
               skip_param_grads: bool = False, prep: dict = None) -> List[torch.Tensor]:
    return list(ConvStackFn.apply(x4, tuple(specs), skip_param_grads, prep, *params))
