"""The adversarial training step, i.e. the body of the reference's ``train_one_epoch`` loop
(jqshang/LCT-GAN train.py:165-249), written against the drop-in modules.  The reference's own
``train.py`` drives the same modules unchanged when ``lct-gan_b200`` is first on ``sys.path``; this
module exists so that the step can be benchmarked, graph-captured and data-parallelised without
the reference checkout (which does not travel to the GPU box).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

import losses as L
from models.discriminators import begin_split_forward, finish_split_forward, run_discriminators
from . import config, ops
from . import functional as LF
from .optim import clip_grad_norm_


@dataclass
class StepArgs:
    """The subset of train.py's argparse namespace the loop body reads (train.py:472-484)."""
    gan_loss: str = "ls"
    lambda_fm: float = 1.0
    lambda_mask: float = 1.0
    lambda_adv: float = 1e-2
    grad_clip: float = 5.0
    #: SURVEY.md section 8f N1.  train.py runs the enhancer twice per step on the same input with the same weights
    #: (no_grad at :180-181 for the D step, with grad at :208 for the G step; only D weights change in between).
    #: With this switch one forward (with grad) serves both - bit-identical values, one forward less - and it runs
    #: on a side stream concurrently with the discriminators' forward on the clean batch (no data dependence).
    reuse_enhancer_forward: bool = False
    #: D step: push clean and enhanced through the discriminators as ONE batch of 2B (no BatchNorm / cross-sample op in
    #: the networks, SURVEY 8e, so logits and parameter gradients are identical to two passes): half the launches and
    #: no gradient-accumulation adds; costs the overlap of D(clean) with the enhancer forward.
    batch_d_step: bool = False
    #: G step: do not compute the discriminators' parameter gradients (train.py computes them only because autograd
    #: cannot know they are dead: the next d_opt.zero_grad discards them).  Changes no weight and no
    #: loss, but it does change the (unused) .grad the D parameters hold after the step, hence opt-in, default off.
    skip_dead_d_grads: bool = False
    #: G step: finish the (never read) discriminator parameter gradients on helper streams that are joined at the end of
    #: the G phase instead of inside the discriminators' backward, so they overlap the generator's backward
    #: (lctgan.config.defer_dead_param_grads).  Every .grad holds the reference's value once the phase has ended.
    defer_dead_d_grads: bool = False


def _align_tf_targets(irm_c: torch.Tensor, pred_mask_c: torch.Tensor):
    """train.py:388-413: crop both [B, F, T] tensors to the common number of frames."""
    if irm_c.dim() != 3 or pred_mask_c.dim() != 3:
        raise ValueError(f"Expected irm_c and pred_mask_c to be [B, F, T], got {irm_c.shape}, {pred_mask_c.shape}")
    if irm_c.shape[0] != pred_mask_c.shape[0] or irm_c.shape[1] != pred_mask_c.shape[1]:
        raise ValueError(f"Batch/Freq mismatch: irm_c {irm_c.shape}, pred_mask_c {pred_mask_c.shape}")
    t = min(irm_c.shape[-1], pred_mask_c.shape[-1])
    return irm_c[..., :t], pred_mask_c[..., :t]


def _phase_d(M, noisy, clean, args: StepArgs, st: dict) -> None:
    """TF features + discriminator forward/backward (train.py:171-199)."""
    enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt = M
    d_opt.zero_grad(set_to_none=True)
    batched_logits = None
    if args.reuse_enhancer_forward and args.batch_d_step and noisy.is_cuda:
        g_opt.zero_grad(set_to_none=True)
        # [clean; enhanced] goes through every sub-discriminator as ONE batch of 2B (one autograd node, one backward),
        # but in two parts: the clean half - and the weight preparation - need nothing from the generator, so they run
        # on the side streams while the generator's forward (a chain of small kernels) has the GPU to itself
        nb = clean.shape[0]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream())
        st["enhanced"], st["mask_c"] = enhancer(noisy)      # enqueued first, so that the graph launches it first
        early = begin_split_forward(mpd, msd, clean.contiguous(), after=start)
        st["irm_c"] = tf_features(noisy, clean)["irm_c"]
        pl, _, sl, _ = finish_split_forward(mpd, msd, st["enhanced"], early)
        batched_logits = list(pl) + list(sl)        # rows [0, nb) real, [nb, 2 nb) fake: one loss node, no slicing
    elif args.reuse_enhancer_forward and noisy.is_cuda:
        g_opt.zero_grad(set_to_none=True)
        cur = torch.cuda.current_stream()
        side = config.side_streams(17, noisy.device)[16]
        side.wait_stream(cur)
        with torch.cuda.stream(side):                   # generator forward (with grad), overlapped with D(clean)
            st["enhanced"], st["mask_c"] = enhancer(noisy)
        st["irm_c"] = tf_features(noisy, clean)["irm_c"]
        (mpd_real, _, msd_real, _), = run_discriminators(mpd, msd, [clean])          # all 8 chains at once
        cur.wait_stream(side)
        enhanced_for_d = st["enhanced"].detach()
        # (same 8 streams as the real pass: giving the fake chains their own streams was measured 30 % slower - the
        # autograd engine then has to synchronise the two streams at every shared parameter's AccumulateGrad)
        (mpd_fake, _, msd_fake, _), = run_discriminators(mpd, msd, [enhanced_for_d])
    else:
        st["irm_c"] = tf_features(noisy, clean)["irm_c"]
        with torch.no_grad():
            enhanced_for_d, _ = enhancer(noisy)
        mpd_real, _ = mpd(clean)
        mpd_fake, _ = mpd(enhanced_for_d)
        msd_real, _ = msd(clean)
        msd_fake, _ = msd(enhanced_for_d)
    if batched_logits is not None:
        d_loss = LF.d_loss_batched(batched_logits, nb, args.gan_loss)
    else:
        d_loss = L.discriminator_loss(L._flatten_logits_lists(mpd_real, msd_real),
                                      L._flatten_logits_lists(mpd_fake, msd_fake), args.gan_loss)
    # data parallel: every sub-discriminator hands its finished gradient buffer to the exchange as soon as its own
    # backward is done (the all-reduce of the first ones overlaps the backward of the others)
    config.stack_grad_hook = getattr(st.get("after_d"), "reduce_async", None) if noisy.is_cuda else None
    try:
        d_loss.backward()
    finally:
        config.stack_grad_hook = None
    st["d_loss"] = d_loss.detach()


def _phase_g(M, noisy, clean, args: StepArgs, st: dict, join_dead: bool = True) -> None:
    """Discriminator update + generator forward/backward (train.py:200-245)."""
    enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt = M
    d_opt.step()
    fm_parts = None
    if "enhanced" in st:
        enhanced, mask_c = st.pop("enhanced"), st.pop("mask_c")
    else:
        g_opt.zero_grad(set_to_none=True)
        enhanced, mask_c = enhancer(noisy)
    if args.reuse_enhancer_forward and noisy.is_cuda:
        # the spectral and mask losses need nothing from the discriminators: a dozen small kernels that run on a side
        # stream beside the 16 chains (and their backward beside the D data gradients) instead of after them
        cur = torch.cuda.current_stream()
        side = config.side_streams(20, noisy.device)[19]
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            mr_loss, _ = mrstft_loss(enhanced, clean)
            irm_al, pred_al = _align_tf_targets(st["irm_c"], mask_c[:, 0])
            m_loss = L.mask_mse_loss(pred_al, irm_al)
        # real feature maps (no grad) and the generator's fake pass are independent: 16 chains at once.  The feature-
        # matching term of a sub-discriminator is reduced on the stream of its enhanced chain as soon as both of its
        # chains are done (and differentiated there, at the head of its own backward) instead of in one launch over all
        # 51 maps after the join: 0.13 + 0.24 ms of HBM-bound kernels that had the GPU to themselves
        nd = len(mpd.discriminators) + len(msd.discriminators)
        nmaps = sum(len(d._specs) for d in list(mpd.discriminators) + list(msd.discriminators))
        fm_parts = [None] * nd

        def fm_term(i, outs, stream, done):
            if i < nd:
                return                          # a clean (no-grad) chain: nothing to do yet
            k = i - nd
            if stream is not None and done[k] is not None:
                stream.wait_event(done[k])      # the clean chain of the same sub-discriminator
            real, fake = list(outs[k][1]), list(outs[i][1])
            fm_parts[k] = LF.mt_loss(ops.OP_ABS_DIFF, fake, real, [1.0 / (t.numel() * nmaps) for t in fake])

        (_, mpd_real_f, _, msd_real_f), (mpd_fake_g, mpd_fake_f, msd_fake_g, msd_fake_f) = \
            run_discriminators(mpd, msd, [clean, enhanced], no_grad=[True, False], post=fm_term)
        cur.wait_stream(side)
    else:
        mr_loss, _ = mrstft_loss(enhanced, clean)
        irm_al, pred_al = _align_tf_targets(st["irm_c"], mask_c[:, 0])
        m_loss = L.mask_mse_loss(pred_al, irm_al)
        mpd_fake_g, mpd_fake_f = mpd(enhanced)
        msd_fake_g, msd_fake_f = msd(enhanced)
        with torch.no_grad():
            _, mpd_real_f = mpd(clean)
            _, msd_real_f = msd(clean)
    adv_loss = L.generator_adv_loss(L._flatten_logits_lists(mpd_fake_g, msd_fake_g), args.gan_loss)
    if fm_parts is not None and all(t is not None for t in fm_parts):
        # train.py:240-243 as one kernel (same terms and weights), the feature-matching loss as its per-discriminator terms
        w_fm = args.lambda_adv * args.lambda_fm
        g_loss = LF.weighted_sum([mr_loss, m_loss, adv_loss] + fm_parts,
                                 [1.0, args.lambda_mask, args.lambda_adv] + [w_fm] * len(fm_parts))
        fm_loss = LF.weighted_sum([t.detach() for t in fm_parts], [1.0] * len(fm_parts))
    else:
        fm_loss = L.feature_matching_loss(mpd_real_f + msd_real_f, mpd_fake_f + msd_fake_f)
        if noisy.is_cuda:   # train.py:240-243 as one kernel (same terms and weights)
            g_loss = LF.weighted_sum([mr_loss, m_loss, adv_loss, fm_loss],
                                     [1.0, args.lambda_mask, args.lambda_adv, args.lambda_adv * args.lambda_fm])
        else:
            g_loss = mr_loss + args.lambda_mask * m_loss + args.lambda_adv * (adv_loss + args.lambda_fm * fm_loss)
    config.defer_dead_param_grads = bool(args.defer_dead_d_grads) and noisy.is_cuda
    try:
        g_loss.backward()
    finally:
        config.defer_dead_param_grads = False
        if join_dead:      # (join_dead=False: the caller joins after the generator's optimiser step, which they can overlap)
            config.join_deferred_param_grads()
    st.update(g_loss=g_loss.detach(), mr=mr_loss.detach(), mask=m_loss.detach(), adv=adv_loss.detach(),
              fm=fm_loss.detach())


def _phase_opt_g(M, args: StepArgs, pre_scale: float = 1.0) -> None:
    """Gradient clipping + generator update (train.py:246-249)."""
    enhancer, g_opt = M[0], M[5]
    if args.grad_clip > 0.0:
        if next(enhancer.parameters()).is_cuda:
            # fused global-norm clip: two launches over the flat buffer the generator's backward wrote
            clip_grad_norm_(enhancer.parameters(), args.grad_clip, arena=getattr(enhancer.gen, "_grad_arena", None),
                            pre_scale=pre_scale)
        else:
            torch.nn.utils.clip_grad_norm_(enhancer.parameters(), args.grad_clip)
    elif pre_scale != 1.0:
        # no clip to fold the data-parallel 1 / world_size into: one pass over the gradient buffer instead
        clip_grad_norm_(enhancer.parameters(), float("inf"), arena=getattr(enhancer.gen, "_grad_arena", None),
                        pre_scale=pre_scale)
    g_opt.step()


_OUT_KEYS = ("d_loss", "g_loss", "mr", "mask", "adv", "fm")


def _set_skip_dead(mpd, msd, flag: bool) -> None:
    for d in list(mpd.discriminators) + list(msd.discriminators):
        d.skip_param_grads_when_input_requires_grad = flag


def train_step(enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt, noisy: torch.Tensor,
               clean: torch.Tensor, args: StepArgs, after_d_backward=None,
               after_g_backward=None) -> Dict[str, torch.Tensor]:
    """One D step + one G step.  Returns the loss tensors (on device; no host sync happens here).
    ``after_*_backward`` are the data-parallel hooks (gradient all-reduce) of lctgan.parallel."""
    M = (enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt)
    _set_skip_dead(mpd, msd, args.skip_dead_d_grads)
    st: dict = {"after_d": after_d_backward}
    _phase_d(M, noisy, clean, args, st)
    if after_d_backward is not None:
        after_d_backward()
    try:
        _phase_g(M, noisy, clean, args, st, join_dead=False)
        _exchange_g(enhancer, after_g_backward)
        _phase_opt_g(M, args, _pre_scale(after_g_backward))
    finally:
        config.join_deferred_param_grads()     # the dead D gradients touch nothing the clip / generator update reads
    return {k: st[k] for k in _OUT_KEYS}


def _exchange_g(enhancer, after_g_backward) -> None:
    """Generator gradient exchange: in place on the flat buffer the generator's backward wrote (no pack / unpack)."""
    if after_g_backward is None:
        return
    arena = getattr(getattr(enhancer, "gen", None), "_grad_arena", None)
    if arena is not None and hasattr(after_g_backward, "reduce_async"):
        after_g_backward.reduce_async(arena)
    after_g_backward()


def _pre_scale(after_g_backward) -> float:
    """1 / world_size when the exchange leaves the SUM in .grad for the clip to fold in (FlatGradAllReduce.fold_scale)."""
    return after_g_backward.grad_scale if getattr(after_g_backward, "fold_scale", False) else 1.0


def synthetic_batch(batch: int, samples: int, seed: int = 1234):
    """The synthetic workload of SURVEY.md section 8(d): clean ~ N(0, 0.1^2), noisy = clean + N(0, 0.05^2), drawn on the
    CPU generator (so that every implementation sees the same waveforms).  Returns (noisy, clean) on the host."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.randn(batch, samples, generator=g) * 0.1
    noisy = clean + torch.randn(batch, samples, generator=g) * 0.05
    return noisy, clean


def snapshot_state(modules, optimizers):
    """Copies of every parameter / buffer and of the optimisers' state tensors (see restore_state)."""
    snap = {"mod": [[t.detach().clone() for t in list(m.parameters()) + list(m.buffers())] for m in modules], "opt": []}
    for o in optimizers:
        ts = []
        for group in o.param_groups:
            if isinstance(group.get("_step"), torch.Tensor):
                ts.append(group["_step"])
            for p in group["params"]:
                ts += [v for v in o.state.get(p, {}).values() if isinstance(v, torch.Tensor)]
        snap["opt"].append([t.detach().clone() for t in ts])
    return snap


def restore_state(modules, optimizers, snap) -> None:
    """Write a snapshot_state() back IN PLACE (addresses are unchanged, so a captured GraphedTrainStep keeps working).
    Optimiser state created after the snapshot (first step) is reset to zero = a freshly built optimiser."""
    with torch.no_grad():
        for m, saved in zip(modules, snap["mod"]):
            for t, s in zip(list(m.parameters()) + list(m.buffers()), saved):
                t.copy_(s)
        for o, saved in zip(optimizers, snap["opt"]):
            ts = []
            for group in o.param_groups:
                if isinstance(group.get("_step"), torch.Tensor):
                    ts.append(group["_step"])
                for p in group["params"]:
                    ts += [v for v in o.state.get(p, {}).values() if isinstance(v, torch.Tensor)]
            if len(saved) == len(ts):
                for t, s in zip(ts, saved):
                    t.copy_(s)
            else:
                for t in ts:
                    t.zero_()


def build_models(device, gan_seed: Optional[int] = 42, max_time_context: Optional[int] = 200,
                 capturable: bool = False, fused_optim: bool = False):
    """Construct the five modules in the reference's order (train.py:569-598) so that a given seed
    yields the reference's initial weights, and the two AdamW optimisers (train.py:601-610)."""
    from datasets.tf_features import TFFeatures, TFFeaturesConfig
    from models.discriminators import MultiPeriodDiscriminator, MultiScaleDiscriminator
    from models.generator import LCTEnhancer, LCTGeneratorConfig
    if gan_seed is not None:
        torch.manual_seed(gan_seed)
    enhancer = LCTEnhancer(LCTGeneratorConfig(max_time_context=max_time_context), c=0.3).to(device)
    mpd = MultiPeriodDiscriminator().to(device)
    msd = MultiScaleDiscriminator().to(device)
    tf = TFFeatures(TFFeaturesConfig(n_fft=512, c=0.3, compress_input=False, return_stfts=False)).to(device)
    mr = L.MultiResolutionSTFTLoss(L.MRSTFTLossConfig()).to(device)
    if fused_optim:   # same hyper-parameters and update rule, one multi-tensor kernel
        from .optim import FusedAdamW
        g_opt = FusedAdamW(enhancer.parameters(), lr=2e-4, betas=(0.8, 0.99))
        d_opt = FusedAdamW(list(mpd.parameters()) + list(msd.parameters()), lr=2e-4, betas=(0.8, 0.99))
        return enhancer, mpd, msd, tf, mr, g_opt, d_opt
    g_opt = torch.optim.AdamW(enhancer.parameters(), lr=2e-4, betas=(0.8, 0.99), capturable=capturable)
    d_opt = torch.optim.AdamW(list(mpd.parameters()) + list(msd.parameters()), lr=2e-4, betas=(0.8, 0.99),
                              capturable=capturable)
    return enhancer, mpd, msd, tf, mr, g_opt, d_opt


class GraphedTrainStep:
    """The whole D+G step captured once into a CUDA graph and replayed.

    At batch 8 the step is ~1100 small kernel launches; replaying a graph removes the per-launch host cost
    (Python, ctypes, autograd bookkeeping, allocator) and lets the GPU run the kernels back to back, the
    sub-discriminators as parallel branches.  Inputs live in static device buffers (`noisy`, `clean`): copy new
    data into them, then call the object.  torch optimisers must have been built with ``capturable=True``.

    Data parallel (hooks given): still ONE graph - the NCCL all-reduces of lctgan.parallel are captured inside it on
    the exchange's communication stream, a forked branch that overlaps the rest of the discriminator backward.  If the
    collectives cannot be captured (``capture_collectives=False`` or a capture error) the step is cut into three graphs
    sharing one memory pool - [D forward + backward] | [D update, G forward + backward] | [clip, G update] - with the
    two exchanges launched eagerly in between.
    """

    def __init__(self, enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt, noisy, clean, args: StepArgs,
                 after_d_backward=None, after_g_backward=None, warmup: int = 3, capture_collectives: bool = True):
        from . import _lib
        self.noisy, self.clean = noisy, clean
        self.enhancer = enhancer
        self.hooks = (after_d_backward, after_g_backward)
        self.capture_error = None
        M = (enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt)
        _set_skip_dead(mpd, msd, args.skip_dead_d_grads)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step(*M, self.noisy, self.clean, args, after_d_backward=after_d_backward,
                           after_g_backward=after_g_backward)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.kernel_launches()
        has_hooks = after_d_backward is not None or after_g_backward is not None
        self.graphs = None
        st: dict = {}
        if capture_collectives or not has_hooks:
            try:
                st = {"after_d": after_d_backward}
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    _phase_d(M, self.noisy, self.clean, args, st)
                    if after_d_backward is not None:
                        after_d_backward()
                    _phase_g(M, self.noisy, self.clean, args, st, join_dead=False)
                    _exchange_g(enhancer, after_g_backward)
                    _phase_opt_g(M, args, _pre_scale(after_g_backward))
                    config.join_deferred_param_grads()
                self.graphs = [g]
            except Exception as e:
                if not has_hooks:
                    raise
                self.capture_error = repr(e)          # reported by bench.py; the segmented form below still runs
                torch.cuda.synchronize()
        if self.graphs is None:
            st = {"after_d": None}                    # no collective inside a graph: exchanges run between the graphs
            pool = torch.cuda.graph_pool_handle()
            g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1, pool=pool):
                _phase_d(M, self.noisy, self.clean, args, st)
            if after_d_backward is not None:
                after_d_backward()
            with torch.cuda.graph(g2, pool=pool):
                _phase_g(M, self.noisy, self.clean, args, st)
            _exchange_g(enhancer, after_g_backward)
            with torch.cuda.graph(g3, pool=pool):
                _phase_opt_g(M, args, _pre_scale(after_g_backward))
            self.graphs = [g1, g2, g3]
        self.out = {k: st[k] for k in _OUT_KEYS}
        #: lctgan kernel launches recorded in the graphs (+ the eager gradient gather/scatter) = per replayed step
        self.launches_per_step = _lib.kernel_launches() - n0

    def __call__(self) -> Dict[str, torch.Tensor]:
        if len(self.graphs) == 1:
            self.graphs[0].replay()
        else:
            self.graphs[0].replay()
            if self.hooks[0] is not None:
                self.hooks[0]()
            self.graphs[1].replay()
            _exchange_g(self.enhancer, self.hooks[1])
            self.graphs[2].replay()
        return self.out
