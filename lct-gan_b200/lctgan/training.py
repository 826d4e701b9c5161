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


@dataclass
class StepArgs:
    """The subset of train.py's argparse namespace the loop body reads (train.py:472-484)."""
    gan_loss: str = "ls"
    lambda_fm: float = 1.0
    lambda_mask: float = 1.0
    lambda_adv: float = 1e-2
    grad_clip: float = 5.0


def _align_tf_targets(irm_c: torch.Tensor, pred_mask_c: torch.Tensor):
    """train.py:388-413: crop both [B, F, T] tensors to the common number of frames."""
    if irm_c.dim() != 3 or pred_mask_c.dim() != 3:
        raise ValueError(f"Expected irm_c and pred_mask_c to be [B, F, T], got {irm_c.shape}, {pred_mask_c.shape}")
    if irm_c.shape[0] != pred_mask_c.shape[0] or irm_c.shape[1] != pred_mask_c.shape[1]:
        raise ValueError(f"Batch/Freq mismatch: irm_c {irm_c.shape}, pred_mask_c {pred_mask_c.shape}")
    t = min(irm_c.shape[-1], pred_mask_c.shape[-1])
    return irm_c[..., :t], pred_mask_c[..., :t]


def train_step(enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt, noisy: torch.Tensor,
               clean: torch.Tensor, args: StepArgs, after_d_backward=None,
               after_g_backward=None) -> Dict[str, torch.Tensor]:
    """One D step + one G step.  Returns the loss tensors (on device; no host sync happens here).
    ``after_*_backward`` are the data-parallel hooks (gradient all-reduce) of lctgan.parallel."""
    irm_c = tf_features(noisy, clean)["irm_c"]

    # ---- discriminator step (train.py:177-200)
    d_opt.zero_grad(set_to_none=True)
    with torch.no_grad():
        enhanced_for_d, _ = enhancer(noisy)
    mpd_real, _ = mpd(clean)
    mpd_fake, _ = mpd(enhanced_for_d)
    msd_real, _ = msd(clean)
    msd_fake, _ = msd(enhanced_for_d)
    d_loss = L.discriminator_loss(L._flatten_logits_lists(mpd_real, msd_real),
                                  L._flatten_logits_lists(mpd_fake, msd_fake), args.gan_loss)
    d_loss.backward()
    if after_d_backward is not None:
        after_d_backward()
    d_opt.step()

    # ---- generator step (train.py:205-249)
    g_opt.zero_grad(set_to_none=True)
    enhanced, mask_c = enhancer(noisy)
    mr_loss, _ = mrstft_loss(enhanced, clean)
    irm_al, pred_al = _align_tf_targets(irm_c, mask_c[:, 0])
    m_loss = L.mask_mse_loss(pred_al, irm_al)
    mpd_fake_g, mpd_fake_f = mpd(enhanced)
    msd_fake_g, msd_fake_f = msd(enhanced)
    with torch.no_grad():
        _, mpd_real_f = mpd(clean)
        _, msd_real_f = msd(clean)
    adv_loss = L.generator_adv_loss(L._flatten_logits_lists(mpd_fake_g, msd_fake_g), args.gan_loss)
    fm_loss = L.feature_matching_loss(mpd_real_f + msd_real_f, mpd_fake_f + msd_fake_f)
    g_loss = mr_loss + args.lambda_mask * m_loss + args.lambda_adv * (adv_loss + args.lambda_fm * fm_loss)
    g_loss.backward()
    if after_g_backward is not None:
        after_g_backward()
    if args.grad_clip > 0.0:
        torch.nn.utils.clip_grad_norm_(enhancer.parameters(), args.grad_clip)
    g_opt.step()
    return {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "mr": mr_loss.detach(), "mask": m_loss.detach(),
            "adv": adv_loss.detach(), "fm": fm_loss.detach()}


def build_models(device, gan_seed: Optional[int] = 42, max_time_context: Optional[int] = 200,
                 capturable: bool = False):
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
    g_opt = torch.optim.AdamW(enhancer.parameters(), lr=2e-4, betas=(0.8, 0.99), capturable=capturable)
    d_opt = torch.optim.AdamW(list(mpd.parameters()) + list(msd.parameters()), lr=2e-4, betas=(0.8, 0.99),
                              capturable=capturable)
    return enhancer, mpd, msd, tf, mr, g_opt, d_opt


class GraphedTrainStep:
    """The whole D+G step captured once into a CUDA graph and replayed.

    At batch 8 the step is ~1100 small kernel launches; replaying a graph removes the per-launch host cost
    (Python, ctypes, autograd bookkeeping, allocator) and lets the GPU run the kernels back to back.
    Inputs live in static device buffers (`noisy`, `clean`): copy new data into them, then call the object.
    The optimisers must have been built with ``capturable=True``.
    """

    def __init__(self, enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt, noisy, clean, args: StepArgs,
                 after_d_backward=None, after_g_backward=None, warmup: int = 3):
        self.noisy, self.clean = noisy, clean
        run = lambda: train_step(enhancer, mpd, msd, tf_features, mrstft_loss, g_opt, d_opt, self.noisy, self.clean,
                                 args, after_d_backward=after_d_backward, after_g_backward=after_g_backward)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.kernel_launches()
        with torch.cuda.graph(self.graph):
            self.out = run()
        #: lctgan kernel launches recorded in the graph (= launches per replayed step)
        self.launches_per_step = _lib.kernel_launches() - n0

    def __call__(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.out
