"""LCT-GAN training losses with the reference's API (jqshang/LCT-GAN losses.py), on the lctgan
sm_100a kernels: the multi-resolution STFT loss reduces inside the STFT kernel, and each
adversarial / feature-matching / mask loss is ONE multi-tensor launch instead of one mse/l1
kernel chain per tensor.

Reference anchors: MRSTFTLossConfig :11-19, MultiResolutionSTFTLoss :22-100,
_flatten_logits_lists :103-107, discriminator_loss :110-135, generator_adv_loss :138-151,
feature_matching_loss :154-173, mask_mse_loss :176-181.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from datasets.stft import ComplexSTFT, STFTConfig
from lctgan import functional as LF
from lctgan import ops as _ops


@dataclass
class MRSTFTLossConfig:
    fft_sizes: Tuple[int, ...] = (320, 512, 768)
    hop_factors: Tuple[float, ...] = (0.5, 0.5, 0.5)
    mag_weight: float = 1.0
    complex_weight: float = 1.0
    main_fft_size: int = 512
    main_fft_weight: float = 2.0
    default_weight: float = 1.0


class MultiResolutionSTFTLoss(nn.Module):
    """sum_r w_r * (mag_weight * mse(|Yh|, |Y|) + complex_weight * mean |Yh - Y|^2) / sum_r w_r."""

    def __init__(self, cfg: Optional[MRSTFTLossConfig] = None):
        super().__init__()
        self.cfg = cfg if cfg is not None else MRSTFTLossConfig()
        cfg = self.cfg
        self.stfts = nn.ModuleList()
        self.weights: List[float] = []
        for n_fft, hop_factor in zip(cfg.fft_sizes, cfg.hop_factors):
            self.stfts.append(ComplexSTFT(STFTConfig(n_fft=n_fft, hop_length=int(round(n_fft * hop_factor)),
                                                     win_length=n_fft).finalize()))
            self.weights.append(cfg.main_fft_weight if n_fft == cfg.main_fft_size else cfg.default_weight)

    def forward(self, y_hat: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        if y_hat.dim() != 2 or y.dim() != 2:
            raise ValueError(f"Expected y_hat, y of shape [B, T], got {y_hat.shape}, {y.shape}")
        res = tuple((s.cfg.n_fft, s.cfg.hop_length, float(w)) for s, w in zip(self.stfts, self.weights))
        wins = [s._full_window(y_hat.device) for s in self.stfts]
        total, mag_t, cplx_t = LF.MRSTFTLossFn.apply(y_hat, y, res, float(self.cfg.mag_weight),
                                                     float(self.cfg.complex_weight), 1e-12, *wins)
        details = {"mrstft_total": total.detach(), "mrstft_mag": mag_t, "mrstft_complex": cplx_t}
        return total, details


def _flatten_logits_lists(*logits_lists):
    flat: List[torch.Tensor] = []
    for lst in logits_lists:
        flat.extend(list(lst))
    return flat


def _mean_scales(ts: Sequence[torch.Tensor], denom: int, sign: float = 1.0) -> List[float]:
    return [sign / (t.numel() * denom) for t in ts]


def discriminator_loss(real_logits: Sequence[torch.Tensor], fake_logits: Sequence[torch.Tensor],
                       loss_type: str = "ls"):
    """ls: mean((r-1)^2) + mean(f^2); hinge: mean(relu(1-r)) + mean(relu(1+f)); averaged over discriminators."""
    if len(real_logits) != len(fake_logits):
        raise ValueError("real_logits and fake_logits must have the same length.")
    if loss_type not in ("ls", "hinge"):
        raise ValueError(f"Unknown loss_type: {loss_type}")
    n = max(len(real_logits), 1)
    if len(real_logits) == 0:
        return 0.0
    real, fake = list(real_logits), list(fake_logits)
    if loss_type == "ls":
        lr = LF.mt_loss(_ops.OP_SQ_CONST, real, None, _mean_scales(real, n), k0=1.0)
        lf = LF.mt_loss(_ops.OP_SQ_CONST, fake, None, _mean_scales(fake, n), k0=0.0)
    else:
        lr = LF.mt_loss(_ops.OP_RELU_AFFINE, real, None, _mean_scales(real, n), k0=1.0, k1=-1.0)
        lf = LF.mt_loss(_ops.OP_RELU_AFFINE, fake, None, _mean_scales(fake, n), k0=1.0, k1=1.0)
    return LF.weighted_sum([lr, lf], [1.0, 1.0]) if lr.is_cuda else lr + lf


def generator_adv_loss(fake_logits, loss_type="ls"):
    """ls: mean((f-1)^2); hinge: -mean(f); averaged over discriminators."""
    if loss_type not in ("ls", "hinge"):
        raise ValueError(f"Unknown loss_type: {loss_type}")
    fake = list(fake_logits)
    n = max(len(fake), 1)
    if not fake:
        return 0.0
    if loss_type == "ls":
        return LF.mt_loss(_ops.OP_SQ_CONST, fake, None, _mean_scales(fake, n), k0=1.0)
    return LF.mt_loss(_ops.OP_SUM, fake, None, _mean_scales(fake, n, sign=-1.0))


def feature_matching_loss(real_fmaps, fake_fmaps):
    """Mean over all (discriminator, layer) pairs of mean |fake - real|."""
    if len(real_fmaps) != len(fake_fmaps):
        raise ValueError("real_fmaps and fake_fmaps must have the same outer length.")
    real: List[torch.Tensor] = []
    fake: List[torch.Tensor] = []
    for rs, fs in zip(real_fmaps, fake_fmaps):
        if len(rs) != len(fs):
            raise ValueError("Mismatched feature map list lengths for a discriminator.")
        for r, f in zip(rs, fs):
            if r.shape != f.shape:
                raise ValueError(f"feature map shape mismatch: {tuple(f.shape)} vs {tuple(r.shape)}")
            real.append(r)
            fake.append(f)
    count = len(fake)
    if count == 0:
        return torch.tensor(0.0, device=real_fmaps[0][0].device)
    return LF.mt_loss(_ops.OP_ABS_DIFF, fake, real, _mean_scales(fake, count))


def mask_mse_loss(pred_mask_c, target_mask_c):
    if pred_mask_c.shape != target_mask_c.shape:
        raise ValueError(f"Shape mismatch: pred_mask_c {pred_mask_c.shape} vs target_mask_c {target_mask_c.shape}")
    p, t = pred_mask_c, target_mask_c
    # both usually arrive as [B, F, Tf] views of [B, Tf, F] buffers: reduce in memory order, no copy
    if p.dim() >= 2 and not p.is_contiguous() and p.stride() == t.stride() and p.transpose(-1, -2).is_contiguous():
        p, t = p.transpose(-1, -2), t.transpose(-1, -2)
    return LF.mt_loss(_ops.OP_SQ_DIFF, [p], [t], [1.0 / p.numel()])
