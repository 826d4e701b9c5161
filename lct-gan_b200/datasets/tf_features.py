"""TFFeatures with the reference's API (jqshang/LCT-GAN datasets/tf_features.py:17-146): noisy + clean
waveforms -> {noisy_mag, irm_c, noisy_mag_c[, noisy_stft, clean_stft]}.  One fused lctgan kernel
computes both STFTs (two-for-one real FFT) and all three real outputs.
"""
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn as nn

from datasets.stft import ComplexSTFT, STFTConfig, make_lct_stft
from lctgan import ops as _ops


@dataclass
class TFFeaturesConfig:
    n_fft: int = 512
    hop_length: Optional[int] = None
    win_length: Optional[int] = None
    c: float = 0.3                 # magnitude compression exponent
    compress_input: bool = False   # expose |X|^c instead of |X| under "noisy_mag"
    return_stfts: bool = True      # also return the two complex spectrograms


class TFFeatures(nn.Module):
    def __init__(self, cfg: Optional[TFFeaturesConfig] = None):
        super().__init__()
        self.cfg = cfg if cfg is not None else TFFeaturesConfig()
        cfg = self.cfg
        if cfg.n_fft == 512 and cfg.hop_length is None and cfg.win_length is None:
            self.stft = make_lct_stft(n_fft=cfg.n_fft)
        else:
            self.stft = ComplexSTFT(STFTConfig(n_fft=cfg.n_fft, hop_length=cfg.hop_length,
                                               win_length=cfg.win_length).finalize())
        self.c = cfg.c

    def forward(self, noisy_wave: torch.Tensor, clean_wave: torch.Tensor) -> Dict[str, torch.Tensor]:
        if noisy_wave.dim() != 2 or clean_wave.dim() != 2:
            raise ValueError(f"Expected noisy_wave and clean_wave of shape [B, T], "
                             f"got {noisy_wave.shape}, {clean_wave.shape}")
        if noisy_wave.shape != clean_wave.shape:
            raise ValueError(f"noisy_wave and clean_wave must have same shape, "
                             f"got {noisy_wave.shape} vs {clean_wave.shape}")
        if not noisy_wave.is_cuda:
            raise RuntimeError("TFFeatures (lctgan) is CUDA only (sm_100a); there is no CPU fallback")
        if noisy_wave.requires_grad or clean_wave.requires_grad:
            raise RuntimeError("TFFeatures: inputs are data; gradients through the features are not implemented")
        sc = self.stft.cfg
        win = self.stft._full_window(noisy_wave.device)
        nmag, irm, nmag_c, ns, cs = _ops.tf_features_fwd(noisy_wave, clean_wave, win, sc.n_fft, sc.hop_length,
                                                         c=self.c, want_specs=self.cfg.return_stfts)
        v = _ops.spec_view
        feats: Dict[str, torch.Tensor] = {
            "noisy_mag": v(nmag_c if self.cfg.compress_input else nmag),
            "irm_c": v(irm),
            "noisy_mag_c": v(nmag_c),
        }
        if self.cfg.return_stfts:
            feats["noisy_stft"] = v(ns)
            feats["clean_stft"] = v(cs)
        return feats
