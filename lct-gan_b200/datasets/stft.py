"""STFT / iSTFT front end with the reference's API (jqshang/LCT-GAN datasets/stft.py), running on the
lctgan sm_100a kernels instead of torch.stft / torch.istft.

Same public names, arguments, return shapes and error behaviour as the reference:
STFTConfig (:10-34), ComplexSTFT.forward/.istft (:37-132), magnitude (:138-160), compress /
decompress (:163-178), compute_compressed_irm (:184-218), decompress_mask (:221-240),
apply_mask (:243-290), make_lct_stft (:293-312).  CUDA tensors only: there is no CPU path.
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from lctgan import functional as LF
from lctgan import ops as _ops


@dataclass
class STFTConfig:
    """Analysis parameters; hop/win default to n_fft // 2 and n_fft after ``finalize()``."""
    n_fft: int = 512
    hop_length: Optional[int] = None
    win_length: Optional[int] = None
    window: str = "hann"
    center: bool = True
    pad_mode: str = "reflect"
    normalized: bool = False
    onesided: bool = True

    def finalize(self) -> "STFTConfig":
        if self.hop_length is None:
            self.hop_length = self.n_fft // 2
        if self.win_length is None:
            self.win_length = self.n_fft
        return self


class ComplexSTFT(nn.Module):
    """Differentiable STFT ([B, T] -> complex [B, F, Tf]) and iSTFT with a periodic Hann window held
    as the buffer ``window`` (state_dict key identical to the reference)."""

    def __init__(self, cfg: STFTConfig):
        super().__init__()
        self.cfg = cfg.finalize()
        if self.cfg.window.lower() != "hann":
            raise ValueError("Only 'hann' window is currently supported.")
        c = self.cfg
        if not (c.center and c.pad_mode == "reflect" and c.onesided and not c.normalized):
            raise ValueError("lctgan ComplexSTFT implements center=True, pad_mode='reflect', onesided=True, "
                             "normalized=False (the only configuration the reference ever builds)")
        self.register_buffer("window", torch.hann_window(self.cfg.win_length))

    def _full_window(self, device) -> torch.Tensor:
        """The window the transform multiplies by: fp32, on `device`, centred-zero-padded to n_fft."""
        w = self.window
        n, wl = self.cfg.n_fft, self.cfg.win_length
        if w.device == device and w.dtype == torch.float32 and wl >= n:
            return w
        # derived window (other device / dtype, or centre-padded): memoised per (buffer address, version, device) so
        # that repeated calls hand the SAME tensor to the kernels (and to the envelope cache keyed on it)
        key = (w.data_ptr(), w._version, str(device))
        memo = self.__dict__.setdefault("_derived_window", {})
        if memo.get("key") != key:
            d = w.to(device=device, dtype=torch.float32)
            if wl < n:
                left = (n - wl) // 2
                d = torch.nn.functional.pad(d, (left, n - wl - left))
            memo.update(key=key, src=w, win=d)
        return memo["win"]

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        if waveform.dim() != 2:
            raise ValueError(f"Expected waveform of shape [B, T], got {waveform.shape}")
        return LF.STFTFn.apply(waveform, self._full_window(waveform.device), self.cfg.n_fft, self.cfg.hop_length)

    def istft(self, stft_matrix: torch.Tensor, length: Optional[int] = None) -> torch.Tensor:
        if not torch.is_complex(stft_matrix):
            raise ValueError("stft_matrix must be a complex tensor.")
        if stft_matrix.dim() != 3:
            raise ValueError(f"Expected stft_matrix of shape [B, F, T], got {stft_matrix.shape}")
        if length is None:   # torch.istft's default for center=True
            length = self.cfg.hop_length * (stft_matrix.shape[-1] - 1)
        return LF.ISTFTFn.apply(stft_matrix, self._full_window(stft_matrix.device), self.cfg.n_fft,
                                self.cfg.hop_length, int(length))


def magnitude(stft_matrix: torch.Tensor, power: float = 1.0, eps: float = 1e-12) -> torch.Tensor:
    """max(|X|, eps) ** power."""
    if not torch.is_complex(stft_matrix):
        raise ValueError("stft_matrix must be a complex tensor.")
    return LF.MagnitudeFn.apply(stft_matrix, float(power), float(eps))


def compress(x: torch.Tensor, c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    """max(x, eps) ** c."""
    return LF.PowClampFn.apply(x, float(c), float(eps))


def decompress(x_c: torch.Tensor, c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    """max(x_c, eps) ** (1 / c)."""
    return LF.PowClampFn.apply(x_c, 1.0 / float(c), float(eps))


def compute_compressed_irm(clean_stft: torch.Tensor, noisy_stft: torch.Tensor, c: float = 0.3,
                           gamma: float = 1e-12, eps: float = 1e-12) -> torch.Tensor:
    """|S|^c / (|X|^c + gamma) with both magnitudes floored at eps; inputs are data (no gradient)."""
    if not (torch.is_complex(clean_stft) and torch.is_complex(noisy_stft)):
        raise ValueError("clean_stft and noisy_stft must be complex tensors.")
    if clean_stft.requires_grad or noisy_stft.requires_grad:
        raise RuntimeError("compute_compressed_irm: gradients are not implemented (targets are data)")
    if clean_stft.dim() == 3:
        out = _ops.irm_fwd(_ops.spec_phys(clean_stft), _ops.spec_phys(noisy_stft), c, gamma, eps)
        return _ops.spec_view(out)
    return _ops.irm_fwd(clean_stft.contiguous(), noisy_stft.contiguous(), c, gamma, eps)


def decompress_mask(mask_c: torch.Tensor, c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    return decompress(mask_c, c=c, eps=eps)


def apply_mask(noisy_stft: torch.Tensor, mask: torch.Tensor, compressed: bool = False, c: float = 0.3,
               eps: float = 1e-12) -> torch.Tensor:
    """noisy_stft * max(mask', 0), mask' = max(mask, eps)^(1/c) when ``compressed``."""
    if not torch.is_complex(noisy_stft):
        raise ValueError("noisy_stft must be a complex tensor.")
    if mask.dim() == 4:
        if mask.size(1) != 1:
            raise ValueError(f"Expected mask shape [B, 1, F, T], got {mask.shape}")
        mask = mask[:, 0, :, :]
    if mask.dim() != 3:
        raise ValueError(f"Expected mask shape [B, F, T] (or [B, 1, F, T]), got {mask.shape}")
    return LF.ApplyMaskFn.apply(noisy_stft, mask, bool(compressed), float(c), float(eps))


def make_lct_stft(n_fft: int = 512, hop_length: Optional[int] = None,
                  win_length: Optional[int] = None) -> ComplexSTFT:
    """The generator's analysis/synthesis pair (512-point, 50 % overlap by default)."""
    return ComplexSTFT(STFTConfig(n_fft=n_fft, hop_length=hop_length, win_length=win_length).finalize())
