"""LCT-GAN mask generator and waveform enhancer with the reference's API
(jqshang/LCT-GAN models/generator.py), executing on the lctgan sm_100a kernels.

The modules below only *hold parameters* in stock torch containers (nn.Conv2d, nn.GRU,
nn.MultiheadAttention, ...), created in the reference's order so that seeds, ``state_dict`` keys,
shapes and parameter iteration order are identical; none of the containers'
own forward methods is ever called.  Forward and backward run through ``lctgan.gen_impl``.

Reference anchors: LCTGeneratorConfig :19-28, GRUblockf :31-145, GRUblockt :148-255, dead classes
DownBlock/UpBlock/Encoder/Decoder :258-437, LCTGenerator :440-632, LCTEnhancer :635-697.
"""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from datasets.stft import ComplexSTFT, STFTConfig, apply_mask, magnitude, make_lct_stft
from lctgan import gen_impl as _gi
from lctgan import functional as LF
from lctgan import ops as _ops


@dataclass
class LCTGeneratorConfig:
    in_channels: int = 1
    out_channels: int = 1
    enc_channels: Tuple[int, int, int] = (16, 32, 64)
    dec_channels: Tuple[int, int, int] = (64, 32, 16)
    num_heads: int = 4                       # stored, never read (as in the reference)
    gru_groups: int = 4                      # stored, never read
    max_time_context: Optional[int] = None   # stored, never read
    output_activation: str = "sigmoid"


class _BlockFn(torch.autograd.Function):
    """A single GRUblockf / GRUblockt on channels-last rows (used when a block is called on its own)."""

    @staticmethod
    def forward(ctx, rows, geom, freq, names, *params):
        B, T, F = geom
        P = dict(zip(names, params))
        S = {} if any(ctx.needs_input_grad) else None     # (released with ctx when no graph is recorded)
        out = _gi._block_fwd(P, "blk", rows.contiguous(), B, T, F, freq, S)
        ctx.S, ctx.geom, ctx.freq, ctx.names = S, geom, freq, names
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, g):
        P = dict(zip(ctx.names, ctx.saved_tensors))
        B, T, F = ctx.geom
        GR = {}
        if ctx.S is None:
            raise RuntimeError("GRU block backward: saved activations already released (second backward through the "
                               "same graph is not supported)")
        dx = _gi._block_bwd(P, "blk", g.contiguous(), B, T, F, ctx.freq, ctx.S, GR)
        ctx.S = None
        return (dx, None, None, None, *[GR.get(n) for n in ctx.names])


class _GRUBlockBase(nn.Module):
    _bidirectional = True

    def __init__(self, channels: int = 64):
        super().__init__()
        assert channels == 64, "This implementation assumes 64 channels."
        self.channels = channels
        self.num_groups = 4
        self.group_dim = channels // self.num_groups
        for i in range(1, 5):
            setattr(self, f"gru{i}", nn.GRU(input_size=self.group_dim, hidden_size=self.group_dim, batch_first=True,
                                            bidirectional=self._bidirectional))
        self.attn = nn.MultiheadAttention(embed_dim=channels, num_heads=4, batch_first=True)
        self.activationtrans = nn.LeakyReLU(0.2, inplace=True)
        self.layernorm1 = nn.LayerNorm(channels)
        self.layernorm2 = nn.LayerNorm(channels)
        self.lin = nn.Linear((2 if self._bidirectional else 1) * channels, channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, 64, T, F] -> [B, 64, T, F]."""
        B, C, T, F = x.shape
        assert C == self.channels
        if not x.is_cuda:
            raise RuntimeError("lctgan GRU blocks are CUDA only (sm_100a); there is no CPU fallback")
        names = _gi.block_param_names("blk", self._bidirectional)
        mine = dict(self.named_parameters())
        params = [mine[n[len("blk."):]] for n in names]
        rows = x.permute(0, 2, 3, 1).reshape(B * T * F, C)
        out = _BlockFn.apply(rows, (B, T, F), self._bidirectional, tuple(names), *params)
        return out.view(B, T, F, C).permute(0, 3, 1, 2)


class GRUblockf(_GRUBlockBase):
    """Frequency block: LN -> 4 bidirectional GRU(16,16) (directions summed) -> +res -> LN -> 4-head
    self-attention over frequency -> cat -> Linear(128,64) -> LeakyReLU(0.2) -> +res."""
    _bidirectional = True


class GRUblockt(_GRUBlockBase):
    """Time block: LN -> 4 unidirectional GRU(16,16) -> +res -> LN -> 4-head (unmasked) self-attention
    over time -> Linear(64,64) -> LeakyReLU(0.2) -> +res."""
    _bidirectional = False


# ---- classes the reference defines but never instantiates (API surface only, plain PyTorch) ----
class DownBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: Tuple[int, int] = (2, 3),
                 stride: Tuple[int, int] = (1, 2), negative_slope: float = 0.03):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=(1, 1))
        self.act = nn.LeakyReLU(negative_slope=negative_slope, inplace=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.act(self.conv(x))


class UpBlock(nn.Module):
    def __init__(self, in_channels: int, skip_channels: int, out_channels: int,
                 kernel_size: Tuple[int, int] = (2, 3), stride: Tuple[int, int] = (1, 2),
                 padding: Tuple[int, int] = (0, 1), output_padding: Tuple[int, int] = (0, 1),
                 negative_slope: float = 0.03):
        super().__init__()
        self.deconv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                         padding=padding, output_padding=output_padding)
        self.conv = nn.Conv2d(out_channels + skip_channels, out_channels, kernel_size=(1, 3), stride=(1, 1),
                              padding=(0, 1))
        self.act = nn.LeakyReLU(negative_slope=negative_slope, inplace=True)

    @staticmethod
    def _align(x: torch.Tensor, skip: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        t = min(x.shape[2], skip.shape[2])
        f = min(x.shape[3], skip.shape[3])
        return x[:, :, :t, :f], skip[:, :, :t, :f]

    def forward(self, x: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        x, skip = self._align(self.deconv(x), skip)
        return self.act(self.conv(torch.cat([x, skip], dim=1)))


class Encoder(nn.Module):
    def __init__(self, in_channels: int = 1, channels: Tuple[int, int, int] = (16, 32, 64)):
        super().__init__()
        c1, c2, c3 = channels
        self.blocks = nn.ModuleList([DownBlock(in_channels, c1), DownBlock(c1, c2), DownBlock(c2, c3)])

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        skips: List[torch.Tensor] = []
        for blk in self.blocks:
            x = blk(x)
            skips.append(x)
        return x, skips


class Decoder(nn.Module):
    def __init__(self, out_channels: int, channels: Tuple[int, int, int] = (64, 32, 16)):
        super().__init__()
        c3, c2, c1 = channels
        self.up3 = UpBlock(in_channels=c3, skip_channels=c3, out_channels=c2)
        self.up2 = UpBlock(in_channels=c2, skip_channels=c2, out_channels=c1)
        self.up1 = UpBlock(in_channels=c1, skip_channels=c1, out_channels=out_channels)

    def forward(self, x: torch.Tensor, skips: List[torch.Tensor]) -> torch.Tensor:
        s1, s2, s3 = skips
        return self.up1(self.up2(self.up3(x, s3), s2), s1)


class LCTGenerator(nn.Module):
    """noisy_mag [B, 1, F, T] -> compressed mask [B, 1, F, T] (sigmoid output in [0.5, 1))."""

    def __init__(self, cfg: LCTGeneratorConfig):
        super().__init__()
        self.cfg = cfg
        in_ch, out_ch = cfg.in_channels, cfg.out_channels
        e1, e2, e3 = cfg.enc_channels
        d3, d2, d1 = cfg.dec_channels
        assert in_ch == 1 and out_ch == 1, "FTFNet is defined for 1→1 masks."
        enc = dict(kernel_size=(2, 3), stride=(1, 2), padding=(1, 1))
        self.conv1 = nn.Conv2d(in_ch, e1, **enc)
        self.conv2 = nn.Conv2d(e1, e2, **enc)
        self.conv3 = nn.Conv2d(e2, e3, **enc)
        self.skip2 = nn.Conv2d(in_ch, e3, kernel_size=1)
        self.skip3 = nn.Conv2d(in_ch, e2, kernel_size=1)
        self.skip4 = nn.Conv2d(in_ch, e1, kernel_size=1)
        self.GRUf1 = GRUblockf(channels=e3)
        self.GRUt1 = GRUblockt(channels=e3)
        self.GRUf2 = GRUblockf(channels=e3)
        dec = dict(kernel_size=(2, 3), stride=(1, 2), padding=(1, 1), output_padding=(0, 1))
        self.deconv2 = nn.ConvTranspose2d(e3, e2, **dec)
        self.deconv3 = nn.ConvTranspose2d(e2, e1, **dec)
        self.deconv4 = nn.ConvTranspose2d(e1, out_ch, **dec)
        self.activation = nn.LeakyReLU(0.2, inplace=True)
        self.activations = nn.Sigmoid()
        self.layernorm = nn.LayerNorm(e3)
        self.pad = nn.ConstantPad2d((0, 0, 0, 0), 0.0)
        self.act_final = nn.ReLU()
        self._names: Optional[Tuple[str, ...]] = None

    @staticmethod
    def _align(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        t = min(a.shape[2], b.shape[2])
        f = min(a.shape[3], b.shape[3])
        return a[:, :, :t, :f], b[:, :, :t, :f]

    def _flat_params(self):
        if self._names is None:
            self._names = tuple(n for n, _ in self.named_parameters())
        return self._names, [p for _, p in self.named_parameters()]

    def forward_phys(self, mag_phys: torch.Tensor) -> torch.Tensor:
        """mag_phys: [B, T, F] (the STFT kernels' layout) -> mask [B, T, F]."""
        names, params = self._flat_params()
        return _gi.GeneratorFn.apply(mag_phys, self.cfg.output_activation == "sigmoid", names, self,
                                     torch.is_grad_enabled(), *params)

    def forward(self, noisy_mag: torch.Tensor) -> torch.Tensor:
        if noisy_mag.dim() != 4 or noisy_mag.size(1) != 1:
            raise ValueError(f"Expected noisy_mag [B, 1, F, T], got {noisy_mag.shape}")
        mag_phys = noisy_mag[:, 0].transpose(1, 2)
        if not mag_phys.is_contiguous():
            mag_phys = mag_phys.contiguous()
        mask = self.forward_phys(mag_phys)
        return mask.transpose(1, 2).unsqueeze(1)


class LCTEnhancer(nn.Module):
    """noisy waveform [B, T] -> (enhanced waveform [B, T], mask_c [B, 1, F, Tf]):
    STFT -> |.| -> LCTGenerator -> compressed-mask application -> iSTFT."""

    def __init__(self, gen_cfg: LCTGeneratorConfig, c: float = 0.3, stft_cfg: Optional[STFTConfig] = None):
        super().__init__()
        self.gen = LCTGenerator(gen_cfg)
        self.c = c
        self.stft = make_lct_stft(n_fft=512) if stft_cfg is None else ComplexSTFT(stft_cfg)

    def forward(self, noisy_wave: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if noisy_wave.dim() != 2:
            raise ValueError(f"Expected noisy_wave [B, T], got {noisy_wave.shape}")
        if noisy_wave.requires_grad:
            # general (rare) path: gradient w.r.t. the input waveform, composed from the unfused operators
            noisy_stft = self.stft(noisy_wave)
            mask_c = self.gen(magnitude(noisy_stft).unsqueeze(1))
            enhanced = self.stft.istft(apply_mask(noisy_stft, mask_c, compressed=True, c=self.c),
                                       length=noisy_wave.shape[-1])
            return enhanced, mask_c
        if not noisy_wave.is_cuda:
            raise RuntimeError("LCTEnhancer (lctgan) is CUDA only (sm_100a); there is no CPU fallback")
        sc = self.stft.cfg
        win = self.stft._full_window(noisy_wave.device)
        # fused front: STFT + magnitude in one kernel; fused tail: mask decompression + iFFT + OLA + envelope
        spec, mag = _ops.stft_fwd(noisy_wave, win, sc.n_fft, sc.hop_length, want_mag=True)
        mask = self.gen.forward_phys(mag)
        enhanced = LF.MaskedISTFTFn.apply(spec, mask, win, sc.n_fft, sc.hop_length, noisy_wave.shape[-1],
                                          float(self.c), 1e-12)
        return enhanced, mask.transpose(1, 2).unsqueeze(1)
