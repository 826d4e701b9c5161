"""Seeded initial weights of the three LCT-GAN networks, built from STOCK torch containers only.  TEST INFRASTRUCTURE
(part of the CPU oracle): used by ``bench.py --impl reference`` / ``cpu_baseline`` so that the reference arm never
imports the product package, and by tests/test_oracle_golden.py (pinned against the parameter checksums and
``state_dict`` keys dumped from the unmodified reference, tests/golden/golden_v1.pt).

The reference creates its parameters in ``__init__`` order with the default torch initialisers
(jqshang/LCT-GAN models/generator.py:449-536 LCTGenerator, :31-82 GRUblockf, :148-198 GRUblockt, :635-657 LCTEnhancer;
models/discriminators.py:30-67 PeriodDiscriminator, :106-126 MultiPeriodDiscriminator, :160-196 ScaleDiscriminator,
:227-255 MultiScaleDiscriminator; seeding train.py:32-38, construction order train.py:569-598).  Holding the same
containers in the same order under the same seed reproduces the same random stream, hence the same weights.
Nothing here has a forward pass: these modules only hold parameters.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from .lct_oracle import MPD_LAYERS, MPD_PERIODS, MSD_LAYERS, hann_window


def _wn(m):
    from torch.nn.utils import weight_norm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return weight_norm(m)


class _GRUBlock(nn.Module):
    def __init__(self, bidirectional: bool, channels: int = 64):
        super().__init__()
        gd = channels // 4
        for i in range(1, 5):
            setattr(self, f"gru{i}", nn.GRU(input_size=gd, hidden_size=gd, batch_first=True, bidirectional=bidirectional))
        self.attn = nn.MultiheadAttention(embed_dim=channels, num_heads=4, batch_first=True)
        self.layernorm1 = nn.LayerNorm(channels)
        self.layernorm2 = nn.LayerNorm(channels)
        self.lin = nn.Linear((2 if bidirectional else 1) * channels, channels)


class _Generator(nn.Module):
    def __init__(self, e=(16, 32, 64)):
        super().__init__()
        e1, e2, e3 = e
        enc = dict(kernel_size=(2, 3), stride=(1, 2), padding=(1, 1))
        self.conv1 = nn.Conv2d(1, e1, **enc)
        self.conv2 = nn.Conv2d(e1, e2, **enc)
        self.conv3 = nn.Conv2d(e2, e3, **enc)
        self.skip2 = nn.Conv2d(1, e3, kernel_size=1)
        self.skip3 = nn.Conv2d(1, e2, kernel_size=1)
        self.skip4 = nn.Conv2d(1, e1, kernel_size=1)
        self.GRUf1 = _GRUBlock(True, e3)
        self.GRUt1 = _GRUBlock(False, e3)
        self.GRUf2 = _GRUBlock(True, e3)
        dec = dict(kernel_size=(2, 3), stride=(1, 2), padding=(1, 1), output_padding=(0, 1))
        self.deconv2 = nn.ConvTranspose2d(e3, e2, **dec)
        self.deconv3 = nn.ConvTranspose2d(e2, e1, **dec)
        self.deconv4 = nn.ConvTranspose2d(e1, 1, **dec)
        self.layernorm = nn.LayerNorm(e3)


class _STFT(nn.Module):
    def __init__(self, n_fft=512):
        super().__init__()
        self.register_buffer("window", hann_window(n_fft))


class _Enhancer(nn.Module):
    def __init__(self):
        super().__init__()
        self.gen = _Generator()
        self.stft = _STFT(512)


class _Period(nn.Module):
    def __init__(self):
        super().__init__()
        convs, cin = [], 1
        for cout, k, s, g in MPD_LAYERS:
            convs.append(_wn(nn.Conv2d(cin, cout, kernel_size=(k, 1), stride=(s, 1), padding=(k // 2, 0), groups=g)))
            cin = cout
        self.convs = nn.ModuleList(convs)
        self.conv_post = _wn(nn.Conv2d(cin, 1, kernel_size=(3, 1), stride=(1, 1), padding=(1, 0)))


class _Scale(nn.Module):
    def __init__(self):
        super().__init__()
        convs, cin = [], 1
        for cout, k, s, g in MSD_LAYERS:
            convs.append(_wn(nn.Conv1d(cin, cout, kernel_size=k, stride=s, padding=k // 2, groups=g)))
            cin = cout
        self.convs = nn.ModuleList(convs)
        self.conv_post = _wn(nn.Conv1d(cin, 1, kernel_size=3, stride=1, padding=1))


class _Multi(nn.Module):
    def __init__(self, subs):
        super().__init__()
        self.discriminators = nn.ModuleList(subs)


def init_state_dicts(seed: int = 42) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(enhancer, mpd, msd) ``state_dict``s as the reference's ``set_seed(seed)`` + train.py:569-587 produce them."""
    torch.manual_seed(seed)
    enh = _Enhancer()
    mpd = _Multi([_Period() for _ in MPD_PERIODS])
    msd = _Multi([_Scale() for _ in range(3)])
    cp = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    return cp(enh), cp(mpd), cp(msd)


def param_order(sd: Dict[str, torch.Tensor]) -> List[str]:
    """``named_parameters`` order = ``state_dict`` order without the buffers (the STFT windows)."""
    return [k for k in sd if not k.endswith("window")]
