"""CPU ORACLE for the LCT-GAN adversarial training step.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement of the reference's hot path
(jqshang/LCT-GAN: datasets/stft.py, datasets/tf_features.py, models/generator.py,
models/discriminators.py, losses.py and the loop body of train.py:165-249).  It is
the checker for the CUDA path; it is never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.

Where the arithmetic lives.  The reference owns *composition*; every number is
produced by PyTorch (third party, not under /root/reference, not pinned by the
reference -- no requirements file; this image ships torch 2.11.0+cu128).  The oracle
therefore restates the composition over the same published operators
(``torch.stft``/``torch.istft``, ``conv1d``/``conv2d``/``conv_transpose2d``,
``layer_norm``, the GRU and multi-head-attention equations written out explicitly,
``avg_pool1d``, reflect ``pad``) on plain tensors held in a flat ``dict`` keyed by the
reference's ``state_dict`` names, so that autograd supplies the gradient oracle.
``stft_explicit``/``istft_explicit`` additionally restate the STFT definition in
float64 from first principles (no FFT library) and are used to pin ``torch.stft``.

Pinning.  The reference ships no tests or golden vectors, so
parity is pinned against outputs of the reference itself, executed in the build
container by ``tests/golden/make_golden.py`` (committed) and stored as small
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function here against them.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LRELU = 0.2
MPD_PERIODS = (2, 3, 5, 7, 11)
# (out_channels, kernel, stride, groups) -- discriminators.py:37-44 and :166-174
MPD_LAYERS = ((32, 5, 3, 1), (128, 5, 3, 4), (512, 5, 3, 16), (1024, 5, 3, 64), (1024, 5, 1, 64))
MSD_LAYERS = ((16, 15, 1, 1), (64, 41, 4, 4), (256, 41, 4, 16), (1024, 41, 4, 64),
              (1024, 41, 4, 256), (1024, 5, 1, 1))


# --------------------------------------------------------------------------------------
# Front end: datasets/stft.py
# --------------------------------------------------------------------------------------

def hann_window(n: int) -> torch.Tensor:
    """Periodic Hann as registered by ComplexSTFT.__init__ (stft.py:56-57): torch.hann_window evaluates
    0.5 - 0.5 cos(2 pi k / n) in fp32, and the reference uses exactly that buffer (golden: window_*)."""
    return torch.hann_window(n)


def stft(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """ComplexSTFT.forward (stft.py:59-88): centred, reflect-padded, one-sided, unnormalised."""
    if x.dim() != 2:
        raise ValueError(f"Expected waveform of shape [B, T], got {tuple(x.shape)}")
    return torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=n_fft,
                      window=window.to(x.dtype), center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)


def istft(spec: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int,
          length: Optional[int]) -> torch.Tensor:
    """ComplexSTFT.istft (stft.py:90-132)."""
    if not torch.is_complex(spec):
        raise ValueError("stft_matrix must be a complex tensor.")
    if spec.dim() != 3:
        raise ValueError(f"Expected stft_matrix of shape [B, F, T], got {tuple(spec.shape)}")
    return torch.istft(spec, n_fft=n_fft, hop_length=hop, win_length=n_fft,
                       window=window.to(spec.real.dtype), center=True, normalized=False,
                       onesided=True, length=length)


def stft_explicit(x: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """The definition behind stft.py:75-86, in float64 without an FFT:
    X[b,k,m] = sum_n w[n] * xp[b, m*hop+n] * exp(-2*pi*i*k*n/N), xp = reflect-pad(x, N/2)."""
    xd = x.to(torch.float64)
    half = n_fft // 2
    xp = F.pad(xd.unsqueeze(1), (half, half), mode="reflect").squeeze(1)
    frames = xp.unfold(-1, n_fft, hop)                       # [B, Tf, N]
    frames = frames * window.to(torch.float64)
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(half + 1, dtype=torch.float64)
    ang = -2.0 * math.pi * torch.outer(k, n) / n_fft          # [F, N]
    basis = torch.complex(torch.cos(ang), torch.sin(ang))
    out = torch.einsum("btn,fn->bft", frames.to(torch.complex128), basis)
    return out


def istft_explicit(spec: torch.Tensor, window: torch.Tensor, n_fft: int, hop: int,
                   length: int) -> torch.Tensor:
    """The definition behind stft.py:120-130, float64, no FFT:
    y = OLA(w * irfft(S)) / OLA(w^2), then keep [N/2 : N/2+length] (zero-fill if short)."""
    s = spec.to(torch.complex128)
    b, f, tf = s.shape
    half = n_fft // 2
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(half + 1, dtype=torch.float64)
    ang = 2.0 * math.pi * torch.outer(n, k) / n_fft           # [N, F]
    cosb, sinb = torch.cos(ang), torch.sin(ang)
    wgt = torch.full((half + 1,), 2.0, dtype=torch.float64)
    wgt[0] = 1.0
    wgt[-1] = 1.0
    re = s.real * wgt[None, :, None]
    im = s.imag * wgt[None, :, None]
    im[:, 0] = 0.0
    im[:, -1] = 0.0
    frames = (torch.einsum("nf,bft->btn", cosb, re) - torch.einsum("nf,bft->btn", sinb, im)) / n_fft
    w = window.to(torch.float64)
    frames = frames * w
    total = n_fft + hop * (tf - 1)
    y = torch.zeros(b, total, dtype=torch.float64)
    env = torch.zeros(total, dtype=torch.float64)
    for m in range(tf):
        y[:, m * hop:m * hop + n_fft] += frames[:, m]
        env[m * hop:m * hop + n_fft] += w * w
    y = y[:, half:]
    env = env[half:]
    keep = min(length, y.shape[1])
    out = torch.zeros(b, length, dtype=torch.float64)
    out[:, :keep] = y[:, :keep] / env[:keep]
    return out


def magnitude(spec: torch.Tensor, power: float = 1.0, eps: float = 1e-12) -> torch.Tensor:
    """stft.py:138-160."""
    if not torch.is_complex(spec):
        raise ValueError("stft_matrix must be a complex tensor.")
    m = spec.abs().clamp_min(eps)
    return m if power == 1.0 else m ** power


def compress(x: torch.Tensor, c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    """stft.py:163-169."""
    return x.clamp_min(eps) ** c


def decompress(x: torch.Tensor, c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    """stft.py:172-178 (and decompress_mask :221-240)."""
    return x.clamp_min(eps) ** (1.0 / c)


def compressed_irm(clean_spec: torch.Tensor, noisy_spec: torch.Tensor, c: float = 0.3,
                   gamma: float = 1e-12, eps: float = 1e-12) -> torch.Tensor:
    """compute_compressed_irm, stft.py:184-218."""
    if not (torch.is_complex(clean_spec) and torch.is_complex(noisy_spec)):
        raise ValueError("clean_stft and noisy_stft must be complex tensors.")
    num = clean_spec.abs().clamp_min(eps) ** c
    den = noisy_spec.abs().clamp_min(eps) ** c + gamma
    return num / den


def apply_mask(noisy_spec: torch.Tensor, mask: torch.Tensor, compressed: bool = False,
               c: float = 0.3, eps: float = 1e-12) -> torch.Tensor:
    """stft.py:243-290."""
    if not torch.is_complex(noisy_spec):
        raise ValueError("noisy_stft must be a complex tensor.")
    if mask.dim() == 4:
        if mask.size(1) != 1:
            raise ValueError(f"Expected mask shape [B, 1, F, T], got {tuple(mask.shape)}")
        mask = mask[:, 0]
    if mask.dim() != 3:
        raise ValueError(f"Expected mask shape [B, F, T] (or [B, 1, F, T]), got {tuple(mask.shape)}")
    if compressed:
        mask = decompress(mask, c=c, eps=eps)
    mask = mask.clamp_min(0.0)
    return noisy_spec * mask.to(noisy_spec.dtype)


def tf_features(noisy: torch.Tensor, clean: torch.Tensor, window: torch.Tensor,
                n_fft: int = 512, hop: int = 256, c: float = 0.3,
                compress_input: bool = False, return_stfts: bool = True) -> Dict[str, torch.Tensor]:
    """TFFeatures.forward, tf_features.py:85-146."""
    if noisy.dim() != 2 or clean.dim() != 2:
        raise ValueError("Expected noisy_wave and clean_wave of shape [B, T]")
    if noisy.shape != clean.shape:
        raise ValueError("noisy_wave and clean_wave must have same shape")
    ns = stft(noisy, window, n_fft, hop)
    cs = stft(clean, window, n_fft, hop)
    nmag = magnitude(ns)
    irm = compressed_irm(cs, ns, c=c)
    nmag_c = compress(nmag, c=c)
    out = {"noisy_mag": nmag_c if compress_input else nmag, "irm_c": irm, "noisy_mag_c": nmag_c}
    if return_stfts:
        out["noisy_stft"] = ns
        out["clean_stft"] = cs
    return out


# --------------------------------------------------------------------------------------
# Generator: models/generator.py
# --------------------------------------------------------------------------------------

def gru_direction(x: torch.Tensor, w_ih: torch.Tensor, w_hh: torch.Tensor, b_ih: torch.Tensor,
                  b_hh: torch.Tensor, reverse: bool = False) -> torch.Tensor:
    """One direction of a single-layer batch-first GRU with h0 = 0, written out
    (torch.nn.GRU documentation; gate order r, z, n; b_hn sits inside r * (...)).
    x: [S, L, I] -> [S, L, H].  Backs generator.py:104 and :219."""
    s, l, _ = x.shape
    hdim = w_hh.shape[1]
    gi_all = F.linear(x, w_ih, b_ih)                          # [S, L, 3H]
    h = x.new_zeros(s, hdim)
    outs: List[torch.Tensor] = [None] * l                    # type: ignore[list-item]
    order = range(l - 1, -1, -1) if reverse else range(l)
    for t in order:
        gi = gi_all[:, t]
        gh = F.linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, dim=-1)
        h_r, h_z, h_n = gh.chunk(3, dim=-1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1.0 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def gru_direction_aten(x, w_ih, w_hh, b_ih, b_hh, reverse=False):
    """Same quantity through aten::gru (the operator nn.GRU dispatches to); used when the
    oracle is *timed* as the CPU baseline so the per-step cost is the library's."""
    if reverse:
        x = x.flip(1)
    h0 = x.new_zeros(1, x.shape[0], w_hh.shape[1])
    # (train flag: no dropout here, so it changes nothing on the CPU; cuDNN's GRU only records what its backward needs
    # in training mode - the oracle also runs on stock CUDA operators as bench.py's `--impl reference-cuda` comparator)
    y, _ = torch._VF.gru(x, h0, [w_ih, w_hh, b_ih, b_hh], True, 1, 0.0, torch.is_grad_enabled(), False, True)
    return y.flip(1) if reverse else y


def multihead_self_attention(x: torch.Tensor, in_w: torch.Tensor, in_b: torch.Tensor,
                             out_w: torch.Tensor, out_b: torch.Tensor, heads: int = 4) -> torch.Tensor:
    """nn.MultiheadAttention(64, 4, batch_first=True)(x, x, x)[0] written out
    (generator.py:133, :245): softmax(q k^T / sqrt(d)) v per head, no mask, no dropout."""
    s, l, e = x.shape
    d = e // heads
    qkv = F.linear(x, in_w, in_b)
    q, k, v = qkv.chunk(3, dim=-1)

    def split(t):
        return t.reshape(s, l, heads, d).permute(0, 2, 1, 3)   # [S, H, L, d]

    q, k, v = split(q), split(k), split(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(s, l, e)
    return F.linear(o, out_w, out_b)


def _gru_bank(P: Params, pre: str, seq: torch.Tensor, bidirectional: bool, aten: bool) -> torch.Tensor:
    run = gru_direction_aten if aten else gru_direction
    outs = []
    for gi, chunk in enumerate(torch.chunk(seq, 4, dim=-1), 1):
        g = f"{pre}gru{gi}."
        y = run(chunk, P[g + "weight_ih_l0"], P[g + "weight_hh_l0"], P[g + "bias_ih_l0"],
                P[g + "bias_hh_l0"], False)
        if bidirectional:
            y = y + run(chunk, P[g + "weight_ih_l0_reverse"], P[g + "weight_hh_l0_reverse"],
                        P[g + "bias_ih_l0_reverse"], P[g + "bias_hh_l0_reverse"], True)
        outs.append(y)
    return torch.cat(outs, dim=-1)


def gru_block_f(P: Params, pre: str, x: torch.Tensor, aten_gru: bool = False) -> torch.Tensor:
    """GRUblockf.forward, generator.py:113-145.  x: [B, 64, T, F]."""
    b, c, t, f = x.shape
    seq = x.permute(0, 2, 3, 1).reshape(b * t, f, c)
    g = _gru_bank(P, pre, F.layer_norm(seq, (c,), P[pre + "layernorm1.weight"], P[pre + "layernorm1.bias"]),
                  True, aten_gru)
    seq = seq + g
    a = multihead_self_attention(
        F.layer_norm(seq, (c,), P[pre + "layernorm2.weight"], P[pre + "layernorm2.bias"]),
        P[pre + "attn.in_proj_weight"], P[pre + "attn.in_proj_bias"],
        P[pre + "attn.out_proj.weight"], P[pre + "attn.out_proj.bias"])
    mix = F.leaky_relu(F.linear(torch.cat([g, a], dim=-1), P[pre + "lin.weight"], P[pre + "lin.bias"]), LRELU)
    seq = seq + mix
    return seq.reshape(b, t, f, c).permute(0, 3, 1, 2)


def gru_block_t(P: Params, pre: str, x: torch.Tensor, aten_gru: bool = False) -> torch.Tensor:
    """GRUblockt.forward, generator.py:225-255.  x: [B, 64, T, F]."""
    b, c, t, f = x.shape
    seq = x.permute(0, 3, 2, 1).reshape(b * f, t, c)
    g = _gru_bank(P, pre, F.layer_norm(seq, (c,), P[pre + "layernorm1.weight"], P[pre + "layernorm1.bias"]),
                  False, aten_gru)
    seq = seq + g
    a = multihead_self_attention(
        F.layer_norm(seq, (c,), P[pre + "layernorm2.weight"], P[pre + "layernorm2.bias"]),
        P[pre + "attn.in_proj_weight"], P[pre + "attn.in_proj_bias"],
        P[pre + "attn.out_proj.weight"], P[pre + "attn.out_proj.bias"])
    mix = F.leaky_relu(F.linear(a, P[pre + "lin.weight"], P[pre + "lin.bias"]), LRELU)
    seq = seq + mix
    return seq.reshape(b, f, t, c).permute(0, 3, 2, 1)


def _crop_pair(a: torch.Tensor, b: torch.Tensor):
    """LCTGenerator._align, generator.py:538-548: crop both to the common low-index corner."""
    t = min(a.shape[2], b.shape[2])
    f = min(a.shape[3], b.shape[3])
    return a[:, :, :t, :f], b[:, :, :t, :f]


def generator_forward(P: Params, mag: torch.Tensor, pre: str = "", output_activation: str = "sigmoid",
                      aten_gru: bool = False) -> torch.Tensor:
    """LCTGenerator.forward, generator.py:550-632.  mag: [B,1,F,T] -> [B,1,F,T]."""
    if mag.dim() != 4 or mag.size(1) != 1:
        raise ValueError(f"Expected noisy_mag [B, 1, F, T], got {tuple(mag.shape)}")
    f_in, t_in = mag.shape[2], mag.shape[3]
    x = mag.transpose(2, 3)
    sk = {n: F.conv2d(x, P[f"{pre}skip{n}.weight"], P[f"{pre}skip{n}.bias"]) for n in (2, 3, 4)}
    h = x
    for n in (1, 2, 3):
        h = F.leaky_relu(F.conv2d(h, P[f"{pre}conv{n}.weight"], P[f"{pre}conv{n}.bias"],
                                  stride=(1, 2), padding=(1, 1)), LRELU)
    h = F.layer_norm(h.permute(0, 2, 3, 1), (h.shape[1],), P[pre + "layernorm.weight"],
                     P[pre + "layernorm.bias"]).permute(0, 3, 1, 2)
    h = gru_block_f(P, pre + "GRUf1.", h, aten_gru)
    h = gru_block_t(P, pre + "GRUt1.", h, aten_gru)
    h = gru_block_f(P, pre + "GRUf2.", h, aten_gru)
    for n, last in ((2, False), (3, False), (4, True)):
        s, h = _crop_pair(sk[n], h)
        h = F.conv_transpose2d(h + s, P[f"{pre}deconv{n}.weight"], P[f"{pre}deconv{n}.bias"],
                               stride=(1, 2), padding=(1, 1), output_padding=(0, 1))
        h = F.relu(h) if last else F.leaky_relu(h, LRELU)
    h = h[:, :, :t_in, :f_in]
    if h.shape[2] < t_in or h.shape[3] < f_in:
        h = F.pad(h, (0, f_in - h.shape[3], 0, t_in - h.shape[2]))
    out = h.transpose(2, 3)
    if output_activation == "sigmoid":
        out = torch.sigmoid(out)
    return out


def enhancer_forward(P: Params, noisy: torch.Tensor, c: float = 0.3, n_fft: int = 512, hop: int = 256,
                     aten_gru: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """LCTEnhancer.forward, generator.py:659-697."""
    if noisy.dim() != 2:
        raise ValueError(f"Expected noisy_wave [B, T], got {tuple(noisy.shape)}")
    w = P["stft.window"]
    spec = stft(noisy, w, n_fft, hop)
    mask_c = generator_forward(P, magnitude(spec).unsqueeze(1), pre="gen.", aten_gru=aten_gru)
    enh = istft(apply_mask(spec, mask_c, compressed=True, c=c), w, n_fft, hop, noisy.shape[-1])
    return enh, mask_c


# --------------------------------------------------------------------------------------
# Discriminators: models/discriminators.py
# --------------------------------------------------------------------------------------

def weight_norm_weight(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but 0."""
    dims = tuple(range(1, v.dim()))
    return v * (g / v.norm(2, dim=dims, keepdim=True))


def period_indices(t: int, period: int) -> torch.Tensor:
    """Integer framing of PeriodDiscriminator (discriminators.py:84-91): source index of every
    element of the [T'/P, P] grid after right reflect-padding to a multiple of P."""
    pad = (period - t % period) % period
    idx = torch.arange(t + pad)
    over = idx >= t
    idx = torch.where(over, 2 * (t - 1) - idx, idx)
    return idx.view((t + pad) // period, period)


def period_disc_forward(P: Params, pre: str, x: torch.Tensor, period: int):
    """PeriodDiscriminator.forward, discriminators.py:69-103."""
    if x.dim() == 2:
        x = x.unsqueeze(1)
    b, c, t = x.shape
    assert c == 1
    if t % period:
        x = F.pad(x, (0, period - t % period), mode="reflect")
    x = x.view(b, 1, -1, period)
    fmaps = []
    for i, (_, k, s, g) in enumerate(MPD_LAYERS):
        w = weight_norm_weight(P[f"{pre}convs.{i}.weight_g"], P[f"{pre}convs.{i}.weight_v"])
        x = F.leaky_relu(F.conv2d(x, w, P[f"{pre}convs.{i}.bias"], stride=(s, 1), padding=(k // 2, 0), groups=g), LRELU)
        fmaps.append(x)
    w = weight_norm_weight(P[pre + "conv_post.weight_g"], P[pre + "conv_post.weight_v"])
    x = F.conv2d(x, w, P[pre + "conv_post.bias"], padding=(1, 0))
    fmaps.append(x)
    return x, fmaps


def mpd_forward(P: Params, x: torch.Tensor, pre: str = ""):
    """MultiPeriodDiscriminator.forward, discriminators.py:125-147."""
    logits, fmaps = [], []
    for i, p in enumerate(MPD_PERIODS):
        lg, fm = period_disc_forward(P, f"{pre}discriminators.{i}.", x, p)
        logits.append(lg)
        fmaps.append(fm)
    return logits, fmaps


def scale_disc_forward(P: Params, pre: str, x: torch.Tensor):
    """ScaleDiscriminator.forward, discriminators.py:199-224."""
    if x.dim() == 2:
        x = x.unsqueeze(1)
    assert x.shape[1] == 1
    fmaps = []
    for i, (_, k, s, g) in enumerate(MSD_LAYERS):
        w = weight_norm_weight(P[f"{pre}convs.{i}.weight_g"], P[f"{pre}convs.{i}.weight_v"])
        x = F.leaky_relu(F.conv1d(x, w, P[f"{pre}convs.{i}.bias"], stride=s, padding=k // 2, groups=g), LRELU)
        fmaps.append(x)
    w = weight_norm_weight(P[pre + "conv_post.weight_g"], P[pre + "conv_post.weight_v"])
    x = F.conv1d(x, w, P[pre + "conv_post.bias"], padding=1)
    fmaps.append(x)
    return x, fmaps


def msd_pool(x: torch.Tensor) -> torch.Tensor:
    """AvgPool1d(4, 2, padding=2, count_include_pad=False), discriminators.py:252-255."""
    return F.avg_pool1d(x, 4, 2, 2, ceil_mode=False, count_include_pad=False)


def msd_forward(P: Params, x: torch.Tensor, pre: str = "", num_scales: int = 3):
    """MultiScaleDiscriminator.forward, discriminators.py:257-286."""
    if x.dim() == 2:
        x = x.unsqueeze(1)
    logits, fmaps = [], []
    for i in range(num_scales):
        lg, fm = scale_disc_forward(P, f"{pre}discriminators.{i}.", x)
        logits.append(lg)
        fmaps.append(fm)
        x = msd_pool(x)
    return logits, fmaps


# --------------------------------------------------------------------------------------
# Losses: losses.py
# --------------------------------------------------------------------------------------

MR_FFT_SIZES = (320, 512, 768)


def mrstft_loss(y_hat: torch.Tensor, y: torch.Tensor, windows: Sequence[torch.Tensor],
                fft_sizes: Sequence[int] = MR_FFT_SIZES, hop_factors: Sequence[float] = (0.5, 0.5, 0.5),
                mag_weight: float = 1.0, complex_weight: float = 1.0, main_fft_size: int = 512,
                main_fft_weight: float = 2.0, default_weight: float = 1.0):
    """MultiResolutionSTFTLoss.forward, losses.py:54-100."""
    if y_hat.dim() != 2 or y.dim() != 2:
        raise ValueError("Expected y_hat, y of shape [B, T]")
    total = mag_t = cplx_t = 0.0
    wsum = 0.0
    for n_fft, hf, win in zip(fft_sizes, hop_factors, windows):
        hop = int(round(n_fft * hf))
        w = main_fft_weight if n_fft == main_fft_size else default_weight
        a = stft(y_hat, win, n_fft, hop)
        b = stft(y, win, n_fft, hop)
        ml = F.mse_loss(magnitude(a), magnitude(b))
        d = a - b
        cl = (d.real.pow(2) + d.imag.pow(2)).mean()
        total = total + w * (mag_weight * ml + complex_weight * cl)
        mag_t = mag_t + w * ml
        cplx_t = cplx_t + w * cl
        wsum += w
    if wsum > 0:
        total, mag_t, cplx_t = total / wsum, mag_t / wsum, cplx_t / wsum
    return total, {"mrstft_total": total.detach(), "mrstft_mag": mag_t.detach(),
                   "mrstft_complex": cplx_t.detach()}


def discriminator_loss(real_logits, fake_logits, loss_type: str = "ls"):
    """losses.py:110-135."""
    if len(real_logits) != len(fake_logits):
        raise ValueError("real_logits and fake_logits must have the same length.")
    loss = 0.0
    for r, f in zip(real_logits, fake_logits):
        if loss_type == "ls":
            loss = loss + ((r - 1.0) ** 2).mean() + (f ** 2).mean()
        elif loss_type == "hinge":
            loss = loss + F.relu(1.0 - r).mean() + F.relu(1.0 + f).mean()
        else:
            raise ValueError(f"Unknown loss_type: {loss_type}")
    return loss / max(len(real_logits), 1)


def generator_adv_loss(fake_logits, loss_type: str = "ls"):
    """losses.py:138-151."""
    loss = 0.0
    for f in fake_logits:
        if loss_type == "ls":
            loss = loss + ((f - 1.0) ** 2).mean()
        elif loss_type == "hinge":
            loss = loss - f.mean()
        else:
            raise ValueError(f"Unknown loss_type: {loss_type}")
    return loss / max(len(fake_logits), 1)


def feature_matching_loss(real_fmaps, fake_fmaps):
    """losses.py:154-173."""
    if len(real_fmaps) != len(fake_fmaps):
        raise ValueError("real_fmaps and fake_fmaps must have the same outer length.")
    loss, count = 0.0, 0
    for rs, fs in zip(real_fmaps, fake_fmaps):
        if len(rs) != len(fs):
            raise ValueError("Mismatched feature map list lengths for a discriminator.")
        for r, f in zip(rs, fs):
            loss = loss + (f - r).abs().mean()
            count += 1
    if count == 0:
        return torch.tensor(0.0)
    return loss / count


def mask_mse_loss(pred: torch.Tensor, target: torch.Tensor):
    """losses.py:176-181."""
    if pred.shape != target.shape:
        raise ValueError(f"Shape mismatch: pred_mask_c {tuple(pred.shape)} vs target_mask_c {tuple(target.shape)}")
    return ((pred - target) ** 2).mean()


# --------------------------------------------------------------------------------------
# The training step: train.py:165-249
# --------------------------------------------------------------------------------------

class StepState:
    """Parameters of the three networks as flat dicts of leaf tensors + the two AdamW
    optimisers (train.py:601-610: lr 2e-4, betas (0.8, 0.99), torch defaults otherwise)."""

    def __init__(self, enhancer: Params, mpd: Params, msd: Params, lr_g: float = 2e-4, lr_d: float = 2e-4,
                 betas=(0.8, 0.99), buffers=("stft.window",), order_g=None, order_d=None):
        def leafify(d):
            out = {}
            for k, v in d.items():
                t = v.detach().clone()
                if k not in buffers:
                    t.requires_grad_(True)
                out[k] = t
            return out

        self.enh, self.mpd, self.msd = leafify(enhancer), leafify(mpd), leafify(msd)
        gp = [self.enh[k] for k in (order_g or self.enh) if k not in buffers]
        dp = [self.mpd[k] for k in (order_d[0] if order_d else self.mpd)] + \
             [self.msd[k] for k in (order_d[1] if order_d else self.msd)]
        self.g_params, self.d_params = gp, dp
        self.g_opt = torch.optim.AdamW(gp, lr=lr_g, betas=tuple(betas))
        self.d_opt = torch.optim.AdamW(dp, lr=lr_d, betas=tuple(betas))


def train_step(st: StepState, noisy: torch.Tensor, clean: torch.Tensor, mr_windows, gan_loss: str = "ls",
               lambda_fm: float = 1.0, lambda_mask: float = 1.0, lambda_adv: float = 1e-2,
               grad_clip: float = 5.0, c: float = 0.3, aten_gru: bool = False) -> Dict[str, float]:
    """One iteration of the loop body of train_one_epoch (train.py:165-249)."""
    feats = tf_features(noisy, clean, st.enh["stft.window"], c=c, return_stfts=False)
    irm_c = feats["irm_c"]
    # ---- D step (train.py:177-200)
    st.d_opt.zero_grad(set_to_none=True)
    with torch.no_grad():
        fake_d, _ = enhancer_forward(st.enh, noisy, c=c, aten_gru=aten_gru)
    pr, _ = mpd_forward(st.mpd, clean)
    pf, _ = mpd_forward(st.mpd, fake_d)
    sr, _ = msd_forward(st.msd, clean)
    sf, _ = msd_forward(st.msd, fake_d)
    d_loss = discriminator_loss(list(pr) + list(sr), list(pf) + list(sf), gan_loss)
    d_loss.backward()
    st.d_opt.step()
    # ---- G step (train.py:205-249)
    st.g_opt.zero_grad(set_to_none=True)
    enh, mask_c = enhancer_forward(st.enh, noisy, c=c, aten_gru=aten_gru)
    mr, _ = mrstft_loss(enh, clean, mr_windows)
    pred = mask_c[:, 0]
    tmin = min(irm_c.shape[-1], pred.shape[-1])                # _align_tf_targets, train.py:388-413
    m_loss = mask_mse_loss(pred[..., :tmin], irm_c[..., :tmin])
    pfl, pff = mpd_forward(st.mpd, enh)
    sfl, sff = msd_forward(st.msd, enh)
    with torch.no_grad():
        _, prf = mpd_forward(st.mpd, clean)
        _, srf = msd_forward(st.msd, clean)
    adv = generator_adv_loss(list(pfl) + list(sfl), gan_loss)
    fm = feature_matching_loss(prf + srf, pff + sff)
    g_loss = mr + lambda_mask * m_loss + lambda_adv * (adv + lambda_fm * fm)
    g_loss.backward()
    gnorm = None
    if grad_clip > 0.0:
        gnorm = torch.nn.utils.clip_grad_norm_(st.g_params, grad_clip)
    st.g_opt.step()
    return {"d_loss": float(d_loss), "g_loss": float(g_loss), "mr": float(mr), "mask": float(m_loss),
            "adv": float(adv), "fm": float(fm), "g_grad_norm": float(gnorm) if gnorm is not None else float("nan")}


# --------------------------------------------------------------------------------------
# Callers on either side of the path
# --------------------------------------------------------------------------------------

def si_sdr(reference: torch.Tensor, estimate: torch.Tensor, eps: float = 1e-8) -> float:
    """train.py:261-282 `_si_sdr_torch` for one (reference, estimate) pair of 1-D tensors."""
    n = min(reference.shape[-1], estimate.shape[-1])
    reference, estimate = reference[..., :n], estimate[..., :n]
    reference = reference - reference.mean()
    estimate = estimate - estimate.mean()
    scale = torch.sum(reference * estimate) / (torch.sum(reference ** 2) + eps)
    s_target = scale * reference
    e_noise = estimate - s_target
    return float(10.0 * torch.log10((torch.sum(s_target ** 2) + eps) / (torch.sum(e_noise ** 2) + eps)))


def crop_pair(noisy: torch.Tensor, clean: torch.Tensor, segment_length: Optional[int], random_segment: bool,
              generator: Optional[torch.Generator] = None):
    """datasets/datasets.py:131-156 `LCTScpDataset._crop_pair`: same start for both waveforms; items no longer than the
    segment are returned whole (the reference draws from the global generator; `generator` makes the draw explicit)."""
    if segment_length is None:
        return noisy, clean
    m = min(noisy.shape[-1], clean.shape[-1])
    if m <= segment_length:
        return noisy, clean
    max_start = m - segment_length
    start = int(torch.randint(low=0, high=max_start + 1, size=(1,), generator=generator).item()) if random_segment \
        else max_start // 2
    return noisy[..., start:start + segment_length], clean[..., start:start + segment_length]


def collate(items):
    """datasets/datasets.py:187-230 `collate_fn` on a list of (noisy, clean) pairs: zero padding to the longest waveform
    of the batch (either side), "lengths" = the noisy lengths."""
    ln = torch.tensor([n.shape[-1] for n, _ in items], dtype=torch.long)
    lc = torch.tensor([c.shape[-1] for _, c in items], dtype=torch.long)
    T = int(max(ln.max(), lc.max()))
    pn, pc = torch.zeros(len(items), T), torch.zeros(len(items), T)
    for i, (n, c) in enumerate(items):
        pn[i, :n.shape[-1]] = n
        pc[i, :c.shape[-1]] = c
    return {"noisy": pn, "clean": pc, "lengths": ln}


def synthetic_batch(batch: int, samples: int, seed: int = 1234):
    """clean ~ N(0, 0.1^2), noisy = clean + N(0, 0.05^2), CPU generator."""
    g = torch.Generator().manual_seed(seed)
    clean = torch.randn(batch, samples, generator=g) * 0.1
    noisy = clean + torch.randn(batch, samples, generator=g) * 0.05
    return noisy, clean
