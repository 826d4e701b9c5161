"""GPU parity of the STFT / iSTFT / TF-feature / spectral-loss kernels against the CPU oracle
(oracle/lct_oracle.py, which restates datasets/stft.py, datasets/tf_features.py and losses.py of the
reference).  Tolerance for the floating-point front end: 1e-5 relative to the largest reference
value (BASELINE.json north_star: "STFT/iSTFT within 1e-5 relative in fp32")."""
import pytest
import torch

from util import oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
RES = [(512, 256), (320, 160), (768, 384)]


def _wave(B, T, seed=0, scale=0.1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, T, generator=g) * scale


@pytest.mark.parametrize("n_fft,hop", RES + [(256, 64), (400, 100)])
@pytest.mark.parametrize("T", [4000, 4097])
def test_stft_forward(dev, n_fft, hop, T):
    from lctgan import ops
    O = oracle()
    x = _wave(3, T, seed=1)
    w = O.hann_window(n_fft)
    spec, mag = ops.stft_fwd(x.to(dev), w.to(dev), n_fft, hop, want_mag=True)
    ref = O.stft(x, w, n_fft, hop)                        # [B, F, Tf]
    ref64 = O.stft_explicit(x, w, n_fft, hop)             # float64 definition, no FFT library
    got = ops.spec_view(spec)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    assert rel_err(got, ref64) < TOL
    assert rel_err(ops.spec_view(mag), O.magnitude(ref)) < TOL


@pytest.mark.parametrize("n_fft,hop", RES + [(256, 64)])
@pytest.mark.parametrize("T,length", [(4096, 4096), (4000, 4000), (4000, 3000), (4096, 4300)])
def test_istft_forward(dev, n_fft, hop, T, length):
    from lctgan import ops
    O = oracle()
    x = _wave(2, T, seed=2)
    w = O.hann_window(n_fft)
    ref_spec = O.stft(x, w, n_fft, hop)
    ref_spec = ref_spec * (1.0 + 0.3 * torch.randn(ref_spec.shape, generator=torch.Generator().manual_seed(5)))
    phys = ref_spec.transpose(1, 2).contiguous().to(dev)
    if length <= hop * (ref_spec.shape[-1] - 1) + n_fft // 2:
        ref = O.istft(ref_spec, w, n_fft, hop, length)
    else:
        ref = O.istft_explicit(ref_spec, w, n_fft, hop, length).float()
    got = ops.istft_fwd(phys, w.to(dev), n_fft, hop, length)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL
    ref64 = O.istft_explicit(ref_spec, w, n_fft, hop, length)
    assert rel_err(got, ref64) < TOL


@pytest.mark.parametrize("n_fft,hop", RES)
def test_stft_istft_round_trip(dev, n_fft, hop):
    from lctgan import ops
    O = oracle()
    T = 32000
    x = _wave(4, T, seed=3).to(dev)
    w = O.hann_window(n_fft).to(dev)
    spec, _ = ops.stft_fwd(x, w, n_fft, hop)
    y = ops.istft_fwd(spec, w, n_fft, hop, T)
    keep = hop * (spec.shape[1] - 1)
    assert rel_err(y[:, :keep], x[:, :keep]) < TOL


@pytest.mark.parametrize("n_fft,hop", RES)
@pytest.mark.parametrize("T", [4000, 4096])
def test_stft_backward(dev, n_fft, hop, T):
    import sys
    from datasets.stft import ComplexSTFT, STFTConfig
    O = oracle()
    x = _wave(2, T, seed=4)
    w = O.hann_window(n_fft)
    xr = x.clone().requires_grad_(True)
    ref = O.stft(xr, w, n_fft, hop)
    g = torch.randn(ref.shape, dtype=torch.complex64, generator=torch.Generator().manual_seed(6))
    (torch.view_as_real(ref) * torch.view_as_real(g)).sum().backward()
    mod = ComplexSTFT(STFTConfig(n_fft=n_fft, hop_length=hop)).to(dev)
    xg = x.to(dev).requires_grad_(True)
    out = mod(xg)
    assert rel_err(out, ref) < TOL
    (torch.view_as_real(out.contiguous()) * torch.view_as_real(g.to(dev))).sum().backward()
    assert rel_err(xg.grad, xr.grad) < TOL


@pytest.mark.parametrize("n_fft,hop", RES)
@pytest.mark.parametrize("T,length", [(4096, 4096), (4000, 4000)])
def test_istft_backward(dev, n_fft, hop, T, length):
    from datasets.stft import ComplexSTFT, STFTConfig
    O = oracle()
    w = O.hann_window(n_fft)
    spec0 = O.stft(_wave(2, T, seed=7), w, n_fft, hop).detach()
    sr = spec0.clone().requires_grad_(True)
    ref = O.istft(sr, w, n_fft, hop, length)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(8))
    (ref * g).sum().backward()
    mod = ComplexSTFT(STFTConfig(n_fft=n_fft, hop_length=hop)).to(dev)
    sg = spec0.to(dev).requires_grad_(True)
    out = mod.istft(sg, length=length)
    assert rel_err(out, ref) < TOL
    (out * g.to(dev)).sum().backward()
    # torch's c2r backward leaves an (ignored-input) imaginary gradient of exactly 0 at DC/Nyquist, as do we
    assert rel_err(sg.grad, sr.grad) < TOL


def test_stft_errors(dev):
    from datasets.stft import ComplexSTFT, STFTConfig, magnitude, apply_mask
    mod = ComplexSTFT(STFTConfig()).to(dev)
    with pytest.raises(ValueError):
        mod(torch.zeros(2, 3, 4000, device=dev))
    with pytest.raises(ValueError):
        mod.istft(torch.zeros(2, 257, 10, device=dev))
    with pytest.raises(ValueError):
        mod.istft(torch.zeros(257, 10, dtype=torch.complex64, device=dev))
    with pytest.raises(ValueError):
        ComplexSTFT(STFTConfig(window="hamming"))
    with pytest.raises(ValueError):
        magnitude(torch.zeros(2, 3, device=dev))
    with pytest.raises(ValueError):
        apply_mask(torch.zeros(2, 5, 6, dtype=torch.complex64, device=dev), torch.zeros(2, 2, 5, 6, device=dev))
    with pytest.raises(RuntimeError):
        mod(torch.zeros(2, 4000))   # CPU tensor: no fallback


def test_elementwise_helpers(dev):
    from datasets import stft as S
    O = oracle()
    w = O.hann_window(512)
    a = O.stft(_wave(2, 6000, seed=9), w, 512, 256).detach()
    b = O.stft(_wave(2, 6000, seed=10), w, 512, 256).detach()
    a[0, 3, 4] = 0   # |X| = 0: clamp + sgn(0) paths
    ad, bd = a.to(dev), b.to(dev)
    # magnitude fwd/bwd (power 1 and 2)
    for power in (1.0, 2.0):
        ar = a.clone().requires_grad_(True)
        mr = O.magnitude(ar, power=power)
        g = torch.randn(mr.shape, generator=torch.Generator().manual_seed(11))
        (mr * g).sum().backward()
        ag = ad.clone().requires_grad_(True)
        mg = S.magnitude(ag, power=power)
        assert rel_err(mg, mr) < TOL
        (mg * g.to(dev)).sum().backward()
        assert rel_err(ag.grad, ar.grad) < TOL
    # compress / decompress fwd/bwd
    m = O.magnitude(a)
    for fn_o, fn_g in ((O.compress, S.compress), (O.decompress, S.decompress)):
        xr = (m * 3).clone().requires_grad_(True)
        yr = fn_o(xr)
        g = torch.randn(yr.shape, generator=torch.Generator().manual_seed(12))
        (yr * g).sum().backward()
        xg = (m * 3).to(dev).requires_grad_(True)
        yg = fn_g(xg)
        assert rel_err(yg, yr) < TOL
        (yg * g.to(dev)).sum().backward()
        assert rel_err(xg.grad, xr.grad) < 5e-5
    # compressed IRM
    assert rel_err(S.compute_compressed_irm(ad, bd), O.compressed_irm(a, b)) < TOL
    # apply_mask (plain and compressed), gradient w.r.t. mask and spectrum
    mask = torch.rand(2, 1, 257, a.shape[-1], generator=torch.Generator().manual_seed(13))
    for compressed in (False, True):
        ar = a.clone().requires_grad_(True)
        mk = mask.clone().requires_grad_(True)
        er = O.apply_mask(ar, mk, compressed=compressed)
        g = torch.randn(er.shape, dtype=torch.complex64, generator=torch.Generator().manual_seed(14))
        (torch.view_as_real(er) * torch.view_as_real(g)).sum().backward()
        ag = ad.clone().requires_grad_(True)
        mg = mask.to(dev).requires_grad_(True)
        eg = S.apply_mask(ag, mg, compressed=compressed)
        assert rel_err(eg, er) < TOL
        (torch.view_as_real(eg.contiguous()) * torch.view_as_real(g.to(dev))).sum().backward()
        assert rel_err(mg.grad, mk.grad) < 5e-5
        assert rel_err(ag.grad, ar.grad) < TOL


@pytest.mark.parametrize("compress_input,return_stfts", [(False, False), (True, True)])
def test_tf_features(dev, compress_input, return_stfts):
    from datasets.tf_features import TFFeatures, TFFeaturesConfig
    O = oracle()
    noisy, clean = O.synthetic_batch(3, 8000, seed=21)
    mod = TFFeatures(TFFeaturesConfig(compress_input=compress_input, return_stfts=return_stfts)).to(dev)
    got = mod(noisy.to(dev), clean.to(dev))
    ref = O.tf_features(noisy, clean, O.hann_window(512), compress_input=compress_input, return_stfts=return_stfts)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape, k
        assert rel_err(got[k], ref[k]) < (5e-5 if k == "irm_c" else TOL), k
    with pytest.raises(ValueError):
        mod(noisy.to(dev)[:, :100], clean.to(dev))
    with pytest.raises(ValueError):
        mod(noisy.to(dev)[0], clean.to(dev)[0])


@pytest.mark.parametrize("T", [8000, 16000])
def test_mrstft_loss(dev, T):
    import losses as L
    O = oracle()
    y, yh = O.synthetic_batch(2, T, seed=31)
    wins = [O.hann_window(n) for n in O.MR_FFT_SIZES]
    yr = yh.clone().requires_grad_(True)
    lr, dr = O.mrstft_loss(yr, y, wins)
    (lr * 1.7).backward()
    mod = L.MultiResolutionSTFTLoss().to(dev)
    yg = yh.to(dev).requires_grad_(True)
    lg, dg = mod(yg, y.to(dev))
    assert abs(lg.item() - lr.item()) <= 2e-5 * abs(lr.item())
    for k in dr:
        assert abs(dg[k].item() - dr[k].item()) <= 2e-5 * abs(dr[k].item()), k
    (lg * 1.7).backward()
    assert rel_err(yg.grad, yr.grad) < 5e-5
    with pytest.raises(ValueError):
        mod(yg[0], y.to(dev)[0])


def test_masked_istft_tail(dev):
    """LCTEnhancer's fused tail: apply_mask(compressed) + istft, gradient to the mask."""
    from lctgan import functional as LF, ops
    O = oracle()
    T = 8000
    noisy, _ = O.synthetic_batch(2, T, seed=41)
    w = O.hann_window(512)
    spec = O.stft(noisy, w, 512, 256).detach()
    mask = 0.5 + 0.5 * torch.rand(2, 1, 257, spec.shape[-1], generator=torch.Generator().manual_seed(42))
    mr = mask.clone().requires_grad_(True)
    ref = O.istft(O.apply_mask(spec, mr, compressed=True), w, 512, 256, T)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(43))
    (ref * g).sum().backward()
    phys = spec.transpose(1, 2).contiguous().to(dev)
    mg = mask[:, 0].transpose(1, 2).contiguous().to(dev).requires_grad_(True)
    out = LF.MaskedISTFTFn.apply(phys, mg, w.to(dev), 512, 256, T, 0.3, 1e-12)
    assert rel_err(out, ref) < TOL
    (out * g.to(dev)).sum().backward()
    assert rel_err(mg.grad.transpose(1, 2), mr.grad[:, 0]) < 5e-5
