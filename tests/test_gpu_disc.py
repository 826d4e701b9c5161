"""GPU parity of the discriminator kernels (integer framing bit-exact; convolutions within fp32
accumulation-order tolerance) against the CPU oracle (models/discriminators.py of the reference)."""
import pytest
import torch
import torch.nn.functional as F

from util import cpu_params, leaf_params, oracle, rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.mark.parametrize("T,period", [(32000, 3), (32000, 7), (32000, 11), (1000, 11), (37, 5)])
def test_reflect_pad_bit_exact(dev, T, period):
    from lctgan import functional as LF
    O = oracle()
    x = torch.randn(3, T, generator=torch.Generator().manual_seed(1))
    pad = period - T % period
    ref = F.pad(x.unsqueeze(1), (0, pad), mode="reflect").squeeze(1)
    xg = x.to(dev).requires_grad_(True)
    got = LF.ReflectPadRightFn.apply(xg, pad)
    assert torch.equal(got.cpu(), ref)                       # integer indexing: bit exact
    idx = O.period_indices(T, period).reshape(-1)
    assert torch.equal(got.cpu(), x[:, idx])
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    (got * g.to(dev)).sum().backward()
    xr = x.clone().requires_grad_(True)
    (F.pad(xr.unsqueeze(1), (0, pad), mode="reflect").squeeze(1) * g).sum().backward()
    assert torch.equal(xg.grad.cpu(), xr.grad)


@pytest.mark.parametrize("L", [32000, 16001, 10, 5])
def test_avgpool(dev, L):
    from lctgan import functional as LF
    O = oracle()
    x = torch.randn(2, L, generator=torch.Generator().manual_seed(3))
    xr = x.clone().requires_grad_(True)
    ref = O.msd_pool(xr.unsqueeze(1)).squeeze(1)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(4))
    (ref * g).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    got = LF.AvgPool4Fn.apply(xg)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 1e-6
    (got * g.to(dev)).sum().backward()
    assert rel_err(xg.grad, xr.grad) < 1e-6
    if L == 10:   # the known answer quoted in SURVEY.md section 8a (D4)
        v = LF.AvgPool4Fn.apply(torch.arange(10, dtype=torch.float32, device=dev).view(1, 10)).cpu().view(-1)
        assert torch.equal(v, torch.tensor([0.5, 1.5, 3.5, 5.5, 7.5, 8.5]))


def test_weight_norm(dev):
    from lctgan import ops
    O = oracle()
    g = torch.rand(32, 1, 1, generator=torch.Generator().manual_seed(5)) + 0.5
    v = torch.randn(32, 4, 41, generator=torch.Generator().manual_seed(6))
    gr, vr = g.clone().requires_grad_(True), v.clone().requires_grad_(True)
    ref = O.weight_norm_weight(gr, vr)
    dw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
    (ref * dw).sum().backward()
    w = ops.weight_norm_fwd(g.to(dev), v.to(dev))
    assert rel_err(w, ref) < 1e-6
    dg, dv = ops.weight_norm_bwd(g.to(dev), v.to(dev), dw.to(dev))
    assert rel_err(dg, gr.grad) < TOL
    assert rel_err(dv, vr.grad) < TOL


# (Cin, Cout, K, S, G, L, P): every distinct layer shape of the MPD / MSD stacks (short L) + odd cases
CONV_CASES = [
    (1, 32, 5, 3, 1, 600, 2), (32, 128, 5, 3, 4, 200, 3), (128, 512, 5, 3, 16, 67, 5), (512, 1024, 5, 3, 64, 23, 7),
    (1024, 1024, 5, 1, 64, 8, 11), (1024, 1, 3, 1, 1, 8, 11),
    (1, 16, 15, 1, 1, 2000, 1), (16, 64, 41, 4, 4, 2000, 1), (64, 256, 41, 4, 16, 500, 1),
    (256, 1024, 41, 4, 64, 125, 1), (1024, 1024, 41, 4, 256, 32, 1), (1024, 1024, 5, 1, 1, 8, 1),
    (1024, 1, 3, 1, 1, 8, 1),
    (6, 9, 4, 2, 3, 701, 2), (4, 4, 7, 1, 1, 1300, 1), (8, 8, 3, 3, 2, 5, 1),
    # conv_post shapes beyond one 512-position tile, with few channels (fewer than one per warp), and a 5-tap kernel
    (1024, 1, 3, 1, 1, 198, 2), (1024, 1, 3, 1, 1, 300, 2), (64, 1, 3, 1, 1, 1100, 1), (5, 1, 3, 1, 1, 37, 11),
    (24, 1, 5, 1, 1, 50, 3),
]


@pytest.mark.parametrize("Cin,Cout,K,S,G,L,P", CONV_CASES)
def test_conv1d_fwd_dgrad_wgrad(dev, Cin, Cout, K, S, G, L, P):
    from lctgan import ops
    gen = torch.Generator().manual_seed(Cin * 7 + Cout + K + L)
    B = 2
    x = torch.randn(B, Cin, L, P, generator=gen)
    w = torch.randn(Cout, Cin // G, K, generator=gen) / (Cin // G * K) ** 0.5
    b = torch.randn(Cout, generator=gen)
    pad = K // 2
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    pre = F.conv2d(xr, wr.unsqueeze(-1), br, stride=(S, 1), padding=(pad, 0), groups=G)
    ref = F.leaky_relu(pre, 0.2)
    gy = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()

    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    y = ops.conv1d_fwd(xd, wd, bd, G, S, pad, act=ops.ACT_LRELU, slope=0.2)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL
    dpre = ops.act_bwd(y, gy.to(dev), ops.ACT_LRELU, 0.2)
    dx = ops.conv1d_dgrad(dpre, wd, x.shape, G, S, pad)
    assert rel_err(dx, xr.grad) < TOL
    dw, db = ops.conv1d_wgrad(xd, dpre, w.shape, G, S, pad)
    assert rel_err(dw, wr.grad) < 5e-5
    assert rel_err(db, br.grad) < 5e-5
    # fused epilogue: (dgrad + gextra) * lrelu'(xact)
    xact = torch.randn(x.shape, generator=gen)
    gextra = torch.randn(x.shape, generator=gen)
    dx2 = ops.conv1d_dgrad(dpre, wd, x.shape, G, S, pad, gextra=gextra.to(dev), xact=xact.to(dev),
                           act=ops.ACT_LRELU, slope=0.2)
    ref2 = (xr.grad + gextra) * torch.where(xact > 0, 1.0, 0.2)
    assert rel_err(dx2, ref2) < TOL


def _copy_into(module, P):
    module.load_state_dict(P, strict=True)


@pytest.mark.parametrize("T", [4000, 6001])
def test_mpd_matches_oracle(dev, T):
    from models.discriminators import MultiPeriodDiscriminator
    O = oracle()
    torch.manual_seed(0)
    mpd = MultiPeriodDiscriminator()
    P = leaf_params(cpu_params(mpd))
    mpd = mpd.to(dev)
    x = torch.randn(2, T, generator=torch.Generator().manual_seed(8)) * 0.1
    xr = x.clone().requires_grad_(True)
    lr, fr = O.mpd_forward(P, xr)
    xg = x.to(dev).requires_grad_(True)
    lg, fg = mpd(xg)
    assert len(lg) == 5 and all(len(f) == 6 for f in fg)
    loss_r, loss_g = 0.0, 0.0
    gen = torch.Generator().manual_seed(9)
    for i in range(5):
        assert lg[i] is fg[i][-1]
        for a, b in zip(fg[i], fr[i]):
            assert a.shape == b.shape
            assert rel_err(a, b) < 5e-5
            gw = torch.randn(b.shape, generator=gen) / b.numel() ** 0.5
            loss_r = loss_r + (b * gw).sum()
            loss_g = loss_g + (a * gw.to(dev)).sum()
    loss_r.backward()
    loss_g.backward()
    assert rel_err(xg.grad, xr.grad) < 2e-4
    for k, p in mpd.named_parameters():
        assert rel_err(p.grad, P[k].grad) < 2e-4, k


@pytest.mark.parametrize("T", [8000])
def test_msd_matches_oracle(dev, T):
    from models.discriminators import MultiScaleDiscriminator
    O = oracle()
    torch.manual_seed(1)
    msd = MultiScaleDiscriminator()
    P = leaf_params(cpu_params(msd))
    msd = msd.to(dev)
    x = torch.randn(2, T, generator=torch.Generator().manual_seed(10)) * 0.1
    xr = x.clone().requires_grad_(True)
    lr, fr = O.msd_forward(P, xr)
    xg = x.to(dev).requires_grad_(True)
    lg, fg = msd(xg.unsqueeze(1))
    assert len(lg) == 3 and all(len(f) == 7 for f in fg)
    loss_r, loss_g = 0.0, 0.0
    gen = torch.Generator().manual_seed(11)
    for i in range(3):
        for a, b in zip(fg[i], fr[i]):
            assert a.shape == b.shape
            assert rel_err(a, b) < 5e-5
            gw = torch.randn(b.shape, generator=gen) / b.numel() ** 0.5
            loss_r = loss_r + (b * gw).sum()
            loss_g = loss_g + (a * gw.to(dev)).sum()
    loss_r.backward()
    loss_g.backward()
    assert rel_err(xg.grad, xr.grad) < 2e-4
    for k, p in msd.named_parameters():
        assert rel_err(p.grad, P[k].grad) < 2e-4, k


def test_disc_known_shapes(dev):
    """Shapes quoted in SURVEY.md section 8a/8c for T = 32000: pads 0/1/0/4/10, MSD lengths, 2207 logits/sample."""
    from models.discriminators import MultiPeriodDiscriminator, MultiScaleDiscriminator
    mpd, msd = MultiPeriodDiscriminator().to(dev), MultiScaleDiscriminator().to(dev)
    x = torch.randn(1, 32000, device=dev) * 0.1
    with torch.no_grad():
        pl, pf = mpd(x)
        sl, sf = msd(x)
    assert [f[0].shape[2] for f in pf] == [5334, 3556, 2134, 1524, 970]
    assert [tuple(f[0].shape) for f in sf] == [(1, 16, 32000), (1, 16, 16001), (1, 16, 8001)]
    assert sum(t.numel() for t in pl + sl) == 2207
    assert sum(len(f) for f in pf + sf) == 51
    with pytest.raises(AssertionError):
        mpd(torch.zeros(1, 2, 100, device=dev))
