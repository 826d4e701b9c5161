"""TF32 tensor-core grouped convolutions (conv_mma.cu) against torch on the CPU, for every grouped / first-layer
shape of the MPD and MSD stacks.  TF32 operands (10-bit mantissa, round-to-nearest), fp32 accumulation:
tolerance 3e-3 relative to the largest reference value against fp32, and 2e-4 against the same convolution
evaluated in fp64 on TF32-ROUNDED operands (the arithmetic contract)."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.bf16]


@pytest.fixture(autouse=True, params=["default", "tcgen05_period_dgrad"])
def _dgrad_kernel_choice(request):
    """Every case runs with the shipped per-layer kernel choice and with the period discriminators' data gradient forced
    onto the tcgen05 kernel as well (lctgan.config.tc_dgrad_periods: its staged, cp.async-prefetched epilogue)."""
    from lctgan import config
    old = config.tc_dgrad_periods
    config.tc_dgrad_periods = request.param != "default"
    yield
    config.tc_dgrad_periods = old


def _tf32(t):
    """round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32"""
    i = t.contiguous().view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32).double()


CASES = [
    # (Cin, Cout, K, S, G, L, P)
    (1, 32, 5, 3, 1, 600, 2), (32, 128, 5, 3, 4, 200, 3), (128, 512, 5, 3, 16, 67, 5), (512, 1024, 5, 3, 64, 23, 7),
    (1024, 1024, 5, 1, 64, 8, 11), (1, 16, 15, 1, 1, 2000, 1), (16, 64, 41, 4, 4, 2000, 1), (64, 256, 41, 4, 16, 500, 1),
    (256, 1024, 41, 4, 64, 125, 1), (1024, 1024, 41, 4, 256, 32, 1), (16, 64, 41, 4, 4, 8000, 1), (32, 128, 5, 3, 4, 5334, 2),
]


@pytest.mark.parametrize("Cin,Cout,K,S,G,L,P", CASES)
def test_conv_mma(dev, Cin, Cout, K, S, G, L, P):
    from lctgan import ops, _lib
    assert _lib.call_ret("lct_conv_mma_supported", Cin, Cout, G, K, S, P) == 1
    gen = torch.Generator().manual_seed(Cin * 7 + Cout + K + L)
    B = 3
    x = torch.randn(B, Cin, L, P, generator=gen)
    w = torch.randn(Cout, Cin // G, K, generator=gen) / (Cin // G * K) ** 0.5
    b = torch.randn(Cout, generator=gen)
    pad = K // 2
    conv = lambda xx, ww, bb: F.conv2d(xx, ww.unsqueeze(-1), bb, stride=(S, 1), padding=(pad, 0), groups=G)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.leaky_relu(conv(xr, wr, br), 0.2)
    gy = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    # forward
    y = ops.conv1d_fwd(xd, wd, bd, G, S, pad, act=ops.ACT_LRELU, slope=0.2)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < 3e-3
    ref_q = F.leaky_relu(conv(_tf32(x), _tf32(w), b.double()), 0.2)
    assert rel_err(y.double(), ref_q) < 2e-4
    # data gradient with the fused (+ gextra) * lrelu'(xact) epilogue
    dpre = (gy * torch.where(ref > 0, 1.0, 0.2)).detach()
    xact = torch.randn(x.shape, generator=gen)
    gextra = torch.randn(x.shape, generator=gen) * 0.1
    dx = ops.conv1d_dgrad(dpre.to(dev), wd, x.shape, G, S, pad, gextra=gextra.to(dev), xact=xact.to(dev),
                          act=ops.ACT_LRELU, slope=0.2)
    x64 = _tf32(x).requires_grad_(True)
    w64 = _tf32(w).requires_grad_(True)
    (conv(x64, w64, None) * _tf32(dpre)).sum().backward()
    ref_dx = (x64.grad + gextra.double()) * torch.where(xact > 0, 1.0, 0.2).double()
    assert rel_err(dx.double(), ref_dx) < 2e-4
    assert rel_err(dx, (xr.grad + gextra) * torch.where(xact > 0, 1.0, 0.2)) < 3e-3
    # weight + bias gradient
    dw, db = ops.conv1d_wgrad(xd, dpre.to(dev), w.shape, G, S, pad)
    assert rel_err(dw.double(), w64.grad) < 2e-4
    assert rel_err(dw, wr.grad) < 3e-3
    assert rel_err(db, br.grad) < 1e-4


@pytest.mark.parametrize("has_g,has_x", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("Cin,Cout,K,S,G,L,P", [(16, 64, 41, 4, 4, 1037, 1), (128, 512, 5, 3, 16, 131, 7),
                                               (1024, 1024, 5, 1, 64, 3, 11)])
def test_conv_mma_dgrad_epilogue_variants(dev, Cin, Cout, K, S, G, L, P, has_g, has_x):
    """The coalesced data-gradient pass is specialised on which of (FM gradient, saved activation) exist; ragged map
    lengths exercise the short last tile and the rows of the polyphase grid that fall outside the map."""
    from lctgan import ops
    gen = torch.Generator().manual_seed(L * 3 + P)
    B = 2
    pad = K // 2
    Lout = (L + 2 * pad - K) // S + 1
    w = torch.randn(Cout, Cin // G, K, generator=gen) / (Cin // G * K) ** 0.5
    dy = torch.randn(B, Cout, Lout, P, generator=gen)
    xact = torch.randn(B, Cin, L, P, generator=gen)
    gextra = torch.randn(B, Cin, L, P, generator=gen) * 0.1
    x64 = torch.zeros(B, Cin, L, P, dtype=torch.float64, requires_grad=True)
    y64 = F.conv2d(x64, _tf32(w).unsqueeze(-1), None, stride=(S, 1), padding=(pad, 0), groups=G)
    (y64 * _tf32(dy)).sum().backward()
    ref = x64.grad
    if has_g:
        ref = ref + gextra.double()
    if has_x:
        ref = ref * torch.where(xact > 0, 1.0, 0.2).double()
    dx = ops.conv1d_dgrad(dy.to(dev), w.to(dev), (B, Cin, L, P), G, S, pad, gextra=gextra.to(dev) if has_g else None,
                          xact=xact.to(dev) if has_x else None, act=ops.ACT_LRELU, slope=0.2)
    assert dx.shape == ref.shape
    assert rel_err(dx.double(), ref) < 2e-4
