"""The tcgen05 / TMA dense convolution (MSD convs.5: 1024 -> 1024, k = 5) against torch on the CPU.

bf16 operands, fp32 accumulation.  Two bars:
  * against the same convolution evaluated in fp64 on bf16-ROUNDED operands: 2e-4 relative (only the fp32
    accumulation order differs) - this is the bit-level check of the tensor-core data path;
  * against the fp32 reference on unrounded operands: 2e-2 relative (bf16 has 8 mantissa bits; K = 5120 terms).
"""
import pytest
import torch
import torch.nn.functional as F

from util import cpu_params, leaf_params, oracle, rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.bf16]


def _bf(t):
    return t.to(torch.bfloat16).to(torch.float64)


@pytest.mark.parametrize("B,L,Ci,Co", [(8, 125, 1024, 1024), (2, 63, 1024, 1024), (3, 32, 256, 128), (1, 300, 128, 256)])
def test_dense_conv_fwd_dgrad_wgrad(dev, B, L, Ci, Co):
    from lctgan import ops
    K, pad = 5, 2
    gen = torch.Generator().manual_seed(B * 1000 + L)
    x = torch.randn(B, Ci, L, generator=gen)
    w = torch.randn(Co, Ci, K, generator=gen) / (Ci * K) ** 0.5
    b = torch.randn(Co, generator=gen) * 0.1
    dy = torch.randn(B, Co, L, generator=gen)
    xact = torch.randn(B, Ci, L, generator=gen)
    gextra = torch.randn(B, Ci, L, generator=gen) * 0.1

    xd, wd_, bd = x.to(dev), w.to(dev), b.to(dev)
    wt, wflip = ops.stage_dense_weights(wd_)
    # staging kernels are plain data movement + bf16 rounding: bit exact
    xp = ops.stage_nlc_bf16(xd.unsqueeze(-1), pad)
    ref_xp = F.pad(x.transpose(1, 2), (0, 0, pad, pad)).to(torch.bfloat16)
    assert torch.equal(xp.cpu(), ref_xp)
    assert torch.equal(wt.cpu(), w.permute(2, 0, 1).contiguous().to(torch.bfloat16))
    assert torch.equal(wflip.cpu(), w.flip(2).permute(2, 1, 0).contiguous().to(torch.bfloat16))

    # forward
    y = ops.dense_conv(xp, wt, B, L, Ci, Co, K, bias=bd, act=ops.ACT_LRELU, slope=0.2).squeeze(-1)
    ref_q = F.leaky_relu(F.conv1d(_bf(x), _bf(w), b.double(), padding=pad), 0.2)
    ref_f = F.leaky_relu(F.conv1d(x, w, b, padding=pad), 0.2)
    assert rel_err(y.double(), ref_q) < 2e-4
    assert rel_err(y, ref_f) < 2e-2

    # dgrad with the fused (+ gextra) * lrelu'(xact) epilogue
    dyd = dy.to(dev)
    dx = ops.dense_conv(ops.stage_nlc_bf16(dyd.unsqueeze(-1), pad), wflip, B, L, Co, Ci, K, gextra=gextra.to(dev),
                        xact=xact.to(dev), act=ops.ACT_LRELU, slope=0.2).squeeze(-1)
    base_q = F.conv_transpose1d(_bf(dy), _bf(w), padding=pad)
    ref_q = (base_q + gextra.double()) * torch.where(xact > 0, 1.0, 0.2).double()
    assert rel_err(dx.double(), ref_q) < 2e-4

    # wgrad (+ bias gradient from the staging kernel)
    Lp = L + K - 1
    db = torch.zeros(Co, device=dev)
    dyq = ops.stage_ncl_bf16(dyd.unsqueeze(-1), Lp, 0, rowsum=db)
    xq = ops.stage_ncl_bf16(xd.unsqueeze(-1), Lp, pad, copies=K)
    dw = ops.dense_wgrad(dyq, xq, Co, Ci, K, w.shape)
    xr = _bf(x).requires_grad_(False)
    wr = _bf(w).clone().requires_grad_(True)
    (F.conv1d(xr, wr, None, padding=pad) * _bf(dy)).sum().backward()
    assert rel_err(dw.double(), wr.grad) < 2e-4
    assert rel_err(db, dy.sum((0, 2))) < 1e-4


def test_msd_bf16_matches_oracle(dev):
    """Whole MultiScaleDiscriminator with the tensor-core layer on: forward maps, input gradient and parameter
    gradients against the fp32 CPU oracle at the stated bf16 tolerance (2e-2 relative to each tensor's max)."""
    from models.discriminators import MultiScaleDiscriminator
    O = oracle()
    torch.manual_seed(1)
    msd = MultiScaleDiscriminator()
    P = leaf_params(cpu_params(msd))
    msd = msd.to(dev)
    x = torch.randn(2, 8000, generator=torch.Generator().manual_seed(10)) * 0.1
    xr = x.clone().requires_grad_(True)
    lr, fr = O.msd_forward(P, xr)
    xg = x.to(dev).requires_grad_(True)
    lg, fg = msd(xg)
    loss_r, loss_g = 0.0, 0.0
    gen = torch.Generator().manual_seed(11)
    for i in range(3):
        for a, b in zip(fg[i], fr[i]):
            assert rel_err(a, b) < 2e-2
            gw = torch.randn(b.shape, generator=gen) / b.numel() ** 0.5
            loss_r = loss_r + (b * gw).sum()
            loss_g = loss_g + (a * gw.to(dev)).sum()
    loss_r.backward()
    loss_g.backward()
    assert rel_err(xg.grad, xr.grad) < 2e-2
    # parameter gradients: relative L2 error.  (A max-norm bar is not meaningful here: a bf16-sized change of a
    # pre-activation that sits at ~0 flips LeakyReLU' between 1 and 0.2 for that element, which moves a 64-term
    # bias-gradient sum by several percent - the same happens under torch autocast.)
    for k, p in msd.named_parameters():
        a, b = p.grad.detach().cpu().double(), P[k].grad.double()
        assert ((a - b).norm() / b.norm()).item() < 2e-2, k
