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


def _tf32(t):
    """round-to-nearest to 10 mantissa bits (cvt.rna.tf32.f32)"""
    i = t.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


class _QuantConv(torch.autograd.Function):
    """CPU emulation of a tensor-core layer's arithmetic contract: every GEMM operand (x, w, dy) is rounded by `q`
    (bf16 for the tcgen05 dense layer, TF32 for the mma.sync grouped layers), products are accumulated exactly
    (fp64 here, fp32 on the GPU); bias and its gradient stay fp32."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, groups, q):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, pad, groups, q)
        return (F.conv1d(q(x), q(w), None, stride=stride, padding=pad, groups=groups) +
                b.double()[None, :, None]).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, pad, groups, q = ctx.cfg
        gq = q(g)
        opad = x.shape[2] - ((g.shape[2] - 1) * stride - 2 * pad + w.shape[2])
        dx = F.conv_transpose1d(gq, q(w), stride=stride, padding=pad, output_padding=opad, groups=groups)
        dw = torch.nn.grad.conv1d_weight(q(x), w.shape, gq, stride=stride, padding=pad, groups=groups)
        return dx.to(x.dtype), dw.to(w.dtype), g.sum((0, 2)), None, None, None, None


def _msd_forward_emulated(O, P, x):
    """oracle.msd_forward with the tensor-core layers replaced by the operand-rounding emulation above
    (convs.0-4: TF32, convs.5: bf16; conv_post stays fp32 like on the GPU)."""
    x = x.unsqueeze(1)
    logits, fmaps = [], []
    for i in range(3):
        pre = f"discriminators.{i}."
        h, fm = x, []
        for j, (_, k, s, g) in enumerate(O.MSD_LAYERS):
            w = O.weight_norm_weight(P[f"{pre}convs.{j}.weight_g"], P[f"{pre}convs.{j}.weight_v"])
            b = P[f"{pre}convs.{j}.bias"]
            h = F.leaky_relu(_QuantConv.apply(h, w, b, s, k // 2, g, _bf if j == 5 else _tf32), 0.2)
            fm.append(h)
        w = O.weight_norm_weight(P[pre + "conv_post.weight_g"], P[pre + "conv_post.weight_v"])
        h = F.conv1d(h, w, P[pre + "conv_post.bias"], padding=1)
        fm.append(h)
        logits.append(h)
        fmaps.append(fm)
        x = O.msd_pool(x)
    return logits, fmaps


def test_msd_bf16_matches_oracle(dev):
    """Whole MultiScaleDiscriminator with the tensor-core layer on.
    (1) against the oracle with the tensor-core layers' operands rounded the same way (TF32 for convs.0-4, bf16 for
        convs.5: same arithmetic contract): 2e-3 relative L2
        on every feature map, the input gradient and every parameter gradient;
    (2) against the plain fp32 oracle: 2e-2 relative to the max on the feature maps and the input gradient
        (bf16 has 8 mantissa bits; stated tolerance of the bf16 training configuration)."""
    from models.discriminators import MultiScaleDiscriminator
    O = oracle()
    torch.manual_seed(1)
    msd = MultiScaleDiscriminator()
    P = leaf_params(cpu_params(msd))
    P32 = leaf_params(cpu_params(msd))
    msd = msd.to(dev)
    x = torch.randn(2, 8000, generator=torch.Generator().manual_seed(10)) * 0.1
    xr = x.clone().requires_grad_(True)
    x32 = x.clone().requires_grad_(True)
    lr, fr = _msd_forward_emulated(O, P, xr)
    l32, f32 = O.msd_forward(P32, x32)
    xg = x.to(dev).requires_grad_(True)
    lg, fg = msd(xg)
    loss_r, loss_g, loss_32 = 0.0, 0.0, 0.0
    gen = torch.Generator().manual_seed(11)
    l2 = lambda a, b: ((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm()).item()
    for i in range(3):
        for a, b, c in zip(fg[i], fr[i], f32[i]):
            assert l2(a, b) < 2e-3
            assert rel_err(a, c) < 2e-2
            gw = torch.randn(b.shape, generator=gen) / b.numel() ** 0.5
            loss_r = loss_r + (b * gw).sum()
            loss_32 = loss_32 + (c * gw).sum()
            loss_g = loss_g + (a * gw.to(dev)).sum()
    loss_r.backward()
    loss_32.backward()
    loss_g.backward()
    assert l2(xg.grad, xr.grad) < 2e-3
    # vs plain fp32: relative L2 (a reduced-precision pre-activation that lands on the other side of 0 flips
    # LeakyReLU' between 1 and 0.2 for that element, so the max-norm of a gradient is not a meaningful bar)
    assert l2(xg.grad, x32.grad) < 2e-2
    for k, p in msd.named_parameters():
        assert l2(p.grad, P[k].grad) < 5e-3, k     # (bias-gradient sums see the occasional LeakyReLU' flip)


def test_mpd_tensor_core_mode_matches_oracle(dev):
    """MultiPeriodDiscriminator with the TF32 mma.sync grouped convolutions on, against the fp32 CPU oracle:
    feature maps 5e-3 relative to each map's max, input gradient 1e-2 and parameter gradients 2e-2 relative L2
    (TF32 = 10 mantissa bits; stated tolerance of the tensor-core configuration)."""
    from models.discriminators import MultiPeriodDiscriminator
    O = oracle()
    torch.manual_seed(0)
    mpd = MultiPeriodDiscriminator()
    P = leaf_params(cpu_params(mpd))
    mpd = mpd.to(dev)
    x = torch.randn(2, 6001, generator=torch.Generator().manual_seed(8)) * 0.1
    xr = x.clone().requires_grad_(True)
    lr, fr = O.mpd_forward(P, xr)
    xg = x.to(dev).requires_grad_(True)
    lg, fg = mpd(xg)
    loss_r, loss_g = 0.0, 0.0
    gen = torch.Generator().manual_seed(9)
    for i in range(5):
        for a, b in zip(fg[i], fr[i]):
            assert a.shape == b.shape
            assert rel_err(a, b) < 5e-3
            gw = torch.randn(b.shape, generator=gen) / b.numel() ** 0.5
            loss_r = loss_r + (b * gw).sum()
            loss_g = loss_g + (a * gw.to(dev)).sum()
    loss_r.backward()
    loss_g.backward()
    a, b = xg.grad.detach().cpu().double(), xr.grad.double()
    assert ((a - b).norm() / b.norm()).item() < 1e-2          # relative L2 (see test_msd_bf16_matches_oracle)
    for k, p in mpd.named_parameters():
        a, b = p.grad.detach().cpu().double(), P[k].grad.double()
        assert ((a - b).norm() / b.norm()).item() < 2e-2, k


@pytest.mark.parametrize("M,N,K", [(1000, 48, 16), (777, 192, 64), (130, 64, 128), (3000, 16, 96)])
def test_gemm_tf32_layouts(dev, M, N, K):
    """lct_gemm on the TF32 mma.sync kernel: NT + epilogue, NN, TN split-K.  2e-4 vs TF32-rounded operands in fp64."""
    from lctgan import ops, _lib
    _lib.call_ret("lct_set_tensor_core_gemm", 1)
    try:
        _gemm_tf32_checks(dev, ops, M, N, K)
    finally:
        _lib.call_ret("lct_set_tensor_core_gemm", 0)


def _gemm_tf32_checks(dev, ops, M, N, K):
    gen = torch.Generator().manual_seed(M + N)
    A, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen), torch.randn(N, generator=gen)
    res = torch.randn(M, N, generator=gen)
    C = torch.empty(M, N, device=dev)
    out2 = torch.empty(M, N, device=dev)
    ops.gemm(A.to(dev), W.to(dev), C, M, N, K, lda=K, ldb=K, ldc=N, bias=b.to(dev), act=ops.ACT_LRELU, slope=0.2,
             res=res.to(dev), ldr=N, out2=out2, ldo=N)
    ref = F.leaky_relu(_tf32(A) @ _tf32(W).t() + b.double(), 0.2)
    assert rel_err(C.double(), ref) < 2e-4
    assert rel_err(out2.double(), ref + res.double()) < 2e-4
    dY = torch.randn(M, N, generator=gen)
    dX = torch.empty(M, K, device=dev)
    ops.gemm(dY.to(dev), W.to(dev), dX, M, K, N, lda=N, ldb=K, ldc=K, tb=True)
    assert rel_err(dX.double(), _tf32(dY) @ _tf32(W)) < 2e-4
    dW = torch.zeros(N, K, device=dev)
    ops.gemm(dY.to(dev), A.to(dev), dW, N, K, M, lda=N, ldb=K, ldc=K, ta=True, tb=True, ksplit=3)
    assert rel_err(dW.double(), _tf32(dY).t() @ _tf32(A)) < 2e-4
