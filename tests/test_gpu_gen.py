"""GPU parity of the generator kernels and of the assembled LCTGenerator / LCTEnhancer against the
CPU oracle (models/generator.py of the reference).  fp32 throughout; tolerances are relative to
the largest reference value and stated per check."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import cpu_params, leaf_params, oracle, rel_err

pytestmark = pytest.mark.gpu


def test_layernorm(dev):
    from lctgan import ops
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(1000, 64, generator=gen) * 2 + 0.3
    gam, bet = torch.randn(64, generator=gen), torch.randn(64, generator=gen)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, gam, bet))
    ref = F.layer_norm(xr, (64,), gr, br)
    gy = torch.randn(ref.shape, generator=gen)
    dres = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()
    y, mean, rstd = ops.layernorm_fwd(x.to(dev), gam.to(dev), bet.to(dev))
    assert rel_err(y, ref) < 1e-5
    dg = torch.zeros(64, device=dev)
    db = torch.zeros(64, device=dev)
    dx = ops.layernorm_bwd(gy.to(dev), x.to(dev), gam.to(dev), mean, rstd, dg, db, dres=dres.to(dev))
    assert rel_err(dx, xr.grad + dres) < 2e-5
    assert rel_err(dg, gr.grad) < 5e-5
    assert rel_err(db, br.grad) < 5e-5


@pytest.mark.parametrize("tc", [False, pytest.param(True, marks=pytest.mark.bf16)])   # SIMT fp32 | 3xTF32 tensor-core kernels
@pytest.mark.parametrize("M,N,K", [(1000, 48, 16), (777, 192, 64), (130, 64, 128), (64, 64, 64)])
def test_gemm_layouts(dev, M, N, K, tc):
    from lctgan import ops
    gen = torch.Generator().manual_seed(M + N)
    A, W, b = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen), torch.randn(N, generator=gen)
    res = torch.randn(M, N, generator=gen)
    Ad, Wd = A.to(dev), W.to(dev)
    # NT + bias + lrelu + residual second output
    C = torch.empty(M, N, device=dev)
    out2 = torch.empty(M, N, device=dev)
    ops.gemm(Ad, Wd, C, M, N, K, lda=K, ldb=K, ldc=N, bias=b.to(dev), act=ops.ACT_LRELU, slope=0.2, res=res.to(dev),
             ldr=N, out2=out2, ldo=N)
    ref = F.leaky_relu(A @ W.t() + b, 0.2)
    assert rel_err(C, ref) < 2e-5
    assert rel_err(out2, ref + res) < 2e-5
    # NN: dX = dY @ W
    dY = torch.randn(M, N, generator=gen)
    dX = torch.empty(M, K, device=dev)
    ops.gemm(dY.to(dev), Wd, dX, M, K, N, lda=N, ldb=K, ldc=K, tb=True)
    assert rel_err(dX, dY @ W) < 2e-5
    # TN split-K: dW = dY^T @ A
    dW = torch.zeros(N, K, device=dev)
    ops.gemm(dY.to(dev), Ad, dW, N, K, M, lda=N, ldb=K, ldc=K, ta=True, tb=True, ksplit=3)
    assert rel_err(dW, dY.t() @ A) < 5e-5
    cs = torch.zeros(N, device=dev)
    ops.colsum(dY.to(dev), cs, M, N, N)
    assert rel_err(cs, dY.sum(0)) < 5e-5


def _gru_ref(x, wih, whh, bih, bhh, D, O):
    """x [S, L, 64] -> per (g, d) hidden sequences stacked [S, L, G*D, 16] via the oracle's explicit GRU."""
    outs = []
    for g in range(4):
        for d in range(D):
            gd = g * D + d
            outs.append(O.gru_direction(x[..., g * 16:(g + 1) * 16], wih[gd], whh[gd], bih[gd], bhh[gd], reverse=(d == 1)))
    return torch.stack(outs, dim=2)


@pytest.mark.parametrize("freq", [True, False])
def test_gru_recurrence(dev, freq):
    """lct_gru_fwd / lct_gru_bwd against the written-out GRU equations, in both stride geometries."""
    from lctgan import gen_impl, ops
    O = oracle()
    B, T, Fq = 2, 7, 9
    D = 2 if freq else 1
    GD = 4 * D
    gen = torch.Generator().manual_seed(3 + D)
    x = torch.randn(B, T, Fq, 64, generator=gen)
    wih = torch.randn(GD, 48, 16, generator=gen) * 0.3
    whh = torch.randn(GD, 48, 16, generator=gen) * 0.3
    bih = torch.randn(GD, 48, generator=gen) * 0.1
    bhh = torch.randn(GD, 48, generator=gen) * 0.1
    leaves = [t.clone().requires_grad_(True) for t in (x, wih, whh, bih, bhh)]
    xr = leaves[0]
    seqs = xr.reshape(B * T, Fq, 64) if freq else xr.permute(0, 2, 1, 3).reshape(B * Fq, T, 64)
    hs_ref = _gru_ref(seqs, *leaves[1:], D, O)                      # [S, L, GD, 16]
    summed = hs_ref.reshape(*hs_ref.shape[:2], 4, D, 16).sum(3).reshape(*hs_ref.shape[:2], 64)
    gy = torch.randn(summed.shape, generator=gen)
    (summed * gy).sum().backward()

    M = B * T * Fq
    geo = gen_impl._geom(B, T, Fq, freq)
    xd = x.reshape(M, 64).to(dev)
    gi = torch.empty(M, GD, 48, device=dev)
    ops.gemm(xd, wih.to(dev), gi, M, 48, 16, lda=64, ldb=16, ldc=GD * 48, bias=bih.to(dev), nbatch=GD, a_div=D, sA=16,
             sB=48 * 16, sC=48, sBias=48)
    hs = torch.empty(M, GD, 16, device=dev)
    gsave = torch.empty(M, GD, 64, device=dev)
    hprev = torch.empty(M, GD, 16, device=dev)
    ops.call("lct_gru_fwd", gi, whh.to(dev), bhh.to(dev), hs, gsave, hprev, geo[0], geo[1], GD, D, geo[2], geo[3], geo[4],
             geo[5])
    hs_inf = torch.empty(M, GD, 16, device=dev)          # inference form: no saved gates, same states
    ops.call("lct_gru_fwd", gi, whh.to(dev), bhh.to(dev), hs_inf, None, None, geo[0], geo[1], GD, D, geo[2], geo[3],
             geo[4], geo[5])
    assert torch.equal(hs, hs_inf)
    # bring the reference into [B,T,F] row order
    if freq:
        hs_ref_rows = hs_ref.reshape(M, GD, 16)
        gy_rows = gy.reshape(M, 64)
    else:
        hs_ref_rows = hs_ref.reshape(B, Fq, T, GD, 16).permute(0, 2, 1, 3, 4).reshape(M, GD, 16)
        gy_rows = gy.reshape(B, Fq, T, 64).permute(0, 2, 1, 3).reshape(M, 64)
    assert rel_err(hs, hs_ref_rows) < 2e-5
    dgi = torch.empty(M, GD, 48, device=dev)
    dgh = torch.empty(M, GD, 48, device=dev)
    dwhh = torch.zeros(GD, 48, 16, device=dev)
    dbih = torch.zeros(GD, 48, device=dev)
    dbhh = torch.zeros(GD, 48, device=dev)
    ops.call("lct_gru_bwd", gsave, hprev, whh.to(dev), gy_rows.contiguous().to(dev), 64, dgi, dgh, geo[0], geo[1], GD, D,
             geo[2], geo[3], geo[4], geo[5])
    ops.gemm(dgh, hprev, dwhh, 48, 16, M, lda=GD * 48, ldb=GD * 16, ldc=16, ta=True, tb=True, ksplit=2, nbatch=GD,
             a_div=1, b_div=1, sA=48, sB=16, sC=48 * 16)
    ops.colsum(dgi, dbih, M, GD * 48, GD * 48)
    ops.colsum(dgh, dbhh, M, GD * 48, GD * 48)
    assert rel_err(dwhh, leaves[2].grad) < 1e-4
    assert rel_err(dbih, leaves[3].grad) < 1e-4
    assert rel_err(dbhh, leaves[4].grad) < 1e-4
    dwih = torch.zeros(GD, 48, 16, device=dev)
    ops.gemm(dgi, xd, dwih, 48, 16, M, lda=GD * 48, ldb=64, ldc=16, ta=True, tb=True, ksplit=2, nbatch=GD, a_div=1,
             b_div=D, sA=48, sB=16, sC=48 * 16)
    assert rel_err(dwih, leaves[1].grad) < 1e-4
    dx = torch.empty(M, 64, device=dev)
    ops.gemm(dgi, wih.to(dev), dx, M, 16, D * 48, lda=GD * 48, ldb=16, ldc=64, tb=True, nbatch=4, sA=D * 48,
             sB=D * 48 * 16, sC=16)
    assert rel_err(dx, leaves[0].grad.reshape(M, 64)) < 1e-4


@pytest.mark.parametrize("freq,T,Fq", [(True, 5, 33), (False, 129, 3), (False, 300, 2)])
def test_attention(dev, freq, T, Fq):
    from lctgan import gen_impl, ops
    O = oracle()
    B = 2
    gen = torch.Generator().manual_seed(17)
    M = B * T * Fq
    x = torch.randn(B, T, Fq, 64, generator=gen)
    in_w, in_b = torch.randn(192, 64, generator=gen) * 0.2, torch.randn(192, generator=gen) * 0.1
    eye, zero = torch.eye(64), torch.zeros(64)
    qkv = F.linear(x, in_w, in_b)                               # [B,T,F,192]
    qr = qkv.clone().requires_grad_(True)

    def core(qkv_btf):
        seqs = qkv_btf.reshape(B * T, Fq, 192) if freq else qkv_btf.permute(0, 2, 1, 3).reshape(B * Fq, T, 192)
        q, k, v = seqs.chunk(3, dim=-1)
        S, Lq, _ = q.shape
        sp = lambda t: t.reshape(S, Lq, 4, 16).permute(0, 2, 1, 3)
        att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 4.0, dim=-1)
        o = (att @ sp(v)).permute(0, 2, 1, 3).reshape(S, Lq, 64)
        return o.reshape(B, T, Fq, 64) if freq else o.reshape(B, Fq, T, 64).permute(0, 2, 1, 3)

    ref = core(qr)
    gy = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()
    geo = gen_impl._geom(B, T, Fq, freq)
    qd = qkv.reshape(M, 192).to(dev)
    out = torch.empty(M, 64, device=dev)
    lse = torch.empty(M, 4, device=dev)
    ops.call("lct_attn_fwd", qd, out, lse, 4, geo[0], geo[1], geo[2], geo[3], geo[4], geo[5])
    assert rel_err(out, ref.reshape(M, 64)) < 2e-5
    dq = torch.empty(M, 192, device=dev)
    ops.call("lct_attn_bwd", qd, out, lse, gy.reshape(M, 64).contiguous().to(dev), dq, 4, geo[0], geo[1], geo[2],
             geo[3], geo[4], geo[5])
    assert rel_err(dq, qr.grad.reshape(M, 192)) < 1e-4
    # and the oracle's own MHA agrees with this formulation (guards the test itself)
    seqs = x.reshape(B * T, Fq, 64) if freq else x.permute(0, 2, 1, 3).reshape(B * Fq, T, 64)
    mha = O.multihead_self_attention(seqs, in_w, in_b, eye, zero)
    mha = mha.reshape(B, T, Fq, 64) if freq else mha.reshape(B, Fq, T, 64).permute(0, 2, 1, 3)
    assert rel_err(ref, mha) < 1e-5


def test_attention_forward_long_sequence(dev):
    """The attention forward streams keys / values through shared memory in chunks, so the time block accepts whole
    utterances of any length (infer.py; L = frames + 3): 2500 frames = 40 s at hop 256, three chunks of 1024 rows and
    ten query blocks per (sequence, head).  Round 1 kept all keys in shared memory and refused L > 1600."""
    from lctgan import gen_impl, ops
    B, T, Fq = 1, 2500, 2
    gen = torch.Generator().manual_seed(23)
    M = B * T * Fq
    qkv = torch.randn(B, T, Fq, 192, generator=gen)
    seqs = qkv.permute(0, 2, 1, 3).reshape(B * Fq, T, 192).double()
    q, k, v = seqs.chunk(3, dim=-1)
    sp = lambda t: t.reshape(B * Fq, T, 4, 16).permute(0, 2, 1, 3)
    att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 4.0, dim=-1)
    ref = (att @ sp(v)).permute(0, 2, 1, 3).reshape(B, Fq, T, 64).permute(0, 2, 1, 3).reshape(M, 64)
    geo = gen_impl._geom(B, T, Fq, False)
    out = torch.empty(M, 64, device=dev)
    lse = torch.empty(M, 4, device=dev)
    ops.call("lct_attn_fwd", qkv.reshape(M, 192).to(dev), out, lse, 4, geo[0], geo[1], geo[2], geo[3], geo[4], geo[5])
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("tc", [False, pytest.param(True, marks=pytest.mark.bf16)])   # SIMT fp32 | 3xTF32 tensor-core kernels
@pytest.mark.parametrize("Ci,Co,T,Fq", [(1, 16, 6, 257), (16, 32, 7, 129), (32, 64, 8, 65), (4, 8, 3, 10)])
def test_gconv_conv(dev, Ci, Co, T, Fq, tc):
    from lctgan import ops
    gen = torch.Generator().manual_seed(Ci + Co)
    B = 2
    x = torch.randn(B, Ci, T, Fq, generator=gen)
    w = torch.randn(Co, Ci, 2, 3, generator=gen) * 0.2
    b = torch.randn(Co, generator=gen)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.leaky_relu(F.conv2d(xr, wr, br, stride=(1, 2), padding=(1, 1)), 0.2)
    gy = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()
    xl = x.permute(0, 2, 3, 1).contiguous().to(dev)
    To, Fo = ref.shape[2], ref.shape[3]
    y = ops.gconv(xl, w.to(dev), b.to(dev), (To, Fo), Co, transposed=False, act=ops.ACT_LRELU, slope=0.2)
    assert rel_err(y, ref.permute(0, 2, 3, 1)) < 2e-5
    dpre = ops.act_bwd(y, gy.permute(0, 2, 3, 1).contiguous().to(dev), ops.ACT_LRELU, 0.2)
    dx = ops.gconv(dpre, w.to(dev), None, (T, Fq), Ci, transposed=True)
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-5
    dw = ops.gconv_wgrad(dpre, xl, w.shape)
    assert rel_err(dw, wr.grad) < 5e-5
    db = torch.zeros(Co, device=dev)
    ops.colsum(dpre, db, dpre.numel() // Co, Co, Co)
    assert rel_err(db, br.grad) < 5e-5


@pytest.mark.parametrize("tc", [False, pytest.param(True, marks=pytest.mark.bf16)])   # SIMT fp32 | 3xTF32 tensor-core kernels
@pytest.mark.parametrize("Ci,Co,T,Fq", [(64, 32, 6, 33), (32, 16, 5, 66), (16, 1, 4, 132), (8, 4, 3, 5)])
def test_gconv_deconv(dev, Ci, Co, T, Fq, tc):
    from lctgan import ops
    gen = torch.Generator().manual_seed(Ci * 3 + Co)
    B = 2
    x = torch.randn(B, Ci, T, Fq, generator=gen)
    w = torch.randn(Ci, Co, 2, 3, generator=gen) * 0.2
    b = torch.randn(Co, generator=gen)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.leaky_relu(F.conv_transpose2d(xr, wr, br, stride=(1, 2), padding=(1, 1), output_padding=(0, 1)), 0.2)
    assert ref.shape[2] == T - 1 and ref.shape[3] == 2 * Fq
    gy = torch.randn(ref.shape, generator=gen)
    (ref * gy).sum().backward()
    xl = x.permute(0, 2, 3, 1).contiguous().to(dev)
    y = ops.gconv(xl, w.to(dev), b.to(dev), (T - 1, 2 * Fq), Co, transposed=True, act=ops.ACT_LRELU, slope=0.2)
    assert rel_err(y, ref.permute(0, 2, 3, 1)) < 2e-5
    dpre = ops.act_bwd(y, gy.permute(0, 2, 3, 1).contiguous().to(dev), ops.ACT_LRELU, 0.2)
    dx = ops.gconv(dpre, w.to(dev), None, (T, Fq), Ci, transposed=False)
    assert rel_err(dx, xr.grad.permute(0, 2, 3, 1)) < 2e-5
    dw = ops.gconv_wgrad(xl, dpre, w.shape)
    assert rel_err(dw, wr.grad) < 5e-5


@pytest.mark.parametrize("tc", [False, pytest.param(True, marks=pytest.mark.bf16)])   # SIMT fp32 | 3xTF32 tensor-core kernels
@pytest.mark.parametrize("cls,fn", [("GRUblockf", "gru_block_f"), ("GRUblockt", "gru_block_t")])
def test_gru_block_module(dev, cls, fn, tc):
    import models.generator as MG
    O = oracle()
    torch.manual_seed(5)
    blk = getattr(MG, cls)(64)
    P = leaf_params(cpu_params(blk))
    P = {"b." + k: v for k, v in P.items()}
    blk = blk.to(dev)
    x = torch.randn(2, 64, 6, 9, generator=torch.Generator().manual_seed(6))
    xr = x.clone().requires_grad_(True)
    ref = getattr(O, fn)(P, "b.", xr)
    gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
    (ref * gy).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    out = blk(xg)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 5e-5
    (out * gy.to(dev)).sum().backward()
    assert rel_err(xg.grad, xr.grad) < 2e-4
    for k, p in blk.named_parameters():
        assert rel_err(p.grad, P["b." + k].grad) < 2e-4, k


def _tc_params():
    """(T, tensor-core mode) cases: fp32 SIMT kernels with the tight bars, and the default product mode in which the
    generator's GEMMs / convolutions run as 3xTF32 (error-compensated) mma.sync - same forward bar, gradient bar 2e-3
    (the tensor cores accumulate with truncation, so K = 64..384 long sums sit a little above fp32 FMA accuracy)."""
    return [pytest.param(8000, False), pytest.param(5000, False),
            pytest.param(8000, True, marks=pytest.mark.bf16), pytest.param(5000, True, marks=pytest.mark.bf16)]


@pytest.mark.parametrize("T,tc", _tc_params())
def test_generator_and_enhancer(dev, T, tc):
    from models.generator import LCTEnhancer, LCTGeneratorConfig
    from lctgan import config
    assert config.gconv_tensor_cores is tc
    gbar = 2e-3 if tc else 5e-4
    O = oracle()
    torch.manual_seed(42)
    enh = LCTEnhancer(LCTGeneratorConfig(max_time_context=200), c=0.3)
    P = leaf_params(cpu_params(enh))
    enh = enh.to(dev)
    noisy, clean = O.synthetic_batch(2, T, seed=99)
    er, mr = O.enhancer_forward(P, noisy)
    gen = torch.Generator().manual_seed(8)
    gw, gm = torch.randn(er.shape, generator=gen), torch.randn(mr.shape, generator=gen) * 0.01
    ((er * gw).sum() + (mr * gm).sum()).backward()
    eg, mg = enh(noisy.to(dev))
    assert eg.shape == er.shape and mg.shape == mr.shape
    assert rel_err(mg, mr) < 5e-5
    assert rel_err(eg, er) < 5e-5
    # reference quirks: mask >= 0.5 and the last 3 frames are exactly 0.5
    assert mg.min().item() >= 0.5
    assert torch.all(mg[..., -3:] == 0.5)
    ((eg * gw.to(dev)).sum() + (mg * gm.to(dev)).sum()).backward()
    # float64 run of the oracle = ground truth.  Gradients of the attention q/k projections are sums with
    # heavy cancellation (softmax shift invariance: dL/db_k == 0 analytically), so the bar is "within 5e-4
    # of the truth, or no worse than 4x the fp32 CPU oracle's own distance from the truth".
    P64 = {k: (v.detach().double().requires_grad_(v.requires_grad)) for k, v in P.items()}
    e64, m64 = O.enhancer_forward(P64, noisy.double())
    ((e64 * gw.double()).sum() + (m64 * gm.double()).sum()).backward()
    for k, p in enh.named_parameters():
        assert p.grad is not None, k
        truth = P64[k].grad
        e_mine = rel_err(p.grad.double(), truth)
        e_ref = rel_err(P[k].grad.double(), truth)
        assert e_mine < max(gbar, 4.0 * e_ref), (k, e_mine, e_ref)
    # LCTGenerator on its own, through the reference's [B,1,F,T] interface
    mag = O.magnitude(O.stft(noisy, P["stft.window"], 512, 256)).unsqueeze(1)
    with torch.no_grad():
        m2 = enh.gen(mag.to(dev))
    assert rel_err(m2, mr) < 5e-5
    with pytest.raises(ValueError):
        enh.gen(mag.to(dev)[:, 0])
    with pytest.raises(ValueError):
        enh(noisy.to(dev)[0])


@pytest.mark.parametrize("B,T", [(1, 160000), (2, 16000), (3, 40123)])
def test_enhancer_full_utterance_inference(dev, B, T):
    """infer.py:147 path (BASELINE configs[1]): eval + no_grad enhancement of 1 - 10 s utterances (time attention
    over up to 629 frames) against the CPU oracle."""
    from models.generator import LCTEnhancer, LCTGeneratorConfig
    O = oracle()
    torch.manual_seed(3)
    enh = LCTEnhancer(LCTGeneratorConfig(), c=0.3).eval()
    P = cpu_params(enh)
    enh = enh.to(dev)
    noisy, _ = O.synthetic_batch(B, T, seed=T)
    with torch.no_grad():
        er, mr = O.enhancer_forward(P, noisy, aten_gru=True)
        eg, mg = enh(noisy.to(dev))
    assert eg.shape == (B, T) and mg.shape == mr.shape
    assert rel_err(eg, er) < 5e-5
    assert rel_err(mg, mr) < 5e-5
