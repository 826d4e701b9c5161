"""GPU parity of the multi-tensor losses and of the whole adversarial training step against the CPU
oracle (losses.py and train.py:165-249 of the reference)."""
import pytest
import torch

from util import cpu_params, oracle, rel_err

pytestmark = pytest.mark.gpu


def _logits(seed, n=8):
    g = torch.Generator().manual_seed(seed)
    shapes = [(2, 1, 50 + 7 * i, (2, 3, 5, 7, 11)[i % 5]) for i in range(5)] + [(2, 1, 125), (2, 1, 63), (2, 1, 32)]
    return [torch.randn(s, generator=g) for s in shapes[:n]]


@pytest.mark.parametrize("loss_type", ["ls", "hinge"])
def test_gan_losses(dev, loss_type):
    import losses as L
    O = oracle()
    real, fake = _logits(1), _logits(2)
    rr = [t.clone().requires_grad_(True) for t in real]
    fr = [t.clone().requires_grad_(True) for t in fake]
    dref = O.discriminator_loss(rr, fr, loss_type)
    (dref * 1.3).backward()
    rg = [t.to(dev).requires_grad_(True) for t in real]
    fg = [t.to(dev).requires_grad_(True) for t in fake]
    dgot = L.discriminator_loss(rg, fg, loss_type)
    assert dgot.dim() == 0
    assert abs(dgot.item() - dref.item()) < 1e-5 * max(1.0, abs(dref.item()))
    (dgot * 1.3).backward()
    for a, b in zip(rg + fg, rr + fr):
        assert rel_err(a.grad, b.grad) < 1e-5
    fr2 = [t.clone().requires_grad_(True) for t in fake]
    gref = O.generator_adv_loss(fr2, loss_type)
    gref.backward()
    fg2 = [t.to(dev).requires_grad_(True) for t in fake]
    ggot = L.generator_adv_loss(fg2, loss_type)
    assert abs(ggot.item() - gref.item()) < 1e-5 * max(1.0, abs(gref.item()))
    ggot.backward()
    for a, b in zip(fg2, fr2):
        assert rel_err(a.grad, b.grad) < 1e-5
    with pytest.raises(ValueError):
        L.discriminator_loss(rg, fg[:-1], loss_type)
    with pytest.raises(ValueError):
        L.discriminator_loss(rg, fg, "wgan")
    with pytest.raises(ValueError):
        L.generator_adv_loss(fg, "wgan")


def test_feature_matching_and_mask(dev):
    import losses as L
    O = oracle()
    g = torch.Generator().manual_seed(3)
    shapes = [[(2, 32, 100, 2), (2, 128, 34, 2), (2, 1, 34, 2)], [(2, 16, 1000), (2, 64, 250), (2, 1024, 17), (2, 1, 17)]]
    real = [[torch.randn(s, generator=g) for s in d] for d in shapes]
    fake = [[torch.randn(s, generator=g) for s in d] for d in shapes]
    fake[0][0][0, 0, :5] = real[0][0][0, 0, :5]     # exact ties: sign(0) = 0
    fr = [[t.clone().requires_grad_(True) for t in d] for d in fake]
    ref = O.feature_matching_loss(real, fr)
    (ref * 0.7).backward()
    fg = [[t.to(dev).requires_grad_(True) for t in d] for d in fake]
    rg = [[t.to(dev) for t in d] for d in real]
    got = L.feature_matching_loss(rg, fg)
    assert abs(got.item() - ref.item()) < 1e-5 * abs(ref.item())
    (got * 0.7).backward()
    for dg_, dr_ in zip(fg, fr):
        for a, b in zip(dg_, dr_):
            assert rel_err(a.grad, b.grad) < 1e-5
    with pytest.raises(ValueError):
        L.feature_matching_loss(rg[:1], fg)
    with pytest.raises(ValueError):
        L.feature_matching_loss([rg[0][:2], rg[1]], fg)
    # mask MSE on [B, F, T] views of [B, T, F] buffers and on plain tensors
    p = torch.rand(2, 30, 257, generator=g)
    t = torch.rand(2, 30, 257, generator=g) * 2
    pr = p.clone().requires_grad_(True)
    mref = O.mask_mse_loss(pr.transpose(1, 2), t.transpose(1, 2))
    mref.backward()
    pg = p.to(dev).requires_grad_(True)
    mgot = L.mask_mse_loss(pg.transpose(1, 2), t.to(dev).transpose(1, 2))
    assert abs(mgot.item() - mref.item()) < 1e-5 * abs(mref.item())
    mgot.backward()
    assert rel_err(pg.grad, pr.grad) < 1e-5
    with pytest.raises(ValueError):
        L.mask_mse_loss(pg, pg[:, :10])


@pytest.mark.parametrize("gan_loss", ["ls", "hinge"])
def test_training_step_matches_oracle(dev, gan_loss):
    """Two consecutive D+G steps on one synthetic batch, identical seeds and weights: losses, the clipped
    enhancer gradient norm and the post-step weights must agree with the CPU oracle (fp32 tolerance
    2e-3 relative on losses after two optimiser steps; 1e-3 on the first)."""
    from lctgan.training import StepArgs, build_models, train_step
    O = oracle()
    B, T = 2, 8000
    enh, mpd, msd, tf, mr, g_opt, d_opt = build_models(dev, gan_seed=42)
    st = O.StepState(cpu_params(enh), cpu_params(mpd), cpu_params(msd),
                     order_g=[k for k, _ in enh.named_parameters()],
                     order_d=([k for k, _ in mpd.named_parameters()], [k for k, _ in msd.named_parameters()]))
    wins = [O.hann_window(n) for n in O.MR_FFT_SIZES]
    noisy, clean = O.synthetic_batch(B, T, seed=1234)
    nd, cd = noisy.to(dev), clean.to(dev)
    args = StepArgs(gan_loss=gan_loss)
    for step in range(2):
        ref = O.train_step(st, noisy, clean, wins, gan_loss=gan_loss)
        got = train_step(enh, mpd, msd, tf, mr, g_opt, d_opt, nd, cd, args)
        tol = 1e-3 if step == 0 else 2e-3
        for k in ("d_loss", "g_loss", "mr", "mask", "adv", "fm"):
            r, g = ref[k], got[k].item()
            assert abs(g - r) <= tol * max(abs(r), 1e-3), (step, k, g, r)
    # post-step weights.  AdamW turns a gradient of any size into a step of ~lr, so entries whose gradient is
    # pure rounding noise (e.g. the attention key bias, analytically zero) can differ by a fraction of
    # lr * steps (a sign flip of a ~0 gradient moves a weight by up to 2 * lr).  Bound: at most 0.1 % of a
    # tensor's entries may differ by more than 5 % of lr * steps, none by more than 2 * lr * steps.
    lr, nsteps = 2e-4, 2
    for mod, ref in ((enh, st.enh), (msd, st.msd), (mpd, st.mpd)):
        for k, p in mod.named_parameters():
            diff = (p.detach().cpu() - ref[k].detach()).abs()
            assert diff.max().item() <= 2.0 * lr * nsteps, (k, diff.max().item())
            bad = (diff > 0.05 * lr * nsteps).float().mean().item()
            assert bad <= 1e-3, (k, bad)


def test_fused_adamw_matches_torch(dev):
    """lctgan.optim.FusedAdamW against torch.optim.AdamW (the optimiser train.py:601-610 builds): same updates."""
    from lctgan.optim import FusedAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(1024, 4, 41), (64,), (3, 5, 2, 1), (1,), (100001,)] * 12       # > 48 tensors: several launches
    pa = [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    # + a parameter that starts 4 bytes off a 16-byte boundary (the kernel's scalar form) on both sides
    base = torch.randn(5003, generator=g).to(dev)
    pa.append(base.clone()[1:].detach().requires_grad_(True))
    pb.append(base.clone()[1:].detach().requires_grad_(True))
    assert pb[-1].data_ptr() % 16 == 4
    oa = torch.optim.AdamW(pa, lr=2e-4, betas=(0.8, 0.99))
    ob = FusedAdamW(pb, lr=2e-4, betas=(0.8, 0.99))
    for step in range(4):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(dev) * (10.0 ** (step - 2))
            a.grad = gr.clone()
            b.grad = gr.clone()
        oa.step()
        ob.step()
    for a, b in zip(pa, pb):
        assert rel_err(b, a) < 2e-6
    st = ob.state[pb[0]]
    assert set(st) == {"step", "exp_avg", "exp_avg_sq"} and float(st["step"]) == 4.0
    assert rel_err(st["exp_avg"], oa.state[pa[0]]["exp_avg"]) < 2e-6
    assert rel_err(st["exp_avg_sq"], oa.state[pa[0]]["exp_avg_sq"]) < 2e-6
    with pytest.raises(ValueError):
        FusedAdamW(pb, lr=-1.0)
