"""Generate the golden fixtures in tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE (jqshang/LCT-GAN at
/root/reference or $LCT_REF) on CPU.  Run in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors of its own, so these files are what pins the
oracle (oracle/lct_oracle.py): tests/test_oracle_golden.py checks every oracle function against them.

Shims: the reference root goes first on sys.path (its `datasets` package is shadowed by the
HuggingFace one otherwise) and `pesq` / `pystoi`, hard-imported by train.py but absent here, are stubbed.
"""
import argparse
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    ref = os.environ.get("LCT_REF", "/root/reference")
    if not os.path.isdir(ref):
        raise SystemExit(f"reference checkout not found at {ref}")
    for k in [k for k in sys.modules if k.split(".")[0] in ("datasets", "models", "losses", "train")]:
        del sys.modules[k]
    sys.path.insert(0, ref)
    for name, attr in (("pesq", "pesq"), ("pystoi", "stoi")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            setattr(m, attr, None)
            sys.modules[name] = m
    import datasets.stft as R_stft
    import datasets.tf_features as R_tf
    import losses as R_losses
    import models.discriminators as R_disc
    import models.generator as R_gen
    import train as R_train
    return R_stft, R_tf, R_losses, R_disc, R_gen, R_train


def batch(b, t, seed=1234):
    g = torch.Generator().manual_seed(seed)
    clean = torch.randn(b, t, generator=g) * 0.1
    noisy = clean + torch.randn(b, t, generator=g) * 0.05
    return noisy, clean


def sub(t, n=4096):
    """A deterministic sub-sample of a tensor (fixtures stay small)."""
    f = t.detach().reshape(-1)
    if f.numel() <= n:
        return f.clone()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return f[idx].clone()


def train_logs(G, R_train, R_gen, R_disc, R_tf, R_losses, noisy, clean):
    """Two unmodified train_one_epoch steps (ls and hinge) on one batch: printed losses, gradient norms after the first
    backward passes and post-step weight checksums."""
    import argparse as _ap
    import contextlib
    import io
    import re
    for gan_loss in ("ls", "hinge"):
        R_train.set_seed(42)
        enh = R_gen.LCTEnhancer(R_gen.LCTGeneratorConfig(max_time_context=200), c=0.3)
        mpd = R_disc.MultiPeriodDiscriminator()
        msd = R_disc.MultiScaleDiscriminator()
        tfm = R_tf.TFFeatures(R_tf.TFFeaturesConfig(n_fft=512, c=0.3, compress_input=False, return_stfts=False))
        mr = R_losses.MultiResolutionSTFTLoss(R_losses.MRSTFTLossConfig())
        g_opt = torch.optim.AdamW(enh.parameters(), lr=2e-4, betas=(0.8, 0.99))
        d_opt = torch.optim.AdamW(list(mpd.parameters()) + list(msd.parameters()), lr=2e-4, betas=(0.8, 0.99))
        ns = _ap.Namespace(gan_loss=gan_loss, lambda_fm=1.0, lambda_mask=1.0, lambda_adv=1e-2, grad_clip=5.0,
                           log_interval=1)
        logs = []
        # gradient statistics at the moment each optimiser steps (optimizer pre-hooks: the reference code is untouched):
        # D gradients of d_loss.backward() alone, enhancer gradients after clip_grad_norm_
        gstats = {"d": [], "g": []}

        def _stat(key, params):
            def hook(opt, a, k):
                gs = [p.grad.double() for p in params if p.grad is not None]
                gstats[key].append({"l2": float(torch.sqrt(sum((g * g).sum() for g in gs))),
                                    "sum": float(sum(g.sum() for g in gs)), "abs": float(sum(g.abs().sum() for g in gs))})
            return hook
        d_opt.register_step_pre_hook(_stat("d", list(mpd.parameters()) + list(msd.parameters())))
        g_opt.register_step_pre_hook(_stat("g", list(enh.parameters())))
        for step in range(2):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                R_train.train_one_epoch(1, {"train": [{"noisy": noisy, "clean": clean}]}, enh, mpd, msd, tfm, mr, g_opt,
                                        d_opt, torch.device("cpu"), ns)
            vals = {k: float(v) for k, v in re.findall(r"(\w+)=(-?[\d.]+)", buf.getvalue())}
            logs.append(vals)
        G[f"train_{gan_loss}"] = {"logs": logs, "grad_stats": gstats,
                                  "enh_checksum": float(sum(p.double().sum() for p in enh.parameters())),
                                  "msd_checksum": float(sum(p.double().sum() for p in msd.parameters())),
                                  "mpd_checksum": float(sum(p.double().sum() for p in mpd.parameters()))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(HERE, "golden_v1.pt"))
    # golden_v1.pt: the defaults.  golden_v2.pt (ragged lengths, odd batch, iSTFT length = the ragged signal length: the tail comes from the last half window):
    #   python tests/golden/make_golden.py --out tests/golden/golden_v2.pt --front-len 5003 --istft-len 5003 \
    #          --model-len 12345 --batch 3 --front-seed 11 --model-seed 4321
    ap.add_argument("--front-len", type=int, default=4000)
    ap.add_argument("--istft-len", type=int, default=3900)
    ap.add_argument("--model-len", type=int, default=8000)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--front-seed", type=int, default=7)
    ap.add_argument("--model-seed", type=int, default=1234)
    # golden_v3.pt: the BASELINE shape (batch 8 x 32000 samples), logs and checksums only (a few KB):
    #   python tests/golden/make_golden.py --out tests/golden/golden_v3.pt --logs-only --model-len 32000 --batch 8
    ap.add_argument("--logs-only", action="store_true", help="store only the two-step train_one_epoch logs, the "
                    "parameter checksums and the recipe of the inputs (not the inputs themselves)")
    args = ap.parse_args()
    R_stft, R_tf, R_losses, R_disc, R_gen, R_train = import_reference()
    torch.set_num_threads(8)
    G = {"torch": torch.__version__, "istft_length": args.istft_len}
    if args.logs_only:
        G = {"torch": torch.__version__, "logs_only": True,
             "input_recipe": {"batch": args.batch, "samples": args.model_len, "seed": args.model_seed,
                              "formula": "g=Generator().manual_seed(seed); clean=randn(b,t,generator=g)*0.1; "
                                         "noisy=clean+randn(b,t,generator=g)*0.05"}}
        noisy, clean = batch(args.batch, args.model_len, seed=args.model_seed)
        G["input_checksum"] = (float(noisy.double().sum()), float(clean.double().sum()))
        train_logs(G, R_train, R_gen, R_disc, R_tf, R_losses, noisy, clean)
        torch.save(G, args.out)
        print("wrote", args.out, os.path.getsize(args.out), "bytes")
        for k in ("train_ls", "train_hinge"):
            print(k, G[k]["logs"])
        return

    # ---- front end: STFT / iSTFT / helpers at the three loss resolutions + the generator's
    noisy, clean = batch(args.batch, args.front_len, seed=args.front_seed)
    for n_fft, hop in ((512, 256), (320, 160), (768, 384)):
        m = R_stft.ComplexSTFT(R_stft.STFTConfig(n_fft=n_fft, hop_length=hop).finalize())
        s = m(noisy)
        G[f"stft_{n_fft}"] = torch.view_as_real(s).clone()
        G[f"istft_{n_fft}"] = m.istft(s * 0.7, length=args.istft_len).clone()
        G[f"window_{n_fft}"] = m.window.clone()
    s512 = R_stft.make_lct_stft()(noisy)
    c512 = R_stft.make_lct_stft()(clean)
    G["magnitude"] = R_stft.magnitude(s512).clone()
    G["compress"] = R_stft.compress(R_stft.magnitude(s512)).clone()
    G["irm_c"] = R_stft.compute_compressed_irm(c512, s512).clone()
    mk = torch.rand(args.batch, 1, 257, s512.shape[-1], generator=torch.Generator().manual_seed(3))
    G["mask_in"] = mk
    G["apply_mask_c"] = torch.view_as_real(R_stft.apply_mask(s512, mk, compressed=True)).clone()
    tf = R_tf.TFFeatures(R_tf.TFFeaturesConfig(return_stfts=False))(noisy, clean)
    G["tf_features"] = {k: v.clone() for k, v in tf.items()}
    G["front_inputs"] = (noisy, clean)

    # ---- models at seed 42, in train.py's construction order
    R_train.set_seed(42)
    enh = R_gen.LCTEnhancer(R_gen.LCTGeneratorConfig(max_time_context=200), c=0.3)
    mpd = R_disc.MultiPeriodDiscriminator()
    msd = R_disc.MultiScaleDiscriminator()
    tfm = R_tf.TFFeatures(R_tf.TFFeaturesConfig(n_fft=512, c=0.3, compress_input=False, return_stfts=False))
    mr = R_losses.MultiResolutionSTFTLoss(R_losses.MRSTFTLossConfig())
    G["state_keys"] = {"enh": list(enh.state_dict().keys()), "mpd": list(mpd.state_dict().keys()),
                       "msd": list(msd.state_dict().keys())}
    G["state_shapes"] = {n: {k: tuple(v.shape) for k, v in m.state_dict().items()}
                         for n, m in (("enh", enh), ("mpd", mpd), ("msd", msd))}
    G["param_checksum"] = {n: float(sum(p.double().abs().sum() for p in m.parameters()))
                           for n, m in (("enh", enh), ("mpd", mpd), ("msd", msd))}
    noisy, clean = batch(args.batch, args.model_len, seed=args.model_seed)
    G["model_inputs"] = (noisy, clean)
    with torch.no_grad():
        e, mask = enh(noisy)
        G["enhanced"] = e.clone()
        G["mask_c_sub"] = sub(mask)
        G["mask_c_tail"] = mask[..., -3:].clone()
        pl, pf = mpd(clean)
        sl, sf = msd(clean)
        G["mpd_logits"] = [t.clone() for t in pl]
        G["msd_logits"] = [t.clone() for t in sl]
        G["mpd_fmap_shapes"] = [[tuple(t.shape) for t in f] for f in pf]
        G["msd_fmap_shapes"] = [[tuple(t.shape) for t in f] for f in sf]
        G["mpd_fmap_sub"] = [[sub(t, 512) for t in f] for f in pf]
        G["msd_fmap_sub"] = [[sub(t, 512) for t in f] for f in sf]
        fl, ff = mpd(e)
        G["fm_loss"] = float(R_losses.feature_matching_loss(pf, ff))
        G["d_loss_ls"] = float(R_losses.discriminator_loss(pl + sl, fl + msd(e)[0], "ls"))
        G["d_loss_hinge"] = float(R_losses.discriminator_loss(pl + sl, fl + msd(e)[0], "hinge"))
        G["g_adv_ls"] = float(R_losses.generator_adv_loss(fl, "ls"))
        G["g_adv_hinge"] = float(R_losses.generator_adv_loss(fl, "hinge"))
        mrl, det = mr(e, clean)
        G["mrstft"] = (float(mrl), {k: float(v) for k, v in det.items()})

    train_logs(G, R_train, R_gen, R_disc, R_tf, R_losses, noisy, clean)
    torch.save(G, args.out)
    print("wrote", args.out, os.path.getsize(args.out), "bytes")
    for k in ("train_ls", "train_hinge"):
        print(k, G[k]["logs"])


if __name__ == "__main__":
    main()
