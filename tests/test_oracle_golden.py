"""Pins the CPU oracle (oracle/lct_oracle.py) against golden vectors produced by EXECUTING THE UNMODIFIED
REFERENCE (tests/golden/make_golden.py -> tests/golden/golden_v1.pt).  The reference ships no tests or
fixtures of its own, so these vectors are the anchor of every parity claim; the GPU parity tests then compare
the CUDA path with this oracle (and test_gpu_golden.py compares it with the same vectors directly).

Tolerances: the oracle runs the same torch CPU operators as the reference, so agreement is to fp32 rounding
(1e-6 relative); training-step losses are compared at the 4 decimals train.py prints.
"""
import os

import pytest
import torch

from util import oracle, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.pt")


# golden_v1: batch 2, 4000 / 8000 samples;  golden_v2: batch 3, ragged 5003 / 12345 samples (not multiples of any hop or
# period: reflect padding, partial last frames, iSTFT tail from the last half window) - both written by tests/golden/make_golden.py from the unmodified reference
@pytest.fixture(scope="module", params=["golden_v1.pt", "golden_v2.pt"])
def G(request):
    return torch.load(os.path.join(os.path.dirname(GOLD), request.param), weights_only=False)


def _sub(t, n=4096):
    f = t.detach().reshape(-1)
    if f.numel() <= n:
        return f.clone()
    return f[torch.linspace(0, f.numel() - 1, n).long()].clone()


def _seeded_models():
    """Same construction order and seed as train.py:569-598 -> the reference's initial weights."""
    import sys
    from lctgan.training import build_models
    enh, mpd, msd, tf, mr, _, _ = build_models("cpu", gan_seed=42)
    cp = lambda m: {k: v.detach().clone() for k, v in m.state_dict().items()}
    return enh, mpd, msd, cp(enh), cp(mpd), cp(msd)


def test_front_end_matches_reference(G):
    O = oracle()
    noisy, clean = G["front_inputs"]
    for n_fft, hop in ((512, 256), (320, 160), (768, 384)):
        w = O.hann_window(n_fft)
        assert torch.equal(w, G[f"window_{n_fft}"])            # fp32-rounded periodic Hann, bit exact
        s = O.stft(noisy, w, n_fft, hop)
        assert rel_err(torch.view_as_real(s), G[f"stft_{n_fft}"]) < 1e-6
        assert rel_err(O.stft_explicit(noisy, w, n_fft, hop), torch.view_as_complex(G[f"stft_{n_fft}"])) < 1e-5
        y = O.istft(s * 0.7, w, n_fft, hop, G.get("istft_length", 3900))
        assert rel_err(y, G[f"istft_{n_fft}"]) < 1e-6
        assert rel_err(O.istft_explicit(s * 0.7, w, n_fft, hop, G.get("istft_length", 3900)), G[f"istft_{n_fft}"]) < 1e-5
    w = O.hann_window(512)
    s, c = O.stft(noisy, w, 512, 256), O.stft(clean, w, 512, 256)
    assert rel_err(O.magnitude(s), G["magnitude"]) < 1e-6
    assert rel_err(O.compress(O.magnitude(s)), G["compress"]) < 1e-6
    assert rel_err(O.compressed_irm(c, s), G["irm_c"]) < 1e-6
    assert rel_err(torch.view_as_real(O.apply_mask(s, G["mask_in"], compressed=True)), G["apply_mask_c"]) < 1e-6
    tf = O.tf_features(noisy, clean, w, return_stfts=False)
    assert set(tf) == set(G["tf_features"])
    for k, v in G["tf_features"].items():
        assert rel_err(tf[k], v) < 1e-6, k


def test_state_dict_layout_and_seeded_init_match_reference(G):
    enh, mpd, msd, Pe, Pp, Ps = _seeded_models()
    for name, mod in (("enh", enh), ("mpd", mpd), ("msd", msd)):
        sd = mod.state_dict()
        assert list(sd.keys()) == G["state_keys"][name]
        assert {k: tuple(v.shape) for k, v in sd.items()} == G["state_shapes"][name]
        assert all(v.dtype == torch.float32 for v in sd.values())
        chk = float(sum(p.double().abs().sum() for p in mod.parameters()))
        assert abs(chk - G["param_checksum"][name]) <= 1e-9 * G["param_checksum"][name]   # identical seeded init
    assert len(G["state_keys"]["enh"]) == 131 and len(G["state_keys"]["mpd"]) == 90 and len(G["state_keys"]["msd"]) == 63


def test_models_and_losses_match_reference(G):
    O = oracle()
    _, _, _, Pe, Pp, Ps = _seeded_models()
    noisy, clean = G["model_inputs"]
    with torch.no_grad():
        e, mask = O.enhancer_forward(Pe, noisy)
        assert rel_err(e, G["enhanced"]) < 2e-5       # explicit GRU / attention equations vs aten::gru / MHA
        assert rel_err(_sub(mask), G["mask_c_sub"]) < 2e-5
        assert torch.equal(mask[..., -3:], G["mask_c_tail"]) and torch.all(mask[..., -3:] == 0.5)
        e2, _ = O.enhancer_forward(Pe, noisy, aten_gru=True)
        assert rel_err(e2, G["enhanced"]) < 2e-6
        pl, pf = O.mpd_forward(Pp, clean)
        sl, sf = O.msd_forward(Ps, clean)
        for a, b in zip(pl + sl, G["mpd_logits"] + G["msd_logits"]):
            assert a.shape == b.shape and rel_err(a, b) < 1e-5
        assert [[tuple(t.shape) for t in f] for f in pf] == G["mpd_fmap_shapes"]
        assert [[tuple(t.shape) for t in f] for f in sf] == G["msd_fmap_shapes"]
        for fa, fb in zip(pf + sf, G["mpd_fmap_sub"] + G["msd_fmap_sub"]):
            for a, b in zip(fa, fb):
                assert rel_err(_sub(a, 512), b) < 1e-5
        eg = G["enhanced"]
        fl, ff = O.mpd_forward(Pp, eg)
        fsl, _ = O.msd_forward(Ps, eg)
        assert abs(float(O.feature_matching_loss(pf, ff)) - G["fm_loss"]) < 1e-6
        assert abs(float(O.discriminator_loss(pl + sl, fl + fsl, "ls")) - G["d_loss_ls"]) < 1e-5
        assert abs(float(O.discriminator_loss(pl + sl, fl + fsl, "hinge")) - G["d_loss_hinge"]) < 1e-5
        assert abs(float(O.generator_adv_loss(fl, "ls")) - G["g_adv_ls"]) < 1e-5
        assert abs(float(O.generator_adv_loss(fl, "hinge")) - G["g_adv_hinge"]) < 1e-5
        mrl, det = O.mrstft_loss(eg, clean, [O.hann_window(n) for n in O.MR_FFT_SIZES])
        assert abs(float(mrl) - G["mrstft"][0]) < 1e-5
        for k, v in G["mrstft"][1].items():
            assert abs(float(det[k]) - v) < 1e-5


@pytest.mark.parametrize("gan_loss", ["ls", "hinge"])
def test_training_step_matches_reference_train_one_epoch(G, gan_loss):
    """Two D+G steps of the oracle vs two calls of the reference's own train_one_epoch (same seed, same batch)."""
    O = oracle()
    enh, mpd, msd, Pe, Pp, Ps = _seeded_models()
    st = O.StepState(Pe, Pp, Ps, order_g=[k for k, _ in enh.named_parameters()],
                     order_d=([k for k, _ in mpd.named_parameters()], [k for k, _ in msd.named_parameters()]))
    noisy, clean = G["model_inputs"]
    wins = [O.hann_window(n) for n in O.MR_FFT_SIZES]
    names = {"D_loss": "d_loss", "G_loss": "g_loss", "MR": "mr", "Mask": "mask", "Adv": "adv", "FM": "fm"}
    for step in range(2):
        got = O.train_step(st, noisy, clean, wins, gan_loss=gan_loss, aten_gru=True)
        for ref_k, k in names.items():
            assert abs(got[k] - G[f"train_{gan_loss}"]["logs"][step][ref_k]) <= 1.01e-4, (step, k)   # 4 printed decimals
    ref = G[f"train_{gan_loss}"]
    assert abs(float(sum(p.double().sum() for p in st.g_params)) - ref["enh_checksum"]) < 1e-3
    # AdamW's first steps move every parameter by ~ +-lr whatever the size of its gradient, so a parameter whose
    # gradient is rounding noise flips direction between two fp32 summation orders (4e-4 on the checksum each): 0.1
    # allows a few hundred of the 17.7 M discriminator parameters to do so (hinge saturates many logits: D_loss = 2.0)
    assert abs(float(sum(st.msd[k].double().sum() for k in st.msd)) - ref["msd_checksum"]) < 1e-1
    assert abs(float(sum(st.mpd[k].double().sum() for k in st.mpd)) - ref["mpd_checksum"]) < 1e-1


def test_oracle_stock_init_matches_reference_weights(G):
    """oracle/ref_init.py (stock torch containers only, no product import - what `bench.py --impl reference` starts
    from) reproduces the reference's seeded initial weights: keys, shapes and the checksum dumped from the reference."""
    from oracle import ref_init
    dicts = dict(zip(("enh", "mpd", "msd"), ref_init.init_state_dicts(42)))
    for n, d in dicts.items():
        assert list(d.keys()) == G["state_keys"][n]
        assert {k: tuple(v.shape) for k, v in d.items()} == G["state_shapes"][n]
        chk = float(sum(v.double().abs().sum() for k, v in d.items() if not k.endswith("window")))
        assert abs(chk - G["param_checksum"][n]) <= 1e-9 * G["param_checksum"][n], n


def test_oracle_step_at_baseline_shape_matches_golden_v3():
    """BASELINE configs[0]/[2] shape (batch 8 x 32000 samples, SURVEY.md section 8c known answers): the first oracle step
    from the stock-container initial weights reproduces the reference's train_one_epoch log (golden_v3.pt, written by
    make_golden.py --logs-only) and the discriminator / clipped enhancer gradient norms at the optimiser steps."""
    O = oracle()
    from oracle import ref_init
    G3 = torch.load(os.path.join(os.path.dirname(GOLD), "golden_v3.pt"), weights_only=False)
    r = G3["input_recipe"]
    assert (r["batch"], r["samples"], r["seed"]) == (8, 32000, 1234)
    noisy, clean = O.synthetic_batch(r["batch"], r["samples"], seed=r["seed"])
    assert abs(float(noisy.double().sum()) - G3["input_checksum"][0]) < 1e-6
    # SURVEY 8c known answers are exactly what the fixture holds
    assert G3["train_ls"]["logs"][0] == {"D_loss": 0.9975, "G_loss": 2.9402, "MR": 2.715, "Mask": 0.2173, "Adv": 0.7964,
                                         "FM": 0.005}
    assert G3["train_hinge"]["logs"][0]["D_loss"] == 2.0 and G3["train_hinge"]["logs"][0]["Adv"] == -0.0025
    Pe, Pp, Ps = ref_init.init_state_dicts(42)
    st = O.StepState(Pe, Pp, Ps, order_g=ref_init.param_order(Pe),
                     order_d=(ref_init.param_order(Pp), ref_init.param_order(Ps)))
    got = O.train_step(st, noisy, clean, [O.hann_window(n) for n in O.MR_FFT_SIZES], gan_loss="ls", aten_gru=True)
    names = {"D_loss": "d_loss", "G_loss": "g_loss", "MR": "mr", "Mask": "mask", "Adv": "adv", "FM": "fm"}
    for ref_k, k in names.items():
        assert abs(got[k] - G3["train_ls"]["logs"][0][ref_k]) <= 1.01e-4, k
    assert abs(got["g_grad_norm"] - 0.0) > 5.0          # clip active: pre-clip norm above the 5.0 threshold
    assert abs(G3["train_ls"]["grad_stats"]["g"][0]["l2"] - 5.0) < 1e-5


def test_known_answers_from_survey():
    """integer framing of the discriminators and the AvgPool known answer."""
    O = oracle()
    assert [(p - 32000 % p) % p for p in O.MPD_PERIODS] == [0, 1, 0, 4, 10]
    x = torch.arange(10, dtype=torch.float32).view(1, 1, 10)
    assert torch.equal(O.msd_pool(x).view(-1), torch.tensor([0.5, 1.5, 3.5, 5.5, 7.5, 8.5]))
    L = 32000
    lens = [L]
    for _ in range(2):
        lens.append(lens[-1] // 2 + 1)
    assert lens == [32000, 16001, 8001]
    idx = O.period_indices(32000, 7)
    assert idx.shape == (4572, 7) and idx[-1].tolist() == [31997, 31998, 31999, 31998, 31997, 31996, 31995]
    g, v = torch.rand(8, 1, 1) + 0.5, torch.randn(8, 4, 5)
    w = O.weight_norm_weight(g, v)
    assert torch.allclose(w.flatten(1).norm(dim=1), g.flatten(), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(os.environ.get("LCT_REF", "/root/reference")),
                    reason="reference checkout not present (it does not travel to the GPU box)")
def test_oracle_against_live_reference():
    """When the reference checkout is reachable, run it directly: different seed and length than the fixtures."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("datasets", "models", "losses", "train")}
    path0 = list(sys.path)
    try:
        R_stft, R_tf, R_losses, R_disc, R_gen, R_train = mg.import_reference()
        O = oracle()
        torch.manual_seed(7)
        enh = R_gen.LCTEnhancer(R_gen.LCTGeneratorConfig(), c=0.3)
        msd = R_disc.MultiScaleDiscriminator(num_scales=2)
        noisy, clean = O.synthetic_batch(1, 5000, seed=99)
        with torch.no_grad():
            e_ref, m_ref = enh(noisy)
            sl_ref, _ = msd(clean)
            P = {k: v.detach().clone() for k, v in enh.state_dict().items()}
            e, m = O.enhancer_forward(P, noisy)
            sl, _ = O.msd_forward({k: v.detach().clone() for k, v in msd.state_dict().items()}, clean, num_scales=2)
        assert rel_err(e, e_ref) < 2e-5 and rel_err(m, m_ref) < 2e-5
        for a, b in zip(sl, sl_ref):
            assert rel_err(a, b) < 1e-5
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("datasets", "models", "losses", "train")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = path0


@pytest.mark.skipif(not os.path.isdir(os.environ.get("LCT_REF", "/root/reference")),
                    reason="reference checkout not present (it does not travel to the GPU box)")
def test_oracle_tail_and_loader_helpers_against_live_reference():
    """The oracle restatements of the callers either side of the path against the reference's own
    code: `_si_sdr_torch` (train.py:261-282), `LCTScpDataset._crop_pair` (datasets.py:131-156), `collate_fn` (:187-230)."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLD), "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("datasets", "models", "losses", "train")}
    path0 = list(sys.path)
    try:
        R_stft, R_tf, R_losses, R_disc, R_gen, R_train = mg.import_reference()
        import datasets.datasets as R_data
        O = oracle()
        g = torch.Generator().manual_seed(5)
        ref, est = torch.randn(30000, generator=g) * 0.1 + 0.01, torch.randn(29000, generator=g) * 0.1
        est = est + 0.7 * ref[:29000]
        assert abs(O.si_sdr(ref, est) - R_train._si_sdr_torch(ref, est)) < 1e-4
        ds = R_data.LCTScpDataset.__new__(R_data.LCTScpDataset)            # no files: only the cropping method is used
        items = [(torch.randn(50000, generator=g), torch.randn(50000, generator=g)),
                 (torch.randn(20000, generator=g), torch.randn(20000, generator=g)),
                 (torch.randn(40000, generator=g), torch.randn(39000, generator=g))]
        for random_segment in (True, False):
            ds.segment_length, ds.random_segment = 32000, random_segment
            torch.manual_seed(9)
            want = [ds._crop_pair(a, c) for a, c in items]
            torch.manual_seed(9)
            got = [O.crop_pair(a, c, 32000, random_segment) for a, c in items]
            for (wa, wc), (ga, gc) in zip(want, got):
                assert torch.equal(wa, ga) and torch.equal(wc, gc)
        batch = R_data.collate_fn([{"id": str(i), "noisy": a, "clean": c, "sr": 16000} for i, (a, c) in enumerate(want)])
        mine = O.collate(want)
        for k in ("noisy", "clean", "lengths"):
            assert torch.equal(batch[k], mine[k]), k
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("datasets", "models", "losses", "train")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = path0
