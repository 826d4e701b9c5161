"""Callers on either side of the hot path on the GPU against the CPU oracle
restatements (oracle.si_sdr / crop_pair / collate, pinned against the reference's own functions by
tests/test_oracle_golden.py::test_oracle_tail_and_loader_helpers_against_live_reference).
Integer work (cropping, padding, bucketing) is bit exact; SI-SDR within 1e-3 dB."""
import pytest
import torch

from util import cpu_params, oracle, rel_err

pytestmark = pytest.mark.gpu


def test_si_sdr_batched_matches_reference_formula(dev):
    from lctgan.inference import si_sdr
    O = oracle()
    g = torch.Generator().manual_seed(2)
    B, T = 5, 40000
    ref = torch.randn(B, T, generator=g) * 0.1 + 0.003                    # a small DC offset: the means matter
    est = 0.6 * ref + torch.randn(B, T, generator=g) * torch.tensor([0.01, 0.05, 0.2, 1.0, 0.0]).view(B, 1)
    lengths = torch.tensor([40000, 12345, 1, 39999, 20000])
    got = si_sdr(ref.to(dev), est.to(dev), lengths.to(dev)).cpu()
    for b in range(B):
        want = O.si_sdr(ref[b, :lengths[b]], est[b, :lengths[b]])
        assert abs(got[b].item() - want) < 1e-3 * max(1.0, abs(want) / 50), (b, got[b].item(), want)
    # no lengths + different widths: compared over the common prefix (train.py:267-269)
    got2 = si_sdr(ref.to(dev), est[:, :30000].contiguous().to(dev)).cpu()
    for b in range(B):
        assert abs(got2[b].item() - O.si_sdr(ref[b], est[b, :30000])) < 1e-3 * max(1.0, abs(got2[b].item()) / 50)


@pytest.mark.parametrize("random_segment", [True, False])
def test_segment_sampler_equals_reference_crop_and_collate(dev, random_segment):
    """One crop kernel over the device cache == collate_fn([dataset[i] ...]) of the reference: bit exact, same draws."""
    from lctgan.pipeline import SegmentSampler, UtteranceCache
    O = oracle()
    g = torch.Generator().manual_seed(4)
    lens = [(50000, 50000), (20000, 20000), (40000, 39000), (32000, 32000), (32001, 32007), (90001, 90001)]
    noisy = [torch.randn(n, generator=g) for n, _ in lens]
    clean = [torch.randn(m, generator=g) for _, m in lens]
    cache = UtteranceCache(noisy, clean, dev)
    for indices in ([0, 1, 2, 3], [5, 4, 2], [1, 3], [0, 5, 0, 5, 2, 2, 4, 4]):
        g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
        got = SegmentSampler(cache, 32000, random_segment, g1).next_batch(indices)
        want = O.collate([O.crop_pair(noisy[i], clean[i], 32000, random_segment, g2) for i in indices])
        assert torch.equal(got["noisy"].cpu(), want["noisy"]) and torch.equal(got["clean"].cpu(), want["clean"])
        assert torch.equal(got["lengths"], want["lengths"])
    # no cropping configured: plain collate_fn padding
    got = SegmentSampler(cache, None).next_batch([2, 1])
    want = O.collate([(noisy[2], clean[2]), (noisy[1], clean[1])])
    assert torch.equal(got["noisy"].cpu(), want["noisy"]) and torch.equal(got["clean"].cpu(), want["clean"])


def test_bucketed_enhancement_equals_padded_batches(dev):
    """enhance_utterances: every bucket is a collate_fn batch (zero padded to its longest utterance, results cropped):
    identical to calling the enhancer on that padded batch, close to the CPU oracle on it, and every utterance comes
    back once, in input order, at its true length."""
    from lctgan.inference import enhance_utterances, length_buckets
    from models.generator import LCTEnhancer, LCTGeneratorConfig
    O = oracle()
    torch.manual_seed(3)
    enh = LCTEnhancer(LCTGeneratorConfig(), c=0.3).eval()
    P = cpu_params(enh)
    enh = enh.to(dev)
    g = torch.Generator().manual_seed(9)
    lens = [9000, 30000, 9500, 29000, 16000, 8800]
    waves = [torch.randn(n, generator=g) * 0.1 for n in lens]
    outs = enhance_utterances(enh, waves, max_batch=4, max_pad_ratio=1.1)
    assert [o.shape[0] for o in outs] == lens
    for bucket in length_buckets(lens, 4, 1.1):
        T = max(lens[i] for i in bucket)
        x = torch.zeros(len(bucket), T)
        for r, i in enumerate(bucket):
            x[r, :lens[i]] = waves[i]
        with torch.no_grad():
            y = enh(x.to(dev))[0].cpu()
            yo, _ = O.enhancer_forward(P, x, aten_gru=True)
        for r, i in enumerate(bucket):
            assert torch.equal(outs[i], y[r, :lens[i]])
            assert rel_err(outs[i], yo[r, :lens[i]]) < 5e-5


def test_validate_mirror_and_prefetcher(dev):
    from lctgan.inference import si_sdr, validate
    from lctgan.pipeline import PinnedPrefetcher
    from lctgan.training import build_models
    O = oracle()
    enh, _, _, _, mr, _, _ = build_models(dev, gan_seed=1)
    g = torch.Generator().manual_seed(12)
    batches = []
    for B, T in ((3, 8000), (2, 12000)):
        clean = torch.randn(B, T, generator=g) * 0.1
        noisy = clean + torch.randn(B, T, generator=g) * 0.05
        batches.append({"noisy": noisy, "clean": clean, "lengths": torch.tensor([T - 100 * b for b in range(B)])})
    got = validate(enh, mr, PinnedPrefetcher(batches, dev))
    # the same quantities composed by hand: sample-weighted MR-STFT, mean of the per-utterance SI-SDRs (train.py:313-334)
    tot_mr, tot_sdr, n = 0.0, 0.0, 0
    with torch.no_grad():
        for b in batches:
            e, _ = enh.eval()(b["noisy"].to(dev))
            tot_mr += mr(e, b["clean"].to(dev))[0].item() * e.shape[0]
            for r in range(e.shape[0]):
                L = int(b["lengths"][r])
                tot_sdr += O.si_sdr(b["clean"][r, :L], e[r, :L].cpu())
            n += e.shape[0]
    assert abs(got["val_mrstft"] - tot_mr / n) < 1e-5 * max(1.0, tot_mr / n)
    assert abs(got["val_si_sdr"] - tot_sdr / n) < 1e-3
    assert got["val_pesq"] != got["val_pesq"] and got["val_stoi"] != got["val_stoi"]        # NaN: third-party, absent
    # the prefetcher hands out the loader's batches unchanged, in order, on the device
    for want, have in zip(batches, PinnedPrefetcher(batches, dev)):
        assert have["noisy"].is_cuda and torch.equal(have["noisy"].cpu(), want["noisy"])
        assert torch.equal(have["clean"].cpu(), want["clean"]) and torch.equal(have["lengths"], want["lengths"])


def test_memset_zero_any_alignment_and_length(dev):
    """lct_memset_zero (the library's own clearing kernel): every byte of [p, p + bytes), nothing outside, for unaligned
    starts and lengths around its 16-byte store width."""
    from lctgan import _lib
    buf = torch.empty(4096 + 64, dtype=torch.uint8, device=dev)
    for off in (0, 1, 3, 4, 15, 16, 17):
        for n in (1, 2, 15, 16, 17, 31, 33, 255, 1024, 4000):
            buf.fill_(0xAB)
            _lib.call("lct_memset_zero", buf.data_ptr() + off, n)
            h = buf.cpu()
            assert int(h[off:off + n].max()) == 0, (off, n)
            assert bool((h[:off] == 0xAB).all()) and bool((h[off + n:] == 0xAB).all()), (off, n)
    big = torch.full((3 * 1024 * 1024 + 5,), 7.0, device=dev)
    _lib.call("lct_memset_zero", big[1:], (big.numel() - 1) * 4)
    assert float(big[0]) == 7.0 and float(big[1:].abs().max()) == 0.0


def test_axpby_scalar_and_vector_forms(dev):
    """lct_axpby (out = ka a + kb b; b optional; out may alias an input): the 16-byte form for large aligned buffers, the
    element form otherwise - exact against torch in fp32."""
    from lctgan import _lib
    g = torch.Generator().manual_seed(5)
    for n in (1, 7, 4095, 4096, 4099, 100003):
        for off in (0, 1):
            a = torch.randn(n + 4, generator=g).to(dev)[off:off + n]
            b = torch.randn(n + 4, generator=g).to(dev)[off:off + n]
            want = 0.5 * a - 2.0 * b
            out = torch.empty(n + 4, device=dev)[off:off + n]
            _lib.call("lct_axpby", a.data_ptr(), b.data_ptr(), out.data_ptr(), n, 0.5, -2.0)
            assert torch.equal(out, want), (n, off)
            _lib.call("lct_axpby", a.data_ptr(), b.data_ptr(), a.data_ptr(), n, 0.5, -2.0)      # in place
            assert torch.equal(a, want), (n, off)
            _lib.call("lct_axpby", b.data_ptr(), None, out.data_ptr(), n, 3.0, 0.0)
            assert torch.equal(out, 3.0 * b), (n, off)
