"""CPU-side checks (no GPU needed): the C-ABI shared library loads and exports every symbol declared in
include/lctgan.h, the host-side mirror of the reference interface behaves like the reference (constructors,
argument validation, state_dict layout), nothing silently falls back to the CPU, and the data-parallel
gradient exchange is correct on a world_size-2 gloo group."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lctgan import _lib
    L = _lib.lib()
    assert os.path.exists(_lib.LIB_PATH)
    assert len(L.decls) >= 55
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in L.decls:
        assert hasattr(cdll, name), name
    # ... and nothing is exported that the header does not declare (the header is the contract)
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lct_\w+)", out))
    assert exported == set(L.decls), exported ^ set(L.decls)
    assert _lib.call_ret("lct_version") >= 1
    assert _lib.call_ret("lct_mt_max_segments") == 64
    assert _lib.call_ret("lct_fft_supported", 512) == 1 and _lib.call_ret("lct_fft_supported", 320) == 1
    assert _lib.call_ret("lct_fft_supported", 768) == 1 and _lib.call_ret("lct_fft_supported", 514) == 0


def test_header_signatures_parse():
    from lctgan import _lib
    decls = _lib.parse_header()
    assert [t for t, _ in decls["lct_stft_fwd"]][:5] == ["const float*"] * 3 + ["float*"] * 2
    assert decls["lct_stft_fwd"][-1][0] == "cudaStream_t"
    assert decls["lct_version"] == []
    for name, args in decls.items():
        for t, _ in args:
            assert "*" in t or t.replace("const ", "") in ("int", "int64_t", "float", "cudaStream_t"), (name, t)


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments with a negative code before touching the device."""
    from lctgan import _lib
    fn = _lib.lib().fns["lct_stft_fwd"][0]
    assert fn(None, None, None, None, None, 1, 4000, 512, 256, 1e-12, None) == -1
    fn = _lib.lib().fns["lct_conv1d_fwd"][0]
    assert fn(None, None, None, None, 1, 1, 1, 1, 3, 1, 1, 10, 1, 0, 0.2, None) == -1
    fn = _lib.lib().fns["lct_gemm"][0]
    assert fn(*([None] * 6), 0, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0.2, 1.0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, None) == -1


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors (there is no PyTorch/CPU fallback)."""
    from datasets.stft import ComplexSTFT, STFTConfig
    from datasets.tf_features import TFFeatures
    from models.discriminators import MultiPeriodDiscriminator, MultiScaleDiscriminator
    from models.generator import LCTEnhancer, LCTGeneratorConfig
    import losses as L
    x = torch.zeros(1, 4000)
    with pytest.raises(RuntimeError):
        ComplexSTFT(STFTConfig())(x)
    with pytest.raises(RuntimeError):
        TFFeatures()(x, x)
    with pytest.raises(RuntimeError):
        LCTEnhancer(LCTGeneratorConfig())(x)
    with pytest.raises(RuntimeError):
        MultiPeriodDiscriminator()(x)
    with pytest.raises(RuntimeError):
        MultiScaleDiscriminator()(x)
    with pytest.raises(RuntimeError):
        L.mask_mse_loss(torch.zeros(2, 3, 4), torch.zeros(2, 3, 4))
    with pytest.raises(RuntimeError):
        L.MultiResolutionSTFTLoss()(x, x)


def test_reference_interface_mirror():
    """Constructors, defaults and validation errors follow the reference."""
    from datasets.stft import ComplexSTFT, STFTConfig, make_lct_stft
    from datasets.tf_features import TFFeatures, TFFeaturesConfig
    from models.generator import (Decoder, DownBlock, Encoder, GRUblockf, GRUblockt, LCTEnhancer, LCTGenerator,
                                  LCTGeneratorConfig, UpBlock)
    from models.discriminators import (MultiPeriodDiscriminator, MultiScaleDiscriminator, PeriodDiscriminator,
                                       ScaleDiscriminator)
    import losses as L
    c = STFTConfig().finalize()
    assert (c.n_fft, c.hop_length, c.win_length, c.window, c.center, c.pad_mode, c.normalized, c.onesided) == \
        (512, 256, 512, "hann", True, "reflect", False, True)
    assert make_lct_stft().window.shape == (512,)
    assert list(ComplexSTFT(STFTConfig(n_fft=320)).state_dict().keys()) == ["window"]
    with pytest.raises(ValueError):
        ComplexSTFT(STFTConfig(window="hamming"))
    t = TFFeaturesConfig()
    assert (t.n_fft, t.c, t.compress_input, t.return_stfts) == (512, 0.3, False, True)
    g = LCTGeneratorConfig()
    assert (g.enc_channels, g.dec_channels, g.num_heads, g.gru_groups, g.max_time_context, g.output_activation) == \
        ((16, 32, 64), (64, 32, 16), 4, 4, None, "sigmoid")
    gen = LCTGenerator(LCTGeneratorConfig(num_heads=8, gru_groups=2, max_time_context=64))   # dead config accepted
    assert sum(p.numel() for p in gen.parameters()) == 135425
    assert sum(p.numel() for p in MultiPeriodDiscriminator().parameters()) == 785770
    assert sum(p.numel() for p in MultiScaleDiscriminator().parameters()) == 16924086
    with pytest.raises(AssertionError):
        GRUblockf(32)
    with pytest.raises(AssertionError):
        MultiScaleDiscriminator(num_scales=0)
    for cls in (DownBlock(1, 4), UpBlock(4, 4, 2), Encoder(), Decoder(1), GRUblockt(64), PeriodDiscriminator(3),
                ScaleDiscriminator()):
        assert isinstance(cls, torch.nn.Module)
    m = L.MRSTFTLossConfig()
    assert (m.fft_sizes, m.hop_factors, m.main_fft_size, m.main_fft_weight) == ((320, 512, 768), (.5, .5, .5), 512, 2.0)
    mr = L.MultiResolutionSTFTLoss()
    assert mr.weights == [1.0, 2.0, 1.0] and list(mr.state_dict().keys()) == [f"stfts.{i}.window" for i in range(3)]
    assert [s.cfg.hop_length for s in mr.stfts] == [160, 256, 384]
    # errors raised before any kernel is involved
    with pytest.raises(ValueError):
        L.discriminator_loss([torch.zeros(1)], [], "ls")
    with pytest.raises(ValueError):
        L.feature_matching_loss([[torch.zeros(1)]], [])
    with pytest.raises(ValueError):
        L.mask_mse_loss(torch.zeros(2, 3), torch.zeros(2, 4))
    assert L._flatten_logits_lists([1, 2], [3]) == [1, 2, 3]
    with pytest.raises(ValueError):
        LCTEnhancer(LCTGeneratorConfig())(torch.zeros(4000))
    # the generator's parameter order is the reference's registration order
    names = [n for n, _ in gen.named_parameters()]
    assert names[:6] == ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"]
    assert names[-2:] == ["layernorm.weight", "layernorm.bias"]
    d = PeriodDiscriminator(2)
    assert [n for n, _ in d.named_parameters()][:3] == ["convs.0.bias", "convs.0.weight_g", "convs.0.weight_v"]


def test_precision_switch():
    from lctgan import config
    config.set_precision("fp32")
    assert config.dense_tensor_cores is False
    config.set_precision("bf16")
    assert config.dense_tensor_cores is True
    with pytest.raises(ValueError):
        config.set_precision("fp8")


_DP_WORKER = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch, torch.distributed as dist
from lctgan.parallel import FlatGradAllReduce, broadcast_parameters
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(100 + rank)                      # different initial weights per rank ...
net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
broadcast_parameters([net])                        # ... made identical by the broadcast
ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
ref.load_state_dict(net.state_dict())
g = torch.Generator().manual_seed(7)
X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 1, generator=g)
xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]          # this rank's shard of the batch
((net(xs) - ys) ** 2).mean().backward()
net[2].bias.grad = None                              # a parameter without gradient still takes part
sync = FlatGradAllReduce(net.parameters())
sync()
((ref(X) - Y) ** 2).mean().backward()                 # single-process full batch = mean over ranks
for (n, p), q in zip(net.named_parameters(), ref.parameters()):
    if n == "2.bias":
        assert torch.allclose(p.grad, torch.zeros_like(p.grad)), n
    else:
        assert torch.allclose(p.grad, q.grad, atol=1e-6), (n, p.grad, q.grad)
w = [torch.zeros_like(net[0].weight) for _ in range(world)]
dist.all_gather(w, net[0].weight.data)
assert torch.equal(w[0], w[1])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_data_parallel_gradient_exchange_gloo_world2(tmp_path):
    """N ranks x B/N samples == 1 rank x B samples: averaged gradients are identical."""
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER % (os.path.join(ROOT, "lct-gan_b200"), ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


def test_length_buckets_and_crop_starts_host_logic():
    """Host logic of the N3 / N4 rows: buckets cover every utterance once with bounded padding; the crop starts follow
    `_crop_pair` (datasets/datasets.py:131-156) draw for draw - checked against the oracle restatement, which
    tests/test_oracle_golden.py pins against the reference's own class."""
    import torch
    from lctgan.inference import length_buckets
    from lctgan.pipeline import crop_starts
    from util import oracle
    O = oracle()
    lens = [16000, 160000, 48000, 47000, 16001, 90000, 91000, 30000, 15999]
    b = length_buckets(lens, max_batch=3, max_pad_ratio=1.1)
    assert sorted(i for g in b for i in g) == list(range(len(lens)))
    for g in b:
        assert len(g) <= 3 and max(lens[i] for i in g) <= 1.1 * min(lens[i] for i in g)
    assert length_buckets([5, 5, 5, 5], max_batch=16) == [[0, 1, 2, 3]]
    # default policy: minimum of (padded samples processed + a fixed cost per enhancer call).  With a huge call cost
    # everything that fits goes into one batch; with none, equal lengths still share a batch and nothing is padded
    one = length_buckets(lens, max_batch=16, call_overhead=10 ** 9)
    assert len(one) == 1 and sorted(one[0]) == list(range(len(lens)))
    free = length_buckets(lens + [16000], max_batch=16, call_overhead=0)
    assert sorted(i for g in free for i in g) == list(range(len(lens) + 1))
    assert all(len({(lens + [16000])[i] for i in g}) == 1 for g in free)
    # optimality against brute force over all contiguous splits of the sorted order (small case)
    import itertools
    ls = [9, 7, 7, 4, 3, 3, 1]
    def cost(groups, ov):
        return sum(ov + len(g) * max(ls[i] for i in g) for g in groups)
    for ov in (0, 2, 5, 50):
        got = length_buckets(ls, max_batch=4, call_overhead=ov)
        bestc = min(cost([list(range(a, b)) for a, b in zip((0,) + cuts, cuts + (len(ls),))], ov)
                    for r in range(len(ls)) for cuts in itertools.combinations(range(1, len(ls)), r)
                    if all(b - a <= 4 for a, b in zip((0,) + cuts, cuts + (len(ls),))))
        assert cost(got, ov) == bestc and all(len(g) <= 4 for g in got), (ov, got)
    # crop starts: same generator stream as the reference's per-item draws
    gen = torch.Generator().manual_seed(11)
    items = [(torch.randn(n, generator=gen), torch.randn(m, generator=gen))
             for n, m in ((50000, 50000), (20000, 20000), (40000, 39000), (32000, 32000), (32001, 32005))]
    ln = torch.tensor([a.shape[-1] for a, _ in items]); lc = torch.tensor([c.shape[-1] for _, c in items])
    for random_segment in (True, False):
        g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
        starts = crop_starts(ln, lc, 32000, random_segment, g1)
        for i, (a, c) in enumerate(items):
            ca, cc = O.crop_pair(a, c, 32000, random_segment, g2)
            s = int(starts[i])
            if min(a.shape[-1], c.shape[-1]) <= 32000:
                assert s == 0 and ca.shape == a.shape
            else:
                assert torch.equal(ca, a[s:s + 32000]) and torch.equal(cc, c[s:s + 32000])
    assert crop_starts(ln, lc, None, True).tolist() == [0] * 5


_DP_HOOK_WORKER = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch, torch.distributed as dist
from lctgan.parallel import BackwardEndExchange, broadcast_parameters
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(5)
G = torch.nn.Linear(6, 6)                       # stands for the enhancer
D = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))      # a discriminator
broadcast_parameters([G, D])
Gr, Dr = torch.nn.Linear(6, 6), torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
Gr.load_state_dict(G.state_dict()); Dr.load_state_dict(D.state_dict())
ex = BackwardEndExchange(G, [D])
gen = torch.Generator().manual_seed(3)
X = torch.randn(8, 6, generator=gen); Y = torch.randn(8, 6, generator=gen)
xs, ys = X[rank * 4:(rank + 1) * 4], Y[rank * 4:(rank + 1) * 4]
# --- the loop body of train_one_epoch, unmodified in structure (train.py:177-249): D step ...
with torch.no_grad():
    fake = G(xs)
d_loss = ((D(ys) - 1) ** 2).mean() + (D(fake) ** 2).mean()
d_loss.backward()                                # <- the exchange fires here, at the end of this backward
assert ex.exchanges == {"g": 0, "d": 1}, ex.exchanges
with torch.no_grad():
    fr = Gr(X)
(((Dr(Y) - 1) ** 2).mean() + (Dr(fr) ** 2).mean()).backward()
for p, q in zip(D.parameters(), Dr.parameters()):
    assert torch.allclose(p.grad, q.grad, atol=1e-6)
d_before = [p.grad.clone() for p in D.parameters()]
# --- ... G step: the backward reaches G AND D; only G is exchanged (the D gradients it produces are dead)
g_loss = ((D(G(xs)) - 1) ** 2).mean()
g_loss.backward()
assert ex.exchanges == {"g": 1, "d": 1}, ex.exchanges
((Dr(Gr(X)) - 1) ** 2).mean().backward()
for p, q in zip(G.parameters(), Gr.parameters()):
    assert torch.allclose(p.grad, q.grad, atol=1e-6)            # averaged BEFORE clip_grad_norm_ would run
torch.nn.utils.clip_grad_norm_(G.parameters(), 5.0)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_backward_end_exchange_for_the_unmodified_loop_gloo_world2(tmp_path):
    """lctgan.parallel.BackwardEndExchange: data parallelism for the reference's unmodified train_one_epoch - the
    exchange fires from an end-of-backward autograd callback (after d_loss.backward(), and after g_loss.backward()
    before the clip), generator passes never exchange the dead discriminator gradients."""
    script = tmp_path / "dp_hook_worker.py"
    script.write_text(_DP_HOOK_WORKER % (os.path.join(ROOT, "lct-gan_b200"), ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
