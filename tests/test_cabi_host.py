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
