"""Bring-up diagnostic: warp-FFT kernels vs the generic shared-memory kernels and the oracle, per configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops, _lib
from oracle import lct_oracle as O
dev = torch.device("cuda:0")
def rel(a, b): return float((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max())
for n_fft, hop in ((512, 256), (320, 160), (768, 384)):
    for T, length in ((4096, 4096), (4000, 4000), (4000, 3000), (4096, 4300), (32000, 32000)):
        g = torch.Generator().manual_seed(2)
        x = torch.randn(2, T, generator=g) * 0.1
        w = O.hann_window(n_fft)
        spec = O.stft(x, w, n_fft, hop)
        phys = spec.transpose(1, 2).contiguous().to(dev)
        _lib.call_ret("lct_fft_force_generic", 1)
        ref = ops.istft_fwd(phys, w.to(dev), n_fft, hop, length)
        _lib.call_ret("lct_fft_force_generic", 0)
        got = ops.istft_fwd(phys, w.to(dev), n_fft, hop, length)
        d = (got - ref).abs().max(dim=0).values
        bad = torch.nonzero(d > 1e-5 * ref.abs().max()).flatten()
        print(f"istft N={n_fft} T={T} len={length}: rel {rel(got, ref):.2e}  bad idx: {bad[:6].tolist()} .. {bad[-6:].tolist()} n={bad.numel()}", flush=True)
