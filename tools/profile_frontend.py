"""One launch of every front-end kernel at the top of the BASELINE sweep (B = 128, 4 s) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops
dev = torch.device("cuda:0")
B, T = int(os.environ.get("FB", 128)), int(os.environ.get("FT", 64000))
x = torch.randn(B, T, device=dev) * 0.1; y = torch.randn(B, T, device=dev) * 0.1
n_fft, hop = 512, 256
w = torch.hann_window(n_fft, device=dev)
Tf, F = 1 + T // hop, n_fft // 2 + 1
mask = torch.rand(B, Tf, F, device=dev) * 0.5 + 0.5
acc = torch.zeros(2, 64, device=dev)
gy = torch.randn(B, T, device=dev)
for it in range(2):
    spec, mag = ops.stft_fwd(x, w, n_fft, hop, want_mag=True)
    ops.tf_features_fwd(x, y, w, n_fft, hop)
    ops.istft_fwd(spec, w, n_fft, hop, T, mask_c=mask)
    ops.istft_bwd(gy, w, n_fft, hop, Tf, xspec=spec, mask_c=mask, want_gspec=False)
    for n, h in ((320, 160), (512, 256), (768, 384)):
        ops.mrstft_sums(x, y, torch.hann_window(n, device=dev), n, h, acc)
    torch.cuda.synchronize()
print("done")
