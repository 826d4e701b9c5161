"""ncu driver: enhancer forward + MR-STFT / mask loss + backward (B = 8, 2 s), a few iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan.training import build_models
import losses as L
from lctgan import training as O          # synthetic_batch lives with the product
dev = torch.device("cuda:0")
enh, mpd, msd, tf, mr, g_opt, d_opt = build_models(dev, 42)
noisy, clean = (t.to(dev) for t in O.synthetic_batch(8, 32000))
for it in range(3):
    for p in enh.parameters(): p.grad = None
    e, m = enh(noisy)
    l, _ = mr(e, clean)
    (l + L.mask_mse_loss(m[:, 0], m[:, 0].detach() * 0.9)).backward()
torch.cuda.synchronize()
print("ok")
