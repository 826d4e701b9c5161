"""Host-side profile (cProfile, top functions by own time) of the drop-in modules driven eagerly, the way the unmodified
train.py drives them (literal schedule, no CUDA graph): where the ~40 ms per step of the `api_eager` bench figure go.
    python tools/profile_eager_host.py"""
import sys, os, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan.training import StepArgs, build_models, synthetic_batch, train_step
dev = torch.device("cuda:0")
noisy, clean = (t.to(dev) for t in synthetic_batch(8, 32000, seed=1234))
M = build_models(dev, gan_seed=42)
run = lambda: train_step(*M, noisy, clean, StepArgs(gan_loss="ls"))
for _ in range(3): run()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35); print(s.getvalue()[:6000])
