"""Time the building blocks of the step as individually captured CUDA graphs (B = 8, 2 s): where does the step go?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import _lib
from lctgan.training import build_models, StepArgs, _phase_d, _phase_g, _phase_opt_g
import losses as L
from lctgan import training as O

dev = torch.device("cuda:0")
enh, mpd, msd, tf, mr, g_opt, d_opt = build_models(dev, 42, capturable=True)
noisy, clean = (t.to(dev) for t in O.synthetic_batch(8, 32000))


def graph_time(fn, name, iters=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    n0 = _lib.kernel_launches()
    with torch.cuda.graph(g): fn()
    n = _lib.kernel_launches() - n0
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / iters:8.3f} ms   ({n} lctgan launches)", flush=True)


def zero(ms):
    for m in ms:
        for p in m.parameters(): p.grad = None

with torch.no_grad():
    fake, _ = enh(noisy)
fake = fake.detach()

def f_enh_fwd():
    with torch.no_grad(): enh(noisy)
def f_enh_fwd_bwd():
    zero([enh]); e, m = enh(noisy); (e.sum() + m.sum()).backward()
def f_enh_fwd_mr_bwd():
    zero([enh]); e, m = enh(noisy); l, _ = mr(e, clean); (l + L.mask_mse_loss(m[:, 0], m[:, 0].detach() * 0.9)).backward()
def f_mpd_fwd():
    with torch.no_grad(): mpd(clean)
def f_msd_fwd():
    with torch.no_grad(): msd(clean)
def f_mpd_fwd_bwd():
    zero([mpd]); lg, _ = mpd(clean); L.generator_adv_loss(lg, "ls").backward()
def f_msd_fwd_bwd():
    zero([msd]); lg, _ = msd(clean); L.generator_adv_loss(lg, "ls").backward()
xg = fake.clone().requires_grad_(True)
def f_d_gstep():
    zero([mpd, msd]); xg.grad = None
    with torch.no_grad():
        _, rf1 = mpd(clean); _, rf2 = msd(clean)
    l1, f1 = mpd(xg); l2, f2 = msd(xg)
    (L.generator_adv_loss(l1 + l2, "ls") + L.feature_matching_loss(rf1 + rf2, f1 + f2)).backward()
def f_tf():
    tf(noisy, clean)
def f_opt():
    d_opt.step(); g_opt.step()

graph_time(f_tf, "TFFeatures")
graph_time(f_enh_fwd, "enhancer fwd (no grad)")
graph_time(f_enh_fwd_bwd, "enhancer fwd+bwd (sum loss)")
graph_time(f_enh_fwd_mr_bwd, "enhancer fwd + MRSTFT + mask loss + bwd")
graph_time(f_mpd_fwd, "MPD fwd (no grad)")
graph_time(f_msd_fwd, "MSD fwd (no grad)")
graph_time(f_mpd_fwd_bwd, "MPD fwd+bwd (params)")
graph_time(f_msd_fwd_bwd, "MSD fwd+bwd (params)")
graph_time(f_d_gstep, "G-step D part: real fwd x2, fake fwd+bwd, FM")
f_msd_fwd_bwd(); f_mpd_fwd_bwd(); f_enh_fwd_bwd()
graph_time(f_opt, "AdamW D + G step")
M = (enh, mpd, msd, tf, mr, g_opt, d_opt)
args = StepArgs(reuse_enhancer_forward=True)
st = {}
def f_pd(): _phase_d(M, noisy, clean, args, st)
def f_all():
    _phase_d(M, noisy, clean, args, st); _phase_g(M, noisy, clean, args, st); _phase_opt_g(M, args)
graph_time(f_pd, "phase D (tf + enh fwd || D fwd x4 + bwd)")
graph_time(f_all, "whole step")
