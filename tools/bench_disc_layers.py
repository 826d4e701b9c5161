"""Grouped discriminator convolutions timed layer by layer from CUDA graphs of 20 launches, on the shapes of the D step
(batch 2B = 16) or the G step (B = 8):  python tools/bench_disc_layers.py [B] [tc|mma]
(tc = tcgen05 kernels of conv_tc.cu, the default; mma = the round-1 mma.sync kernels of conv_mma.cu)

Prints per layer: forward / data-gradient (with the fused FM-gradient + LeakyReLU' epilogue) / weight-gradient time and
the achieved GB/s on the algorithmic bytes (fwd: x + y; dgrad: dy + gextra + xact + dx; wgrad: x + dy)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import config, ops
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
config.grouped_conv_tcgen05 = not (len(sys.argv) > 2 and sys.argv[2] == "mma")
config.tc_dgrad_periods = len(sys.argv) > 2 and sys.argv[2] == "tcall"
print("kernels:", "tcgen05 (conv_tc.cu)" if config.grouped_conv_tcgen05 else "mma.sync (conv_mma.cu)", flush=True)


def gtime(run, n=20):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


# (name, Cin, Cout, K, S, groups, Lin, P)
LAYERS = [
    ("MSD0.convs0", 1, 16, 15, 1, 1, 32000, 1),
    ("MPD2.convs0", 1, 32, 5, 3, 1, 16000, 2),
    ("MPD11.convs0", 1, 32, 5, 3, 1, 2910, 11),
    ("MSD0.convs1", 16, 64, 41, 4, 4, 32000, 1),
    ("MSD0.convs2", 64, 256, 41, 4, 16, 8000, 1),
    ("MSD0.convs3", 256, 1024, 41, 4, 64, 2000, 1),
    ("MSD0.convs4", 1024, 1024, 41, 4, 256, 500, 1),
    ("MSD1.convs1", 16, 64, 41, 4, 4, 16001, 1),
    ("MPD2.convs1", 32, 128, 5, 3, 4, 5334, 2),
    ("MPD2.convs2", 128, 512, 5, 3, 16, 1778, 2),
    ("MPD2.convs3", 512, 1024, 5, 3, 64, 593, 2),
    ("MPD2.convs4", 1024, 1024, 5, 1, 64, 198, 2),
    ("MPD11.convs1", 32, 128, 5, 3, 4, 970, 11),
    ("MPD11.convs2", 128, 512, 5, 3, 16, 324, 11),
    ("MPD11.convs3", 512, 1024, 5, 3, 64, 108, 11),
    ("MPD2.conv_post", 1024, 1, 3, 1, 1, 198, 2),
    ("MPD11.conv_post", 1024, 1, 3, 1, 1, 36, 11),
    ("MSD0.conv_post", 1024, 1, 3, 1, 1, 125, 1),
]
tot = [0.0, 0.0, 0.0]
for name, Cin, Cout, K, S, G, Lin, P in LAYERS:
    pad = K // 2
    x = torch.randn(B, Cin, Lin, P, device=dev)
    w = torch.randn(Cout, Cin // G, K, device=dev) * 0.05
    b = torch.zeros(Cout, device=dev)
    gw = torch.ones(Cout, 1, 1, device=dev)
    _, imf, imd = ops.mt_weight_norm_fwd([gw], [w], [(K, S, pad, G)], P)          # staged weight images, as in the step
    act = ops.ACT_NONE if Cout == 1 else ops.ACT_LRELU          # conv_post has no activation
    y = ops.conv1d_fwd(x, w, b, G, S, pad, act=act, wimg=imf[0])
    dy = torch.randn_like(y)
    dw, db = torch.zeros_like(w), torch.zeros(Cout, device=dev)
    t_f = gtime(lambda: ops.conv1d_fwd(x, w, b, G, S, pad, act=act, wimg=imf[0]))
    t_d = gtime(lambda: ops.conv1d_dgrad(dy, w, x.shape, G, S, pad, gextra=x, xact=x, act=ops.ACT_LRELU, wimg=imd[0]))
    t_w = gtime(lambda: ops.conv1d_wgrad(x, dy, w.shape, G, S, pad, dw=dw, db=db))
    bx, by = x.numel() * 4, y.numel() * 4
    tot[0] += t_f; tot[1] += t_d; tot[2] += t_w
    print(f"{name:19s} B={B:2d} x {bx/1e6:6.1f} MB y {by/1e6:6.1f} MB: fwd {t_f:6.1f} us ({(bx+by)/t_f/1e3:5.0f} GB/s)  "
          f"dgrad {t_d:6.1f} us ({(3*bx+by)/t_d/1e3:5.0f} GB/s)  wgrad {t_w:6.1f} us ({(bx+by)/t_w/1e3:5.0f} GB/s)", flush=True)
print(f"sum: fwd {tot[0]:.1f} us  dgrad {tot[1]:.1f} us  wgrad {tot[2]:.1f} us")
