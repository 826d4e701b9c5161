"""conv_post kernels (1024 -> 1 channel) timed from CUDA graphs of 20 launches on the step's shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops
dev = torch.device("cuda:0")
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
w = torch.randn(1, 1024, 3, device=dev) * 0.03
b = torch.zeros(1, device=dev)
for B, L, P in ((16, 198, 2), (16, 125, 1), (16, 36, 11), (16, 63, 1), (16, 32, 1), (8, 198, 2), (8, 32, 1)):
    x = torch.randn(B, 1024, L, P, device=dev)
    dy = torch.randn(B, 1, L, P, device=dev)
    t_f = gtime(lambda: ops.conv1d_fwd(x, w, b, 1, 1, 1))
    t_d = gtime(lambda: ops.conv1d_dgrad(dy, w, x.shape, 1, 1, 1, gextra=x, xact=x, act=ops.ACT_LRELU))
    t_w = gtime(lambda: ops.conv1d_wgrad(x, dy, w.shape, 1, 1, 1))
    mb = x.numel() * 4 / 1e6
    print(f"B={B:2d} L={L:3d} P={P:2d} ({mb:5.1f} MB): fwd {t_f:6.1f} us  dgrad {t_d:6.1f} us  wgrad {t_w:6.1f} us", flush=True)
