"""tcgen05 dense conv (MSD convs.5 forward) timed from a CUDA graph of 20 launches, per tile width and batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops, _lib
dev = torch.device("cuda:0")
C, K = 1024, 5
w = torch.randn(C, C, K, device=dev) * 0.02
bias = torch.randn(C, device=dev)
wt, _ = ops.stage_dense_weights(w, want_wt=True, want_wd=False)
for B, L in ((8, 125), (16, 125), (8, 63), (16, 32)):
    x = torch.randn(B, C, L, 1, device=dev)
    xs = ops.stage_nlc_bf16(x, K // 2)
    ref = None
    for bn in (0,):       # tile width is chosen by the library (128 x 128 when that still gives >= 100 tiles)
        run = lambda: ops.dense_conv(xs, wt, B, L, C, C, K, bias=bias, act=ops.ACT_LRELU, slope=0.2)
        y = run(); torch.cuda.synchronize()
        if ref is None: ref = y
        err = ((y - ref).abs().max() / ref.abs().max()).item()
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): run()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): run()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        fl = 2.0 * B * L * C * C * K
        print(f"B={B:2d} L={L:3d} BN={bn:3d}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s   max diff vs BN=64 {err:.2e}", flush=True)

# weight gradient (all taps per CTA)
for B, L in ((8, 125), (16, 125), (16, 63), (16, 32)):
    x = torch.randn(B, C, L, 1, device=dev)
    dy = torch.randn(B, C, L, 1, device=dev)
    Lp = L + K - 1
    dyq = ops.stage_ncl_bf16(dy, Lp, 0)
    xq = ops.stage_ncl_bf16(x, Lp, K // 2, copies=K)
    out = torch.zeros(C, C, K, device=dev)
    run = lambda: ops.dense_wgrad(dyq, xq, C, C, K, (C, C, K), out=out)
    for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * B * L * C * C * K
    print(f"wgrad B={B:2d} L={L:3d}: {us:6.1f} us  {fl / us / 1e6:6.1f} TFLOP/s", flush=True)
