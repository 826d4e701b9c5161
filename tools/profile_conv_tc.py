"""ncu target: the tcgen05 grouped convolutions on the heaviest MSD / MPD layer shapes of the D step (batch 2B = 16).
    ncu --set full --import-source on -k regex:conv_tc_kernel -s 8 -c 4 python tools/profile_conv_tc.py
(two passes per layer: the first warms the instruction cache and the launch-configuration caches)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops
dev = torch.device("cuda:0")
# (B, Cin, Cout, K, S, G, Lin, P): bench.py's roofline instance (MSD convs.1 at B = 8), the same layer at the D step's 2B = 16
# and an MPD layer (forward on tcgen05, data gradient on mma.sync)
LAYERS = ((8, 16, 64, 41, 4, 4, 32000, 1), (16, 16, 64, 41, 4, 4, 32000, 1), (16, 512, 1024, 5, 3, 64, 593, 2))
state = []
for (B, Cin, Cout, K, S, G, Lin, P) in LAYERS:
    pad = K // 2
    x = torch.randn(B, Cin, Lin, P, device=dev)
    w = torch.randn(Cout, Cin // G, K, device=dev) * 0.05
    b = torch.zeros(Cout, device=dev)
    gw = torch.ones(Cout, 1, 1, device=dev)
    _, imf, imd = ops.mt_weight_norm_fwd([gw], [w], [(K, S, pad, G)], P)
    y = ops.conv1d_fwd(x, w, b, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0])
    state.append((x, w, b, imf, imd, torch.randn_like(y), torch.randn_like(x), (G, S, pad)))
for rep in range(2):          # first pass: warm-up (instruction cache, launch-configuration caches); second: profiled
    for (x, w, b, imf, imd, dy, ge, (G, S, pad)) in state:
        y = ops.conv1d_fwd(x, w, b, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0])
        dx = ops.conv1d_dgrad(dy, w, x.shape, G, S, pad, gextra=ge, xact=x, act=ops.ACT_LRELU, wimg=imd[0])
        dw, db = ops.conv1d_wgrad(x, dy, w.shape, G, S, pad)
torch.cuda.synchronize()
print("done")
