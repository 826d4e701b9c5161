"""ncu target: the tcgen05 grouped convolutions on the heaviest MSD / MPD layer shapes of the D step (batch 2B = 16).
    ncu --set full --import-source on -k regex:conv_tc_kernel -s 8 -c 4 python tools/profile_conv_tc.py
(two passes per layer: the first warms the instruction cache and the launch-configuration caches)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops
dev = torch.device("cuda:0")
B = 16
LAYERS = ((16, 64, 41, 4, 4, 32000, 1), (512, 1024, 5, 3, 64, 593, 2))
state = []
for (Cin, Cout, K, S, G, Lin, P) in LAYERS:
    pad = K // 2
    x = torch.randn(B, Cin, Lin, P, device=dev)
    w = torch.randn(Cout, Cin // G, K, device=dev) * 0.05
    b = torch.zeros(Cout, device=dev)
    gw = torch.ones(Cout, 1, 1, device=dev)
    _, imf, imd = ops.mt_weight_norm_fwd([gw], [w], [(K, S, pad, G)], P)
    y = ops.conv1d_fwd(x, w, b, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0])
    state.append((x, w, b, imf, imd, torch.randn_like(y), (G, S, pad)))
for rep in range(2):          # launches 0-3: warm-up pass; 4-7 (+4 of the set-up above = skip 8): profiled
    for (x, w, b, imf, imd, dy, (G, S, pad)) in state:
        y = ops.conv1d_fwd(x, w, b, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0])
        dx = ops.conv1d_dgrad(dy, w, x.shape, G, S, pad, gextra=x, xact=x, act=ops.ACT_LRELU, wimg=imd[0])
torch.cuda.synchronize()
print("done")
