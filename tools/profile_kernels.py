"""Small driver for ncu: launches the kernels bench.py reports rooflines for (and the other heavy discriminator
kernels) on their real B = 8, 2 s shapes a few times.

    ncu --set full --clock-control none --import-source on -k regex:"dense_kernel|conv_mma" -s 14 -c 7 \
        -o gpurun_out/prof python tools/profile_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200"))
import torch  # noqa: E402

from lctgan import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B = 8
# MSD convs.5 (dense, tcgen05): x [8,1024,125]
x = torch.randn(B, 1024, 125, 1, generator=g).to(dev)
w = (torch.randn(1024, 1024, 5, generator=g) / 72).to(dev)
bias = torch.zeros(1024, device=dev)
wt, wd = ops.stage_dense_weights(w)
xp = ops.stage_nlc_bf16(x, 2)
dyq = ops.stage_ncl_bf16(x, 129, 0)
xq = ops.stage_ncl_bf16(x, 129, 2, copies=5)
# MSD convs.1 (grouped 16 -> 64, k 41, s 4, g 4): x [8,16,32000]
x1 = torch.randn(B, 16, 32000, 1, generator=g).to(dev)
w1 = (torch.randn(64, 4, 41, generator=g) / 13).to(dev)
b1 = torch.zeros(64, device=dev)
ge1 = torch.randn(B, 16, 32000, 1, generator=g).to(dev)
_, imf1, imd1 = ops.mt_weight_norm_fwd([torch.ones(64, 1, 1, device=dev)], [w1], [(41, 4, 20, 4)], 1)
# MPD period 2 convs.1 (grouped 32 -> 128, k 5, s 3, g 4): x [8,32,5334,2]
x2 = torch.randn(B, 32, 5334, 2, generator=g).to(dev)
w2 = (torch.randn(128, 8, 5, generator=g) / 6).to(dev)
b2 = torch.zeros(128, device=dev)
_, imf2, imd2 = ops.mt_weight_norm_fwd([torch.ones(128, 1, 1, device=dev)], [w2], [(5, 3, 2, 4)], 2)
for _ in range(3):
    y = ops.dense_conv(xp, wt, B, 125, 1024, 1024, 5, bias=bias, act=ops.ACT_LRELU)                       # 1 dense fwd
    ops.dense_wgrad(dyq, xq, 1024, 1024, 5, w.shape)                                                      # 2 dense wgrad
    y1 = ops.conv1d_fwd(x1, w1, b1, 4, 4, 20, act=ops.ACT_LRELU, wimg=imf1[0])                            # 3 fwd MSD
    ops.conv1d_wgrad(x1, y1, w1.shape, 4, 4, 20)                                                          # 4 wgrad MSD
    ops.conv1d_dgrad(y1, w1, x1.shape, 4, 4, 20, gextra=ge1, xact=x1, act=ops.ACT_LRELU, wimg=imd1[0])    # 5 dgrad MSD
    y2 = ops.conv1d_fwd(x2, w2, b2, 4, 3, 2, act=ops.ACT_LRELU, wimg=imf2[0])                             # 6 fwd MPD
    ops.conv1d_dgrad(y2, w2, x2.shape, 4, 3, 2, xact=x2, act=ops.ACT_LRELU, wimg=imd2[0])                 # 7 dgrad MPD
torch.cuda.synchronize()
print("ok")
