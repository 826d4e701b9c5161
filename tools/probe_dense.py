"""Bring-up probe for the tcgen05 dense kernels: runs each bisection stage in its own process."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from lctgan import ops, _lib
stage = int(sys.argv[1]); K = int(sys.argv[2])
dev = torch.device("cuda:0")
B, L, Ci, Co = 2, 63, 256, 128
x = torch.randn(B, Ci, L, 1, device=dev); w = torch.randn(Co, Ci, K, device=dev) * 0.03
dy = torch.randn(B, Co, L, 1, device=dev)
dyq = ops.stage_ncl_bf16(dy, L + K - 1, 0); xq = ops.stage_ncl_bf16(x, L + K - 1, K // 2, copies=K)
torch.cuda.synchronize(); print("staging ok", flush=True)
_lib.call_ret("lct_dense_debug", stage)
dw = ops.dense_wgrad(dyq, xq, Co, Ci, K, w.shape); torch.cuda.synchronize(); print("wgrad stage", stage, "K", K, "ok", flush=True)
if stage == 0:
    import torch.nn.functional as F
    xr = x.squeeze(-1).bfloat16().float(); wr = w.clone().requires_grad_(True)
    (F.conv1d(xr, wr, padding=K // 2) * dy.squeeze(-1).bfloat16().float()).sum().backward()
    print("wgrad rel err", ((dw - wr.grad).abs().max() / wr.grad.abs().max()).item(), flush=True)
''' % (os.path.join(ROOT, "lct-gan_b200"), ROOT)
for st, K in ((2, 5), (0, 5)):
    r = subprocess.run([sys.executable, "-c", CHILD, str(st), str(K)], capture_output=True, text=True, timeout=120)
    print(f"--- stage {st} K {K}: rc={r.returncode}\n{r.stdout[-300:]}\n{r.stderr[-300:]}", flush=True)
