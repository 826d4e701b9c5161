"""Micro-benchmark of lct_gemm (SIMT fp32 vs TF32 mma) on the generator's shapes (M = 34056 rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops, _lib
dev = torch.device("cuda:0")
M = 34056
def timeit(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
x64 = torch.randn(M, 64, device=dev); x128 = torch.randn(M, 128, device=dev); x192 = torch.randn(M, 192, device=dev)
x384 = torch.randn(M, 384, device=dev)
cases = {
 "NT qkv  N192 K64": lambda: ops.gemm(x64, torch.randn(192, 64, device=dev), x192, M, 192, 64, lda=64, ldb=64, ldc=192),
 "NT out  N64 K64 ": lambda: ops.gemm(x64, w6464, y64, M, 64, 64, lda=64, ldb=64, ldc=64),
 "NT lin  N64 K128": lambda: ops.gemm(x128, w64128, y64, M, 64, 128, lda=128, ldb=128, ldc=64),
 "NT gi   N48 K16 b8": lambda: ops.gemm(x64, wih, x384, M, 48, 16, lda=64, ldb=16, ldc=384, nbatch=8, a_div=2, sA=16, sB=768, sC=48),
 "NN dsn  N64 K192": lambda: ops.gemm(x192, w19264, y64, M, 64, 192, lda=192, ldb=64, ldc=64, tb=True),
 "NN dxn  N16 K96 b4": lambda: ops.gemm(x384, wih, y64, M, 16, 96, lda=384, ldb=16, ldc=64, tb=True, nbatch=4, sA=96, sB=1536, sC=16),
 "TN dwi  M192 N64 ks67": lambda: ops.gemm(x192, x64, dw19264, 192, 64, M, lda=192, ldb=64, ldc=64, ta=True, tb=True, ksplit=67),
 "TN dlin M64 N128 ks67": lambda: ops.gemm(x64, x128, dw64128, 64, 128, M, lda=64, ldb=128, ldc=128, ta=True, tb=True, ksplit=67),
 "TN dwih M48 N16 b8 ks67": lambda: ops.gemm(x384, x64, dwih, 48, 16, M, lda=384, ldb=64, ldc=16, ta=True, tb=True, ksplit=67, nbatch=8, a_div=1, b_div=2, sA=48, sB=16, sC=768),
}
w6464 = torch.randn(64, 64, device=dev); w64128 = torch.randn(64, 128, device=dev); w19264 = torch.randn(192, 64, device=dev)
wih = torch.randn(8, 48, 16, device=dev); y64 = torch.empty(M, 64, device=dev)
dw19264 = torch.zeros(192, 64, device=dev); dw64128 = torch.zeros(64, 128, device=dev); dwih = torch.zeros(8, 48, 16, device=dev)
for name, fn in cases.items():
    _lib.call_ret("lct_set_tensor_core_gemm", 0); a = timeit(fn)
    _lib.call_ret("lct_set_tensor_core_gemm", 1); b = timeit(fn)
    print(f"{name:26s} simt {a:7.1f} us   tf32 {b:7.1f} us", flush=True)
