"""Bring-up / regression check of the tcgen05 grouped convolutions (csrc/conv_tc.cu) against torch on the same GPU
(fp64 on tf32-rounded operands = the arithmetic contract), for every grouped / first-layer shape of the MPD and MSD
stacks.  Prints one line per (case, pass); with --probe also decodes which input element each output row multiplied
(delta weights, ramp inputs), which is what one needs when a descriptor field is misread."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from lctgan import ops

dev = torch.device("cuda:0")


def tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


def rel(a, b):
    den = b.abs().max().item() or 1.0
    return (a.double() - b.double()).abs().max().item() / den


CASES = [
    # (Cin, Cout, K, S, G, L, P)
    (16, 64, 41, 4, 4, 2000, 1), (64, 256, 41, 4, 16, 500, 1), (256, 1024, 41, 4, 64, 125, 1),
    (1024, 1024, 41, 4, 256, 32, 1), (1, 16, 15, 1, 1, 2000, 1), (16, 64, 41, 4, 4, 8001, 1), (16, 64, 41, 4, 4, 1037, 1),
    (1, 32, 5, 3, 1, 600, 2), (32, 128, 5, 3, 4, 200, 3), (128, 512, 5, 3, 16, 67, 5), (512, 1024, 5, 3, 64, 23, 7),
    (1024, 1024, 5, 1, 64, 8, 11), (32, 128, 5, 3, 4, 5334, 2), (128, 512, 5, 3, 16, 131, 7), (1024, 1024, 5, 1, 64, 3, 11),
    (1, 32, 5, 3, 1, 2910, 11),
]


def run_case(Cin, Cout, K, S, G, L, P, B=3, verbose=True):
    gen = torch.Generator().manual_seed(Cin * 7 + Cout + K + L)
    pad = K // 2
    x = torch.randn(B, Cin, L, P, generator=gen).to(dev)
    w = (torch.randn(Cout, Cin // G, K, generator=gen) / (Cin // G * K) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=gen).to(dev)
    conv = lambda xx, ww, bb: F.conv2d(xx, ww.unsqueeze(-1), bb, stride=(S, 1), padding=(pad, 0), groups=G)
    img_f, img_d = ops.conv_tc_images([w], [(K, S, pad, G)], P)
    y = ops.conv_tc_fwd(x, img_f[0], b, Cout, G, K, S, pad, act=ops.ACT_LRELU, slope=0.2)
    ref = F.leaky_relu(conv(tf32(x), tf32(w), b.double()), 0.2)
    e_f = rel(y, ref)
    gy = torch.randn(ref.shape, generator=gen).to(dev)
    xact = torch.randn(x.shape, generator=gen).to(dev)
    gextra = (torch.randn(x.shape, generator=gen) * 0.1).to(dev)
    x64 = tf32(x).requires_grad_(True)
    w64 = tf32(w).requires_grad_(True)
    (conv(x64, w64, None) * tf32(gy)).sum().backward()
    ref_dx = (x64.grad + gextra.double()) * torch.where(xact > 0, 1.0, 0.2).double()
    dx = ops.conv_tc_dgrad(gy, img_d[0], (B, Cin, L, P), Cout, G, K, S, pad, gextra=gextra, xact=xact, act=ops.ACT_LRELU,
                           slope=0.2)
    e_d = rel(dx, ref_dx)
    dw, db = ops.conv1d_wgrad(x, gy, w.shape, G, S, pad)          # (TF32 mma.sync kernel: see conv_tc.cu header)
    e_w = rel(dw, w64.grad)
    e_b = rel(db, gy.double().sum(dim=(0, 2, 3)))
    ok = max(e_f, e_d, e_w) < 2e-4 and e_b < 1e-4
    if verbose:
        print(f"{'OK ' if ok else 'BAD'} Cin={Cin:4d} Cout={Cout:4d} K={K:2d} S={S} G={G:3d} L={L:5d} P={P:2d}: "
              f"fwd {e_f:.2e}  dgrad {e_d:.2e}  wgrad {e_w:.2e}  db {e_b:.2e}", flush=True)
    return ok, (e_f, e_d, e_w, e_b)


def probe(Cin, Cout, K, S, G, L, P):
    """delta weights x ramp input: which input position does output row l read for tap k0 / which column gets channel n"""
    pad = K // 2
    cig = Cin // G
    x = torch.zeros(1, Cin, L, P, device=dev)
    x[0, 0, :, 0] = torch.arange(L, device=dev).float() % 1000 + 1        # exact in tf32
    for (n0, c0, k0) in ((0, 0, 0), (0, 0, 1), (0, 0, min(4, K - 1)), (min(5, Cout // G - 1), 0, pad), (0, 0, K - 1)):
        w = torch.zeros(Cout, cig, K, device=dev)
        w[n0, c0, k0] = 1.0
        img_f, img_d = ops.conv_tc_images([w], [(K, S, pad, G)], P)
        y = ops.conv_tc_fwd(x, img_f[0], None, Cout, G, K, S, pad)
        ref = F.conv2d(x, w.unsqueeze(-1), None, stride=(S, 1), padding=(pad, 0), groups=G)
        got = y[0, n0, :10, 0].tolist()
        want = ref[0, n0, :10, 0].tolist()
        nz = (y[0, :, :, 0].abs().sum(dim=1) > 0).nonzero().flatten().tolist()
        print(f"  probe fwd n0={n0} k0={k0}: got {[int(v) for v in got]} want {[int(v) for v in want]} nonzero out-channels {nz[:8]}",
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    bad = 0
    for i, c in enumerate(CASES):
        if args.only >= 0 and i != args.only:
            continue
        errs = (1.0, 1.0, 1.0, 1.0)
        try:
            ok, errs = run_case(*c)
        except Exception as e:      # keep going: one line per case
            print(f"ERR {c}: {e!r}", flush=True)
            ok = False
        if not ok:
            bad += 1
            if args.probe:
                try:
                    if max(errs[0], errs[1]) >= 2e-4:
                        probe(*c)
                except Exception as e:
                    print(f"  probe failed: {e!r}", flush=True)
    print(f"{len(CASES) - bad} / {len(CASES)} cases within 2e-4", flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
