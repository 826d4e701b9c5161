"""HBM-roofline check of the STFT / iSTFT / loss kernels at the top of the BASELINE sweep (B = 128, 4 s) and at B = 8, 2 s.
Algorithmic bytes per SURVEY.md section 8d; CUDA events, L2 flushed between launches."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch
from lctgan import ops
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def t(run, iters=10):
    for _ in range(3): run()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)
for B, T in ((8, 32000), (128, 64000)):
    x = torch.randn(B, T, device=dev) * 0.1; y = torch.randn(B, T, device=dev) * 0.1
    w = torch.hann_window(512, device=dev); n_fft, hop = 512, 256
    Tf, F = 1 + T // hop, 257
    spec, mag = ops.stft_fwd(x, w, n_fft, hop, want_mag=True)
    mask = torch.rand(B, Tf, F, device=dev) * 0.5 + 0.5
    acc = torch.zeros(2, 64, device=dev)
    gy = torch.randn(B, T, device=dev)
    cases = {
        "tf_features (2 STFT + mag + IRM^c + mag^c)": (lambda: ops.tf_features_fwd(x, y, w, n_fft, hop), 4.0 * (2 * B * T + 3 * B * Tf * F)),
        "stft + magnitude (enhancer front)": (lambda: ops.stft_fwd(x, w, n_fft, hop, want_mag=True), 4.0 * (B * T + 3 * B * Tf * F)),
        "mask-apply + istft (enhancer tail)": (lambda: ops.istft_fwd(spec, w, n_fft, hop, T, mask_c=mask), 4.0 * (3 * B * Tf * F + B * T)),
        "istft backward + mask gradient": (lambda: ops.istft_bwd(gy, w, n_fft, hop, Tf, xspec=spec, mask_c=mask, want_gspec=False), 4.0 * (B * T + 4 * B * Tf * F)),
        "mrstft sums 512 (loss, no spectrogram)": (lambda: ops.mrstft_sums(x, y, w, n_fft, hop, acc), 4.0 * 2 * B * T),
    }
    fm_a = [torch.randn(B, 128, 1778, 2, device=dev), torch.randn(B, 512, 593, 2, device=dev)]
    fm_b = [torch.randn_like(a) for a in fm_a]
    cases["feature-matching L1 (2 maps)"] = (lambda: ops.mt_reduce(fm_a, fm_b, [1.0, 1.0], ops.OP_ABS_DIFF), 4.0 * 2 * sum(a.numel() for a in fm_a))
    for name, (run, nbytes) in cases.items():
        ms = t(run)
        gbs = nbytes / (ms * 1e-3) / 1e9
        print(f"B={B:3d} T={T:5d}  {name:44s} {ms*1e3:8.1f} us  {nbytes/1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs/peak:5.2f} of HBM peak", flush=True)
