#!/usr/bin/env python
"""bench.py -- LCT-GAN adversarial training step on B200 (metric of BASELINE.json: train samples/s,
2 s @ 16 kHz segments).

    python bench.py --gpus 1 --steps 20 --warmup 5            # this framework (lctgan sm_100a kernels)
    python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU path (oracle port), host cores
    python bench.py --impl reference-cuda --steps 10          # the same port on stock torch CUDA ops (cuDNN/cuFFT/cuBLAS)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" is one full D+G iteration of the reference's train_one_epoch loop body (train.py:165-249)
on one synthetic batch of 8 segments per GPU (configs[2] of BASELINE.json; weak scaling).  Rank 0
prints ONE JSON line.  See DESIGN.md section "Measurement" for the definition of every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "lct-gan_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "train samples/sec (2 s @16 kHz)"
UNIT = "samples/s"
BATCH = 8
SEGMENT = 32000
WORKLOAD = "LCT-GAN D+G training step, batch 8 per GPU, 2.0 s @ 16 kHz segments, LS loss, reference hyperparameters (BASELINE configs[2])"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tensor": float(p["bf16_tflops"]), "tensor_sustained":
                float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the reference arms: the oracle port of the reference's step on stock torch operators.  Nothing here imports the
# product package (lctgan): initial weights come from oracle/ref_init.py (stock torch containers, pinned against
# the reference's own parameter checksums by tests/test_oracle_golden.py).
# --------------------------------------------------------------------------------------------------
def _config(gan_loss, world):
    """The workload description shared verbatim by every arm (ours, reference, reference-cuda)."""
    return {"workload": WORKLOAD if gan_loss == "ls" else WORKLOAD.replace("LS loss", "hinge loss").replace(
        "configs[2]", "configs[3]"), "gan_loss": gan_loss, "batch_per_gpu": BATCH, "segment_samples": SEGMENT,
        "parallelism": f"dp{world}", "l2": "no explicit flush: one step streams > 2 GB of activations (>> 126 MB L2)"}


def _oracle_state(device):
    import torch
    from oracle import lct_oracle as O
    from oracle import ref_init
    Pe, Pp, Ps = ref_init.init_state_dicts(42)
    mv = lambda d: {k: v.to(device) for k, v in d.items()}
    st = O.StepState(mv(Pe), mv(Pp), mv(Ps), order_g=ref_init.param_order(Pe),
                     order_d=(ref_init.param_order(Pp), ref_init.param_order(Ps)))
    wins = [O.hann_window(n).to(device) for n in O.MR_FFT_SIZES]
    return O, st, wins


def reference_steps(steps, warmup, gan_loss="ls", device="cpu", batch=BATCH):
    """Time `steps` D+G steps of the oracle port (reference algorithm on stock torch operators) after `warmup` untimed
    ones, batch 8 x 2 s exactly like the product arm.  Returns (samples_per_s, ms_per_step, losses of the first step)."""
    import torch
    O, st, wins = _oracle_state(device)
    noisy, clean = O.synthetic_batch(batch, SEGMENT, seed=1234)
    noisy, clean = noisy.to(device), clean.to(device)
    sync = torch.cuda.synchronize if str(device).startswith("cuda") else (lambda: None)
    first = None
    for _ in range(warmup):
        out = O.train_step(st, noisy, clean, wins, gan_loss=gan_loss, aten_gru=True)
        first = first or out
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = O.train_step(st, noisy, clean, wins, gan_loss=gan_loss, aten_gru=True)
        first = first or out
    sync()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, first


def reference_rtf(device="cpu"):
    """BASELINE.md section 3 inference baseline: the reference enhancer (oracle port, eval / no_grad) on the utterances
    of measure_enhance_rtf (1-10 s, zero-padded to the batch maximum like datasets.py:210-221), batch 1 and 16."""
    import torch
    O, st, _ = _oracle_state(device)
    out = {}
    gl = torch.Generator().manual_seed(7)
    for bs in (1, 16):
        x, lens = _rtf_batch(bs, gl)
        x = x.to(device)
        with torch.no_grad():
            O.enhancer_forward(st.enh, x[:, :16000], aten_gru=True)
            t0 = time.perf_counter()
            y, _ = O.enhancer_forward(st.enh, x, aten_gru=True)
            y = y.cpu()
            dt = time.perf_counter() - t0
        out[f"batch{bs}"] = {"rtf": dt / (sum(lens) / 16000.0), "ms": dt * 1e3, "audio_s": sum(lens) / 16000.0}
    return out


def _rtf_batch(bs, gl):
    import torch
    lens = torch.randint(16000, 160001, (bs,), generator=gl).tolist()
    x = torch.zeros(bs, max(lens))
    for i, n in enumerate(lens):
        x[i, :n] = torch.randn(n, generator=gl) * 0.1
    return x, lens


def run_reference(args):
    """`--impl reference` (host CPU, all threads) and `--impl reference-cuda` (stock torch CUDA operators: cuDNN / cuFFT /
    cuBLAS on the same B200; fp32, or TF32-allowed with --allow-tf32).  Rank 0 alone works; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cuda = args.impl == "reference-cuda"
    if cuda:
        if not torch.cuda.is_available():
            print(json.dumps({"impl": "reference-cuda", "unavailable": "no CUDA device"}), flush=True)
            return
        torch.backends.cuda.matmul.allow_tf32 = bool(args.allow_tf32)
        torch.backends.cudnn.allow_tf32 = bool(args.allow_tf32)
        device, cores = "cuda:0", 0
    else:
        torch.set_num_threads(os.cpu_count() or 1)        # torchrun exports OMP_NUM_THREADS=1: use every host core
        device, cores = "cpu", torch.get_num_threads()
    if args.what == "rtf":
        print(json.dumps({"impl": args.impl, "what": "enhance_rtf", "cores": cores, "rtf": reference_rtf(device)}),
              flush=True)
        return
    sps, ms, first = reference_steps(args.steps, args.warmup, args.gan_loss, device)
    where = "host CPU" if not cuda else ("stock torch CUDA ops (cuDNN/cuFFT/cuBLAS), " +
                                         ("TF32 allowed" if args.allow_tf32 else "fp32"))
    sample = (f"{args.steps} timed D+G steps of {BATCH} x 2 s segments after {args.warmup} warm-up steps (oracle port of "
              f"train.py:165-249, {where}" + (f", {cores} threads)" if not cuda else ")"))
    line = {"impl": args.impl, "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if not (cuda and args.allow_tf32) else "tf32", "data": "synthetic",
            "config": _config(args.gan_loss, world), "device": where,
            "cpu_baseline": {"value": sps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "losses_first_step": {k: first[k] for k in ("d_loss", "g_loss", "mr", "mask", "adv", "fm")}}
    print(json.dumps(line), flush=True)


def _sub_bench(extra, timeout=600):
    """Run another arm of this script in a subprocess (so that the product process never imports the oracle) and return
    its JSON line."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS")}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + extra, capture_output=True, text=True,
                           timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as e:      # a missing baseline must not lose the product's measurement
        return {"error": repr(e)}


# --------------------------------------------------------------------------------------------------
# roofline of the dominant kernel, timed live
# --------------------------------------------------------------------------------------------------
def _time_kernel(run, flush, iters=10):
    """Average CUDA-event duration (ms) of one launch, L2 flushed (256 MB write) before every timed launch.

    The GPU is parked on a ~200 us spin kernel while the host enqueues [event, launch, event]: without it the event pair
    also spans the host-side cost of the launch (torch.empty + ctypes, 30-60 us - as long as the kernels themselves) on
    a box whose CPU is slower than the 256 MB flush, which made the reported fractions jump between runs."""
    import torch
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(400_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return sum(times) / len(times)


def _time_kernel_stream(run, nsets, iters=5):
    """Average CUDA-event duration (ms) of one launch inside a stream of launches: `run(i)` works on the i-th of `nsets`
    DISTINCT operand sets whose total size exceeds the 126 MB L2 several times, so every launch finds its inputs in HBM
    without a flush between launches; one event pair brackets the nsets launches (GPU parked on a spin kernel while the
    host enqueues them).  This is how the kernel runs inside the step: launch latency and the ramp-down of the previous
    grid are hidden by the next launch, which the launch-alone number of _time_kernel pays in full."""
    import torch
    for i in range(nsets):
        run(i)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nsets):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / nsets)
    return sum(times) / len(times)


#: dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel (ncu --set full of the B = 8 instance
#: with distinct FM-gradient / activation buffers: profiles/ncu_full_r2_conv_tc_v2_raw.csv, 49.24 MB read + 0.60 MB
#: written back before the kernel ends)
ROOFLINE_TRAFFIC_BYTES = 49.84e6


def measure_roofline(dev, peaks):
    """Rooflines of the kernels that matter, each timed alone with CUDA events on the launching stream:

    roofline        - the kernel family that dominates the step's GPU time (profiles/): the grouped discriminator
                      convolutions; as in round 1 the instance is the data gradient of MSD convs.1 (16 -> 64 channels,
                      k 41, stride 4, 4 groups, B = 8, L = 32000), now on tcgen05 (conv_tc.cu).  HBM bound (AI ~ 40
                      flop/B): algorithmic bytes = dY + FM gradient + saved activation read, dX written = 4 x 16.4 MB.
    roofline_family - forward / data gradient / weight gradient of the same layer at B = 8 and at the D step's 2B = 16,
                      and of an MPD layer, with the kernel that runs each (tcgen05 or mma.sync, lctgan/ops.py).
    roofline_tensor - the one contraction above the ridge, MSD convs.5 (1024 -> 1024, k 5) on tcgen05:
                      algorithmic FLOPs = 2 B L Cin Cout K = 10.49 GFLOP.
    """
    import torch
    from lctgan import ops
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    family = {}
    roof = None
    for tag, (Bq, Cin, Cout, K, S, G, L, P) in (("msd_convs1_B8", (BATCH, 16, 64, 41, 4, 4, 32000, 1)),
                                               ("msd_convs1_2B16", (2 * BATCH, 16, 64, 41, 4, 4, 32000, 1)),
                                               ("mpd2_convs2_2B16", (2 * BATCH, 128, 512, 5, 3, 16, 1778, 2))):
        pad = K // 2
        Lout = (L + 2 * pad - K) // S + 1
        x = torch.randn(Bq, Cin, L, P, generator=g).to(dev)
        dy = torch.randn(Bq, Cout, Lout, P, generator=g).to(dev)
        w = (torch.randn(Cout, Cin // G, K, generator=g) / (Cin // G * K) ** 0.5).to(dev)
        bias = torch.zeros(Cout, device=dev)
        gextra = torch.randn(Bq, Cin, L, P, generator=g).to(dev)
        gw = torch.ones(Cout, 1, 1, device=dev)
        _, imf, imd = ops.mt_weight_norm_fwd([gw], [w], [(K, S, pad, G)], P)      # staged weight images, as in the step
        dw, db = torch.zeros_like(w), torch.zeros(Cout, device=dev)
        runs = {
            "fwd": (lambda: ops.conv1d_fwd(x, w, bias, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0]),
                    4.0 * (x.numel() + dy.numel()), "tcgen05" if ops._use_tc(Cin, Cout, K, G, S, pad, P) else "mma.sync"),
            "dgrad": (lambda: ops.conv1d_dgrad(dy, w, (Bq, Cin, L, P), G, S, pad, gextra=gextra, xact=x,
                                               act=ops.ACT_LRELU, wimg=imd[0]),
                      4.0 * (dy.numel() + 3 * x.numel()),
                      "tcgen05" if ops._use_tc_dgrad(Cin, Cout, K, G, S, pad, P) else "mma.sync"),
            "wgrad": (lambda: ops.conv1d_wgrad(x, dy, w.shape, G, S, pad, dw=dw, db=db), 4.0 * (x.numel() + dy.numel()),
                      "mma.sync"),
        }
        family[tag] = {}
        # the same three passes over NS distinct operand sets (NS x the layer's bytes >> L2), launched back to back
        NS = 8 if Bq == BATCH else 4
        xs = [x] + [torch.randn_like(x) for _ in range(NS - 1)]
        dys = [dy] + [torch.randn_like(dy) for _ in range(NS - 1)]
        ges = [gextra] + [torch.randn_like(gextra) for _ in range(NS - 1)]
        streams = {
            "fwd": lambda i: ops.conv1d_fwd(xs[i], w, bias, G, S, pad, act=ops.ACT_LRELU, wimg=imf[0]),
            "dgrad": lambda i: ops.conv1d_dgrad(dys[i], w, (Bq, Cin, L, P), G, S, pad, gextra=ges[i], xact=xs[i],
                                                act=ops.ACT_LRELU, wimg=imd[0]),
            "wgrad": lambda i: ops.conv1d_wgrad(xs[i], dys[i], w.shape, G, S, pad, dw=dw, db=db),
        }
        for name, (run, nbytes, kern) in runs.items():
            ms_alone = _time_kernel(run, flush)
            ms = _time_kernel_stream(streams[name], NS)
            gbs = nbytes / (ms * 1e-3) / 1e9
            family[tag][name] = {"kernel": kern, "us_per_launch": ms * 1e3, "achieved": gbs, "frac": gbs / peaks["hbm"],
                                 "algorithmic_bytes": nbytes, "us_launched_alone": ms_alone * 1e3,
                                 "frac_launched_alone": nbytes / (ms_alone * 1e-3) / 1e9 / peaks["hbm"]}
            if tag == "msd_convs1_B8" and name == "dgrad":
                roof = {"kernel": "conv_tc_kernel<dgrad> (MSD convs.1 data gradient: 64->16 ch, k=41, s=4, groups=4, B=8, "
                                  "L=32000; tcgen05.mma kind::tf32 implicit GEMM over channel-quad planes in shared memory, "
                                  "accumulators in TMEM, fused FM-gradient / LeakyReLU' epilogue with one float4 per row and "
                                  "channel)",
                        "bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                        "traffic": ROOFLINE_TRAFFIC_BYTES, "traffic_note": "dram__bytes_read+write per launch, ncu --set full "
                        "(profiles/ncu_full_r2_conv_tc_v2_raw.csv); the 16.4 MB output is still in the 126 MB L2 when the "
                        "kernel ends",
                        "algorithmic_bytes": nbytes, "ms_per_launch": ms, "peak_source": peaks["src"] + " (STREAM-style copy)",
                        "timing": f"CUDA events around {NS} back-to-back launches over {NS} distinct operand sets "
                                  f"({NS * nbytes / 1e6:.0f} MB >> 126 MB L2, no flush needed), as the kernel runs inside the "
                                  "step; launched alone after a 256 MB L2 flush: ms_launched_alone",
                        "ms_launched_alone": ms_alone, "frac_launched_alone": nbytes / (ms_alone * 1e-3) / 1e9 / peaks["hbm"],
                        "bound_note": "22 K = 8 MMAs of N = 16 per 128-row tile: tools/umma_rate.cu measures 48.6 cycles per "
                                      "such MMA (32 of them the 4 KB A-operand fetch from shared memory), i.e. a tensor-pipe "
                                      "floor of 7.4 us for this launch against 10 us of HBM time, see DESIGN.md section 6"}
    # ---- dense conv forward on tcgen05 (MSD convs.5): the D step pushes clean + enhanced through as one batch of 2B
    C, K = 1024, 5
    w = (torch.randn(C, C, K, generator=g) / (C * K) ** 0.5).to(dev)
    bias = torch.zeros(C, device=dev)
    wt, _ = ops.stage_dense_weights(w, want_wd=False)
    per_shape = {}
    for Bd in (2 * BATCH, BATCH):
        L = 125
        x = torch.randn(Bd, C, L, 1, generator=g).to(dev)
        xp = ops.stage_nlc_bf16(x, K // 2)
        flops = 2.0 * Bd * L * C * C * K
        ms = _time_kernel(lambda: ops.dense_conv(xp, wt, Bd, L, C, C, K, bias=bias, act=ops.ACT_LRELU), flush)
        per_shape[Bd] = (flops / (ms * 1e-3) / 1e12, ms, flops)
    tf, ms, flops = per_shape[2 * BATCH]
    roof_t = {"kernel": "dense_kernel<128,5,conv> (MSD convs.5 forward: 1024->1024, k=5, the D step's batch of 2B=16, L=125, "
                        "tcgen05 bf16, 128x128 tiles)",
              "bound": "tensor", "achieved": tf, "peak": peaks["tensor"], "unit": "TFLOP/s",
              "frac": tf / peaks["tensor"], "traffic": None, "ms_per_launch": ms, "algorithmic_flops": flops,
              "g_step_shape_B8": {"achieved": per_shape[BATCH][0], "ms_per_launch": per_shape[BATCH][1],
                                  "frac": per_shape[BATCH][0] / peaks["tensor"], "tiles": "128x64 (144 CTAs)"},
              "note": "bound by the per-SM TMA / L2 feed (~37 B/clk/SM measured), not by the tensor pipe: see DESIGN.md section 6",
              "peak_source": peaks["src"] + " (burst: kernel timed alone)"}
    roof["family"] = family
    return roof, roof_t


def measure_frontend_rooflines(dev, peaks):
    """HBM rooflines of the STFT / iSTFT / spectral-loss kernels (BASELINE north_star: STFT front end and losses profiled
    separately) at the top of the BASELINE sweep (B = 128, 4 s; at B = 8 x 2 s they are launch-latency sized, 8-15 us).
    Algorithmic bytes per SURVEY.md section 8d; every launch timed alone with CUDA events, L2 flushed before it."""
    import torch
    from lctgan import ops
    B, T, n_fft, hop = 128, 64000, 512, 256
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    x = torch.randn(B, T, device=dev) * 0.1
    y = torch.randn(B, T, device=dev) * 0.1
    w = torch.hann_window(n_fft, device=dev)
    Tf, F = 1 + T // hop, n_fft // 2 + 1
    spec, _ = ops.stft_fwd(x, w, n_fft, hop, want_mag=True)
    mask = torch.rand(B, Tf, F, device=dev) * 0.5 + 0.5
    acc = torch.zeros(2, 64, device=dev)
    gy = torch.randn(B, T, device=dev)
    fm_a = [torch.randn(B, 128, 1778, 2, device=dev) for _ in range(2)]
    cases = [
        ("stft_magnitude", "wf::r2c_warp_kernel<8,STFT> (ComplexSTFT.forward + magnitude)",
         lambda: ops.stft_fwd(x, w, n_fft, hop, want_mag=True), 4.0 * (B * T + 3 * B * Tf * F)),
        ("tf_features", "wf::r2c_warp_kernel<8,TFF> (2 STFT + |.| + IRM^c + mag^c)",
         lambda: ops.tf_features_fwd(x, y, w, n_fft, hop), 4.0 * (2 * B * T + 3 * B * Tf * F)),
        ("mask_istft", "wf::c2r_warp_kernel<8,ISTFT> (mask decompression + iSTFT)",
         lambda: ops.istft_fwd(spec, w, n_fft, hop, T, mask_c=mask), 4.0 * (3 * B * Tf * F + B * T)),
        ("istft_adjoint_mask_grad", "wf::r2c_warp_kernel<8,ISTFT_BWD> (iSTFT backward + mask gradient)",
         lambda: ops.istft_bwd(gy, w, n_fft, hop, Tf, xspec=spec, mask_c=mask, want_gspec=False),
         4.0 * (B * T + 4 * B * Tf * F)),
        ("mrstft_sums_512", "wf::r2c_warp_kernel<8,MRLOSS> (spectral-loss sums, no spectrogram written)",
         lambda: ops.mrstft_sums(x, y, w, n_fft, hop, acc), 4.0 * 2 * B * T),
        ("feature_matching_l1", "mt_kernel<reduce> (L1 over two 58 M-element feature-map pairs)",
         lambda: ops.mt_reduce(fm_a[:1], fm_a[1:], [1.0], ops.OP_ABS_DIFF), 4.0 * 2 * fm_a[0].numel()),
    ]
    out = {"shape": "B=128 x 4 s @ 16 kHz, n_fft 512 / hop 256", "bound": "hbm", "peak": peaks["hbm"], "unit": "GB/s"}
    for key, kernel, run, nbytes in cases:
        ms = _time_kernel(run, flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[key] = {"kernel": kernel, "achieved": gbs, "frac": gbs / peaks["hbm"], "us_per_launch": ms * 1e3,
                    "algorithmic_bytes": nbytes}
    return out


def measure_enhance_rtf(dev):
    """BASELINE.json configs[1]: full-utterance enhancement, synthetic 16 kHz utterances of 1-10 s zero-padded to the
    batch maximum like datasets.py:210-221, batch 1 and 16, eval + no_grad; real-time factor = (enhancer wall time
    incl. the H2D copy of the batch and the D2H copy of the waveforms) / (seconds of audio in the batch)."""
    import torch
    from lctgan.training import build_models
    enh = build_models(dev, gan_seed=42)[0].eval()
    gl = torch.Generator().manual_seed(7)
    out = {}
    for bs in (1, 16):
        x, lens = _rtf_batch(bs, gl)
        xh = x.pin_memory()
        with torch.no_grad():
            for _ in range(2):
                enh(xh.to(dev, non_blocking=True))[0].cpu()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                y = enh(xh.to(dev, non_blocking=True))[0].cpu()
            dt = (time.perf_counter() - t0) / reps
        out[f"batch{bs}"] = {"rtf": dt / (sum(lens) / 16000.0), "ms": dt * 1e3, "audio_s": sum(lens) / 16000.0,
                             "padded_to_s": x.shape[1] / 16000.0}
        if bs == 16:
            # the same 16 utterances at their true lengths through the repo's inference helper: length buckets (no
            # utterance padded by more than 10 %), pinned staging and D2H overlapped with the enhancer
            # (lctgan.inference.enhance_utterances, the N3 form of infer.py:142-157); host tensors in, host tensors out
            from lctgan.inference import enhance_utterances
            waves = [x[i, :n].clone() for i, n in enumerate(lens)]
            with torch.no_grad():
                enhance_utterances(enh, waves, max_batch=16)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    ys = enhance_utterances(enh, waves, max_batch=16)
                dt = (time.perf_counter() - t0) / reps
            assert [int(y.shape[-1]) for y in ys] == list(lens)
            out["batch16_length_buckets"] = {"rtf": dt / (sum(lens) / 16000.0), "ms": dt * 1e3,
                                             "audio_s": sum(lens) / 16000.0}
    return out


GOLDEN_V3 = os.path.join(ROOT, "tests", "golden", "golden_v3.pt")
_LOG_KEYS = {"D_loss": "d_loss", "G_loss": "g_loss", "MR": "mr", "Mask": "mask", "Adv": "adv", "FM": "fm"}


def parity_check(losses, gan_loss, world, tol_rel):
    """Step-1 losses of the benchmarked configuration against the reference's own train_one_epoch log at this exact shape
    (tests/golden/golden_v3.pt, written by tests/golden/make_golden.py --logs-only from the unmodified reference).
    With N > 1 the discriminator update averages gradients over ranks that hold different data, so only the quantities
    computed before that update are comparable (d_loss, mr, mask)."""
    import torch
    if not os.path.exists(GOLDEN_V3):
        return {"checked": False, "why": "tests/golden/golden_v3.pt missing"}
    ref = torch.load(GOLDEN_V3, weights_only=False)[f"train_{gan_loss}"]["logs"][0]
    keys = list(_LOG_KEYS) if world == 1 else ["D_loss", "MR", "Mask"]
    worst, bad = 0.0, []
    for rk in keys:
        got, want = losses[_LOG_KEYS[rk]], ref[rk]
        tol = 1.01e-4 + tol_rel * abs(want)          # 4 printed decimals + the stated tensor-core tolerance
        if gan_loss == "hinge" and rk == "Adv":
            # -mean(fake logits) ~ -0.0025 after the first D update, whose direction is the SIGN (AdamW) of a gradient
            # that is the 0.5 % residual of two cancelling sums: TF32/bf16 operand rounding moves it by ~1e-3 absolute
            # (tests/test_gpu_golden.py::test_bench_config_tensor_core_graph...; the fp32 kernels match at 1e-4)
            tol += 2e-3
        worst = max(worst, abs(got - want) / tol)
        if abs(got - want) > tol:
            bad.append((rk, got, want))
    return {"checked": True, "ok": not bad, "fixture": "tests/golden/golden_v3.pt (reference train_one_epoch, step 1)",
            "tolerance": f"|got - ref| <= 1.01e-4 + {tol_rel:g} * |ref|", "worst_over_tolerance": worst,
            "compared": keys, "got": {k: losses[_LOG_KEYS[k]] for k in keys}, "ref": {k: ref[k] for k in keys},
            "mismatch": bad}


def measure_variant(dev, mode, steps=10, warmup=3):
    """Throughput of the same step through other paths of this repo (context for the headline):
    "api_eager"  - the drop-in modules driven the way the unmodified train.py drives them: literal schedule (two enhancer
                   forwards, four discriminator passes in the D step), torch.optim.AdamW, clip_grad_norm_, no CUDA graph;
    "fp32"       - the benchmarked schedule (graph, fused AdamW) with every tensor-core path off (fp32 SIMT kernels);
    "skip_dead"  - the benchmarked schedule without the discriminator parameter gradients of the G step, which train.py
                   computes and never reads (StepArgs.skip_dead_d_grads: same weights and losses at every step).  NOT
                   the headline - the headline computes them like the reference does - but it says what they cost."""
    import torch
    from lctgan import config
    from lctgan.training import GraphedTrainStep, StepArgs, build_models, synthetic_batch, train_step
    noisy, clean = (t.to(dev) for t in synthetic_batch(BATCH, SEGMENT, seed=1234))
    try:
        if mode == "api_eager":
            M = build_models(dev, gan_seed=42)
            run = lambda: train_step(*M, noisy, clean, StepArgs(gan_loss="ls"))
        else:
            if mode == "fp32":
                config.set_precision("fp32")
            M = build_models(dev, gan_seed=42, fused_optim=True)
            g = GraphedTrainStep(*M, noisy, clean, StepArgs(gan_loss="ls", reuse_enhancer_forward=True, batch_d_step=True,
                                                            skip_dead_d_grads=mode == "skip_dead",
                                                            defer_dead_d_grads=mode != "skip_dead"), warmup=3)
            run = g
        for _ in range(warmup):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": BATCH / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps}
    finally:
        config.set_precision("bf16")


def _leave(world, code):
    """End of a data-parallel rank.  The step's CUDA graph holds captured NCCL kernels; tearing the communicator down
    under it (dist.destroy_process_group, or the interpreter's own shutdown order) was measured to block for minutes
    after the result line had been printed.  Everything is flushed and the device idle, so the rank leaves at once."""
    if world <= 1:
        return
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(code)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--what", default="step", choices=["step", "rtf"], help="reference arms: time the training step or "
                    "the enhancement real-time factor")
    ap.add_argument("--allow-tf32", action="store_true", help="reference-cuda: allow TF32 in cuDNN / cuBLAS")
    ap.add_argument("--gan_loss", default="ls", choices=["ls", "hinge"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-comparators", action="store_true", help="skip the torch-CUDA / api_eager / fp32 comparator legs")
    ap.add_argument("--no-reuse", action="store_true", help="run the enhancer forward twice per step like train.py does "
                    "(default: one forward serves the D and the G step: identical values, SURVEY 8f N1)")
    ap.add_argument("--no-batch-d", action="store_true", help="D step: clean and enhanced as two discriminator passes "
                    "instead of one batch of 2B")
    ap.add_argument("--torch-optim", action="store_true", help="use torch.optim.AdamW (as train.py builds it) instead of "
                    "the fused multi-tensor AdamW kernel (same update rule; SURVEY 8f N2)")
    ap.add_argument("--skip-dead-d-grads", action="store_true", help="do not compute the discriminator weight gradients "
                    "of the G step (the reference computes and discards them); off by default = the reference's work")
    ap.add_argument("--no-defer-dead-d-grads", action="store_true", help="join the helper streams that finish the (never read) "
                    "discriminator parameter gradients of the G step inside the discriminators' backward instead of at the "
                    "end of the G phase")
    ap.add_argument("--grouped-convs", default="tcgen05", choices=["tcgen05", "mma.sync"], help="A/B: grouped discriminator "
                    "convolutions on the tcgen05 kernels (conv_tc.cu, default) or on the round-1 mma.sync kernels")
    ap.add_argument("--no-capture-nccl", action="store_true", help="N > 1: keep the NCCL all-reduces out of the CUDA graph "
                    "(three graphs with eager exchanges in between) instead of capturing them inside the single graph")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--dp-same-data", action="store_true", help="validation aid: every rank gets rank 0's batch, so the "
                    "averaged gradients - and losses_last_step - must equal the single-GPU run's")
    ap.add_argument("--timeline", default=None, metavar="TRACE.json", help="record 2 steps with torch.profiler (CUPTI kernel "
                    "activity: start, duration and stream of every kernel of the replayed graph) into a chrome trace and "
                    "exit; tools/timeline_summary.py reads it")
    ap.add_argument("--profile-range", action="store_true", help="bracket the timed steps with cudaProfilerStart/Stop "
                    "(for `ncu --profile-from-start off`: the launch list of exactly the timed steps; never a bench value)")
    args = ap.parse_args()
    if args.impl != "ours":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the lctgan kernels have no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from lctgan import _lib, config
    config.grouped_conv_tcgen05 = args.grouped_convs == "tcgen05"
    from lctgan.parallel import FlatGradAllReduce, broadcast_parameters
    from lctgan.training import (GraphedTrainStep, StepArgs, build_models, restore_state, snapshot_state,
                                 synthetic_batch, train_step)

    use_graph = not args.no_graph
    enh, mpd, msd, tf, mr, g_opt, d_opt = build_models(dev, gan_seed=42, capturable=use_graph,
                                                       fused_optim=not args.torch_optim)
    if world > 1:
        broadcast_parameters([enh, mpd, msd])
    # data parallel: in-place all-reduce (SUM) of the gradient arenas the backward passes write, started per
    # sub-discriminator as soon as its backward is done; the 1 / world_size is folded into the consumers (fused AdamW
    # for the discriminators, the global-norm clip for the enhancer) - lctgan/parallel.py
    fold = world > 1 and not args.torch_optim
    sync_d = FlatGradAllReduce(list(mpd.parameters()) + list(msd.parameters()), fold_scale=fold) if world > 1 else None
    sync_g = FlatGradAllReduce(list(enh.parameters()), fold_scale=fold) if world > 1 else None
    if fold:
        d_opt.grad_scale = sync_d.grad_scale
    sargs = StepArgs(gan_loss=args.gan_loss, reuse_enhancer_forward=not args.no_reuse,
                     batch_d_step=not args.no_reuse and not args.no_batch_d,
                     skip_dead_d_grads=args.skip_dead_d_grads, defer_dead_d_grads=not args.no_defer_dead_d_grads)

    noisy_h, clean_h = synthetic_batch(BATCH, SEGMENT, seed=1234 + (0 if args.dp_same_data else rank))
    noisy_h, clean_h = noisy_h.pin_memory(), clean_h.pin_memory()
    noisy_d, clean_d = noisy_h.to(dev), clean_h.to(dev)
    res_h = torch.empty(6, dtype=torch.float32).pin_memory()
    snap = snapshot_state([enh, mpd, msd], [g_opt, d_opt])

    graphed = None
    if use_graph:
        # the whole D+G step (forward, backward, clip, both AdamW updates) as one CUDA graph over static buffers
        # (with N > 1: three graphs with the two NCCL gradient all-reduces launched eagerly in between)
        graphed = GraphedTrainStep(enh, mpd, msd, tf, mr, g_opt, d_opt, noisy_d, clean_d, sargs, after_d_backward=sync_d,
                                   after_g_backward=sync_g, warmup=3, capture_collectives=not args.no_capture_nccl)

    def step(n, c):
        if graphed is not None:
            if n is not noisy_d:
                noisy_d.copy_(n, non_blocking=True)
                clean_d.copy_(c, non_blocking=True)
            return graphed()
        return train_step(enh, mpd, msd, tf, mr, g_opt, d_opt, n, c, sargs, after_d_backward=sync_d,
                          after_g_backward=sync_g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    # ---- parity gate: rewind to the seeded initial state (in place: the captured graph keeps its addresses) and compare
    # the first step of exactly this configuration with the reference's own log at this shape
    restore_state([enh, mpd, msd], [g_opt, d_opt], snap)
    first = {k: float(v) for k, v in step(noisy_d, clean_d).items()}
    parity = parity_check(first, args.gan_loss, world, tol_rel=5e-3) if rank == 0 else None

    for _ in range(args.warmup):
        out = step(noisy_d, clean_d)
    barrier()

    if args.timeline:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step(noisy_d, clean_d)
            torch.cuda.synchronize()
        prof.export_chrome_trace(args.timeline)
        return

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_kernel_launches()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for _ in range(args.steps):
        out = step(noisy_d, clean_d)
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.cudart().cudaProfilerStop()
        return
    launches = _lib.kernel_launches() if graphed is None else graphed.launches_per_step * args.steps
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = BATCH * world * args.steps / (ms_total * 1e-3)

    # ---- end to end: pinned host inputs -> H2D -> step -> D2H of the step's losses, every step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if graphed is not None:
            out = step(noisy_h, clean_h)              # H2D from pinned memory into the graph's static inputs
        else:
            out = step(noisy_h.to(dev, non_blocking=True), clean_h.to(dev, non_blocking=True))
        res = torch.stack([out[k] for k in ("d_loss", "g_loss", "mr", "mask", "adv", "fm")])
        res_h.copy_(res, non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = BATCH * world * args.steps / (ms_e2e * 1e-3)
    losses = {k: float(v) for k, v in zip(("d_loss", "g_loss", "mr", "mask", "adv", "fm"), res_h.tolist())}

    if rank != 0:
        _leave(world, 0)
        return
    peaks = _peaks()
    roof, roof_t = (None, None) if args.no_roofline else measure_roofline(dev, peaks)
    roof_family = roof.pop("family", None) if roof else None
    rtf = None if args.no_roofline else measure_enhance_rtf(dev)
    roof_fe = None if args.no_roofline else measure_frontend_rooflines(dev, peaks)
    cpu = torch_cuda = variants = None
    if world == 1 and not args.no_cpu_baseline:
        r = _sub_bench(["--impl", "reference", "--steps", "2", "--warmup", "1", "--gan_loss", args.gan_loss])
        cpu = dict(r.get("cpu_baseline", {}), ms_per_step=r.get("ms_per_step")) if "error" not in r else r
        if rtf is not None:
            rr = _sub_bench(["--impl", "reference", "--what", "rtf"])
            rtf["cpu_reference"] = rr.get("rtf", rr)
            rtf["cpu_reference_cores"] = rr.get("cores")
    if world == 1 and not args.no_comparators:
        # the practical bar: the same algorithm on stock torch CUDA operators on this very GPU
        torch.cuda.empty_cache()
        torch_cuda = {}
        for key, extra in (("fp32", []), ("tf32", ["--allow-tf32"])):
            r = _sub_bench(["--impl", "reference-cuda", "--steps", "10", "--warmup", "3", "--gan_loss", args.gan_loss] + extra)
            torch_cuda[key] = {k: r.get(k) for k in ("value", "unit", "ms_per_step", "device", "losses_first_step")} \
                if "error" not in r else r
            if "value" in r and r["value"]:
                torch_cuda[key]["ours_over_this"] = value / r["value"]
        variants = {m: measure_variant(dev, m) for m in ("api_eager", "fp32", "skip_dead")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 (tcgen05 dense contraction) / tf32 (grouped discriminator convolutions) / 3xtf32 (generator GEMMs and convolutions) tensor-core operands with fp32 accumulation; f32 elsewhere",
        "data": "synthetic",
        "config": _config(args.gan_loss, world),
        "schedule": {"cuda_graph": graphed is not None, "graphs": len(graphed.graphs) if graphed is not None else 0,
                     "nccl_captured": bool(graphed is not None and world > 1 and len(graphed.graphs) == 1),
                     "capture_error": getattr(graphed, "capture_error", None), "reuse_enhancer_forward": sargs.reuse_enhancer_forward,
                     "batch_d_step": sargs.batch_d_step, "skip_dead_d_grads": sargs.skip_dead_d_grads,
                     "defer_dead_d_grads": sargs.defer_dead_d_grads, "fused_adamw": not args.torch_optim},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(noisy_h.numel() * 4 * 2), "d2h_bytes_per_step": int(res_h.numel() * 4)},
        "gpu_launches": launches,
        "parity_check": parity,
        "roofline": roof,
        "roofline_family": roof_family,
        "roofline_tensor": roof_t,
        "roofline_frontend": roof_fe,
        "enhance_rtf": rtf,
        "cpu_baseline": cpu,
        "torch_cuda_baseline": torch_cuda,
        "api_eager": variants["api_eager"] if variants else None,
        "fp32_mode": variants["fp32"] if variants else None,
        "without_dead_d_grads": variants["skip_dead"] if variants else None,
        "losses_last_step": losses,
    }
    print(json.dumps(line), flush=True)
    bad = bool(parity and parity.get("checked") and not parity["ok"])
    if bad:
        print(f"parity check failed: {parity['mismatch']}", file=sys.stderr, flush=True)
    _leave(world, 1 if bad else 0)
    if bad:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
