"""BASELINE configs[4]: throughput sweep of the D+G training step over segment length x per-GPU batch on ONE GPU
(the step replayed from a CUDA graph, device-resident synthetic inputs, same switches as bench.py).

    python tools/sweep.py [--steps 10] [--cases 1x8,2x8,4x8,2x32,4x32,2x128]

Prints one JSON line per case: {"segment_s", "batch", "ms_per_step", "samples_per_s", "audio_s_per_s", "peak_mem_GB",
"losses"}.  A case that does not fit (kernel index range or memory) is reported with "error" instead of numbers."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lct-gan_b200")); sys.path.insert(0, ROOT)
import torch


def run_case(seg_s, batch, steps, dev):
    from lctgan.training import GraphedTrainStep, StepArgs, build_models
    from lctgan import training as O          # synthetic_batch lives with the product
    T = int(round(seg_s * 16000))
    torch.cuda.reset_peak_memory_stats(dev)
    enh, mpd, msd, tf, mr, g_opt, d_opt = build_models(dev, gan_seed=42, capturable=True, fused_optim=True)
    sargs = StepArgs(gan_loss="ls", reuse_enhancer_forward=True, batch_d_step=True, defer_dead_d_grads=True)
    noisy, clean = (t.to(dev) for t in O.synthetic_batch(batch, T, seed=1234))
    graphed = GraphedTrainStep(enh, mpd, msd, tf, mr, g_opt, d_opt, noisy, clean, sargs, warmup=3)
    for _ in range(3):
        out = graphed()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = graphed()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"segment_s": seg_s, "batch": batch, "ms_per_step": ms, "samples_per_s": batch / (ms * 1e-3),
            "audio_s_per_s": batch * seg_s / (ms * 1e-3),
            "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
            "losses": {k: float(out[k]) for k in ("d_loss", "g_loss")}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cases", default="1x8,2x8,4x8,2x32,4x32,2x128")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.zeros(1, device=dev)
    for c in args.cases.split(","):
        s, b = c.split("x")
        try:
            r = run_case(float(s), int(b), args.steps, dev)
        except Exception as e:                              # report and go on with the next case
            import traceback; traceback.print_exc()
            r = {"segment_s": float(s), "batch": int(b), "error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
