#!/usr/bin/env python
"""Data-parallel launcher for the reference's UNMODIFIED train.py (SURVEY.md section 8e "Launcher").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/train_dp.py --ref /path/to/LCT-GAN -- <train.py arguments ...>

One process per GPU.  This script only arranges the environment and then calls the reference's own `train.main()`:
  * `lct-gan_b200/` goes first on sys.path, so `from models.generator import LCTEnhancer ...` (train.py:17-29) resolves
    to the sm_100a drop-in modules; `datasets.datasets` (file I/O, out of scope) is forwarded to the reference checkout;
  * `--device` is forced to this rank's GPU, the seed is offset by the rank (each replica shuffles and crops its own
    stream of segments: per-GPU batch = train.py's --batch_size, weak scaling);
  * `train.train_one_epoch` is wrapped - not edited - so that on its first call the replicas are made identical
    (broadcast from rank 0) and `lctgan.parallel.BackwardEndExchange` is attached to the models it received: gradients
    are then averaged over NCCL from an end-of-backward autograd callback, i.e. after `d_loss.backward()`
    (train.py:199) and after `g_loss.backward()` (train.py:245) before `clip_grad_norm_` (train.py:246-248);
  * validation, logging and checkpoints run on every rank on the same weights; only rank 0 keeps its checkpoint files
    (other ranks write to a scratch directory).
The host logic is covered by tests/test_cabi_host.py::test_backward_end_exchange_for_the_unmodified_loop_gloo_world2;
the loop itself needs the reference checkout and audio files (torchaudio + torchcodec), which this image lacks.
"""
import argparse
import os
import sys
import tempfile


def main():
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--ref", default=os.environ.get("LCT_REF", "/root/reference"))
    known, rest = ap.parse_known_args()
    if rest and rest[0] == "--":
        rest = rest[1:]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    os.environ["LCT_REF"] = known.ref
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.path.insert(0, known.ref)                          # train.py, infer.py, metrics.py
    sys.path.insert(0, os.path.join(root, "lct-gan_b200"))  # ... but models / datasets / losses from the drop-in

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import train as R                                       # the reference's module, unmodified
    from lctgan.parallel import BackwardEndExchange, broadcast_parameters

    inner = R.train_one_epoch
    state = {}

    def train_one_epoch(epoch, loaders, enhancer, mpd, msd, *a, **k):
        if "exchange" not in state and world > 1:
            broadcast_parameters([enhancer, mpd, msd])
            state["exchange"] = BackwardEndExchange(enhancer, [mpd, msd])
        return inner(epoch, loaders, enhancer, mpd, msd, *a, **k)

    R.train_one_epoch = train_one_epoch
    seed0 = R.set_seed

    def set_seed(seed):
        seed0(seed + rank)                                  # every replica draws its own segments

    R.set_seed = set_seed
    argv = [a for a in rest]
    if "--device" in argv:
        i = argv.index("--device")
        del argv[i:i + 2]
    argv += ["--device", f"cuda:{local}"]
    if rank != 0:                                           # one set of run directories / checkpoint files (train.py:541-546)
        if "--expr_root" in argv:
            i = argv.index("--expr_root")
            del argv[i:i + 2]
        argv += ["--expr_root", tempfile.mkdtemp(prefix=f"lctgan_rank{rank}_")]
    sys.argv = ["train.py"] + argv
    try:
        R.main()
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
