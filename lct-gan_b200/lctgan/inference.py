"""Validation / inference tail of the hot path.

Reference anchors: train.py:261-282 `_si_sdr_torch` (one utterance at a time, a `.item()` host sync each),
train.py:285-385 `validate_and_compute_metrics`, infer.py:130-160 `run_inference`, datasets/datasets.py:187-230
`collate_fn` (zero-pads every batch to its longest utterance).

* `si_sdr`             - the whole batch in ONE kernel launch with per-row valid lengths, result stays on the device.
* `length_buckets`     - batches of similar length: the reference pads a batch to its longest member, so a batch that
                         mixes 1 s and 10 s utterances runs the enhancer on mostly zeros (BASELINE configs[1]: batch 16 of
                         1-10 s utterances carries 98.7 s of audio in 16 x 9.93 s of samples).  Sorting by length and
                         cutting buckets at a bounded padding ratio removes that waste; every bucket is still exactly a
                         `collate_fn` batch (zero padding at the end, results cropped to the true lengths).
* `enhance_utterances` - bucketed enhancement with pinned staging: the H2D copy of bucket k+1 and the D2H copy of bucket
                         k-1 overlap the enhancer on bucket k (copy stream + events, no host sync inside the loop).
* `validate`           - mirror of validate_and_compute_metrics: MR-STFT loss + SI-SDR, one host sync at the very end
                         (PESQ / STOI are third-party CPU metrics: out of scope, reported as NaN like the reference
                         does when the packages are missing).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from ._lib import call


def si_sdr(reference: torch.Tensor, estimate: torch.Tensor, lengths: Optional[torch.Tensor] = None,
           eps: float = 1e-8) -> torch.Tensor:
    """SI-SDR in dB per row: reference [B, Tr], estimate [B, Te] (compared over min(Tr, Te, lengths[b]) samples, both
    made zero-mean over that span - train.py:261-282).  Returns a float32 [B] tensor on the device (no host sync)."""
    if reference.dim() == 1:
        reference, estimate = reference.unsqueeze(0), estimate.unsqueeze(0)
    if reference.dim() != 2 or estimate.dim() != 2 or reference.shape[0] != estimate.shape[0]:
        raise ValueError(f"Expected reference / estimate of shape [B, T], got {reference.shape}, {estimate.shape}")
    if not (reference.is_cuda and estimate.is_cuda):
        raise RuntimeError("lctgan si_sdr is CUDA only (sm_100a); there is no CPU fallback")
    reference = reference.detach().float().contiguous()
    estimate = estimate.detach().float().contiguous()
    out = torch.empty(reference.shape[0], dtype=torch.float32, device=reference.device)
    if lengths is not None:
        lengths = lengths.to(device=reference.device, dtype=torch.int64).contiguous()
    call("lct_si_sdr", reference, estimate, _i64(lengths), out, reference.shape[0], reference.shape[1],
         estimate.shape[1], float(eps))
    return out


def _i64(t):
    """int64 device tensor -> raw pointer for the C ABI (the float-only `ptr` helper refuses other dtypes)."""
    import ctypes
    if t is None:
        return None
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise RuntimeError("expected a contiguous CUDA int64 tensor")
    return ctypes.c_void_p(t.data_ptr())


def length_buckets(lengths: Sequence[int], max_batch: int = 16, max_pad_ratio: float = 1.1) -> List[List[int]]:
    """Indices grouped into batches: sorted by length (longest first), a bucket is closed when it holds `max_batch`
    utterances or when the next (shorter) utterance would be padded by more than `max_pad_ratio` x its own length."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    buckets: List[List[int]] = []
    cur: List[int] = []
    for i in order:
        if cur and (len(cur) >= max_batch or int(lengths[cur[0]]) > max_pad_ratio * max(int(lengths[i]), 1)):
            buckets.append(cur)
            cur = []
        cur.append(i)
    if cur:
        buckets.append(cur)
    return buckets


@torch.no_grad()
def enhance_utterances(enhancer, waves: Sequence[torch.Tensor], max_batch: int = 16, max_pad_ratio: float = 1.1,
                       device=None) -> List[torch.Tensor]:
    """Enhance variable-length utterances (1-D float tensors on the host) and return the enhanced waveforms (host, true
    lengths, input order).  infer.py:142-157 per bucket: pad -> enhancer -> crop; buckets by `length_buckets`."""
    device = device if device is not None else next(enhancer.parameters()).device
    lens = [int(w.shape[-1]) for w in waves]
    buckets = length_buckets(lens, max_batch, max_pad_ratio)
    copy = torch.cuda.Stream(device=device)
    cur = torch.cuda.current_stream(device)
    out: List[Optional[torch.Tensor]] = [None] * len(waves)

    def stage(bucket):                      # host: collate_fn's zero padding into pinned memory; device copy on `copy`
        T = lens[bucket[0]]
        h = torch.zeros(len(bucket), T, dtype=torch.float32).pin_memory()
        for r, i in enumerate(bucket):
            h[r, :lens[i]] = waves[i].reshape(-1).float()
        with torch.cuda.stream(copy):
            d = h.to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        return h, d, ev

    pending = []                            # (bucket, pinned result, event)
    nxt = stage(buckets[0]) if buckets else None
    for k, bucket in enumerate(buckets):
        h, d, ev = nxt
        nxt = stage(buckets[k + 1]) if k + 1 < len(buckets) else None        # overlaps the enhancer below
        cur.wait_event(ev)
        y, _ = enhancer(d)
        done = torch.cuda.Event()
        done.record(cur)
        res = torch.empty(y.shape, dtype=torch.float32).pin_memory()
        with torch.cuda.stream(copy):
            copy.wait_event(done)
            res.copy_(y, non_blocking=True)
            ev2 = torch.cuda.Event()
            ev2.record(copy)
        y.record_stream(copy)
        pending.append((bucket, res, ev2))
    for bucket, res, ev2 in pending:
        ev2.synchronize()
        for r, i in enumerate(bucket):
            out[i] = res[r, :lens[i]].clone()
    return out            # type: ignore[return-value]


@torch.no_grad()
def validate(enhancer, mrstft_loss, batches: Iterable[Dict], device=None) -> Dict[str, float]:
    """validate_and_compute_metrics (train.py:285-385) without per-utterance host work: `batches` yields collate_fn
    dictionaries ("noisy", "clean" [B, T] and optionally "lengths" [B]).  Sample-weighted MR-STFT loss and mean SI-SDR;
    the sums stay on the device and are read once at the end."""
    device = device if device is not None else next(enhancer.parameters()).device
    was_training = enhancer.training
    enhancer.eval()
    tot_mr = torch.zeros((), dtype=torch.float64, device=device)
    tot_sdr = torch.zeros((), dtype=torch.float64, device=device)
    count = 0
    for batch in batches:
        noisy = batch["noisy"].to(device, non_blocking=True)
        clean = batch["clean"].to(device, non_blocking=True)
        lengths = batch.get("lengths")
        enhanced, _ = enhancer(noisy)
        mr, _ = mrstft_loss(enhanced, clean)
        B = noisy.shape[0]
        tot_mr += mr.double() * B
        tot_sdr += si_sdr(clean, enhanced, lengths).double().sum()
        count += B
    if was_training:
        enhancer.train()
    n = max(count, 1)
    mr_v, sdr_v = (tot_mr / n).item(), (tot_sdr / n).item()
    return {"val_mrstft": float(mr_v), "val_si_sdr": float(sdr_v), "val_pesq": float("nan"), "val_stoi": float("nan")}
