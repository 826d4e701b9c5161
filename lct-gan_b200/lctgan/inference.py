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


#: What one more enhancer call costs, in samples of audio it could have processed instead: on B200 a call is ~0.9 ms of
#: launch-latency-bound fixed time plus ~0.07 ms per second of (padded) audio in the batch (bench.py enhance_rtf: 1.6 ms for
#: 1 x 9.7 s, 11.8 ms for 16 x 9.9 s), i.e. ~13 s of audio at 16 kHz.  Splitting a batch pays only if it saves more padding.
CALL_OVERHEAD_SAMPLES = 13 * 16000


def length_buckets(lengths: Sequence[int], max_batch: int = 16, max_pad_ratio: float = None,
                   call_overhead: int = CALL_OVERHEAD_SAMPLES) -> List[List[int]]:
    """Indices grouped into batches, longest first, minimising  sum over batches of (call_overhead + rows x longest row)
    - the samples the enhancer processes, padding included, plus the fixed cost of every call - by dynamic programming
    over the length-sorted order (an optimal grouping is contiguous in it).  At most `max_batch` utterances per batch;
    `max_pad_ratio` (optional) additionally forbids padding an utterance beyond that multiple of its own length."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    L = [int(lengths[i]) for i in order]
    n = len(L)
    INF = float("inf")
    best = [0.0] + [INF] * n          # best[i]: cost of the first i (longest) utterances
    cut = [0] * (n + 1)
    for i in range(1, n + 1):
        for j in range(max(0, i - max_batch), i):             # batch = sorted positions j .. i-1, longest = L[j]
            if max_pad_ratio is not None and L[j] > max_pad_ratio * max(L[i - 1], 1):
                continue
            c = best[j] + call_overhead + (i - j) * L[j]
            if c < best[i]:
                best[i], cut[i] = c, j
    buckets: List[List[int]] = []
    i = n
    while i > 0:
        buckets.append(order[cut[i]:i])
        i = cut[i]
    return buckets[::-1]


_PINNED: dict = {}
_DEVICE: dict = {}


def _device_slot(device, slot: int, rows: int, cols: int) -> torch.Tensor:
    key = (str(device), slot)
    buf = _DEVICE.get(key)
    if buf is None or buf.numel() < rows * cols:
        buf = torch.empty(max(rows * cols, 1), dtype=torch.float32, device=device)
        _DEVICE[key] = buf
    return buf[:rows * cols].view(rows, cols)



def _pinned(tag: str, rows: int, cols: int) -> torch.Tensor:
    """A pinned [rows, cols] fp32 staging buffer, grown on demand and kept (cudaHostAlloc costs ~1 ms a call)."""
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < rows * cols:
        buf = torch.empty(max(rows * cols, 1), dtype=torch.float32).pin_memory()
        _PINNED[tag] = buf
    return buf[:rows * cols].view(rows, cols)


@torch.no_grad()
def enhance_utterances(enhancer, waves: Sequence[torch.Tensor], max_batch: int = 16, max_pad_ratio: float = None,
                       device=None) -> List[torch.Tensor]:
    """Enhance variable-length utterances (1-D float tensors on the host) and return the enhanced waveforms (host, true
    lengths, input order).  infer.py:142-157 per bucket: pad -> enhancer -> crop; buckets by `length_buckets`."""
    device = device if device is not None else next(enhancer.parameters()).device
    lens = [int(w.shape[-1]) for w in waves]
    buckets = length_buckets(lens, max_batch, max_pad_ratio)
    copy = torch.cuda.Stream(device=device)
    cur = torch.cuda.current_stream(device)
    out: List[Optional[torch.Tensor]] = [None] * len(waves)

    in_events = [None, None]
    used_events = [None, None]              # enhancer done with the device slot

    def stage(k):                           # host: collate_fn's zero padding into pinned memory; device copy on `copy`
        bucket, slot = buckets[k], k % 2
        if in_events[slot] is not None:
            in_events[slot].synchronize()   # the copy that last read this pinned slot (two buckets ago) is done
        T = lens[bucket[0]]
        h = _pinned(f"in{slot}", len(bucket), T)
        h.zero_()
        for r, i in enumerate(bucket):
            h[r, :lens[i]] = waves[i].reshape(-1).float()
        d = _device_slot(device, slot, len(bucket), T)      # (kept between calls: no allocator traffic per bucket)
        with torch.cuda.stream(copy):
            if used_events[slot] is not None:
                copy.wait_event(used_events[slot])
            d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        in_events[slot] = ev
        return d, ev

    # one pinned result buffer for all buckets (read on the host only after the last copy has landed)
    sizes = [len(bk) * lens[bk[0]] for bk in buckets]
    outbuf = _pinned("out", 1, sum(sizes)).view(-1)
    pending = []                            # (bucket, pinned result view, event)
    nxt = stage(0) if buckets else None
    off = 0
    for k, bucket in enumerate(buckets):
        d, ev = nxt
        nxt = stage(k + 1) if k + 1 < len(buckets) else None                 # overlaps the enhancer below
        cur.wait_event(ev)
        y, _ = enhancer(d)
        done = torch.cuda.Event()
        done.record(cur)
        used_events[k % 2] = done
        if y.shape[1] == lens[bucket[0]]:
            res = outbuf[off:off + sizes[k]].view(len(bucket), lens[bucket[0]])
        else:                               # (an enhancer that changes the length: its own pinned buffer)
            res = torch.empty(y.shape, dtype=torch.float32).pin_memory()
        off += sizes[k]
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
