"""Input side of the hot path: what feeds `train_one_epoch` its {"noisy", "clean"}
batches once the step runs at thousands of segments per second.

Reference anchors: datasets/datasets.py:131-156 `_crop_pair` (same random start for the noisy and the clean waveform,
centred crop when not random, utterances no longer than the segment are left as they are), :187-230 `collate_fn`
(zero padding to the longest item, "lengths"), train.py:97-142 (the DataLoader around them).  File decoding / resampling
(`torchaudio.load`, datasets.py:112-129) stays on the CPU and out of scope: `UtteranceCache` takes decoded waveforms.

* `UtteranceCache`   - decoded utterances packed into two flat device buffers (180 GB of HBM hold ~700 hours of 16 kHz
                       fp32 audio pairs): one H2D copy per utterance for the whole training run instead of one per epoch.
* `SegmentSampler`   - a training batch = ONE crop kernel over the cache: B (utterance, start) pairs drawn on the host
                       exactly like `_crop_pair` does (torch.randint(0, max_start + 1) per item, same generator
                       stream), cropping / zero padding on the GPU - integer indexing, bit exact against the reference.
* `PinnedPrefetcher` - for data that does come from a host loader: double-buffered pinned staging, the H2D copy of batch
                       k+1 runs on a copy stream while step k computes; batches are handed over with stream events.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

from ._lib import call


def _p64(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


class UtteranceCache:
    def __init__(self, noisy: Sequence[torch.Tensor], clean: Sequence[torch.Tensor], device, ids: Optional[Sequence[str]] = None):
        if len(noisy) != len(clean) or not noisy:
            raise ValueError("UtteranceCache needs the same (non-zero) number of noisy and clean waveforms")
        self.device = torch.device(device)
        self.ids = list(ids) if ids is not None else [str(i) for i in range(len(noisy))]
        self.len_n = torch.tensor([int(w.shape[-1]) for w in noisy], dtype=torch.int64)
        self.len_c = torch.tensor([int(w.shape[-1]) for w in clean], dtype=torch.int64)
        self.off_n = torch.cumsum(self.len_n, 0) - self.len_n
        self.off_c = torch.cumsum(self.len_c, 0) - self.len_c
        pack = lambda ws: torch.cat([w.reshape(-1).float() for w in ws]).pin_memory().to(self.device, non_blocking=True)
        self.noisy, self.clean = pack(noisy), pack(clean)
        self.d_len_n, self.d_len_c = self.len_n.to(self.device), self.len_c.to(self.device)
        self.d_off_n, self.d_off_c = self.off_n.to(self.device), self.off_c.to(self.device)

    def __len__(self) -> int:
        return len(self.ids)


def crop_starts(len_n: torch.Tensor, len_c: torch.Tensor, segment: Optional[int], random_segment: bool,
                generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Start index per item exactly as `_crop_pair` chooses it (datasets.py:131-156): 0 when no cropping applies
    (segment None or min(len) <= segment), else randint(0, max_start + 1) (one draw per cropped item, in order) or the
    centred start max_start // 2."""
    n = len_n.numel()
    starts = torch.zeros(n, dtype=torch.int64)
    if segment is None:
        return starts
    for i in range(n):
        m = int(min(len_n[i], len_c[i]))
        if m <= segment:
            continue
        max_start = m - segment
        starts[i] = int(torch.randint(low=0, high=max_start + 1, size=(1,), generator=generator).item()) \
            if random_segment else max_start // 2
    return starts


class SegmentSampler:
    """Batches of cropped segments straight from the device cache.  `next_batch(indices)` reproduces, for those dataset
    indices, `collate_fn([dataset[i] for i in indices])` of the reference with the same random stream."""

    def __init__(self, cache: UtteranceCache, segment_length: Optional[int], random_segment: bool = True,
                 generator: Optional[torch.Generator] = None):
        self.cache = cache
        self.segment = segment_length
        self.random = random_segment
        self.gen = generator

    def next_batch(self, indices: Sequence[int]) -> Dict:
        c = self.cache
        idx = torch.as_tensor(list(indices), dtype=torch.int64)
        ln, lc = c.len_n[idx], c.len_c[idx]
        starts = crop_starts(ln, lc, self.segment, self.random, self.gen)
        if self.segment is not None:
            crop = torch.minimum(ln, lc) > self.segment
            out_n = torch.where(crop, torch.full_like(ln, self.segment), ln)      # uncropped items keep their length
            out_c = torch.where(crop, torch.full_like(lc, self.segment), lc)
        else:
            out_n, out_c = ln, lc
        T = int(max(out_n.max(), out_c.max()))                                     # collate_fn: pad to the longest
        B = len(idx)
        dev = c.device
        meta = torch.stack([c.off_n[idx], c.off_c[idx], starts + out_n, starts + out_c, starts]).pin_memory()
        meta = meta.to(dev, non_blocking=True)            # ONE small H2D copy: offsets, crop ends, starts
        noisy = torch.empty(B, T, dtype=torch.float32, device=dev)
        clean = torch.empty(B, T, dtype=torch.float32, device=dev)
        call("lct_crop_segments", c.noisy, c.clean, _p64(meta[0]), _p64(meta[1]), _p64(meta[2]), _p64(meta[3]),
             _p64(meta[4]), noisy, clean, B, T)
        return {"id": [c.ids[i] for i in idx.tolist()], "noisy": noisy, "clean": clean, "lengths": out_n.clone(),
                "sr": None}


class PinnedPrefetcher:
    """Wrap a host iterator of {"noisy", "clean", ...} batches: two pinned staging slots per key, H2D copies on a copy
    stream one batch ahead of the consumer (reference train.py:165-169 copies synchronously inside the step)."""

    def __init__(self, loader: Iterable[Dict], device, keys: Tuple[str, ...] = ("noisy", "clean")):
        self.loader = loader
        self.device = torch.device(device)
        self.keys = keys
        self.copy = torch.cuda.Stream(device=self.device)
        self._slots: List[Dict[str, torch.Tensor]] = [{}, {}]

    def _stage(self, batch: Dict, slot: int):
        out = dict(batch)
        with torch.cuda.stream(self.copy):
            for k in self.keys:
                src = batch[k]
                pin = self._slots[slot].get(k)
                if pin is None or pin.shape != src.shape or pin.dtype != src.dtype:
                    pin = torch.empty(src.shape, dtype=src.dtype).pin_memory()
                    self._slots[slot][k] = pin
                pin.copy_(src)
                out[k] = pin.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        return out, ev

    def __iter__(self) -> Iterator[Dict]:
        it = iter(self.loader)
        slot = 0
        try:
            nxt = self._stage(next(it), slot)
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            slot ^= 1
            try:
                # the slot being refilled was consumed two iterations ago; its device tensor was handed out, the pinned
                # source is safe to overwrite once that copy's event has completed
                if self._slots[slot]:
                    self.copy.synchronize()
                nxt = self._stage(next(it), slot)
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for k in self.keys:
                cur[k].record_stream(torch.cuda.current_stream(self.device))
            yield cur
