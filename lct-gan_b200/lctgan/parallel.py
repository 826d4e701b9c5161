"""Data parallelism for the adversarial step: one process per GPU, identical replicas, one gradient exchange per
optimiser group per step over NCCL / NVLink.  The reference is single-device
(train.py:552-553); this layer is new.

* after ``d_loss.backward()``: the 17.7 M discriminator gradients (70.8 MB; 89 % of it the three MSD convs.5 weights)
* after ``g_loss.backward()`` and before gradient clipping: the 135 k enhancer gradients (0.54 MB).  The
  discriminator gradients the generator backward also produces are dead (the reference's next ``zero_grad`` discards
  them) and are never exchanged.

Zero-copy exchange.  The backward passes of this build write the gradients of one layer stack (a sub-discriminator)
or of the whole generator into ONE flat buffer ("arena") and hand ``.grad`` views of it to autograd, so the
collective runs in place on those buffers: no gather / scatter copies.  ``reduce_async(arena)`` is called by the
stack's backward the moment its arena is complete and launches the all-reduce (SUM) on a communication stream, so
the exchange of the sub-discriminators that finish first overlaps the backward of the others; ``__call__`` joins.
The 1 / world_size is folded into the consumer (FusedAdamW.grad_scale, clip_grad_norm_(pre_scale=...)) when
``fold_scale`` is set, else applied in place by one kernel per arena.  Gradients that do not live in an arena (a
parameter driven through another code path, the gloo / CPU tests) take the packed path: one multi-tensor copy into a
flat buffer, one collective, one copy back.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatGradAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, fold_scale: bool = False):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradAllReduce: no parameter requires grad")
        self.group = group
        self.fold_scale = bool(fold_scale)
        self.numel = sum(p.numel() for p in self.params)
        self.device = self.params[0].device
        self.flat: Optional[torch.Tensor] = None        # packed path only (allocated on first use)
        self.views: List[torch.Tensor] = []
        self._pending: List[torch.Tensor] = []          # arenas whose all-reduce is in flight on the comm stream
        self._comm: Optional[torch.cuda.Stream] = None

    # ---------------------------------------------------------------------------------------------- helpers
    def _world(self) -> int:
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    @property
    def grad_scale(self) -> float:
        """What the consumer must multiply the exchanged gradients by when ``fold_scale`` is set."""
        return 1.0 / self._world()

    def _comm_stream(self) -> torch.cuda.Stream:
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=self.device, priority=-2)
        return self._comm

    def _scale(self, t: torch.Tensor, w: int) -> None:
        if self.fold_scale or w == 1:
            return
        if t.is_cuda:
            ops.call("lct_axpby", t, t, t, t.numel(), 1.0 / w, 0.0)
        else:
            t.div_(w)

    # ---------------------------------------------------------------------------------------------- zero-copy path
    def reduce_async(self, arena: torch.Tensor) -> None:
        """Start the in-place all-reduce (SUM) of a flat buffer that holds finished gradients of this group; returns
        at once.  Must be followed by ``__call__`` before anything reads the gradients."""
        w = self._world()
        if w == 1 or arena is None or arena.numel() == 0:
            return
        if not arena.is_cuda:
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.group)
            self._scale(arena, w)
            self._pending.append(arena)
            return
        cur = torch.cuda.current_stream(arena.device)
        comm = self._comm_stream()
        comm.wait_stream(cur)
        with torch.cuda.stream(comm):
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.group)
            self._scale(arena, w)
        self._pending.append(arena)      # (kept alive here and through the .grad views until the join in __call__)

    def _covered(self, g: torch.Tensor) -> bool:
        a, b = g.data_ptr(), g.data_ptr() + g.numel() * g.element_size()
        for t in self._pending:
            lo = t.data_ptr()
            if lo <= a and b <= lo + t.numel() * t.element_size():
                return True
        return False

    # ---------------------------------------------------------------------------------------------- the exchange
    def __call__(self) -> None:
        """Make every ``.grad`` of this group the sum (``fold_scale``) or the mean over ranks."""
        w = self._world()
        if w == 1:
            self._pending.clear()
            return
        if self._pending and self._comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm)
        rest = [p for p in self.params if p.grad is None or not self._covered(p.grad)]
        self._pending.clear()
        if rest:
            self._packed(rest, w)

    def _packed(self, params: List[torch.nn.Parameter], w: int) -> None:
        if self.flat is None:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        views, o = [], 0
        for p in params:
            views.append(self.flat[o:o + p.numel()])
            o += p.numel()
        flat = self.flat[:o]
        srcs, dsts = [], []
        for p, v in zip(params, views):
            if p.grad is None:
                v.zero_()              # a parameter without gradient on this rank still takes part in the average
            else:
                srcs.append(p.grad.contiguous())
                dsts.append(v)
        if srcs:
            if flat.is_cuda:
                ops.mt_copy(srcs, dsts)
            else:                      # gloo / CPU path used by the world_size-2 host-logic tests
                for s, d in zip(srcs, dsts):
                    d.copy_(s.reshape(-1))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self._scale(flat, w)
        back_src, back_dst = [], []
        for p, v in zip(params, views):
            if p.grad is None or not p.grad.is_contiguous():
                p.grad = v.view_as(p).clone()
            else:
                back_src.append(v)
                back_dst.append(p.grad)
        if back_src:
            if flat.is_cuda:
                ops.mt_copy(back_src, back_dst)
            else:
                for s, d in zip(back_src, back_dst):
                    d.view(-1).copy_(s)


def broadcast_parameters(modules: Iterable[torch.nn.Module], src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s weights and buffers."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)


# --------------------------------------------------------------------------------------------------------------
# Data parallelism for the UNMODIFIED reference loop (train.py:145-258 `train_one_epoch`), which has no call site for
# an exchange: autograd hooks queue an end-of-backward callback (the mechanism DDP uses), so the gradients are
# averaged after `d_loss.backward()` (train.py:199) and after `g_loss.backward()` (train.py:245) - i.e. before
# `clip_grad_norm_` (train.py:246-248), which an optimiser pre-hook would miss.
# --------------------------------------------------------------------------------------------------------------
class BackwardEndExchange:
    """attach(enhancer, [mpd, msd]): after every backward pass, average the gradients of the group that pass was FOR:
    a backward that reached the enhancer's parameters is the generator step (its discriminator gradients are dead:
    not exchanged); a backward that reached only discriminator parameters is the discriminator step."""

    def __init__(self, enhancer: torch.nn.Module, discriminators: Iterable[torch.nn.Module], group=None):
        self.sync_g = FlatGradAllReduce(enhancer.parameters(), group=group)
        self.sync_d = FlatGradAllReduce([p for m in discriminators for p in m.parameters()], group=group)
        self._queued = False
        self._saw_g = False
        self.exchanges = {"g": 0, "d": 0}
        self._handles = []
        for p in self.sync_g.params:
            self._handles.append(p.register_hook(self._make_hook(True)))
        for p in self.sync_d.params:
            self._handles.append(p.register_hook(self._make_hook(False)))

    def _make_hook(self, is_g: bool):
        def hook(grad):
            if is_g:
                self._saw_g = True
            if not self._queued:
                self._queued = True
                torch.autograd.Variable._execution_engine.queue_callback(self._finish)
            return None
        return hook

    def _finish(self) -> None:
        saw_g, self._queued, self._saw_g = self._saw_g, False, False
        if saw_g:
            self.sync_g()
            self.exchanges["g"] += 1
        else:
            self.sync_d()
            self.exchanges["d"] += 1

    def detach(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles.clear()
