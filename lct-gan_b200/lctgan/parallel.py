"""Data parallelism for the adversarial step: one process per GPU, identical replicas, and exactly
one NCCL all-reduce (average) over a flat fp32 gradient buffer per optimiser group per step
(SURVEY.md section 8e).  The reference is single-device; this layer is new.

* after ``d_loss.backward()``: all-reduce of the 17.7 M discriminator gradients (70.8 MB)
* after ``g_loss.backward()`` and before gradient clipping: all-reduce of the 135 k enhancer
  gradients (0.54 MB).  The discriminator gradients that the generator backward also produces
  are discarded by the reference's next ``zero_grad`` and are therefore never exchanged.

Gradients are gathered into / scattered from the flat buffer with one multi-tensor copy launch per
64 tensors (lct_mt_copy), so the exchange is a single collective whatever the parameter count.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatGradAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGradAllReduce: no parameter requires grad")
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views: List[torch.Tensor] = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()])
            o += p.numel()
        self.numel = n

    def __call__(self) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        srcs, dsts, missing = [], [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                missing.append(v)
            else:
                srcs.append(p.grad.contiguous())
                dsts.append(v)
        for v in missing:      # a parameter without gradient on this rank still takes part in the average
            v.zero_()
        if self.flat.is_cuda:
            ops.mt_copy(srcs, dsts)
        else:                  # gloo / CPU path used by the world_size-2 host-logic tests
            for s, d in zip(srcs, dsts):
                d.copy_(s.reshape(-1))
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(dist.get_world_size(self.group))
        back_src, back_dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.view_as(p).clone()
            else:
                back_src.append(v)
                back_dst.append(p.grad if p.grad.is_contiguous() else None)
                if back_dst[-1] is None:
                    p.grad = v.view_as(p).clone()
                    back_src.pop()
                    back_dst.pop()
        if back_src:
            if self.flat.is_cuda:
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
