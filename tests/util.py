"""Shared helpers for the parity tests."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import lct_oracle
    return lct_oracle


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b| (the "relative to the largest reference value" error used for tolerances)."""
    a = a.detach().cpu()
    b = b.detach().cpu()
    if torch.is_complex(a) or torch.is_complex(b):
        a = torch.view_as_real(a.to(torch.complex128).contiguous())
        b = torch.view_as_real(b.to(torch.complex128).contiguous())
    a = a.double()
    b = b.double()
    assert a.shape == b.shape, (a.shape, b.shape)
    den = b.abs().max().item()
    if den == 0.0:
        den = 1.0
    return (a - b).abs().max().item() / den


def rel_err_where(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> float:
    """rel_err restricted to the elements selected by `mask` (still relative to the largest reference value overall)."""
    a = a.detach().cpu().double()
    b = b.detach().cpu().double()
    mask = mask.detach().cpu()
    assert a.shape == b.shape == mask.shape, (a.shape, b.shape, mask.shape)
    den = b.abs().max().item() or 1.0
    d = (a - b).abs()[mask]
    return (d.max().item() if d.numel() else 0.0) / den


def cpu_params(module: torch.nn.Module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def leaf_params(P, buffers=("stft.window",)):
    out = {}
    for k, v in P.items():
        t = v.detach().clone()
        if k not in buffers and not k.endswith("window"):
            t.requires_grad_(True)
        out[k] = t
    return out
