"""ctypes binding of liblctgan_sm100.so.

The C ABI in ``include/lctgan.h`` is the single source of truth: this module parses the header
and derives every ctypes signature from it, so a declaration that is missing from the library
(or the reverse) fails loudly at import.  There is deliberately no fallback: if the shared
library has not been built (``python lct-gan_b200/csrc/build.py`` or ``__graft_entry__.build()``)
importing the product raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
LIB_PATH = os.path.join(_HERE, "liblctgan_sm100.so")
HEADER_PATH = os.path.join(_REPO, "include", "lctgan.h")

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER_PATH) -> Dict[str, List[Tuple[str, str]]]:
    """Return {function name: [(ctype spelling, arg name), ...]} for every LCT_API declaration."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    decls: Dict[str, List[Tuple[str, str]]] = {}
    for m in re.finditer(r"LCT_API\s+int\s+(\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out: List[Tuple[str, str]] = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                out.append((mm.group(1).strip(), mm.group(2)))
        decls[name] = out
    return decls


def _to_ctype(spelling: str):
    if "*" in spelling:
        return ctypes.c_void_p
    return _CTYPES[spelling.replace("const ", "").strip()]


class _Library:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python lct-gan_b200/csrc/build.py` "
                "(there is no CPU or PyTorch fallback for the lctgan kernels)")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.decls = parse_header()
        self.fns = {}
        for name, args in self.decls.items():
            try:
                fn = getattr(self.cdll, name)
            except AttributeError as e:
                raise ImportError(f"{LIB_PATH} does not export {name} declared in {HEADER_PATH}") from e
            fn.restype = ctypes.c_int
            fn.argtypes = [_to_ctype(t) for t, _ in args]
            self.fns[name] = (fn, bool(args) and args[-1][0] == "cudaStream_t")


_LIB = None


def lib() -> _Library:
    global _LIB
    if _LIB is None:
        _LIB = _Library()
    return _LIB


_ERRORS = {-1: "invalid argument", -2: "unsupported shape/configuration"}


def ptr(t):
    """Device pointer of a tensor for the C ABI (None -> NULL).  No silent copies, no CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("lctgan kernels are CUDA only (sm_100a); got a CPU tensor - there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError(f"lctgan kernel argument must be contiguous, got strides {t.stride()} for shape {tuple(t.shape)}")
    if t.dtype not in (torch.float32, torch.complex64, torch.bfloat16):
        raise RuntimeError(f"lctgan kernels take float32/complex64/bfloat16 tensors, got {t.dtype}")
    return t.data_ptr()


# the raw handle of the current stream without building a torch.cuda.Stream object per kernel call (the eager path makes
# ~1 100 calls per training step)
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_GET_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def _current_stream_handle() -> int:
    if _RAW_STREAM is not None and _GET_DEVICE is not None:
        return _RAW_STREAM(_GET_DEVICE())
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke ``name`` on the current CUDA stream; tensors are passed as raw device pointers."""
    fn, wants_stream = (_LIB or lib()).fns[name]
    conv = [ptr(a) if isinstance(a, torch.Tensor) else a for a in args]
    if wants_stream:
        conv.append(_current_stream_handle())
    rc = fn(*conv)
    if rc != 0:
        if rc > 0:
            raise RuntimeError(f"{name}: CUDA error {rc}")
        raise RuntimeError(f"{name}: {_ERRORS.get(rc, rc)}")


def call_ret(name: str, *args) -> int:
    """For the few entry points whose return value is data (lct_version, lct_kernel_launches, ...)."""
    fn, _ = lib().fns[name]
    return fn(*args)


def kernel_launches() -> int:
    return call_ret("lct_kernel_launches")


def reset_kernel_launches() -> None:
    call_ret("lct_reset_kernel_launches")
