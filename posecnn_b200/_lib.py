"""ctypes binding of libposecnn_b200.so (include/posecnn_b200.h).

There is NO fallback: if the shared object is missing or a call fails, the ops raise.
PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os

import torch

from .build import LIB

_lib = None
_ws = {}

c_fp = ctypes.c_void_p  # device pointers travel as void*


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(
                f"{LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(posecnn_b200 has no CPU or PyTorch fallback)")
        _lib = ctypes.CDLL(LIB)
        _lib.pcnn_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"posecnn_b200 native call failed ({rc}): {lib().pcnn_last_error().decode()}")


def ptr(t):
    if t is None:
        return c_fp(0)
    return c_fp(t.data_ptr())


def stream():
    return c_fp(torch.cuda.current_stream().cuda_stream)


def f32(x: float):
    return ctypes.c_float(float(x))


def workspace(tag: str, nbytes: int, device) -> torch.Tensor:
    """Grow-only per-(tag, device) scratch buffer; stream-ordered reuse on the current stream."""
    key = (tag, str(device))
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def require_cuda(name: str, t: torch.Tensor, dtype, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (posecnn_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() not in (ndim if isinstance(ndim, (tuple, list)) else (ndim,)):
        raise ValueError(f"{name} must be {ndim}-dimensional")  # OP_REQUIRES rank checks of the reference
    return t.contiguous()
