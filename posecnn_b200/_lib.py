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


def workspace(tag: str, nbytes: int, device, zero: bool = False) -> torch.Tensor:
    """Scratch buffer for one native call, stream-ordered on the current stream.

    Eager calls: grow-only cache keyed by (tag, device, stream) — two streams never share a scratch buffer, and a
    buffer is only ever replaced by a larger one for the SAME stream (the old one returns to the caching allocator,
    which is stream-ordered for that stream).  Under CUDA-graph capture the buffer is allocated fresh from the graph's
    private memory pool and NOT cached: it belongs to the graph for the graph's lifetime, so a later eager call that
    needs a larger scratch can never free (or scribble over) memory a captured graph still replays into.
    `zero=True`: the buffer is zero-filled when it is created (kernels that re-arm their own tickets)."""
    device = torch.device(device)
    n = max(int(nbytes), 256)
    make = torch.zeros if zero else torch.empty
    if torch.cuda.is_current_stream_capturing():
        return make(n, dtype=torch.uint8, device=device)
    key = (tag, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < n:
        buf = make(n, dtype=torch.uint8, device=device)
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
