"""Device plumbing: torch owns device memory, streams and host<->device copies; the kernels
are reached only through the C ABI (``_lib``) with raw device pointers."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("wmf_heavy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def to_dev(a, dtype: torch.dtype | None = None) -> torch.Tensor:
    """numpy / torch -> contiguous CUDA tensor (async copy when the host buffer is pinned)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
        if not a.flags["WRITEABLE"]:
            a = a.copy()
        t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype and not t.is_cuda:
        # narrow on the host only when it shrinks the copy; otherwise convert on the device
        if torch.empty((), dtype=dtype).element_size() < t.element_size():
            t = t.to(dtype)
    if not t.is_cuda:
        t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def workspace(nbytes: int) -> torch.Tensor:
    """Scratch for ONE library call, from torch's caching allocator: stream-ordered (the block is handed out again
    only to later work on the same stream) and private to the call, so operators may be used from several host
    threads and streams at once.  After the first call of a size the allocation is a free-list lookup."""
    dev = require_cuda()
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=dev)


def free_workspaces() -> None:
    torch.cuda.empty_cache()
