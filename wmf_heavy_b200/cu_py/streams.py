"""Drop-in for the raster-sized functions of ``wmf_py/cu_py/streams.py``.

    stream_find(acum, threshold)                     <- streams.py:19-36   (acum >= threshold)
    stream_cut(stream, mask_basin)                   <- streams.py:39-42
    stream_find_to_corr(stream, flowdir, outlet_rc)  <- streams.py:45-83   (stream cells connected to the outlet)

The node / edge segmentation (basin_netxy_*) and the seed search produce small Python dictionaries and stay in
the reference's own Python (out of scope).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import ops
from .._dev import require_cuda, to_dev
from .._lib import ENC_INTERNAL
from .basics import _flowdir_to_dev, _mask_to_dev


def stream_find(acum, threshold: int) -> np.ndarray:
    """uint8 mask of the cells whose accumulation reaches ``threshold`` (works on host or device arrays)."""
    if isinstance(acum, torch.Tensor):
        return (acum >= threshold).to(torch.uint8)
    return (np.asarray(acum) >= threshold).astype(np.uint8)


def stream_cut(stream, mask_basin) -> np.ndarray:
    return (np.asarray(stream).astype(np.uint8) & np.asarray(mask_basin).astype(np.uint8)).astype(np.uint8)


def stream_find_to_corr(stream, flowdir, outlet_rc: Tuple[int, int]) -> np.ndarray:
    """Keep only the stream cells whose flow path reaches ``outlet_rc`` through stream cells."""
    require_cuda()
    ny, nx = np.asarray(stream).shape if not isinstance(stream, torch.Tensor) else tuple(stream.shape)
    r0, c0 = outlet_rc
    if not (0 <= r0 < ny and 0 <= c0 < nx):
        raise ValueError("outlet outside grid")
    s = _mask_to_dev(stream)
    active = s.clone()
    active[r0, c0] = 1                       # the search starts at the outlet whether or not it is a stream cell
    mb, _ = ops.basin_mask(_flowdir_to_dev(flowdir), active, (int(r0), int(c0)), ENC_INTERNAL)
    return (mb & s).cpu().numpy().astype(np.uint8)
