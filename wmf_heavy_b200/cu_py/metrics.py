"""Drop-in for the raster-sized functions of ``wmf_py/cu_py/metrics.py``: same names,
arguments, return types and nan / -1 conventions, computed by ``libwmfb200.so`` (pointer jumping over the
receiver graph with the stream cells as roots).

    hydro_distance_and_receiver(flowdir, stream_mask, mask_basin, dx, dy)  <- metrics.py:189-268
    hand(dem, flowdir, stream_mask, mask_basin)                            <- metrics.py:271-323
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import ops
from .._dev import require_cuda, to_dev
from .._lib import ENC_INTERNAL
from .basics import _flowdir_to_dev, _mask_to_dev


def hydro_distance_and_receiver(flowdir, stream_mask, mask_basin, dx: float, dy: float) -> Tuple[np.ndarray, np.ndarray]:
    """(hdnd_model float64, a_quien int): distance along the flow path to the first stream cell and that cell's
    linear index; nan / -1 outside the basin or when no stream is reached; 0 / own index on stream cells."""
    require_cuda()
    d = _flowdir_to_dev(flowdir)
    aq, hd = ops.stream_receiver(d, _mask_to_dev(mask_basin), _mask_to_dev(stream_mask), dx, dy, ENC_INTERNAL)
    return hd.cpu().numpy(), aq.cpu().numpy().astype(int)


def hand(dem, flowdir, stream_mask, mask_basin) -> np.ndarray:
    """Height Above Nearest Drainage following the D8 flow path (float64; nan outside the basin)."""
    require_cuda()
    d = _flowdir_to_dev(flowdir)
    z = to_dev(np.ascontiguousarray(dem, dtype=np.float64))
    return ops.hand(z, d, _mask_to_dev(mask_basin), _mask_to_dev(stream_mask), ENC_INTERNAL).cpu().numpy()
