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


# ---- principal stream (metrics.py:337-490) ------------------------------------------------------------------
_ppal_stream_profile = None                      # like the reference: _find leaves the profile for _cut


def basin_ppalstream_find(stream, flowdir, dem, dx: float, dy: float, outlet_rc: Tuple[int, int]):
    """(n_cells, start_rc) of the longest stream: the stream cell with the largest distance along the stream network
    to the end of its path (first one in row-major order on ties), and the profile from there down, kept for
    ``basin_ppalstream_cut``.  ``outlet_rc`` is accepted and unused, as in the reference (metrics.py:337-470)."""
    global _ppal_stream_profile
    require_cuda()
    s_host = np.asarray(stream).astype(bool)
    d_host = np.asarray(flowdir)
    ny, nx = s_host.shape
    d = _flowdir_to_dev(flowdir)
    s = _mask_to_dev(s_host)
    dist = ops.path_sum(d, s, None, dx, dy, cycle_nan=False, encoding=ENC_INTERNAL)
    dist = torch.where(s.bool(), dist, torch.zeros_like(dist))             # non-stream cells hold 0.0 (metrics.py:420)
    flat = dist.reshape(-1)
    start_flat = int(torch.nonzero(flat == flat.max())[0, 0])               # np.argmax: the first maximum
    start_rc = divmod(start_flat, nx)
    # the profile is one path: walk it on the host (its cells, not the raster, bound the work)
    off = {0: (0, 1), 1: (1, 1), 2: (1, 0), 3: (1, -1), 4: (0, -1), 5: (-1, -1), 6: (-1, 0), 7: (-1, 1)}
    diag = (dx ** 2 + dy ** 2) ** 0.5
    elev, dst, rows, cols = [], [], [], []
    r, c = start_rc
    cum = 0.0
    seen = 0
    while True:
        elev.append(float(dem[r][c])); dst.append(cum); rows.append(r); cols.append(c)
        o = off.get(int(d_host[r, c]))
        if not o:
            break
        rr, cc = r + o[0], c + o[1]
        if not (0 <= rr < ny and 0 <= cc < nx) or not s_host[rr, cc]:
            break
        cum += dx if o[0] == 0 else (dy if o[1] == 0 else diag)
        r, c = rr, cc
        seen += 1
        if seen > ny * nx:
            raise RuntimeError("the principal stream runs into a drainage cycle (the reference does not terminate)")
    _ppal_stream_profile = np.vstack([np.array(elev, dtype=float), np.array(dst, dtype=float), np.array(rows, dtype=int),
                                      np.array(cols, dtype=int)])
    return _ppal_stream_profile.shape[1], (int(start_rc[0]), int(start_rc[1]))


def basin_ppalstream_cut(*args, **kwargs) -> np.ndarray:
    """(4, N) elevation, cumulative distance, row, col of the principal stream found last (metrics.py:473-490)."""
    if _ppal_stream_profile is None:
        raise RuntimeError("basin_ppalstream_find must be called before cutting")
    return _ppal_stream_profile.copy()


def basin_time_to_out(flowdir, dx: float, dy: float, slope, n_manning) -> np.ndarray:
    """Travel time from every cell to the end of its flow path, Manning with unit hydraulic radius; nan on and behind
    drainage cycles and behind cells without a positive speed (metrics.py:528-599)."""
    require_cuda()
    d = _flowdir_to_dev(flowdir)
    sl = to_dev(np.ascontiguousarray(slope, dtype=np.float64))
    nm = to_dev(np.ascontiguousarray(np.broadcast_to(np.asarray(n_manning, dtype=np.float64), sl.shape)))
    return ops.time_to_out(d, sl, nm, dx, dy, ENC_INTERNAL).cpu().numpy()
