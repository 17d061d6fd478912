"""Drop-in for ``wmf_py/cu_py/basics.py``: same names, argument meaning, return types and
error behaviour, computed by ``libwmfb200.so`` on the GPU (no CPU fallback).

    basin_acum(flowdir, mask)              <- basics.py:124-177
    basin_cut(dem, outlet_rc, flowdir, mask) <- basics.py:206-248
    basin_find(x, y, xll, yll, dx, dy, ncols, nrows) <- basics.py:180-203 (host arithmetic)
    basin_2map / basin_map2basin           <- basics.py:251-301 (linear-index scatter/gather)
    dir_reclass_opentopo / dir_reclass_rwatershed <- basics.py:50-121

Inputs are numpy arrays (any integer dtype for ``flowdir``; codes outside 0..7 mean "no
receiver") or torch tensors already on the device; outputs are numpy arrays like the
reference's unless ``device_out=True``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .. import ops
from .._dev import require_cuda, to_dev
from .._lib import ACC_F64, ENC_INTERNAL, WmfError, E_OUTLET_MASK, E_OUTLET_RANGE

_RECLASS_OPENTOPO = {1: 0, 2: 1, 3: 2, 4: 3, 5: 4, 6: 5, 7: 6, 8: 7}          # basics.py:50-59
_RECLASS_RWATERSHED = {1: 0, 2: 7, 3: 6, 4: 5, 5: 4, 6: 3, 7: 2, 8: 1}        # basics.py:61-70


def _flowdir_to_dev(flowdir) -> torch.Tensor:
    """Any integer raster -> int8 device tensor; values outside int8 cannot be valid codes."""
    if isinstance(flowdir, torch.Tensor):
        t = flowdir if flowdir.is_cuda else flowdir.to(require_cuda(), non_blocking=True)
        if t.dtype != torch.int8:
            t = torch.where((t >= 0) & (t <= 9), t, torch.full_like(t, -1)).to(torch.int8)
        return t.contiguous()
    a = np.asarray(flowdir)
    if a.dtype != np.int8:
        # narrow on the host so the copy is 1 B/cell; out-of-range codes become -1 (no receiver)
        a = np.where((a >= 0) & (a <= 9), a, -1).astype(np.int8)
    return to_dev(a)


def _mask_to_dev(mask):
    if mask is None:
        return None
    if isinstance(mask, torch.Tensor):
        t = mask if mask.is_cuda else mask.to(require_cuda(), non_blocking=True)
        return (t != 0).to(torch.uint8).contiguous()
    a = np.asarray(mask)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    elif a.dtype != np.uint8:
        a = (a != 0).astype(np.uint8)
    return to_dev(a)


def basin_acum(flowdir, mask, *, device_out: bool = False, out=None):
    """Upstream cell accumulation following D8 directions; float64 like the reference."""
    require_cuda()
    shape = tuple(flowdir.shape)
    if len(shape) != 2:
        raise ValueError("flowdir must be 2-D")
    if shape[0] == 0 or shape[1] == 0:
        return np.zeros(shape, dtype=float)
    d = _flowdir_to_dev(flowdir)
    m = _mask_to_dev(mask)
    if m is not None and tuple(m.shape) != shape:
        raise ValueError("mask and flowdir shapes differ")
    acc = ops.d8_accum(d, m, acc_dtype=ACC_F64, encoding=ENC_INTERNAL)
    if device_out:
        return acc
    if out is not None:                       # e.g. a pinned host buffer
        torch.from_numpy(out).copy_(acc, non_blocking=False)
        return out
    return acc.cpu().numpy()


def basin_find(x: float, y: float, xll: float, yll: float, dx: float, dy: float, ncols: int, nrows: int) -> Tuple[int, int]:
    """Map coordinates -> (row, col), origin at the lower-left cell (basics.py:180-203)."""
    col = int((x - xll) / dx)
    row = int((y - yll) / dy)
    if not (0 <= col < ncols and 0 <= row < nrows):
        raise ValueError("point outside raster bounds")
    return row, col


def basin_cut(dem, outlet_rc: Tuple[int, int], flowdir, mask) -> Dict[str, np.ndarray]:
    """Delineate the basin draining to ``outlet_rc``; returns dem/flowdir masked to it and the mask."""
    require_cuda()
    ny, nx = flowdir.shape
    r0, c0 = outlet_rc
    if not (0 <= r0 < ny and 0 <= c0 < nx):
        raise ValueError("outlet outside grid")
    d = _flowdir_to_dev(flowdir)
    m = _mask_to_dev(mask)
    try:
        mb, _ = ops.basin_mask(d, m, (int(r0), int(c0)), ENC_INTERNAL)
    except WmfError as e:
        if e.code == E_OUTLET_RANGE:
            raise ValueError("outlet outside grid") from None
        if e.code == E_OUTLET_MASK:
            raise ValueError("outlet not in mask") from None
        raise
    mask_basin = mb.cpu().numpy().astype(bool)
    dem_cut = np.asarray(dem) * mask_basin
    flowdir_cut = np.asarray(flowdir.cpu() if isinstance(flowdir, torch.Tensor) else flowdir) * mask_basin
    return {"dem": dem_cut, "flowdir": flowdir_cut, "mask": mask_basin}


def basin_2map(values, idx_basin_to_map, map_shape: Tuple[int, int], nodata: float) -> np.ndarray:
    """Scatter basin values into a raster by linear index (basics.py:251-287).  Host-side: the
    data are a few MB and the reference mutates ``idx_basin_to_map`` in place."""
    vals = np.asarray(values).ravel(order="C")
    idx = np.asarray(idx_basin_to_map).ravel(order="C")
    if vals.size != idx.size:
        raise ValueError("size of values and indices must match")
    ny, nx = map_shape
    ncell = ny * nx
    if (idx >= ncell).any() or (idx < 0).any():
        rows = idx // 100
        cols = idx % 100
        idx[:] = rows * nx + cols % nx
    if (idx < 0).any() or (idx >= ncell).any():
        raise ValueError("indices out of range")
    out = np.full(ncell, nodata, dtype=vals.dtype)
    out[idx] = vals
    idx_basin_to_map[:] = idx
    return out.reshape(map_shape)


def basin_map2basin(values_map, idx_basin_to_map, basin_shape) -> np.ndarray:
    """Gather basin values from a raster by linear index (basics.py:290-301)."""
    flat = np.asarray(values_map).ravel(order="C")
    if np.max(idx_basin_to_map) >= flat.size or np.min(idx_basin_to_map) < 0:
        raise ValueError("indices out of range")
    return flat[idx_basin_to_map].reshape(basin_shape)


def _reclass(flowdir: np.ndarray, table: Dict[int, int]) -> np.ndarray:
    """basics.py:73-108: pass internal codes through, map external ones, reject unknown >= 0."""
    if flowdir.size == 0:
        return flowdir.astype(int)
    vals = np.unique(flowdir)
    valid_ext = set(table.keys())
    valid_internal = set(range(8))
    if set(int(v) for v in vals).issubset(valid_internal):
        return flowdir.astype(int, copy=True)
    unknown = {int(v) for v in vals if int(v) not in valid_ext | valid_internal and int(v) >= 0}
    if unknown:
        raise ValueError(f"unknown D8 codes: {sorted(unknown)}")
    out = flowdir.astype(int, copy=True)
    for k, v in table.items():
        out[flowdir == k] = v
    return out


def dir_reclass_opentopo(flowdir: np.ndarray) -> np.ndarray:
    return _reclass(np.asarray(flowdir), _RECLASS_OPENTOPO)


def dir_reclass_rwatershed(flowdir: np.ndarray) -> np.ndarray:
    return _reclass(np.asarray(flowdir), _RECLASS_RWATERSHED)
