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

# external code (1..8) -> internal code (0..7) of the two conventions the reference knows (basics.py:50-70):
# OpenTopo numbers the eight neighbours clockwise from East starting at 1, r.watershed counter-clockwise.
_EXT_TO_INTERNAL = {"opentopo": lambda e: e - 1, "rwatershed": lambda e: (9 - e) % 8}


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
    # The result comes back through page-locked memory from torch's caching host allocator: a pageable 8 B/cell copy runs
    # at a fraction of the PCIe rate (2.1 GB for a 16384^2 raster), and after the first call the pinned block is reused.
    try:
        host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    except RuntimeError:                      # (no page-locked memory left: the slow way)
        return acc.cpu().numpy()
    host.copy_(acc)
    return host.numpy()


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


_STORE = {1: np.uint8, 2: np.int16, 4: np.int32, 8: np.int64}     # torch-copyable integer type of the same width


def _vec_to_dev(a) -> torch.Tensor:
    """Flattened (C order) device copy of a host array, moved as same-width integers (bit pattern preserved)."""
    a = np.ascontiguousarray(a).reshape(-1)
    if a.dtype.kind not in "iufb" or a.dtype.itemsize not in _STORE:
        raise TypeError(f"unsupported dtype {a.dtype}")
    return to_dev(a.view(_STORE[a.dtype.itemsize]))


def basin_2map(values, idx_basin_to_map, map_shape: Tuple[int, int], nodata: float) -> np.ndarray:
    """Basin vector -> raster by linear index (replaces basics.py:251-287): a raster of ``nodata`` with
    ``values`` written at ``idx_basin_to_map`` (device scatter; the last of duplicate indices wins).  Like the
    reference, indices that do not fit the raster are first re-read as ``row*100 + col`` and the repaired indices
    are written back into the caller's array."""
    vals = np.asarray(values)
    idx_in = np.asarray(idx_basin_to_map)
    if vals.size != idx_in.size:
        raise ValueError("size of values and indices must match")
    ny, nx = (int(v) for v in map_shape)
    ncell = ny * nx
    if vals.size == 0:
        return np.full(ncell, nodata, dtype=vals.dtype).reshape(ny, nx)
    idx = to_dev(np.ascontiguousarray(idx_in, dtype=np.int64).reshape(-1))
    if ops.index_out_of_range(idx, ncell):
        idx = idx.clone()
        ops.index_repair(idx, nx)
        if ops.index_out_of_range(idx, ncell):
            raise ValueError("indices out of range")
        if isinstance(idx_basin_to_map, np.ndarray) and idx_basin_to_map.flags.writeable:
            idx_basin_to_map[...] = idx.cpu().numpy().reshape(idx_basin_to_map.shape)
    fill_bits = int(np.array([nodata]).astype(vals.dtype).view(_STORE[vals.dtype.itemsize])[0])
    out = ops.scatter_last(_vec_to_dev(vals), idx, ncell, fill_bits)
    return out.cpu().numpy().view(vals.dtype).reshape(ny, nx)


def basin_map2basin(values_map, idx_basin_to_map, basin_shape) -> np.ndarray:
    """Raster -> basin vector by linear index (replaces basics.py:290-301), a device gather."""
    m = np.asarray(values_map)
    idx = to_dev(np.ascontiguousarray(idx_basin_to_map).reshape(-1).astype(np.int64, copy=False), torch.int64)
    if idx.numel() == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if ops.index_out_of_range(idx, m.size):
        raise ValueError("indices out of range")
    return ops.gather(_vec_to_dev(m), idx).cpu().numpy().view(m.dtype).reshape(basin_shape)


def _reclass(flowdir, convention: str) -> np.ndarray:
    """External D8 codes -> internal 0..7 with a 256-entry lookup table on the device (replaces basics.py:73-108).
    A raster that already holds only 0..7 passes through; codes >= 0 that neither convention knows raise."""
    a = np.asarray(flowdir)
    if a.size == 0:
        return a.astype(int)
    if a.dtype.kind not in "iu" or a.dtype == np.uint64:
        a = a.astype(np.int64)                                # (non-integral floats are truncated first)
    elif a.dtype.kind == "u":
        a = a.astype({1: np.int16, 2: np.int32, 4: np.int64}[a.dtype.itemsize])
    t = to_dev(a)
    present, beyond = ops.dir_reclass_scan(t)
    lut = np.arange(-128, 128, dtype=np.int64)
    if not (beyond or any(v < 0 or v > 7 for v in present)):
        return ops.dir_reclass_apply(t, lut).cpu().numpy().astype(int, copy=False)     # already internal
    unknown = [v for v in present if v > 8]
    if beyond:                                                # error path only: name the far-away codes too
        unknown += [int(v) for v in np.unique(a[a > 127])]
    if unknown:
        raise ValueError(f"unknown D8 codes: {sorted(unknown)}")
    f = _EXT_TO_INTERNAL[convention]
    for e in range(1, 9):
        lut[e + 128] = f(e)
    return ops.dir_reclass_apply(t, lut).cpu().numpy().astype(int, copy=False)


def dir_reclass_opentopo(flowdir: np.ndarray) -> np.ndarray:
    return _reclass(flowdir, "opentopo")


def dir_reclass_rwatershed(flowdir: np.ndarray) -> np.ndarray:
    return _reclass(flowdir, "rwatershed")
