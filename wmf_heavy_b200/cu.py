"""f2py-style namespace ``cu``: the call shapes ``wmf_heavy/wmfv2.py`` uses against the Fortran
module ``cu`` (cuencas_heavy.f90), served by ``libwmfb200.so``.

Conventions mirrored from f2py: rasters are indexed ``(col, fil)`` -- i.e. callers
pass ``Map.T`` of the GIS raster (wmfv2.py:246,253) --, ``integer`` is int32, ``real`` is float32,
intent(out) arrays come back as a tuple, dimension arguments are trailing and optional, and map
geometry lives in MODULE-LEVEL STATE (``cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp,
cu.nodata``; set by wmfv2.read_map_raster :227-241).  Like the Fortran module this namespace is
not re-entrant: ``basin_find`` leaves its result in a module buffer that ``basin_cut`` slices.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._dev import require_cuda, to_dev
from ._lib import ENC_KEYPAD, WmfError
from .cu_py.basics import (basin_2map as _py_basin_2map, basin_map2basin as _py_basin_map2basin,  # noqa: F401
                           dir_reclass_opentopo as _py_reclass_opentopo, dir_reclass_rwatershed as _py_reclass_rw)

# ---- module state (f2py lower-cases Fortran names) : cuencas_heavy.f90:30-60
ncols = 0
nrows = 0
xll = 0.0
yll = 0.0
dx = 30.0
dy = 30.0
dxp = 30.0
nodata = -9999.0
area = 0.0
perimetro = 0.0
pend_media = 0.0
elevacion = 0.0
centrox = 0.0
centroy = 0.0

_basin_temp = None        # device structure of the last basin_find (Fortran: basin_temp)
_dir_dev_cache = None     # (id(DIR), device int8 raster) so Basin + SimuBasin do not re-upload


def _dir_to_dev(DIR) -> torch.Tensor:
    """DIR(col, fil) keypad codes -> GIS-oriented int8 device raster [nrows][ncols]."""
    global _dir_dev_cache
    if isinstance(DIR, torch.Tensor) and DIR.is_cuda and DIR.dtype == torch.int8:
        return DIR.t().contiguous()
    key = id(DIR)
    if _dir_dev_cache is not None and _dir_dev_cache[0] == key:
        return _dir_dev_cache[1]
    a = np.asarray(DIR)
    g = np.ascontiguousarray(a.T)                       # (nrows, ncols), row = fil
    g = np.where((g >= 1) & (g <= 9), g, 0).astype(np.int8)   # nodata / <=0 -> no receiver
    t = to_dev(g)
    _dir_dev_cache = (key, t)
    return t


def coord2fil_col(x, y):
    """cuencas_heavy.f90:355-365 with the module geometry (fp32)."""
    return ops.coord2fil_col(x, y, xll, yll, dx, ncols, nrows)


def basin_find(x, y, DIR, nc=None, nr=None) -> int:
    """cu.basin_find(lat, lon, DIR, cu.ncols, cu.nrows) -> ncells (wmfv2.py:823; f90:507-560).
    Errors are printed, not raised, like the Fortran; on failure the count is 1 (the outlet row only
    exists in the buffer when the coordinates were valid)."""
    global _basin_temp, ncols, nrows
    require_cuda()
    d = _dir_to_dev(DIR)
    if nc is not None and nr is not None and (int(nr), int(nc)) != tuple(d.shape):
        raise ValueError("DIR shape does not match (ncols, nrows)")
    if not ncols or not nrows:
        nrows, ncols = int(d.shape[0]), int(d.shape[1])
    col, fil = coord2fil_col(x, y)
    try:
        _basin_temp = ops.basin_structure(d, col, fil)
    except WmfError as e:
        print(f"Error: basin_find: {e.msg}")
        _basin_temp = None
        return 1
    return int(_basin_temp.shape[1])


def basin_cut(nceldas: int) -> np.ndarray:
    """cu.basin_cut(ncells) -> structure int32 (3, ncells), Fortran-ordered (f90:565-576)."""
    if _basin_temp is None:
        print("Error: basin_temp no está alojada")
        return np.zeros((3, int(nceldas)), dtype=np.int32, order="F")
    st = _basin_temp
    n = st.shape[1]
    return np.asfortranarray(st[:, n - int(nceldas):].cpu().numpy())


def _st_dev(structure) -> torch.Tensor:
    if isinstance(structure, torch.Tensor):
        return structure.to(require_cuda()).to(torch.int32).contiguous()
    return to_dev(np.ascontiguousarray(structure, dtype=np.int32))


def basin_acum(structure, nceldas=None) -> np.ndarray:
    """cu.basin_acum(structure, ncells) -> int32 (ncells,) (wmfv2.py:1868; f90:593-606)."""
    return ops.basin_acum_vec(_st_dev(structure)).cpu().numpy()


def basin_basics(structure, DEMvec, DIRvec, nceldas=None):
    """cu.basin_basics(structure, DEMvec, DIRvec, ncells) -> acum, long_cel, pend, elev
    (wmfv2.py:976; f90:581-627).  Sets cu.area, cu.pend_media, cu.elevacion, cu.centrox/centroy."""
    global area, pend_media, elevacion, centrox, centroy
    st = _st_dev(structure)
    acum, lon, pend, elev, sc = ops.basin_basics(st, to_dev(np.asarray(DEMvec, dtype=np.float32)),
                                                 to_dev(np.asarray(DIRvec).astype(np.int32)), dxp, xll, yll, dx, nrows)
    area, pend_media, elevacion, centrox, centroy = (float(v) for v in sc[:5])
    return acum.cpu().numpy(), lon.cpu().numpy(), pend.cpu().numpy(), elev.cpu().numpy()


def basin_coordxy(structure, nceldas=None):
    """cu.basin_coordxy(structure, ncells) -> X, Y float32 (wmfv2.py:926; f90:639-646)."""
    X, Y = ops.basin_coordxy(_st_dev(structure), xll, yll, dx, nrows)
    return X.cpu().numpy(), Y.cpu().numpy()


def basin_stream_nod(structure, acum, umbral, nceldas=None):
    """cu.basin_stream_nod(structure, acum, umbral, ncells) -> cauce, nodos, trazado, n_nodos, n_cauce
    (f90:780-809)."""
    st = _st_dev(structure)
    c, nd, tr, n_nodos, n_cauce = ops.stream_nod(st, to_dev(np.asarray(acum).astype(np.int32)), int(umbral))
    return c.cpu().numpy(), nd.cpu().numpy(), tr.cpu().numpy(), n_nodos, n_cauce


def basin_map2basin(structure, Map, *_ignored) -> np.ndarray:
    """cu.basin_map2basin(structure, Map, ...) -> vector (wmfv2.py:1167-1169): gather Map(col, fil)
    at each basin cell.  The extra positional arguments of the (mangled) reference call are accepted
    and ignored; geometry comes from the module state."""
    a = np.asarray(Map)
    g = np.ascontiguousarray(a.T)
    if g.dtype.itemsize not in (1, 4, 8):
        g = g.astype(np.float32)
    return ops.basin_map2basin(_st_dev(structure), to_dev(g)).cpu().numpy()


def basin_2map_find(structure, nceldas=None):
    """Size of the raster basin_2map fills (wmfv2.py:1192): the full map."""
    return ncols, nrows


def basin_2map(structure, BasinVar, map_ncols=None, map_nrows=None, nceldas=None):
    """cu.basin_2map(structure, BasinVar, ncols, nrows, ncells) -> M(col, fil), xll, yll (wmfv2.py:1193)."""
    v = np.asarray(BasinVar)
    if v.dtype.itemsize not in (1, 4, 8):
        v = v.astype(np.float32)
    fill = nodata if np.issubdtype(v.dtype, np.floating) else int(nodata)
    M = ops.basin_2map(_st_dev(structure), to_dev(np.ascontiguousarray(v)), nrows, ncols, fill)
    return np.asfortranarray(M.cpu().numpy().T), xll, yll


def dir_reclass_opentopo(Mapa, nc=None, nr=None):
    """opentopo 1..8 -> keypad is done on the wmf_py convention by the reference shim (wmfv2.py:58-59)."""
    return _py_reclass_opentopo(np.asarray(Mapa))


def dir_reclass_rwatershed(Mapa, nc=None, nr=None):
    return _py_reclass_rw(np.asarray(Mapa))
