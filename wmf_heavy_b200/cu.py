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


def _dir_to_dev(DIR) -> torch.Tensor:
    """DIR(col, fil) keypad codes -> GIS-oriented int8 device raster [nrows][ncols].  Uploaded on every call: the
    caller owns DIR and may edit it in place between calls (pass a CUDA int8 tensor to keep it resident)."""
    if isinstance(DIR, torch.Tensor) and DIR.is_cuda and DIR.dtype == torch.int8:
        return DIR.t().contiguous()
    a = np.asarray(DIR)
    g = np.ascontiguousarray(a.T)                       # (nrows, ncols), row = fil
    g = np.where((g >= 1) & (g <= 9), g, 0).astype(np.int8)   # nodata / <=0 -> no receiver
    return to_dev(g)


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


# external code -> keypad code of the two conventions (cuencas_heavy.f90:1086-1119); everything else becomes cu.nodata
_EXT_TO_KEYPAD = {"opentopo": {3: 8, 2: 9, 1: 6, 8: 3, 7: 2, 6: 1, 5: 4, 4: 7},
                  "rwatershed": {3: 7, 2: 8, 1: 9, 8: 6, 7: 3, 6: 2, 5: 1, 4: 4}}


def _reclass_keypad(Mapa, convention: str) -> np.ndarray:
    """Mat_out(nc, nr) int32, Fortran-ordered: a 256-entry lookup table applied on the device."""
    a = np.asarray(Mapa)
    if a.dtype.kind not in "iu" or a.dtype == np.uint64:
        a = a.astype(np.int64)
    elif a.dtype.kind == "u":
        a = a.astype({1: np.int16, 2: np.int32, 4: np.int64}[a.dtype.itemsize])
    if a.size == 0:
        return np.zeros(a.shape, dtype=np.int32, order="F")
    nd = int(nodata)
    lut = np.full(256, nd, dtype=np.int64)
    for ext, key in _EXT_TO_KEYPAD[convention].items():
        lut[ext + 128] = key
    out = ops.dir_reclass_apply(to_dev(a), lut, keep_oor=False, oor_value=nd)
    return np.asfortranarray(out.to(torch.int32).cpu().numpy())


def dir_reclass_opentopo(Mapa, nc=None, nr=None):
    """cu.dir_reclass_opentopo(Mapa.T, cu.ncols, cu.nrows) -> keypad codes, cu.nodata elsewhere (wmfv2.py:250; f90:1105-1119)."""
    return _reclass_keypad(Mapa, "opentopo")


def dir_reclass_rwatershed(Mapa, nc=None, nr=None):
    """cu.dir_reclass_rwatershed(Mapa.T, cu.ncols, cu.nrows) -> keypad codes (wmfv2.py:246; f90:1089-1103)."""
    return _reclass_keypad(Mapa, "rwatershed")
