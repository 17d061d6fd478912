"""ctypes front-end of ``libwmforacle.so`` (see ``wmf_oracle.c``).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwmforacle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (``make -C oracle``)."""
    src = os.path.join(_HERE, "wmf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_f_basin_find.restype = C.c_int64
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


# ----------------------------------------------------------------- family (A): wmf_py
def basin_acum(flowdir: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """wmf_py/cu_py/basics.py:124-177 -> float64 (ny, nx)."""
    fd = _c(flowdir, np.int32)
    m = _c(np.asarray(mask) != 0, np.uint8)
    ny, nx = fd.shape
    acc = np.zeros((ny, nx), dtype=np.float64)
    rc = lib().orc_basin_acum(_p(fd), _p(m), _p(acc), C.c_int64(ny), C.c_int64(nx))
    assert rc == 0
    return acc


def synth_scheidegger(ny: int, nx: int, seed: int = 20260101, row0: int = 0, col0: int = 0, nx_total=None) -> np.ndarray:
    """The benchmark's Scheidegger DIR raster (internal codes), any window of it."""
    out = np.empty((ny, nx), dtype=np.int8)
    f = lib().orc_synth_scheidegger
    f.restype = None
    f(_p(out), C.c_int64(ny), C.c_int64(nx), C.c_int64(row0), C.c_int64(col0), C.c_int64(nx if nx_total is None else nx_total),
      C.c_uint64(seed))
    return out


def basin_cut(dem, outlet_rc, flowdir, mask):
    """wmf_py/cu_py/basics.py:206-248 -> dict(dem, flowdir, mask)."""
    flowdir = np.asarray(flowdir)
    fd = _c(flowdir, np.int32)
    mk = np.asarray(mask)
    m = _c(mk != 0, np.uint8)
    ny, nx = fd.shape
    mb = np.zeros((ny, nx), dtype=np.uint8)
    rc = lib().orc_basin_cut_mask(_p(fd), _p(m), C.c_int64(int(outlet_rc[0])), C.c_int64(int(outlet_rc[1])),
                                  _p(mb), C.c_int64(ny), C.c_int64(nx))
    if rc == 1:
        raise ValueError("outlet outside grid")
    if rc == 2:
        raise ValueError("outlet not in mask")
    assert rc == 0
    mb = mb.astype(bool)
    return {"dem": np.asarray(dem) * mb, "flowdir": flowdir * mb, "mask": mb}


class _Shia3Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "dt", "max_capilar", "max_gravita", "max_aquifer", "drena", "retorno_gr",
        "k_perc_cap_to_grav", "k_perc_grav_to_aqu", "k_baseflow", "max_infil_mm_h", "cell_area")] + [
        ("proportional", C.c_int32), ("pad", C.c_int32)]


def shia3(forcings, grid, params, cap, grav, aqu, nsteps, separate_fluxes=False, save_storage=False):
    """wmf_py/models_py/core.py:64-295 on flat state arrays.  ``params`` is any object with the
    ModelParams attributes (core.py:24-45).  Returns a dict of arrays; raises RuntimeError like
    core.py:256."""
    rain = _c(forcings["rain_mm_h"], np.float64)
    nt, ny, nx = rain.shape
    assert nt >= nsteps
    ncell = ny * nx
    et, et_scale = None, 0.0
    if "et_mm_h" in forcings:
        et, et_scale = _c(forcings["et_mm_h"], np.float64), params.dt / 3600.0
    elif "et_mm_d" in forcings:
        et, et_scale = _c(forcings["et_mm_d"], np.float64), params.dt / 86400.0
    area = float(grid["dx"]) * float(grid["dy"]) if ("dx" in grid and "dy" in grid) else -1.0
    p = _Shia3Params(params.dt, params.max_capilar, params.max_gravita, params.max_aquifer, params.drena,
                     params.retorno_gr, params.k_perc_cap_to_grav, params.k_perc_grav_to_aqu,
                     params.k_baseflow, params.max_infil_mm_h, area,
                     int(params.eta_priority == "proportional"), 0)
    cap = np.array(cap, dtype=np.float64).reshape(ny, nx).copy()
    grav = np.array(grav, dtype=np.float64).reshape(ny, nx).copy()
    aqu = np.array(aqu, dtype=np.float64).reshape(ny, nx).copy()
    out = {
        "q_super_last": np.zeros((ny, nx)), "q_base_last": np.zeros((ny, nx)),
        "q_outlet": np.zeros(nsteps), "mass_error_step": np.zeros(nsteps), "mass_error_cum": np.zeros(nsteps),
    }
    ser = {}
    if separate_fluxes:
        for k in ("q_super_t", "q_base_t", "q_total_t"):
            ser[k] = np.zeros((nsteps, ny, nx))
    if save_storage:
        for k in ("storage_capilar_t", "storage_gravita_t", "storage_aquifer_t"):
            ser[k] = np.zeros((nsteps, ny, nx))
    fail = C.c_int64(-1)
    rc = lib().orc_shia3(C.byref(p), _p(rain), _p(et), C.c_double(et_scale), C.c_int64(nsteps), C.c_int64(ncell),
                         _p(cap), _p(grav), _p(aqu), _p(out["q_super_last"]), _p(out["q_base_last"]),
                         _p(out["q_outlet"]), _p(out["mass_error_step"]), _p(out["mass_error_cum"]),
                         _p(ser.get("q_super_t")), _p(ser.get("q_base_t")), _p(ser.get("q_total_t")),
                         _p(ser.get("storage_capilar_t")), _p(ser.get("storage_gravita_t")),
                         _p(ser.get("storage_aquifer_t")), C.byref(fail))
    if rc == 3:
        raise RuntimeError("cumulative mass balance error >1%")
    assert rc == 0
    out.update(ser)
    out.update({"storage_capilar": cap, "storage_gravita": grav, "storage_aquifer": aqu})
    return out


# ----------------------------------------------------------------- family (B): Fortran cu / models
def f_coord2fil_col(x, y, xll, yll, dx, ncols, nrows):
    """cuencas_heavy.f90:355-365 (fp32) -> (col, fil), 1-based, -999 when out of range."""
    col, fil = C.c_int32(0), C.c_int32(0)
    lib().orc_f_coord2fil_col(C.c_float(x), C.c_float(y), C.c_float(xll), C.c_float(yll), C.c_float(dx),
                              C.c_int32(ncols), C.c_int32(nrows), C.byref(col), C.byref(fil))
    return col.value, fil.value


def f_basin_find(dir_keypad: np.ndarray, col: int, fil: int) -> np.ndarray:
    """cuencas_heavy.f90:507-576.  ``dir_keypad`` is the GIS raster (nrows, ncols) of keypad codes
    (== Fortran DIR(col, fil)).  Returns structure int32 (3, n)."""
    d = _c(dir_keypad, np.int32)
    nrows, ncols = d.shape
    st = np.zeros((3, nrows * ncols), dtype=np.int32)
    # structure is written with stride n, so first find n with a full-capacity buffer
    tmp = np.zeros(3 * nrows * ncols, dtype=np.int32)
    n = lib().orc_f_basin_find(_p(d), C.c_int32(ncols), C.c_int32(nrows), C.c_int32(col), C.c_int32(fil),
                               _p(tmp), C.c_int64(nrows * ncols))
    if n < 0:
        raise ValueError(f"oracle basin_find failed: {n}")
    del st
    return tmp[: 3 * n].reshape(3, n).copy()


def f_basin_coordxy(structure, xll, yll, dx, nrows):
    st = _c(structure, np.int32)
    n = st.shape[1]
    X, Y = np.zeros(n, np.float32), np.zeros(n, np.float32)
    lib().orc_f_basin_coordxy(_p(st), C.c_int64(n), C.c_float(xll), C.c_float(yll), C.c_float(dx),
                              C.c_int32(nrows), _p(X), _p(Y))
    return X, Y


def f_basin_basics(structure, demv, dirv, dxp, xll, yll, dx, nrows):
    """cuencas_heavy.f90:581-627 -> acum, long_cel, pend, elev, scalars[6]."""
    st = _c(structure, np.int32)
    n = st.shape[1]
    dem, dv = _c(demv, np.float32), _c(dirv, np.int32)
    acum = np.zeros(n, np.int32)
    lon, pend, elev = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    sc = np.zeros(6, np.float32)
    lib().orc_f_basin_basics(_p(st), _p(dem), _p(dv), C.c_int64(n), C.c_float(dxp), C.c_float(xll),
                             C.c_float(yll), C.c_float(dx), C.c_int32(nrows), _p(acum), _p(lon), _p(pend),
                             _p(elev), _p(sc))
    return acum, lon, pend, elev, sc


def f_basin_acum(structure):
    st = _c(structure, np.int32)
    n = st.shape[1]
    acum = np.zeros(n, np.int32)
    lib().orc_f_basin_acum(_p(st), C.c_int64(n), _p(acum))
    return acum


def f_stream_nod(structure, acum, umbral):
    """cuencas_heavy.f90:780-809 -> cauce, nodos, trazado, n_nodos, n_cauce."""
    st = _c(structure, np.int32)
    n = st.shape[1]
    ac = _c(acum, np.int32)
    cauce, nodos, traz = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    cnt = np.zeros(2, np.int32)
    lib().orc_f_stream_nod(_p(st), _p(ac), C.c_int64(n), C.c_int32(umbral), _p(cauce), _p(nodos), _p(traz), _p(cnt))
    return cauce, nodos, traz, int(cnt[0]), int(cnt[1])


class _Shia5Params(C.Structure):
    _fields_ = [("dt", C.c_float), ("retorno_gr", C.c_float), ("retorno_aq", C.c_float),
                ("calc_niter", C.c_int32), ("speed_type", C.c_int32 * 3), ("hspeed_given", C.c_int32)]


@dataclass
class Shia5Inputs:
    """Per-cell tables and forcing of the Fortran shia_v1 (Modelos_heavy.f90:169-548), SoA."""
    v_coef: np.ndarray      # (4, N) f32
    h_coef: np.ndarray      # (4, N)
    h_exp: np.ndarray       # (4, N)
    max_capilar: np.ndarray  # (N,)
    max_gravita: np.ndarray
    max_aquifer: np.ndarray
    hill_long: np.ndarray
    elem_area: np.ndarray
    rain_records: np.ndarray  # (n_rec, N) int32, record k at row k-1
    pos_evento: np.ndarray    # (N_reg,) int32, 1-based; 1 == no rain
    evp: np.ndarray           # (N_reg,) f32
    dt: float = 300.0
    retorno_gr: float = 0.0
    retorno_aq: float = 0.0
    calc_niter: int = 5
    speed_type: tuple = (1, 1, 1)


def f_shia_v1(inp: Shia5Inputs, calib, sto_in, hspeed_in=None, show_storage=True, show_mean_speed=True,
              show_mean_retorno=False):
    """Modelos_heavy.f90:169-548 as shipped.  ``sto_in`` (5, N) is the initial StoOut."""
    N = inp.max_capilar.shape[0]
    n_reg = inp.pos_evento.shape[0]
    sto = np.array(sto_in, dtype=np.float32).reshape(5, N).copy()
    given = hspeed_in is not None
    hs = np.array(hspeed_in, dtype=np.float32).reshape(4, N).copy() if given else np.zeros((4, N), np.float32)
    p = _Shia5Params(inp.dt, inp.retorno_gr, inp.retorno_aq, inp.calc_niter, (C.c_int32 * 3)(*inp.speed_type),
                     int(given))
    cal = _c(calib, np.float32)
    assert cal.shape == (11,)
    balance, mean_rain = np.zeros(n_reg, np.float32), np.zeros(n_reg, np.float32)
    acum_rain = np.zeros(N, np.float32)
    ms = np.zeros((5, n_reg), np.float32) if show_storage else None
    msp = np.zeros((4, n_reg), np.float32) if show_mean_speed else None
    mr = np.zeros(n_reg, np.float32) if show_mean_retorno else None
    b64 = np.zeros(n_reg, np.float64)
    f32 = np.float32
    rc = lib().orc_f_shia_v1(C.byref(p), C.c_int64(N), C.c_int64(n_reg), _p(cal),
                             _p(_c(inp.v_coef, f32)), _p(_c(inp.h_coef, f32)), _p(_c(inp.h_exp, f32)),
                             _p(_c(inp.max_capilar, f32)), _p(_c(inp.max_gravita, f32)),
                             _p(_c(inp.max_aquifer, f32)), _p(_c(inp.hill_long, f32)), _p(_c(inp.elem_area, f32)),
                             _p(_c(inp.rain_records, np.int32)), _p(_c(inp.pos_evento, np.int32)),
                             _p(_c(inp.evp, f32)), _p(sto), _p(hs), _p(balance), _p(mean_rain), _p(acum_rain),
                             _p(ms), _p(msp), _p(mr), _p(b64))
    assert rc == 0
    return {"StoOut": sto, "hspeed": hs, "balance": balance, "balance64": b64, "mean_rain": mean_rain, "acum_rain": acum_rain,
            "mean_storage": ms, "mean_speed": msp, "mean_retorno": mr}


def f_shia_route(inp: Shia5Inputs, calib, sto_in, drain, ctrl, n_ctrl):
    """Routing extension (N4; no reference behaviour): the shipped 5-tank update plus a linear-reservoir
    channel store (tank 5) draining along ``drain``.  ``ctrl`` int32 [N]: row of Q for that cell (> 0) or 0; row 0 is
    the outlet.  Returns StoOut (5, N), Q (n_ctrl, N_reg) [m3/s], acum_rain, totals (rain, losses, outflow) [m3]."""
    N = inp.max_capilar.shape[0]
    n_reg = inp.pos_evento.shape[0]
    assert tuple(inp.speed_type) == (1, 1, 1), "the routing extension is defined for speed type 1"
    sto = np.array(sto_in, dtype=np.float32).reshape(5, N).copy()
    p = _Shia5Params(inp.dt, inp.retorno_gr, inp.retorno_aq, inp.calc_niter, (C.c_int32 * 3)(*inp.speed_type), 0)
    cal = _c(calib, np.float32)
    acum = np.zeros(N, np.float32)
    Q = np.zeros((n_ctrl, n_reg), np.float32)
    totals = np.zeros(3, np.float64)
    f32 = np.float32
    rc = lib().orc_f_shia_route(C.byref(p), C.c_int64(N), C.c_int64(n_reg), _p(cal), _p(_c(inp.v_coef, f32)),
                                _p(_c(inp.h_coef, f32)), _p(_c(inp.max_capilar, f32)), _p(_c(inp.max_gravita, f32)),
                                _p(_c(inp.max_aquifer, f32)), _p(_c(inp.hill_long, f32)), _p(_c(inp.elem_area, f32)),
                                _p(_c(inp.rain_records, np.int32)), _p(_c(inp.pos_evento, np.int32)), _p(_c(inp.evp, f32)),
                                _p(_c(drain, np.int32)), _p(_c(ctrl, np.int32)), C.c_int32(n_ctrl), _p(sto), _p(acum), _p(Q),
                                _p(totals))
    assert rc == 0
    return {"StoOut": sto, "Q": Q, "acum_rain": acum, "totals": totals}
