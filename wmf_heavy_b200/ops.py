"""Device-level operators: CUDA tensors in, CUDA tensors out, every one a thin call into the
C ABI of ``libwmfb200.so``.  The two reference-facing shims (``cu_py``/``models_py`` with the
wmf_py signatures, ``cu``/``models`` with the f2py signatures) are built on these; no numerics
happen in Python."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._dev import ptr, require_cuda, stream_ptr, workspace
from ._lib import (ACC_F64, ACC_I32, ACC_I64, ALGO_AUTO, ALGO_SWEEP, ALGO_WALK, ENC_INTERNAL, ENC_KEYPAD, check)

_ACC_TORCH = {ACC_I32: torch.int32, ACC_I64: torch.int64, ACC_F64: torch.float64}


def _chk_dev(t: torch.Tensor, dtype, name: str):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous CUDA tensor of dtype {dtype}")


# --------------------------------------------------------------------------------------- D1
def d8_accum(dir_t: torch.Tensor, mask_t: Optional[torch.Tensor] = None, acc_dtype: int = ACC_I32,
             encoding: int = ENC_INTERNAL, algo: int = ALGO_AUTO, out: Optional[torch.Tensor] = None,
             stats: Optional[dict] = None) -> torch.Tensor:
    """Full-raster D8 accumulation (replaces wmf_py/cu_py/basics.py:124-177).  ``stats`` (a dict) receives
    ``tile_classes``: how many tiles each solver took (diagnostic, synchronises)."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    ny, nx = dir_t.shape
    if mask_t is not None:
        _chk_dev(mask_t, torch.uint8, "mask")
    if out is None:
        out = torch.empty((ny, nx), dtype=_ACC_TORCH[acc_dtype], device=dir_t.device)
    else:
        _chk_dev(out, _ACC_TORCH[acc_dtype], "out")
    lib = _lib.load()
    nb = lib.wmf_d8_accum_workspace(ny, nx, acc_dtype, algo)
    ws = workspace(nb)
    check(lib.wmf_d8_accum(ptr(dir_t), ptr(mask_t), ptr(out), acc_dtype, ny, nx, encoding, algo, ptr(ws), ws.numel(),
                           stream_ptr()))
    if stats is not None and algo != ALGO_WALK and ny * nx <= 0xFFFFFFFF:
        cnt = (C.c_int64 * 8)()
        check(lib.wmf_d8_accum_tile_classes(ptr(ws), ws.numel(), ny, nx, C.addressof(cnt), stream_ptr()))
        c = list(cnt)
        stats["tile_classes"] = {"one_pass": c[0], "jump": c[4], "chain_walk": c[1] + c[3], "chain_walk_unresolved_upstream": c[2]}
        if c[4]:
            stats["jump_rounds"] = {"phase_a_mean": c[5] / c[4], "phase_c_mean": c[6] / c[4], "max": c[7]}
    return out


# --------------------------------------------------------------------------------------- D2
def basin_mask(dir_t: torch.Tensor, mask_t: Optional[torch.Tensor], outlet_rc: Tuple[int, int],
               encoding: int = ENC_INTERNAL) -> Tuple[torch.Tensor, int]:
    """Mask of the cells whose D8 path reaches the outlet (traversal of basics.py:206-248)."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    ny, nx = dir_t.shape
    if mask_t is not None:
        _chk_dev(mask_t, torch.uint8, "mask")
    out = torch.empty((ny, nx), dtype=torch.uint8, device=dir_t.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_basin_mask_workspace(ny, nx))
    n = C.c_int64(0)
    check(lib.wmf_basin_mask(ptr(dir_t), ptr(mask_t), ny, nx, encoding, int(outlet_rc[0]), int(outlet_rc[1]), ptr(out),
                             C.addressof(n), ptr(ws), ws.numel(), stream_ptr()))
    return out, int(n.value)


# --------------------------------------------------------------------------------------- D3
def coord2fil_col(x, y, xll, yll, dx, ncols, nrows) -> Tuple[int, int]:
    """cuencas_heavy.f90:355-365 in fp32 (host)."""
    col, fil = C.c_int32(0), C.c_int32(0)
    _lib.load().wmf_coord2fil_col_host(float(x), float(y), float(xll), float(yll), float(dx), int(ncols), int(nrows),
                                       C.addressof(col), C.addressof(fil))
    return col.value, fil.value


def basin_structure(dir_keypad_t: torch.Tensor, col: int, fil: int) -> torch.Tensor:
    """BFS-ordered basin vector (cu.basin_find + cu.basin_cut, cuencas_heavy.f90:507-576).
    ``dir_keypad_t`` is the GIS-oriented raster int8 [nrows][ncols].  Returns int32 [3][n]."""
    require_cuda()
    _chk_dev(dir_keypad_t, torch.int8, "dir")
    nrows, ncols = dir_keypad_t.shape
    lib = _lib.load()
    ws = workspace(lib.wmf_basin_structure_workspace(nrows, ncols))
    cap = nrows * ncols
    st = torch.empty((3, cap), dtype=torch.int32, device=dir_keypad_t.device)
    n = C.c_int64(0)
    check(lib.wmf_basin_structure(ptr(dir_keypad_t), nrows, ncols, int(col), int(fil), ptr(st), cap, C.addressof(n),
                                  ptr(ws), ws.numel(), stream_ptr()))
    return st[:, : n.value].contiguous()


# --------------------------------------------------------------------------------------- D4-D6
def _st_args(st: torch.Tensor):
    _chk_dev(st, torch.int32, "structure")
    if st.dim() != 2 or st.shape[0] != 3:
        raise ValueError("structure must have shape (3, n)")
    return st.shape[1], st.shape[1]


def basin_acum_vec(st: torch.Tensor) -> torch.Tensor:
    require_cuda()
    n, stride = _st_args(st)
    out = torch.empty(n, dtype=torch.int32, device=st.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_basin_vec_workspace(n))
    check(lib.wmf_basin_acum_vec(ptr(st), stride, n, ptr(out), ptr(ws), ws.numel(), stream_ptr()))
    return out


def basin_basics(st: torch.Tensor, dem_t: torch.Tensor, dirv_t: torch.Tensor, dxp: float, xll: float, yll: float,
                 dx: float, nrows: int):
    """cu.basin_basics (cuencas_heavy.f90:581-627) -> acum, long_cel, pend, elev, scalars[6]."""
    require_cuda()
    n, stride = _st_args(st)
    _chk_dev(dem_t, torch.float32, "DEM")
    _chk_dev(dirv_t, torch.int32, "DIR")
    dev = st.device
    acum = torch.empty(n, dtype=torch.int32, device=dev)
    lon, pend, elev = (torch.empty(n, dtype=torch.float32, device=dev) for _ in range(3))
    geo = (C.c_float * 5)(dxp, xll, yll, dx, float(nrows))
    sc = (C.c_float * 6)()
    lib = _lib.load()
    ws = workspace(lib.wmf_basin_vec_workspace(n))
    check(lib.wmf_basin_basics(ptr(st), stride, n, ptr(dem_t), ptr(dirv_t), C.addressof(geo), ptr(acum), ptr(lon),
                               ptr(pend), ptr(elev), C.addressof(sc), ptr(ws), ws.numel(), stream_ptr()))
    return acum, lon, pend, elev, np.array(list(sc), dtype=np.float32)


def basin_coordxy(st: torch.Tensor, xll: float, yll: float, dx: float, nrows: int):
    require_cuda()
    n, stride = _st_args(st)
    X = torch.empty(n, dtype=torch.float32, device=st.device)
    Y = torch.empty(n, dtype=torch.float32, device=st.device)
    check(_lib.load().wmf_basin_coordxy(ptr(st), stride, n, xll, yll, dx, int(nrows), ptr(X), ptr(Y), stream_ptr()))
    return X, Y


def basin_map2basin(st: torch.Tensor, raster_t: torch.Tensor) -> torch.Tensor:
    """Gather a GIS-oriented raster [nrows][ncols] into the basin vector by (col, fil)."""
    require_cuda()
    n, stride = _st_args(st)
    if not (raster_t.is_cuda and raster_t.is_contiguous() and raster_t.element_size() in (1, 4, 8)):
        raise TypeError("raster must be a contiguous CUDA tensor with 1, 4 or 8-byte elements")
    nrows, ncols = raster_t.shape
    out = torch.zeros(n, dtype=raster_t.dtype, device=st.device)
    check(_lib.load().wmf_basin_map2basin(ptr(st), stride, n, ptr(raster_t), nrows, ncols, raster_t.element_size(),
                                          ptr(out), stream_ptr()))
    return out


def basin_2map(st: torch.Tensor, vec_t: torch.Tensor, nrows: int, ncols: int, nodata) -> torch.Tensor:
    require_cuda()
    n, stride = _st_args(st)
    if not (vec_t.is_cuda and vec_t.is_contiguous() and vec_t.element_size() in (1, 4, 8) and vec_t.numel() == n):
        raise TypeError("vec must be a contiguous CUDA tensor of n elements of 1, 4 or 8 bytes")
    out = torch.full((nrows, ncols), nodata, dtype=vec_t.dtype, device=st.device)
    check(_lib.load().wmf_basin_2map(ptr(st), stride, n, ptr(vec_t), nrows, ncols, vec_t.element_size(), ptr(out),
                                     stream_ptr()))
    return out


def stream_nod(st: torch.Tensor, acum_t: torch.Tensor, umbral: int):
    """cu.basin_stream_nod (cuencas_heavy.f90:780-809) -> cauce, nodos, trazado, n_nodos, n_cauce."""
    require_cuda()
    n, stride = _st_args(st)
    _chk_dev(acum_t, torch.int32, "acum")
    cauce, nodos, traz = (torch.empty(n, dtype=torch.int32, device=st.device) for _ in range(3))
    cnt = (C.c_int32 * 2)()
    check(_lib.load().wmf_stream_nod(ptr(st), stride, n, ptr(acum_t), int(umbral), ptr(cauce), ptr(nodos), ptr(traz),
                                     C.addressof(cnt), stream_ptr()))
    return cauce, nodos, traz, int(cnt[0]), int(cnt[1])


# --------------------------------------------------------------------------------------- S1/S2
def _check_pos_evento(pos_evento: torch.Tensor, n_rec: int) -> None:
    """Every step must name a record of the rain file (1 = no rain).  The Fortran prints 'Se ha tratado de leer un
    valor fuera del rango' and carries on with an undefined RainInt (Modelos_heavy.f90:580-590); here it is an error."""
    lo, hi = (int(v) for v in torch.aminmax(pos_evento))
    if lo < 1 or hi > n_rec:
        raise ValueError(f"posEvento names record {lo if lo < 1 else hi}, the rain file holds records 1..{n_rec} "
                         "(Se ha tratado de leer un valor fuera del rango)")


def shia5_run(cell_static: torch.Tensor, rain_records: torch.Tensor, pos_evento: torch.Tensor, evp: torch.Tensor,
              calib: torch.Tensor, sto: torch.Tensor, hspeed: Optional[torch.Tensor] = None, *, dt: float,
              retorno_gr: float = 0.0, retorno_aq: float = 0.0, calc_niter: int = 5, speed_type=(1, 1, 1),
              show_storage: bool = False, show_mean_speed: bool = False, snap_slot: Optional[torch.Tensor] = None,
              n_snap: int = 0, validate: bool = True, strict: bool = False, show_mean_retorno: bool = False,
              save_vfluxes: bool = False, save_rc: bool = False, with_retorned: bool = False):
    """SHIA 5-tank update as shipped (Modelos_heavy.f90:169-548) for P parameter sets at once.
    ``sto`` [P,5,N] and ``hspeed`` [P,4,N] are advanced IN PLACE.  Returns a dict of device tensors.
    ``snap_slot`` int32 [N_reg] (-1 or a slot in [0, n_snap)): the storages after those steps are returned as
    ``snapshots`` [P, n_snap, 5, N] (save_storage / guarda_cond, Modelos_heavy.f90:496-500).
    ``strict``: the Fortran's own operation forms (true division, powf) instead of the fast ones; see the header.
    Optional outputs of the shipped routine: ``show_mean_retorno`` -> ``mean_retorno`` [P, N_reg]; ``save_vfluxes`` ->
    ``mean_vfluxes`` [P, 4, N_reg] and ``vfluxes`` [P, 4, N] (last step); ``save_rc`` -> ``rc_coef`` [P, 2, N];
    ``with_retorned`` -> ``retorned`` [P, N] (zeros when the mean is shown: the Fortran clears it every step then)."""
    require_cuda()
    _chk_dev(cell_static, torch.float32, "cell_static")
    _chk_dev(rain_records, torch.int32, "rain_records")
    _chk_dev(pos_evento, torch.int32, "pos_evento")
    _chk_dev(evp, torch.float32, "evp")
    _chk_dev(calib, torch.float32, "calib")
    _chk_dev(sto, torch.float32, "sto")
    P = calib.shape[0]
    N = cell_static.shape[1]
    n_reg = pos_evento.shape[0]
    n_rec = rain_records.shape[0]
    if cell_static.shape != (17, N) or calib.shape != (P, 11) or sto.shape != (P, 5, N) or rain_records.shape != (n_rec, N):
        raise ValueError("shape mismatch: cell_static (17,N), calib (P,11), sto (P,5,N), rain_records (n_rec,N)")
    if evp.shape[0] < n_reg:
        raise ValueError("evp must have at least N_reg entries")
    if validate:
        _check_pos_evento(pos_evento, n_rec)
    dev = sto.device
    given = hspeed is not None
    if given:
        _chk_dev(hspeed, torch.float32, "hspeed")
        if hspeed.shape != (P, 4, N):
            raise ValueError("hspeed must be (P,4,N)")
    else:
        hspeed = torch.empty((P, 4, N), dtype=torch.float32, device=dev)
    balance = torch.empty((P, n_reg), dtype=torch.float32, device=dev)
    mean_rain = torch.empty(n_reg, dtype=torch.float32, device=dev)
    acum_rain = torch.empty(N, dtype=torch.float32, device=dev)
    ms = torch.empty((P, 5, n_reg), dtype=torch.float32, device=dev) if show_storage else None
    mv = torch.empty((P, 4, n_reg), dtype=torch.float32, device=dev) if show_mean_speed else None
    prm = _lib.Shia5Params(dt, retorno_gr, retorno_aq, int(calc_niter), (C.c_int32 * 3)(*[int(s) for s in speed_type]),
                           int(given), int(bool(strict)))
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    mret = f32(P, n_reg) if show_mean_retorno else None
    mvf, vfl = (f32(P, 4, n_reg), f32(P, 4, N)) if save_vfluxes else (None, None)
    rcc = f32(P, 2, N) if save_rc else None
    rtd = f32(P, N) if (with_retorned or show_mean_retorno) else None
    diag = any(t is not None for t in (ms, mv, mret, mvf, rcc, rtd))
    lib = _lib.load()
    ws = workspace(lib.wmf_shia5_workspace(N, n_reg, P, int(diag)))
    snaps = None
    if snap_slot is not None:
        _chk_dev(snap_slot, torch.int32, "snap_slot")
        if snap_slot.shape != (n_reg,) or int(n_snap) <= 0:
            raise ValueError("snap_slot must be int32 [N_reg] and n_snap > 0")
        snaps = torch.zeros((P, int(n_snap), 5, N), dtype=torch.float32, device=dev)
    check(lib.wmf_shia5_run_ext(C.addressof(prm), N, n_reg, n_rec, P, ptr(cell_static), ptr(rain_records), ptr(pos_evento),
                                ptr(evp), ptr(calib), ptr(sto), ptr(hspeed), ptr(balance), ptr(mean_rain), ptr(acum_rain),
                                ptr(ms), ptr(mv), ptr(snap_slot), int(n_snap) if snaps is not None else 0, ptr(snaps),
                                ptr(mret), ptr(mvf), ptr(vfl), ptr(rcc), ptr(rtd), int(bool(show_mean_retorno)),
                                ptr(ws), ws.numel(), stream_ptr()))
    return {"StoOut": sto, "hspeed": hspeed, "balance": balance, "mean_rain": mean_rain, "acum_rain": acum_rain,
            "mean_storage": ms, "mean_speed": mv, "snapshots": snaps, "mean_retorno": mret, "mean_vfluxes": mvf,
            "vfluxes": vfl, "rc_coef": rcc, "retorned": rtd}


def path_sum(dir_t: torch.Tensor, active_t: Optional[torch.Tensor], w_t: Optional[torch.Tensor], dx: float, dy: float,
             cycle_nan: bool, encoding: int = ENC_INTERNAL) -> torch.Tensor:
    """Sum of per-cell weights (``w_t`` float64, or the step lengths when None) down every cell's flow path:
    metrics.py:420-432 / :583-597."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    if active_t is not None:
        _chk_dev(active_t, torch.uint8, "active")
    if w_t is not None:
        _chk_dev(w_t, torch.float64, "w")
    ny, nx = dir_t.shape
    out = torch.empty((ny, nx), dtype=torch.float64, device=dir_t.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_path_sum_workspace(ny, nx))
    check(lib.wmf_path_sum(ptr(dir_t), ptr(active_t), ptr(w_t), ny, nx, encoding, float(dx), float(dy), int(bool(cycle_nan)),
                           ptr(out), ptr(ws), ws.numel(), stream_ptr()))
    return out


def time_to_out(dir_t: torch.Tensor, slope_t: torch.Tensor, manning_t: torch.Tensor, dx: float, dy: float,
                encoding: int = ENC_INTERNAL) -> torch.Tensor:
    """Manning travel time to the end of every cell's flow path: metrics.py:528-599."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    _chk_dev(slope_t, torch.float64, "slope")
    _chk_dev(manning_t, torch.float64, "n_manning")
    ny, nx = dir_t.shape
    if slope_t.shape != (ny, nx) or manning_t.shape != (ny, nx):
        raise ValueError("slope and n_manning must have the raster's shape")
    out = torch.empty((ny, nx), dtype=torch.float64, device=dir_t.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_time_to_out_workspace(ny, nx))
    check(lib.wmf_time_to_out(ptr(dir_t), ptr(slope_t), ptr(manning_t), ny, nx, encoding, float(dx), float(dy), ptr(out),
                              ptr(ws), ws.numel(), stream_ptr()))
    return out


def stream_nodes(dir_t: torch.Tensor, stream_t: torch.Tensor, outlet_index: int, encoding: int = ENC_INTERNAL):
    """Network nodes and node-to-node links (streams.py:282-384) -> (node uint8, down_node, down_steps, up_node,
    up_steps) rasters; the node / link fields are linear indices (-1 = none)."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    _chk_dev(stream_t, torch.uint8, "stream")
    ny, nx = dir_t.shape
    dev = dir_t.device
    node = torch.empty((ny, nx), dtype=torch.uint8, device=dev)
    outs = [torch.empty((ny, nx), dtype=torch.int32, device=dev) for _ in range(4)]
    lib = _lib.load()
    ws = workspace(lib.wmf_stream_nodes_workspace(ny, nx))
    check(lib.wmf_stream_nodes(ptr(dir_t), ptr(stream_t), ny, nx, encoding, int(outlet_index), ptr(node), ptr(outs[0]),
                               ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ptr(ws), ws.numel(), stream_ptr()))
    return (node, *outs)


def route_topology(drain: np.ndarray):
    """Host preparation of the routing extension from the structure's drain ids (upstream-first order, receiver of
    cell i = n - drain[i], 0 = outlet): (rank int32 [n], max_rank, don_ptr int32 [n + 1], don_idx int32 [n - roots])."""
    dr = np.ascontiguousarray(drain, dtype=np.int32).reshape(-1)
    n = dr.size
    depth = np.empty(n, dtype=np.int32)
    dmax = int(_lib.load().wmf_basin_depth_host(dr.ctypes.data, n, depth.ctypes.data))
    if dmax < 0:
        raise ValueError("malformed structure: a receiver does not come after its donor")
    rank = (dmax - depth).astype(np.int32)
    has = dr != 0
    recv = (n - dr[has]).astype(np.int64)
    donors = np.nonzero(has)[0].astype(np.int32)
    order = np.argsort(recv, kind="stable")               # donors of a receiver stay in ascending index order
    don_idx = donors[order]
    don_ptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(don_ptr, recv + 1, 1)
    don_ptr = np.cumsum(don_ptr, dtype=np.int64).astype(np.int32)
    return rank, dmax, don_ptr, np.ascontiguousarray(don_idx, dtype=np.int32)


def shia5_route_run(cell_static: torch.Tensor, rain_records: torch.Tensor, pos_evento: torch.Tensor, evp: torch.Tensor,
                    calib: torch.Tensor, sto: torch.Tensor, drain, ctrl=None, n_ctrl: int = 1, *, dt: float,
                    retorno_gr: float = 0.0, retorno_aq: float = 0.0, persistent: bool = True):
    """Routing EXTENSION (N4; the reference has no behaviour to match): the shipped 5-tank update plus a channel
    store (tank 5) draining along ``drain``.  ``sto`` [5, N] advanced in place; ``ctrl`` int32 [N] = row of Q (> 0) or 0.
    Returns {"StoOut", "Q" [n_ctrl, N_reg] m3/s (row 0 = outlet), "acum_rain"}."""
    dev = require_cuda()
    _chk_dev(cell_static, torch.float32, "cell_static")
    _chk_dev(rain_records, torch.int32, "rain_records")
    _chk_dev(pos_evento, torch.int32, "pos_evento")
    _chk_dev(evp, torch.float32, "evp")
    _chk_dev(calib, torch.float32, "calib")
    _chk_dev(sto, torch.float32, "sto")
    N = cell_static.shape[1]
    n_reg, n_rec = pos_evento.shape[0], rain_records.shape[0]
    if cell_static.shape != (17, N) or calib.numel() != 11 or sto.shape != (5, N) or rain_records.shape != (n_rec, N):
        raise ValueError("shape mismatch: cell_static (17,N), calib (11,), sto (5,N), rain_records (n_rec,N)")
    _check_pos_evento(pos_evento, n_rec)
    rank, dmax, don_ptr, don_idx = route_topology(np.asarray(drain))
    if rank.size != N:
        raise ValueError("drain must have one entry per cell")
    c = np.zeros(N, dtype=np.int32) if ctrl is None else np.ascontiguousarray(ctrl, dtype=np.int32).reshape(-1)
    if c.size != N or c.min() < 0 or c.max() >= max(int(n_ctrl), 1):
        raise ValueError("ctrl must hold rows in [0, n_ctrl) for every cell")
    to = lambda a: torch.from_numpy(a).to(dev)
    rank_t, dp_t, di_t, c_t = to(rank), to(don_ptr), to(don_idx if don_idx.size else np.zeros(1, np.int32)), to(c)
    # the structure's BFS order lists the ranks in ascending order: one cooperative launch can then walk all diagonals
    ls_t = None
    if persistent and (np.diff(rank) >= 0).all():
        ls_t = to(np.searchsorted(rank, np.arange(dmax + 2)).astype(np.int32))
    Q = torch.empty((int(n_ctrl), n_reg), dtype=torch.float32, device=dev)
    acum = torch.empty(N, dtype=torch.float32, device=dev)
    prm = _lib.Shia5Params(dt, retorno_gr, retorno_aq, 0, (C.c_int32 * 3)(1, 1, 1), 0, 0)
    lib = _lib.load()
    ws = workspace(lib.wmf_shia5_route_workspace(N))
    check(lib.wmf_shia5_route_run(C.addressof(prm), N, n_reg, n_rec, ptr(cell_static), ptr(rain_records), ptr(pos_evento),
                                  ptr(evp), ptr(calib), ptr(rank_t), int(dmax), ptr(dp_t), ptr(di_t), ptr(c_t), int(n_ctrl),
                                  ptr(ls_t), ptr(sto), ptr(acum), ptr(Q), ptr(ws), ws.numel(), stream_ptr()))
    return {"StoOut": sto, "Q": Q, "acum_rain": acum}


def fitness_nash_peak(qsim: torch.Tensor, qobs: torch.Tensor) -> torch.Tensor:
    """Nash-Sutcliffe efficiency and peak-flow error of P hydrographs (qsim float32 [P, T]) against qobs [T] -> float64 [P, 2]."""
    require_cuda()
    _chk_dev(qsim, torch.float32, "qsim")
    _chk_dev(qobs, torch.float32, "qobs")
    P, T = qsim.shape
    if qobs.shape != (T,):
        raise ValueError("qobs must have one entry per step of qsim")
    out = torch.empty((P, 2), dtype=torch.float64, device=qsim.device)
    check(_lib.load().wmf_fitness_nash_peak(ptr(qsim), ptr(qobs), P, T, ptr(out), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------- S4
def shia3_run(rain: torch.Tensor, et: Optional[torch.Tensor], cap: torch.Tensor, grav: torch.Tensor,
              aqu: torch.Tensor, prm: "_lib.Shia3Params", nsteps: int, separate_fluxes: bool = False,
              save_storage: bool = False):
    """SHIA 3-tank balance (wmf_py/models_py/core.py:64-295); cap/grav/aqu advanced IN PLACE.
    Returns dict with per-step sums [nsteps+1, 6] and optional series."""
    require_cuda()
    for t, nm in ((rain, "rain"), (cap, "cap"), (grav, "grav"), (aqu, "aqu")):
        _chk_dev(t, torch.float64, nm)
    if et is not None:
        _chk_dev(et, torch.float64, "et")
    ncell = cap.numel()
    dev = cap.device
    qsl = torch.empty(ncell, dtype=torch.float64, device=dev)
    qbl = torch.empty(ncell, dtype=torch.float64, device=dev)
    sums = torch.empty((nsteps + 1, 6), dtype=torch.float64, device=dev)
    ser = {}
    if separate_fluxes:
        for k in ("q_super_t", "q_base_t", "q_total_t"):
            ser[k] = torch.empty((nsteps, ncell), dtype=torch.float64, device=dev)
    if save_storage:
        for k in ("storage_capilar_t", "storage_gravita_t", "storage_aquifer_t"):
            ser[k] = torch.empty((nsteps, ncell), dtype=torch.float64, device=dev)
    lib = _lib.load()
    ws = workspace(lib.wmf_shia3_workspace(ncell, nsteps))
    check(lib.wmf_shia3_run(C.addressof(prm), nsteps, ncell, ptr(rain), ptr(et), ptr(cap), ptr(grav), ptr(aqu),
                            ptr(qsl), ptr(qbl), ptr(sums), ptr(ser.get("q_super_t")), ptr(ser.get("q_base_t")),
                            ptr(ser.get("q_total_t")), ptr(ser.get("storage_capilar_t")),
                            ptr(ser.get("storage_gravita_t")), ptr(ser.get("storage_aquifer_t")), ptr(ws), ws.numel(),
                            stream_ptr()))
    out = {"q_super_last": qsl, "q_base_last": qbl, "sums": sums}
    out.update(ser)
    return out


# --------------------------------------------------------------------------------------- T2 / K1, N3
def _int_elem(t: torch.Tensor, name: str) -> int:
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous() and
            t.dtype in (torch.int8, torch.int16, torch.int32, torch.int64)):
        raise TypeError(f"{name} must be a contiguous CUDA tensor of a signed integer dtype")
    return t.element_size()


def dir_reclass_scan(codes: torch.Tensor):
    """Which values does the raster hold?  -> (sorted list of the values in -128..127, flag: values beyond that)."""
    require_cuda()
    eb = _int_elem(codes, "codes")
    res = torch.empty(9, dtype=torch.int32, device=codes.device)         # presence uint32[8] + the flag
    check(_lib.load().wmf_dir_reclass_scan(ptr(codes), eb, codes.numel(), ptr(res), ptr(res) + 32, stream_ptr()))
    r = res.cpu().numpy()
    bits = np.unpackbits(r[:8].view(np.uint8), bitorder="little")
    return [int(v) - 128 for v in np.flatnonzero(bits)], bool(r[8])


def dir_reclass_apply(codes: torch.Tensor, lut, keep_oor: bool = True, oor_value: int = 0) -> torch.Tensor:
    """out[i] = lut[codes[i] + 128] (lut: 256 integers for the values -128..127) -> int64, the raster's shape."""
    require_cuda()
    eb = _int_elem(codes, "codes")
    lut_t = torch.from_numpy(np.ascontiguousarray(lut, dtype=np.int64).reshape(256)).to(codes.device)
    out = torch.empty(codes.shape, dtype=torch.int64, device=codes.device)
    check(_lib.load().wmf_dir_reclass_apply(ptr(codes), eb, codes.numel(), ptr(lut_t), int(bool(keep_oor)), int(oor_value),
                                            ptr(out), stream_ptr()))
    return out


def index_out_of_range(idx: torch.Tensor, ncell: int) -> bool:
    require_cuda()
    _chk_dev(idx, torch.int64, "idx")
    flag = torch.empty(1, dtype=torch.int32, device=idx.device)
    check(_lib.load().wmf_index_scan(ptr(idx), idx.numel(), int(ncell), ptr(flag), stream_ptr()))
    return bool(flag.item())


def index_repair(idx: torch.Tensor, nx: int) -> None:
    """In place: idx = (idx // 100) * nx + (idx % 100) % nx (basics.py:273-276)."""
    require_cuda()
    _chk_dev(idx, torch.int64, "idx")
    check(_lib.load().wmf_index_repair(ptr(idx), idx.numel(), int(nx), stream_ptr()))


def scatter_last(vals: torch.Tensor, idx: torch.Tensor, ncell: int, nodata) -> torch.Tensor:
    """Raster (flat, ncell) filled with nodata, then out[idx[i]] = vals[i]; the last duplicate wins."""
    require_cuda()
    _chk_dev(idx, torch.int64, "idx")
    if not (vals.is_cuda and vals.is_contiguous() and vals.numel() == idx.numel() and vals.element_size() in (1, 2, 4, 8)):
        raise TypeError("vals must be a contiguous CUDA tensor with as many elements as idx")
    out = torch.full((int(ncell),), nodata, dtype=vals.dtype, device=vals.device)
    owner = torch.empty(int(ncell), dtype=torch.int32, device=vals.device)
    check(_lib.load().wmf_scatter_last(ptr(vals), vals.element_size(), ptr(idx), idx.numel(), ptr(out), ptr(owner),
                                       int(ncell), stream_ptr()))
    return out


def gather(flat: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    require_cuda()
    _chk_dev(idx, torch.int64, "idx")
    if not (flat.is_cuda and flat.is_contiguous() and flat.element_size() in (1, 2, 4, 8)):
        raise TypeError("flat must be a contiguous CUDA tensor")
    out = torch.empty(idx.numel(), dtype=flat.dtype, device=flat.device)
    check(_lib.load().wmf_gather(ptr(flat), flat.element_size(), ptr(idx), idx.numel(), ptr(out), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------- synthetic inputs
def synth_scheidegger(ny: int, nx: int, seed: int = 20260101, encoding: int = ENC_INTERNAL, row0: int = 0,
                      col0: int = 0, nx_total: Optional[int] = None) -> torch.Tensor:
    dev = require_cuda()
    out = torch.empty((ny, nx), dtype=torch.int8, device=dev)
    check(_lib.load().wmf_synth_scheidegger(ptr(out), ny, nx, row0, col0, nx if nx_total is None else nx_total,
                                            seed, encoding, stream_ptr()))
    return out


def synth_scheidegger_host(ny: int, nx: int, seed: int = 20260101, encoding: int = ENC_INTERNAL, row0: int = 0,
                           col0: int = 0, nx_total: Optional[int] = None) -> np.ndarray:
    """Same generator on the host (no GPU needed): used by oracles and CPU baselines."""
    out = np.empty((ny, nx), dtype=np.int8)
    _lib.load().wmf_synth_scheidegger_host(out.ctypes.data, ny, nx, row0, col0, nx if nx_total is None else nx_total,
                                           seed, encoding)
    return out


# --------------------------------------------------------------------------------------- forest / stripes
def forest_accum(parent: torch.Tensor, weight: Optional[torch.Tensor] = None, blocked: Optional[torch.Tensor] = None):
    """acc[i] = weight[i] + sum of acc over children (parent int32, -1 root) -> (acc int64, resolved uint8)."""
    require_cuda()
    _chk_dev(parent, torch.int32, "parent")
    n = parent.numel()
    if weight is not None:
        _chk_dev(weight, torch.int64, "weight")
    if blocked is not None:
        _chk_dev(blocked, torch.uint8, "blocked")
    acc = torch.empty(n, dtype=torch.int64, device=parent.device)
    res = torch.empty(n, dtype=torch.uint8, device=parent.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_forest_accum_workspace(n))
    check(lib.wmf_forest_accum(ptr(parent), ptr(weight), ptr(blocked), n, ptr(acc), ptr(res), ptr(ws), ws.numel(),
                               stream_ptr()))
    return acc, res


def d8_stripe_local(dir_t: torch.Tensor, row0: int, ny_total: int, mask_t: Optional[torch.Tensor] = None,
                    encoding: int = ENC_INTERNAL):
    """Phase 1 of the striped accumulation -> (packed boundary records int64 [2, 2, nx], private workspace)."""
    dev = require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    ny, nx = dir_t.shape
    if mask_t is not None:
        _chk_dev(mask_t, torch.uint8, "mask")
    lib = _lib.load()
    ws = torch.empty(lib.wmf_d8_stripe_workspace(ny, nx), dtype=torch.uint8, device=dev)   # lives until finish
    rec = torch.empty((2, 2, nx), dtype=torch.int64, device=dev)
    check(lib.wmf_d8_stripe_local(ptr(dir_t), ptr(mask_t), ny, nx, int(row0), int(ny_total), encoding, ptr(rec),
                                  ptr(ws), ws.numel(), stream_ptr()))
    return rec, ws


def stripe_boundary_inflow(recs_all: torch.Tensor, me: int):
    """Boundary forest of all stripes (recs_all int64 [G, 2, 2, nx]) -> (ext_inflow int64 [2, nx],
    ext_unres uint8 [2, nx]) of stripe `me`."""
    require_cuda()
    _chk_dev(recs_all, torch.int64, "recs_all")
    G, _, _, nx = recs_all.shape
    inflow = torch.empty((2, nx), dtype=torch.int64, device=recs_all.device)
    unres = torch.empty((2, nx), dtype=torch.uint8, device=recs_all.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_stripe_boundary_workspace(G, nx))
    check(lib.wmf_stripe_boundary_inflow(ptr(recs_all), G, nx, int(me), ptr(inflow), ptr(unres), ptr(ws), ws.numel(),
                                         stream_ptr()))
    return inflow, unres


def d8_stripe_finish(dir_t: torch.Tensor, ws: torch.Tensor, row0: int, ny_total: int, ext_inflow: torch.Tensor,
                     ext_unres: Optional[torch.Tensor] = None, mask_t: Optional[torch.Tensor] = None,
                     encoding: int = ENC_INTERNAL, acc_dtype: int = ACC_I64, out: Optional[torch.Tensor] = None):
    """Phase 2 of the striped accumulation: inject the inflow from the other stripes, write the final ACC."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    _chk_dev(ext_inflow, torch.int64, "ext_inflow")
    ny, nx = dir_t.shape
    if out is None:
        out = torch.empty((ny, nx), dtype=_ACC_TORCH[acc_dtype], device=dir_t.device)
    check(_lib.load().wmf_d8_stripe_finish(ptr(dir_t), ptr(mask_t), ptr(out), acc_dtype, ny, nx, int(row0), int(ny_total),
                                           encoding, ptr(ext_inflow), ptr(ext_unres), ptr(ws), ws.numel(), stream_ptr()))
    return out


# --------------------------------------------------------------------------------------- N2: stream receiver / HAND
def stream_receiver(dir_t: torch.Tensor, mask_t: Optional[torch.Tensor], stream_t: torch.Tensor, dx: float, dy: float,
                    encoding: int = ENC_INTERNAL, with_distance: bool = True):
    """First stream cell downstream of every cell (+ path length): metrics.py:189-268."""
    require_cuda()
    _chk_dev(dir_t, torch.int8, "dir")
    _chk_dev(stream_t, torch.uint8, "stream")
    if mask_t is not None:
        _chk_dev(mask_t, torch.uint8, "mask")
    ny, nx = dir_t.shape
    aq = torch.empty((ny, nx), dtype=torch.int32, device=dir_t.device)
    hd = torch.empty((ny, nx), dtype=torch.float64, device=dir_t.device) if with_distance else None
    lib = _lib.load()
    ws = workspace(lib.wmf_stream_receiver_workspace(ny, nx))
    check(lib.wmf_stream_receiver(ptr(dir_t), ptr(mask_t), ptr(stream_t), ny, nx, encoding, float(dx), float(dy), ptr(aq),
                                  ptr(hd), ptr(ws), ws.numel(), stream_ptr()))
    return aq, hd


def hand(dem_t: torch.Tensor, dir_t: torch.Tensor, mask_t: Optional[torch.Tensor], stream_t: torch.Tensor,
         encoding: int = ENC_INTERNAL) -> torch.Tensor:
    """Height above the nearest drainage along the flow path: metrics.py:271-323."""
    require_cuda()
    _chk_dev(dem_t, torch.float64, "dem")
    _chk_dev(dir_t, torch.int8, "dir")
    _chk_dev(stream_t, torch.uint8, "stream")
    if mask_t is not None:
        _chk_dev(mask_t, torch.uint8, "mask")
    ny, nx = dir_t.shape
    out = torch.empty((ny, nx), dtype=torch.float64, device=dir_t.device)
    lib = _lib.load()
    ws = workspace(lib.wmf_hand_workspace(ny, nx))
    check(lib.wmf_hand(ptr(dem_t), ptr(dir_t), ptr(mask_t), ptr(stream_t), ny, nx, encoding, ptr(out), ptr(ws), ws.numel(),
                       stream_ptr()))
    return out
