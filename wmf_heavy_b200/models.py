"""f2py-style namespace ``models``: module-level state + ``shia_v1`` with the call shape
``wmf_heavy/wmfv2.py:2000-2020`` uses against the Fortran module ``models`` (Modelos_heavy.f90).

State attributes are the lower-cased Fortran module variables (Modelos_heavy.f90:20-120) that
``SimuBasin.__init__`` / ``run_shia`` assign (wmfv2.py:1872-1916, 1967-1995) and read back
(``models.acum_rain``, ``models.mean_rain`` :2024; ``models.control`` :2036).  Per-cell tables have
the Fortran shapes ``(k, N)``.

In scope: the shipped 5-tank update (S1/S2) with its diagnostics, and the rain record formats (S5).
Out of scope, rejected loudly: sediments, slides, floods, separate_rain.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import ops
from ._dev import require_cuda, to_dev

# ---- module state -----------------------------------------------------------------------------
drena = None
v_coef = h_coef = v_exp = h_exp = None
max_capilar = max_gravita = max_aquifer = None
hill_long = hill_slope = elem_area = None
storage = None
storage_constant = 0.001
evpserie = None
dt = 300.0
dxp = 30.0
calc_niter = 5
retorno_gr = 0
retorno_aq = 0
verbose = 0
rain_first_point = 1
speed_type = np.ones(3, dtype=np.int32)
control = None
control_h = None
sim_sediments = sim_slides = sim_floods = 0
save_storage = save_speed = save_retorno = save_vfluxes = save_rc = 0
show_storage = show_speed = show_mean_speed = show_mean_retorno = show_area = 0
separate_fluxes = separate_rain = 0
route_channels = 0       # EXTENSION (not in the reference, whose Q is 0 as shipped): 1 routes tank 5 along `drena`
strict_arithmetic = 0    # 1: the Fortran's own operation forms in the kernel (true division, powf): ~2x slower, bit-level parity
guarda_cond = None
# outputs
mean_rain = None
acum_rain = None
mean_storage = None
mean_speed = None
mean_retorno = None      # (N_reg,)   show_mean_retorno (Modelos_heavy.f90:525-527)
mean_vfluxes = None      # (4, N_reg) save_vfluxes (:502-507)
vfluxes = None           # (4, N) of the last step (:429-433)
rc_coef = None           # (2, N) save_rc (:435-438, 543-544)
retorned = None          # (1, N) return flow accumulated per cell (:424)
guarda_vfluxes = None
idevento = None
posevento = None



# ---- rain record formats (Modelos_heavy.f90:580-590, 631-663) -----------------------------------
def rain_read_ascii_table(ruta_hdr: str, n_intervals: int):
    """Header table -> (idEvento, posEvento).  3 lines whose 4th token is an integer (the last one is
    Ntotal), 3 skipped lines, ``rain_first_point`` skipped lines if > 1, then ``idEvento, posEvento, text``."""
    global idevento, posevento
    with open(ruta_hdr, "r") as fh:
        lines = fh.read().splitlines()
    ntotal = 0
    for i in range(3):
        toks = lines[i].replace(",", " ").split()
        ntotal = int(float(toks[3]))
    body = lines[6:]
    if not (n_intervals + rain_first_point <= ntotal + 1):
        raise ValueError("rain header: N_intervals + rain_first_point exceeds the number of records")
    if rain_first_point > 1:
        body = body[rain_first_point:]
    ids = np.zeros(n_intervals, dtype=np.int32)
    pos = np.zeros(n_intervals, dtype=np.int32)
    for i in range(n_intervals):
        toks = body[i].replace(",", " ").split()
        ids[i], pos[i] = int(toks[0]), int(toks[1])
    idevento, posevento = ids, pos
    return ids, pos


def read_int_basin(ruta_bin: str, record: int, n_cel: int) -> np.ndarray:
    """Record ``record`` (1-based) of a direct-access int32 file with RECL = 4*N_cel bytes."""
    with open(ruta_bin, "rb") as fh:
        fh.seek((record - 1) * 4 * n_cel)
        v = np.fromfile(fh, dtype="<i4", count=n_cel)
    if v.size != n_cel:
        print("Error: Se ha tratado de leer un valor fuera del rango")
        v = np.zeros(n_cel, dtype=np.int32)
    return v


def read_rain_records(ruta_bin: str, n_cel: int) -> np.ndarray:
    """All records of the rain file as int32 [n_rec, N_cel]."""
    size = os.path.getsize(ruta_bin)
    n_rec = size // (4 * n_cel)
    return np.fromfile(ruta_bin, dtype="<i4", count=n_rec * n_cel).reshape(n_rec, n_cel)


# ---- basin record files (Modelos_heavy.f90:552-624): Fortran direct-access, RECL = 4 * N_col * N_cel bytes,
# record r at byte (r - 1) * RECL, the (N_col, N_cel) array in column-major order (N_col values per cell).
def _read_record(ruta: str, record: int, n_cel: int, n_col: int, dtype: str):
    res = 0
    v = np.zeros((n_col, n_cel), dtype=dtype, order="F")
    with open(ruta, "rb") as fh:                     # status='old': a missing file is an error, as in Fortran
        if record < 1:
            res = 1
        else:
            fh.seek((record - 1) * 4 * n_cel * n_col)
            raw = np.fromfile(fh, dtype=dtype, count=n_cel * n_col)
            if raw.size != n_cel * n_col:
                res = 1
            else:
                v = np.asfortranarray(raw.reshape(n_cel, n_col).T)
    if res != 0:
        print("Error: Se ha tratado de leer un valor fuera del rango")
    return v, res


def read_float_basin(ruta: str, record: int, n_cel: int):
    """(vect float32 [N_cel], Res) -- Modelos_heavy.f90:552-562.  Res != 0: record out of range."""
    v, res = _read_record(ruta, int(record), int(n_cel), 1, "<f4")
    return v[0].copy(), res


def read_float_basin_ncol(ruta: str, record: int, n_cel: int, n_col: int):
    """(vect float32 (N_col, N_cel), Res) -- Modelos_heavy.f90:566-576."""
    return _read_record(ruta, int(record), int(n_cel), int(n_col), "<f4")


def _write_record(ruta: str, vect, record: int, n_cel: int, n_col: int, dtype: str) -> None:
    a = np.asarray(vect, dtype=dtype).reshape(n_col, n_cel)
    if record == 1:
        mode = "wb"                                  # status='replace'
    else:
        if not os.path.exists(ruta):                 # status='old'
            raise FileNotFoundError(f"{ruta}: record {record} can only go to an existing file (record 1 creates it)")
        mode = "r+b"
    with open(ruta, mode) as fh:
        fh.seek((record - 1) * 4 * n_cel * n_col)
        fh.write(np.ascontiguousarray(a.T).tobytes())


def write_float_basin(ruta: str, vect, record: int, n_cel: int, n_col: int) -> None:
    """Modelos_heavy.f90:594-608: record 1 replaces the file, record <= 0 writes nothing."""
    if int(record) > 0:
        _write_record(ruta, vect, int(record), int(n_cel), int(n_col), "<f4")


def write_int_basin(ruta: str, vect, record: int, n_cel: int, n_col: int) -> None:
    """Modelos_heavy.f90:612-624 (no record > 0 guard there: record 0 is an error)."""
    if int(record) < 1:
        raise ValueError("write_int_basin: record must be >= 1")
    _write_record(ruta, vect, int(record), int(n_cel), int(n_col), "<i4")


def write_rain(ruta_bin: str, ruta_hdr: str, records: np.ndarray, pos_evento: np.ndarray, n_hills: int = 0) -> None:
    """Write a rain binary + header in the format the readers above (and the Fortran) expect."""
    rec = np.ascontiguousarray(records, dtype="<i4")
    rec.tofile(ruta_bin)
    n_cel = rec.shape[1]
    with open(ruta_hdr, "w") as fh:
        fh.write(f"Numero de celdas: {n_cel}\n")
        fh.write(f"Numero de laderas: {n_hills}\n")
        fh.write(f"Numero de registros: {len(pos_evento)}\n")
        fh.write(f"Numero de campos no cero: {rec.shape[0] - 1}\n")
        fh.write("Tipo de interpolacion: synthetic\n")
        fh.write("IDfecha, Record, Lluvia, Fecha\n")
        for i, p in enumerate(pos_evento):
            mean = 0.0 if p == 1 else float(rec[p - 1].mean()) / 1000.0
            fh.write(f"{i + 1}, {int(p)}, {mean:.2f}, step{i + 1}\n")


# ---- the model ------------------------------------------------------------------------------------
def _table(a, k: int, n: int, name: str) -> np.ndarray:
    if a is None:
        raise ValueError(f"models.{name} is not set")
    a = np.asarray(a, dtype=np.float32)
    if a.shape != (k, n):
        raise ValueError(f"models.{name} must have shape ({k}, {n}), got {a.shape}")
    return a


def cell_static_table(n: int) -> np.ndarray:
    """Pack the per-cell module tables into the [17][N] SoA block of the C ABI."""
    cs = np.empty((17, n), dtype=np.float32)
    cs[0:4] = _table(v_coef, 4, n, "v_coef")
    cs[4:8] = _table(h_coef, 4, n, "h_coef")
    cs[8:12] = _table(h_exp, 4, n, "h_exp")
    cs[12] = _table(max_capilar, 1, n, "max_capilar")[0]
    cs[13] = _table(max_gravita, 1, n, "max_gravita")[0]
    cs[14] = _table(max_aquifer, 1, n, "max_aquifer")[0]
    cs[15] = _table(hill_long, 1, n, "hill_long")[0]
    cs[16] = _table(elem_area, 1, n, "elem_area")[0]
    return cs


def _reject_out_of_scope():
    for nm in ("sim_sediments", "sim_slides", "sim_floods", "separate_rain"):
        if int(globals()[nm]) == 1:
            raise NotImplementedError(f"models.{nm} = 1 is outside the B200 hot path (see DESIGN.md, out of scope)")


def _prepare_inputs(calibs, rain_records, pos_evento, n_cel: int, n_reg: int, sto_in, hspeed_in):
    """Device copies of everything one shia_v1 run reads: (cell_static, rain records, posEvento, evp, calib, sto, hspeed
    or None, speed_type).  Packed and uploaded on every call: the tables are module state the caller edits in place between
    runs (a calibration loop scales h_coef / max_capilar ...), and 68 B/cell is small next to the run itself."""
    cal = np.atleast_2d(np.asarray(calibs, dtype=np.float32))
    if cal.shape[1] != 11:
        raise ValueError("calib must have 11 entries (Modelos_heavy.f90:177)")
    P = cal.shape[0]
    cs = to_dev(cell_static_table(n_cel))
    rr = rain_records if isinstance(rain_records, torch.Tensor) else to_dev(np.ascontiguousarray(rain_records, dtype=np.int32))
    pe = pos_evento if isinstance(pos_evento, torch.Tensor) else to_dev(np.asarray(pos_evento, dtype=np.int32)[:n_reg])
    if evpserie is None:
        raise ValueError("models.evpserie must be set (the Fortran dereferences EvpSerie(t), Modelos:396)")
    ev = to_dev(np.asarray(evpserie, dtype=np.float32)[:n_reg])
    if ev.shape[0] < n_reg:
        raise ValueError("models.evpserie is shorter than N_reg")
    # StoOut = StoIn if StoIn(1,1) > 0 else Storage + storage_constant (Modelos:229-233)
    if sto_in is not None and float(np.asarray(sto_in).reshape(-1)[0]) > 0:
        s0 = np.asarray(sto_in, dtype=np.float32)
    else:
        base = np.zeros((5, n_cel), np.float32) if storage is None else np.asarray(storage, dtype=np.float32)
        s0 = base + np.float32(storage_constant)
    s0 = s0.reshape((-1, 5, n_cel))
    if s0.shape[0] != P:
        if s0.shape[0] != 1:
            raise ValueError('StoIn must be (5, N) or (P, 5, N)')
        s0 = np.broadcast_to(s0, (P, 5, n_cel))
    sto = to_dev(np.ascontiguousarray(s0))
    hs = None
    if hspeed_in is not None and float(np.asarray(hspeed_in).reshape(-1)[0]) > 0:
        h0 = np.asarray(hspeed_in, dtype=np.float32)
        h0 = np.broadcast_to(h0.reshape((-1, 4, n_cel)), (P, 4, n_cel))
        hs = to_dev(np.ascontiguousarray(h0))
    st = tuple(np.asarray(speed_type).astype(int).reshape(-1)[:3])
    return cs, rr, pe, ev, to_dev(cal), sto, hs, st


def write_step_records(calib, rain_records, pos_evento, n_cel: int, n_reg: int, sto_in=None, hspeed_in=None,
                       ruta_speed=None, ruta_retorno=None, ruta_vfluxes=None) -> None:
    """The per-step record files of the shipped routine (Modelos_heavy.f90:502-515): with ``save_speed`` record t of
    ``ruta_speed`` holds hspeed (4 columns) after step t, with ``save_retorno`` record t of ``ruta_retorno`` the return
    flow of step t (Retorned is cleared after every step then, :529-531), and with ``save_vfluxes`` record
    ``guarda_vfluxes(t)`` (> 0) of ``ruta_vfluxes`` the vertical fluxes of step t.  The time-fused kernel keeps none of this
    per step, so the records come from a SECOND pass that advances the same kernel one step at a time (state and speeds
    carried on the device); the run's ordinary outputs are those of the single fused run and do not depend on it."""
    want_speed = int(save_speed) == 1 and bool(ruta_speed)
    want_ret = int(save_retorno) == 1 and bool(ruta_retorno)
    gv = None
    if int(save_vfluxes) == 1 and ruta_vfluxes and guarda_vfluxes is not None:
        gv = np.asarray(guarda_vfluxes).astype(np.int64).reshape(-1)
        if gv.size < n_reg:
            raise ValueError("guarda_vfluxes is shorter than N_reg")
        if not np.any(gv[:n_reg] > 0):
            gv = None
    if not (want_speed or want_ret or gv is not None):
        return
    cs, rr, pe, ev, cal, sto, hs, st = _prepare_inputs(np.asarray(calib, dtype=np.float32).reshape(1, -1), rain_records,
                                                        pos_evento, n_cel, n_reg, sto_in, hspeed_in)
    ops._check_pos_evento(pe, rr.shape[0])
    for t in range(n_reg):
        want_vf = gv is not None and gv[t] > 0
        out = ops.shia5_run(cs, rr, pe[t:t + 1], ev[t:t + 1], cal, sto, hs, dt=float(dt), retorno_gr=float(retorno_gr),
                            retorno_aq=float(retorno_aq), calc_niter=int(calc_niter), speed_type=st, validate=False,
                            strict=bool(int(strict_arithmetic)), save_vfluxes=bool(want_vf), with_retorned=want_ret)
        sto, hs = out["StoOut"], out["hspeed"]
        if want_speed:
            write_float_basin(ruta_speed, hs[0].cpu().numpy(), t + 1, n_cel, 4)
        if want_ret:
            write_float_basin(ruta_retorno, out["retorned"][0].cpu().numpy().reshape(1, n_cel), t + 1, n_cel, 1)
        if want_vf:
            write_float_basin(ruta_vfluxes, out["vfluxes"][0].cpu().numpy(), int(gv[t]), n_cel, 4)


def shia_v1_population(calibs, rain_records, pos_evento, n_cel: int, n_reg: int, sto_in=None, hspeed_in=None,
                       device_out: bool = False, save_records=None):
    """P parameter sets x one basin in one launch.  ``calibs`` (P, 11); ``sto_in`` (5, N) or (P, 5, N).
    ``save_records`` int [N_reg] (guarda_cond): steps with a record number > 0 return their storages in
    ``out["snapshots"]`` [P, n_snap, 5, N] together with ``out["snapshot_records"]`` (record number per slot)."""
    global mean_rain, acum_rain, mean_storage, mean_speed, mean_retorno, mean_vfluxes, vfluxes, rc_coef, retorned
    _reject_out_of_scope()
    dev = require_cuda()
    cs_dev, rr, pe, ev, cal_dev, sto, hs, st = _prepare_inputs(calibs, rain_records, pos_evento, n_cel, n_reg, sto_in, hspeed_in)
    P = cal_dev.shape[0]
    snap_slot, snap_recs = None, None
    if save_records is not None:
        g = np.asarray(save_records).astype(np.int64).reshape(-1)
        if g.size < n_reg:
            raise ValueError("guarda_cond is shorter than N_reg (Modelos_heavy.f90:317-328)")
        steps = np.nonzero(g[:n_reg] > 0)[0]
        if steps.size:
            need = P * steps.size * 5 * n_cel * 4
            free = torch.cuda.mem_get_info(dev)[0]
            if need > 0.8 * free:
                raise MemoryError(f"{steps.size} storage snapshots need {need / 2**30:.1f} GiB of device memory; "
                                  "save fewer steps (guarda_cond / Dates2Save)")
            slot = np.full(n_reg, -1, dtype=np.int32)
            slot[steps] = np.arange(steps.size, dtype=np.int32)
            snap_slot, snap_recs = to_dev(slot), g[steps].copy()
    out = ops.shia5_run(cs_dev, rr, pe, ev, cal_dev, sto, hs, dt=float(dt), retorno_gr=float(retorno_gr),
                        retorno_aq=float(retorno_aq), calc_niter=int(calc_niter), speed_type=tuple(st),
                        show_storage=bool(show_storage), show_mean_speed=bool(show_mean_speed), snap_slot=snap_slot,
                        n_snap=0 if snap_recs is None else int(snap_recs.size), strict=bool(int(strict_arithmetic)),
                        show_mean_retorno=bool(int(show_mean_retorno)), save_vfluxes=bool(int(save_vfluxes)),
                        save_rc=bool(int(save_rc)), with_retorned=float(retorno_gr) > 0)
    out["snapshot_records"] = snap_recs
    mean_rain = out["mean_rain"].cpu().numpy().reshape(1, n_reg)
    acum_rain = out["acum_rain"].cpu().numpy().reshape(1, n_cel)
    if out["mean_storage"] is not None:
        mean_storage = out["mean_storage"][0].cpu().numpy()
    if out["mean_speed"] is not None:
        mean_speed = out["mean_speed"][0].cpu().numpy()
    if out["mean_retorno"] is not None:
        mean_retorno = out["mean_retorno"][0].cpu().numpy()
    if out["mean_vfluxes"] is not None:
        mean_vfluxes, vfluxes = out["mean_vfluxes"][0].cpu().numpy(), out["vfluxes"][0].cpu().numpy()
    if out["rc_coef"] is not None:
        rc_coef = out["rc_coef"][0].cpu().numpy()
    if out["retorned"] is not None:
        retorned = out["retorned"][0].cpu().numpy().reshape(1, n_cel)
    if device_out:
        return out
    return {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


def shia_v1(ruta_bin, ruta_hdr, calib, n_cel, n_cont, n_conth, n_reg, stoin=None, hspeedin=None, ruta_storage=None,
            ruta_speed=None, ruta_vfluxes=None, ruta_binconv=None, ruta_binstra=None, ruta_hdrconv=None,
            ruta_hdrstra=None, ruta_retorno=None, ruta_rc=None):
    """models.shia_v1(...) with the positional order of wmfv2.py:2000-2020 and the 11 outputs of
    Modelos_heavy.f90:169-174: Q, Qsed, Qseparated, Hum, St1, St3, balance, speed, AreaControl, StoOut,
    Qsep_byrain.  As shipped, Q == 0 and the control/humidity series are never written."""
    n_cel, n_cont, n_conth, n_reg = int(n_cel), int(n_cont), int(n_conth), int(n_reg)
    _, pos = rain_read_ascii_table(ruta_hdr, n_reg)
    records = read_rain_records(ruta_bin, n_cel)
    cal = np.asarray(calib, dtype=np.float32).reshape(-1)
    save = None
    if int(save_storage) == 1:                       # Modelos_heavy.f90:317-328, 496-500
        if not ruta_storage:
            raise ValueError("models.save_storage = 1 needs ruta_storage")
        save = np.zeros(n_reg, dtype=np.int64) if guarda_cond is None else np.asarray(guarda_cond)
    if int(route_channels) == 1:
        return _shia_v1_routed(cal, records, pos, n_cel, n_cont, n_conth, n_reg, stoin)
    out = shia_v1_population(cal[None, :], records, pos, n_cel, n_reg, stoin, hspeedin, save_records=save)
    write_step_records(cal, records, pos, n_cel, n_reg, stoin, hspeedin, ruta_speed, ruta_retorno, ruta_vfluxes)
    if int(save_rc) == 1 and ruta_rc:                # Modelos_heavy.f90:543-544: one record of two columns
        write_float_basin(ruta_rc, out["rc_coef"][0], 1, n_cel, 2)
    if out.get("snapshot_records") is not None:      # in time order, like the Fortran loop writes them
        for k, rec in enumerate(out["snapshot_records"]):
            write_float_basin(ruta_storage, out["snapshots"][0, k], int(rec), n_cel, 5)
    f32 = np.float32
    Q = np.zeros((n_cont, n_reg), f32, order="F")
    Qsed = np.zeros((n_cont, 3, n_reg), f32, order="F")
    Qsep = np.zeros((n_cont, 3, n_reg), f32, order="F")
    Hum = np.zeros((n_conth, n_reg), f32, order="F")
    St1 = np.zeros((n_conth, n_reg), f32, order="F")
    St3 = np.zeros((n_conth, n_reg), f32, order="F")
    speed = np.zeros((n_cont, n_reg), f32, order="F")
    area_c = np.zeros((n_cont, n_reg), f32, order="F")
    Qsep_rain = np.zeros((n_cont, 2, n_reg), f32, order="F")
    return (Q, Qsed, Qsep, Hum, St1, St3, out["balance"][0], speed, area_c, np.asfortranarray(out["StoOut"][0]),
            Qsep_rain)


def _shia_v1_routed(cal, records, pos, n_cel, n_cont, n_conth, n_reg, stoin):
    """``models.route_channels = 1`` -- the routing EXTENSION: same 11 outputs, with Q [n_cont, N_reg]
    in m3/s filled at the outlet (row 0) and at the cells ``models.control`` marks (rows 1.. in cell order)."""
    global mean_rain, acum_rain
    _reject_out_of_scope()
    if int(save_storage) == 1:
        raise NotImplementedError("models.save_storage with models.route_channels")
    if drena is None:
        raise ValueError("models.drena (the basin structure) must be set for routing")
    if evpserie is None:
        raise ValueError("models.evpserie must be set (the Fortran dereferences EvpSerie(t), Modelos:396)")
    st = np.asarray(speed_type).astype(int).reshape(-1)[:3]
    if tuple(st) != (1, 1, 1):
        raise NotImplementedError("the routing extension is defined for models.speed_type = (1, 1, 1)")
    drain = np.asarray(drena, dtype=np.int32).reshape(-1, n_cel)[0]
    marks = np.zeros(n_cel, dtype=np.int64) if control is None else np.asarray(control).reshape(-1)[:n_cel]
    ctrl = np.where(marks != 0, np.cumsum(marks != 0), 0).astype(np.int32)
    if int(ctrl.max(initial=0)) + 1 > n_cont:
        raise ValueError("N_cont is smaller than the number of control points + 1")
    if stoin is not None and float(np.asarray(stoin).reshape(-1)[0]) > 0:
        s0 = np.asarray(stoin, dtype=np.float32).reshape(5, n_cel)
    else:
        base = np.zeros((5, n_cel), np.float32) if storage is None else np.asarray(storage, dtype=np.float32)
        s0 = base + np.float32(storage_constant)
    rr = to_dev(np.ascontiguousarray(records, dtype=np.int32))
    pe = to_dev(np.asarray(pos, dtype=np.int32)[:n_reg])
    out = ops.shia5_route_run(to_dev(cell_static_table(n_cel)), rr, pe, to_dev(np.asarray(evpserie, dtype=np.float32)[:n_reg]),
                              to_dev(np.ascontiguousarray(cal, dtype=np.float32)), to_dev(np.ascontiguousarray(s0)), drain, ctrl,
                              n_cont, dt=float(dt), retorno_gr=float(retorno_gr), retorno_aq=float(retorno_aq))
    rec_mean = (rr.to(torch.float64).mean(dim=1) / 1000.0).cpu().numpy()
    mean_rain = rec_mean[np.asarray(pos[:n_reg]) - 1].astype(np.float32).reshape(1, n_reg)
    acum_rain = out["acum_rain"].cpu().numpy().reshape(1, n_cel)
    f32 = np.float32
    z = lambda *shape: np.zeros(shape, f32, order="F")
    return (np.asfortranarray(out["Q"].cpu().numpy()), z(n_cont, 3, n_reg), z(n_cont, 3, n_reg), z(n_conth, n_reg),
            z(n_conth, n_reg), z(n_conth, n_reg), np.zeros(n_reg, f32), z(n_cont, n_reg), z(n_cont, n_reg),
            np.asfortranarray(out["StoOut"].cpu().numpy()), z(n_cont, 2, n_reg))
