"""Vectors of the Fortran family from the literal transliteration (fortran_transliteration.py) -- the second restatement
the C oracle is checked against.  Run in the build container:  python tests/golden/make_golden_fortran.py
-> tests/golden/fortran_golden.npz.  The outlet of every basin case is the cell of largest accumulation according to the
UNMODIFIED Python reference (wmf_py basin_acum), so nothing here depends on the oracle or on the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
from wmf_py.cu_py.basics import basin_acum  # noqa: E402
import fortran_transliteration as ft  # noqa: E402
import helpers  # noqa: E402

out = {}
f32 = np.float32

# ------------------------------------------------------------------ basin: find / cut / basics / coordXY / stream_nod
rng = np.random.default_rng(2026)
cases = {
    "steep_28x37": helpers.steepest(28, 37, 501),
    "steep_pits_40x23": helpers.steepest(40, 23, 502, pit_frac=0.03),
    "sch_31x31": rng.choice(np.array([0, 1, 2], np.int8), size=(31, 31), p=[0.25, 0.5, 0.25]),
    "derived_3x3": None,
}
for name, fd in cases.items():
    if fd is None:                                           # SURVEY 8(a) worked example
        key = np.array([[3, 2, 1], [6, 3, 2], [6, 6, 0]], dtype=np.int64)
        r0, c0 = 2, 2
    else:
        acc = basin_acum(fd.astype(np.int64), np.ones(fd.shape, bool))
        r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
        key = helpers.to_keypad(fd).astype(np.int64)
    nrows, ncols = key.shape
    DIR = np.zeros((ncols + 1, nrows + 1), dtype=np.int64)   # DIR(col, fil), 1-based
    DIR[1:, 1:] = key.T
    xll, yll, dx, dxp = 1000.0, 5000.0, 30.0, 28.5
    col, fil = int(c0) + 1, int(r0) + 1
    x, y = xll + dx * (col - 0.5), yll + dx * ((nrows - fil) + 0.5)
    assert ft.coord2fil_col(x, y, xll, yll, dx, ncols, nrows) == (col, fil)
    n, temp = ft.basin_find(col, fil, DIR, ncols, nrows)
    bf = ft.basin_cut(n, temp, ncols, nrows)
    dem = (2000.0 - 0.5 * np.add.outer(np.arange(nrows), np.arange(ncols)) + rng.random((nrows, ncols))).astype(f32)
    DEMv = np.zeros(n + 1, dtype=f32)
    DIRv = np.zeros(n + 1, dtype=np.int64)
    for i in range(1, n + 1):
        DEMv[i] = dem[bf[3, i] - 1, bf[2, i] - 1]
        DIRv[i] = key[bf[3, i] - 1, bf[2, i] - 1]
    acum, lon, pend, elev, sc = ft.basin_basics(bf, DEMv, DIRv, n, dxp, xll, yll, dx, nrows)
    X, Y = ft.basin_coordXY(bf, n, xll, yll, dx, nrows)
    p = f"basin/{name}/"
    out[p + "key"] = key.astype(np.int8)
    out[p + "geo"] = np.array([xll, yll, dx, dxp, x, y], dtype=np.float64)
    out[p + "colfil"] = np.array([col, fil])
    out[p + "dem"] = dem
    out[p + "structure"] = bf[1:, 1:].astype(np.int32)
    out[p + "acum"] = acum[1:].astype(np.int32)
    out[p + "long"], out[p + "pend"], out[p + "elev"] = lon[1:], pend[1:], elev[1:]
    out[p + "scalars"] = np.array(sc, dtype=f32)
    out[p + "X"], out[p + "Y"] = X[1:], Y[1:]
    for umbral in (3, max(4, n // 20)):
        cauce, nodos, traz, n_nodos, n_cauce = ft.basin_stream_nod(bf, acum, n, umbral)
        q = p + f"nod{umbral}/"
        out[q + "cauce"], out[q + "nodos"], out[q + "trazado"] = (a[1:].astype(np.int32) for a in (cauce, nodos, traz))
        out[q + "counts"] = np.array([n_nodos, n_cauce])

# coord2fil_col, incl. out-of-range
pts = [(1000.0 + 30.0 * 3.2, 5000.0 + 30.0 * 7.9), (999.0, 5100.0), (1000.0 + 30 * 40.5, 5100.0), (1100.0, 4990.0),
       (1100.0, 5000.0 + 30 * 28.01), (1000.0, 5000.0), (1000.0 + 30 * 37, 5000.0 + 30 * 28)]
out["coord/pts"] = np.array(pts)
out["coord/colfil"] = np.array([ft.coord2fil_col(x, y, 1000.0, 5000.0, 30.0, 37, 28) for x, y in pts])

# ------------------------------------------------------------------ SHIA 5-tank
def shia_case(name, N, T, seed, speed_type, ret_gr, ret_aq, niter, sto_in=False, hs_in=False, flags=None):
    r = np.random.default_rng(seed)
    n_rec = 6
    rec = np.vstack([np.zeros((1, N), np.int32), (1000 * r.gamma(2.0, 2.0, (n_rec - 1, N))).astype(np.int32)])
    pos = np.where(r.random(T) < 0.5, 1, r.integers(2, n_rec + 1, T)).astype(np.int32)
    inp = dict(
        v_coef=(r.uniform(0.5, 1.5, (4, N)) * np.array([[1e-2], [3e-3], [1e-3], [1e-4]])).astype(f32),
        h_coef=(r.uniform(0.5, 1.5, (4, N)) * np.array([[0.5], [0.05], [0.005], [1.0]])).astype(f32),
        h_exp=r.uniform(0.8, 1.2, (4, N)).astype(f32), Max_capilar=r.uniform(2, 15, N).astype(f32),
        Max_gravita=r.uniform(1, 5, N).astype(f32), Max_aquifer=r.uniform(5, 30, N).astype(f32),
        hill_long=np.full(N, 30.0, f32), elem_area=np.full(N, 900.0, f32), rain_records=rec, posEvento=pos,
        EvpSerie=(0.1 + 0.1 * np.sin(np.arange(T) / 7.0)).astype(f32))
    calib = r.uniform(0.7, 1.4, 11).astype(f32)
    sto = r.uniform(0.5, 4.0, (5, N)).astype(f32) if sto_in else None
    hs = r.uniform(0.01, 0.5, (4, N)).astype(f32) if hs_in else None
    fl = dict(show_storage=1, show_mean_speed=1)
    fl.update(flags or {})
    o = ft.shia_v1(N, T, calib, dt=300.0, speed_type=speed_type, retorno_gr=ret_gr, retorno_aq=ret_aq, calc_niter=niter,
                   StoIn=sto, HspeedIn=hs, **inp, **fl)
    p = f"shia/{name}/"
    for k, v in inp.items():
        out[p + "in/" + k] = v
    out[p + "in/calib"] = calib
    out[p + "in/cfg"] = np.array([300.0, ret_gr, ret_aq, niter, *speed_type], dtype=np.float64)
    if sto is not None:
        out[p + "in/StoIn"] = sto
    if hs is not None:
        out[p + "in/HspeedIn"] = hs
    for k, v in o.items():
        out[p + "out/" + k] = v


shia_case("type1", 160, 36, 601, (1, 1, 1), 0.0, 0.0, 5)
shia_case("type2_returns", 120, 30, 602, (2, 2, 2), 1.0, 1.0, 5, sto_in=True)
shia_case("mixed_flags", 140, 32, 603, (1, 2, 1), 1.0, 0.0, 3, sto_in=True,
          flags=dict(show_mean_retorno=1, save_vfluxes=1, save_rc=1))
shia_case("hspeed_in", 100, 24, 604, (2, 1, 2), 0.0, 1.0, 4, hs_in=True)
np.savez_compressed(os.path.join(HERE, "fortran_golden.npz"), **out)
print(len(out), "arrays")
