"""Generate golden input/output vectors by running the UNMODIFIED Python reference.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes ``tests/golden/wmf_py_golden.npz``.
Reference functions exercised:
  wmf_py/cu_py/basics.py:124-177  basin_acum
  wmf_py/cu_py/basics.py:206-248  basin_cut
  wmf_py/cu_py/basics.py:180-203  basin_find
  wmf_py/models_py/core.py:64-295 shia_v1
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from wmf_py.cu_py.basics import basin_acum, basin_cut, basin_find  # noqa: E402
from wmf_py.models_py.core import ModelParams, ModelState, shia_v1  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


def scheidegger(ny, nx, seed):
    rng = np.random.default_rng(seed)
    return rng.choice(np.array([0, 1, 2], dtype=np.int8), size=(ny, nx), p=[0.25, 0.5, 0.25])


def steepest(ny, nx, seed, pit_frac=0.01):
    """D8 steepest descent of a tilted noisy plane, internal codes 0..7, -1 for pits."""
    rng = np.random.default_rng(seed)
    r, c = np.mgrid[0:ny, 0:nx]
    z = 0.7 * r + 0.4 * c + 3.0 * rng.standard_normal((ny, nx))
    dr = [0, 1, 1, 1, 0, -1, -1, -1]
    dc = [1, 1, 0, -1, -1, -1, 0, 1]
    zp = np.pad(z, 1, constant_values=np.inf)
    best = np.full((ny, nx), -1, dtype=np.int8)
    bestdrop = np.zeros((ny, nx))
    for d in range(8):
        nb = zp[1 + dr[d]:1 + dr[d] + ny, 1 + dc[d]:1 + dc[d] + nx]
        drop = (nb - z) / np.hypot(dr[d], dc[d])   # "lower" here means larger (r,c) -> flows down-right
        take = drop > bestdrop
        best[take] = d
        bestdrop[take] = drop[take]
    best[rng.random((ny, nx)) < pit_frac] = -1
    return best


cases = {
    "sch_40x56": (scheidegger(40, 56, 1), None),
    "steep_48x33": (steepest(48, 33, 2), None),
    "rand_24x31": (np.random.default_rng(3).integers(-1, 9, size=(24, 31)).astype(np.int8), None),
}
rng = np.random.default_rng(4)
cases["steep_mask_37x41"] = (steepest(37, 41, 5), rng.random((37, 41)) > 0.07)
cases["rand_mask_19x23"] = (rng.integers(0, 8, size=(19, 23)).astype(np.int8), rng.random((19, 23)) > 0.2)
cases["row_1x17"] = (np.array([[0, 0, 8, 0, 0, 4, 4, 0, 0, 0, 2, 6, 0, 4, 0, -1, 4]], dtype=np.int8), None)
cases["col_13x1"] = (np.array([[2, 2, 6, 2, 2, 2, 0, 6, 6, 2, 2, 2, 2]], dtype=np.int8).T.copy(), None)

for name, (fd, mask) in cases.items():
    if mask is None:
        mask = np.ones(fd.shape, dtype=bool)
    acc = basin_acum(fd.astype(np.int64), mask)
    assert acc.dtype == np.float64
    out[f"acum/{name}/flowdir"] = fd
    out[f"acum/{name}/mask"] = mask
    out[f"acum/{name}/acc"] = acc.astype(np.int32)
    assert np.array_equal(acc, acc.astype(np.int32))
    # basin_cut from the cell of largest accumulation inside the mask
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    dem = np.arange(fd.size, dtype=np.float64).reshape(fd.shape) + 1.0
    res = basin_cut(dem, (int(r0), int(c0)), fd.astype(np.int64), mask)
    out[f"cut/{name}/outlet"] = np.array([r0, c0], dtype=np.int64)
    out[f"cut/{name}/mask"] = res["mask"]
    out[f"cut/{name}/flowdir"] = res["flowdir"].astype(np.int32)
    out[f"cut/{name}/dem"] = res["dem"]

out["find/args"] = np.array([[60, 60, 0, 0, 30, 30, 3, 3], [12.5, 99.9, 10.0, 50.0, 2.5, 10.0, 7, 5],
                             [0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 4, 4]])
out["find/rc"] = np.array([basin_find(*a[:6], int(a[6]), int(a[7])) for a in out["find/args"]], dtype=np.int64)


def shia_case(name, seed, ny, nx, nt, et_kind, **kw):
    rng = np.random.default_rng(seed)
    rain = np.where(rng.random((nt, ny, nx)) < 0.4, rng.gamma(2.0, 4.0, size=(nt, ny, nx)), 0.0)
    rain[rng.random((nt, ny, nx)) < 0.02] = -1.0       # negative rain is clipped (core.py:152)
    forc = {"rain_mm_h": rain}
    if et_kind == "h":
        forc["et_mm_h"] = rng.random((nt, ny, nx)) * 0.6
    elif et_kind == "d":
        forc["et_mm_d"] = rng.random((nt, ny, nx)) * 9.0
    base = dict(dt=900.0, h_coef=1, h_exp=1, v_coef=1, v_exp=1, max_capilar=60.0, max_gravita=80.0,
                max_aquifer=300.0, drena=0.0, retorno_gr=0.05, storage_constant=0.0,
                k_perc_cap_to_grav=0.08, k_perc_grav_to_aqu=0.05, k_baseflow=0.01, max_infil_mm_h=20.0)
    base.update(kw)
    params = ModelParams(**base)
    st = ModelState(rng.random((ny, nx)) * 30, rng.random((ny, nx)) * 20, rng.random((ny, nx)) * 100)
    grid = {"dx": 30.0, "dy": 30.0} if seed % 2 == 0 else {}
    out[f"shia3/{name}/rain"] = rain
    out[f"shia3/{name}/state0"] = np.stack([st.storage_capilar, st.storage_gravita, st.storage_aquifer])
    try:
        st2, d = shia_v1(forc, grid, params, st, nt)
        out[f"shia3/{name}/raises"] = np.array(0)
    except RuntimeError:                      # core.py:255-256
        out[f"shia3/{name}/raises"] = np.array(1)
        st2, d = None, None
    for k in ("et_mm_h", "et_mm_d"):
        if k in forc:
            out[f"shia3/{name}/{k}"] = forc[k]
    out[f"shia3/{name}/params"] = np.array([base[k] for k in (
        "dt", "max_capilar", "max_gravita", "max_aquifer", "drena", "retorno_gr", "k_perc_cap_to_grav",
        "k_perc_grav_to_aqu", "k_baseflow", "max_infil_mm_h")] + [1.0 if base.get("eta_priority") == "proportional" else 0.0,
                                                             900.0 if grid else -1.0])
    if st2 is None:
        return
    out[f"shia3/{name}/state1"] = np.stack([st2.storage_capilar, st2.storage_gravita, st2.storage_aquifer])
    out[f"shia3/{name}/q_last"] = np.stack([st2.q_super_last, st2.q_base_last])
    out[f"shia3/{name}/q_outlet"] = d["q_outlet"]
    out[f"shia3/{name}/mass_error_step"] = d["diag"]["mass_error_step"]
    out[f"shia3/{name}/mass_error_cum"] = d["diag"]["mass_error_cum"]


shia_case("capfirst_et_h", 10, 9, 11, 30, "h")
shia_case("prop_et_d", 11, 8, 7, 25, "d", eta_priority="proportional")
shia_case("noet", 12, 6, 10, 20, None, retorno_gr=0.0, k_baseflow=0.03)
shia_case("raises_drena", 13, 5, 6, 20, None, drena=-0.02, retorno_gr=0.0)

np.savez_compressed(os.path.join(HERE, "wmf_py_golden.npz"), **out)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "wmf_py_golden.npz")), "bytes")
