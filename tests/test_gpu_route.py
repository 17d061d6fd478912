"""N4, the routing EXTENSION (no reference behaviour exists: Q == 0 as shipped): GPU wavefront kernels against their own
CPU oracle (oracle/wmf_oracle.c orc_f_shia_route), plus mass conservation of the oracle itself in fp64."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from wmf_heavy_b200 import models, ops

pytestmark = pytest.mark.gpu


def _basin(ny, nx, seed):
    """The largest basin of a Scheidegger raster as an upstream-first structure (drain ids, f90:507-576)."""
    fd = ops.synth_scheidegger_host(ny, nx, seed=seed)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    key = helpers.to_keypad(fd)
    st = oracle.f_basin_find(key, int(c0) + 1, int(r0) + 1)
    return st[0].astype(np.int32), int(acc[r0, c0])


@pytest.mark.parametrize("ret", [(0.0, 0.0), (1.0, 1.0)])
def test_route_vs_oracle(ret):
    drain, N = _basin(200, 200, 7)
    assert drain.size == N and drain[-1] == 0
    T = 60
    inp = helpers.shia5_inputs(oracle, N, T, 102)
    inp.retorno_gr, inp.retorno_aq = ret
    inp.h_coef[3] = 0.5 + 0.5 * np.random.default_rng(3).random(N).astype(np.float32)      # channel speed, m/s
    rng = np.random.default_rng(103)
    ctrl = np.zeros(N, np.int32)
    marks = rng.choice(N - 1, 5, replace=False)
    ctrl[np.sort(marks)] = np.arange(1, 6)
    cal = np.ones(11, np.float32)
    sto0 = np.full((5, N), 0.001, np.float32)
    o = oracle.f_shia_route(inp, cal, sto0, drain, ctrl, 6)
    out = ops.shia5_route_run(torch.from_numpy(helpers.cell_static(inp)).cuda(), torch.from_numpy(inp.rain_records).cuda(),
                              torch.from_numpy(inp.pos_evento).cuda(), torch.from_numpy(inp.evp).cuda(),
                              torch.from_numpy(cal).cuda(), torch.from_numpy(sto0.copy()).cuda(), drain, ctrl, 6, dt=inp.dt,
                              retorno_gr=ret[0], retorno_aq=ret[1])
    np.testing.assert_allclose(out["StoOut"].cpu().numpy(), o["StoOut"], rtol=1e-5, atol=8e-5)
    q, eq = out["Q"].cpu().numpy(), o["Q"]
    assert eq[0].max() > 0                                   # water does reach the outlet
    np.testing.assert_allclose(q, eq, rtol=1e-4, atol=1e-6 * float(eq.max()))
    np.testing.assert_allclose(out["acum_rain"].cpu().numpy(), o["acum_rain"], rtol=1e-6)
    # the oracle conserves mass: storage change = rain - losses - outflow (all tanks, m3)
    m3mm = (inp.elem_area / 1000.0).astype(np.float64)
    d_sto = float(((o["StoOut"].astype(np.float64) - sto0) * m3mm).sum())
    rain, loss, outflow = o["totals"]
    assert abs(d_sto - (rain - loss - outflow)) <= 1e-4 * rain


def test_route_topology_and_errors():
    drain, N = _basin(90, 80, 11)
    rank, dmax, ptr, idx = ops.route_topology(drain)
    recv = np.where(drain != 0, N - drain, -1)
    assert rank[-1] == dmax and (rank[recv[:-1][recv[:-1] >= 0]] == rank[:-1][recv[:-1] >= 0] + 1).all()
    for j in (N - 1, N // 2, 0):
        assert sorted(idx[ptr[j]:ptr[j + 1]]) == list(np.nonzero(recv == j)[0])
    with pytest.raises(ValueError):
        ops.route_topology(np.array([3, 0, 0], np.int32))    # the receiver of cell 0 would be cell 0 itself


def test_run_shia_routed_object_api(tmp_path):
    """SimuBasin.run_shia with models.route_channels = 1: Qsim at the outlet and at control points, against the oracle
    run on the same files and the basin's own structure; with the flag off, Qsim is 0 as the reference ships it."""
    from wmf_heavy_b200 import wmfv2
    ny, nx = 120, 150
    fd = ops.synth_scheidegger_host(ny, nx, seed=21)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    key = helpers.to_keypad(fd)
    wmfv2.set_map_properties(ncols=nx, nrows=ny, xll=0.0, yll=0.0, dx=30.0, dxp=30.0)
    x, y = 30.0 * (c0 + 0.5), 30.0 * ((ny - (r0 + 1)) + 0.5)
    DIR = np.asfortranarray(key.T.astype(np.int32))
    DEM = np.asfortranarray((1000 - 0.3 * np.add.outer(np.arange(nx), np.arange(ny))).astype(np.float32))
    sb = wmfv2.SimuBasin(x, y, DEM, DIR, dt=300, threshold=30)
    N, T = sb.ncells, 40
    inp = helpers.shia5_inputs(oracle, N, T, 22)
    inp.hill_long = models.hill_long[0].copy()
    inp.elem_area = models.elem_area[0].copy()
    inp.h_coef[3] = 0.8
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None] for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.speed_type = np.array([1, 1, 1])
    models.control = np.zeros((1, N), np.int32)
    models.control[0, [N // 3, N // 2]] = [7, 9]
    path = str(tmp_path / "lluvia")
    models.write_rain(path + ".bin", path + ".hdr", inp.rain_records, inp.pos_evento)
    cal = np.ones(11, np.float32)
    try:
        models.route_channels = 1
        ret, qdf = sb.run_shia(cal, path, T, EvpSerie=inp.evp)
    finally:
        models.route_channels = 0
    ctrl = np.zeros(N, np.int32)
    ctrl[[N // 3, N // 2]] = [1, 2]
    o = oracle.f_shia_route(inp, cal, np.full((5, N), 0.001, np.float32), np.asarray(sb.structure[0], np.int32), ctrl, 3)
    assert ret["Qsim"].shape == (3, T) and ret["Qsim"][0].max() > 0
    np.testing.assert_allclose(ret["Qsim"], o["Q"], rtol=1e-4, atol=1e-6 * float(o["Q"].max()))
    np.testing.assert_allclose(ret["Storage"], o["StoOut"], rtol=1e-5, atol=8e-5)
    assert list(qdf.columns) == ["7", "9"] and len(qdf) == T
    ret0, _ = sb.run_shia(cal, path, T, EvpSerie=inp.evp)
    assert not ret0["Qsim"].any()


def test_fitness_reduction_vs_numpy():
    """ops.fitness_nash_peak: Nash-Sutcliffe efficiency and peak-flow error of P hydrographs, NaN gaps in the observation
    skipped (fp64 sums on the device; numpy in fp64 as the check)."""
    import torch
    rng = np.random.default_rng(5)
    P, T = 37, 1000
    qobs = (5 + 3 * np.sin(np.arange(T) / 40.0) + rng.random(T)).astype(np.float32)
    qobs[rng.random(T) < 0.05] = np.nan
    qsim = (qobs[None] * rng.uniform(0.5, 1.5, (P, 1)) + rng.normal(0, 0.3, (P, T))).astype(np.float32)
    qsim = np.nan_to_num(qsim, nan=1.0).astype(np.float32)
    got = ops.fitness_nash_peak(torch.from_numpy(qsim).cuda(), torch.from_numpy(qobs).cuda()).cpu().numpy()
    ok = np.isfinite(qobs)
    o, s = qobs[ok].astype(np.float64), qsim[:, ok].astype(np.float64)
    nash = 1 - ((o - s) ** 2).sum(1) / ((o - o.mean()) ** 2).sum()
    peak = np.abs(s.max(1) - o.max()) / o.max()
    np.testing.assert_allclose(got[:, 0], nash, rtol=1e-12)
    np.testing.assert_allclose(got[:, 1], peak, rtol=1e-12)


def test_calib_nsgaii_object_api(tmp_path):
    """SimuBasin.Calib_NSGAII(nsga_el, nodo_eval, ...) with the reference's signature (wmfv2.py:2060-2106): population
    rounded up to a multiple of 4, every generation evaluated in one go, (pop, QsimPar, fitness (2, pop)) returned.  On
    the routed extension the hydrographs depend on the genes, so the best Nash of the last population is at least the
    best of a random one; as shipped (no routing) every hydrograph is zero and the run still completes."""
    from wmf_heavy_b200 import wmfv2
    ny, nx = 60, 70
    fd = ops.synth_scheidegger_host(ny, nx, seed=23)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    key = helpers.to_keypad(fd)
    wmfv2.set_map_properties(ncols=nx, nrows=ny, xll=0.0, yll=0.0, dx=30.0, dxp=30.0)
    x, y = 30.0 * (c0 + 0.5), 30.0 * ((ny - (r0 + 1)) + 0.5)
    DIR = np.asfortranarray(key.T.astype(np.int32))
    DEM = np.asfortranarray((1000 - 0.3 * np.add.outer(np.arange(nx), np.arange(ny))).astype(np.float32))
    sb = wmfv2.SimuBasin(x, y, DEM, DIR, dt=300, threshold=30)
    N, T = sb.ncells, 30
    inp = helpers.shia5_inputs(oracle, N, T, 24)
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None] for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.speed_type = np.array([1, 1, 1])
    models.control = np.zeros((1, N), np.int32)
    models.show_storage = models.show_mean_speed = models.save_storage = 0
    path = str(tmp_path / "lluvia")
    models.write_rain(path + ".bin", path + ".hdr", inp.rain_records, inp.pos_evento)
    np.random.seed(3)
    try:
        models.route_channels = 1
        truth = np.array([0.8, 1.2, 1.1, 0.9, 0.7, 0.6, 0.9, 0.8, 0.9, 1.1], np.float32)
        qobs = sb.run_shia(truth, path, T, EvpSerie=inp.evp, QsimDataFrame=False)["Qsim"][0]
        el = wmfv2.nsgaii_element(path, qobs, T, 1, sb, evp=[0.5, 1.5], infil=[0.5, 1.5], perco=[0.5, 1.5], losses=[0.5, 1.5],
                                  velRun=[0.5, 1.5], velSub=[0.5, 1.5], velSup=[0.5, 1.5], velStream=[0.5, 1.5],
                                  Hu=[0.5, 1.5], Hg=[0.5, 1.5])
        assert np.allclose(el.__evalfunc__(qobs), (1.0, 0.0))
        pop, qsim, fit = sb.Calib_NSGAII(el, 0, pop_size=6, NGEN=2)
        assert len(pop) == 8 and fit.shape == (2, 8) and len(qsim) == 8 and qsim[0].shape == (T,)
        assert np.isfinite(fit).all() and fit[0].max() <= 1.0 and fit[1].min() >= 0.0
        assert all(len(ind) == 1 and len(ind[0]) == 10 for ind in pop)
    finally:
        models.route_channels = 0
    pop0, q0, fit0 = sb.Calib_NSGAII(el, 0, pop_size=4, NGEN=1)
    assert len(pop0) == 4 and not np.asarray(q0).any() and fit0.shape == (2, 4)
