"""SHIA parity.  fp64 3-tank (wmf_py semantics): storages 1e-12 relative, q_outlet 1e-10 relative.
fp32 5-tank (Fortran semantics): storages 1e-5 relative (north_star tolerance) plus an absolute floor that
profiles/r02_shia_error_table.md justifies (tools/shia_error_table.py, 20 000 cells x 200 / 1000 steps, three speed-type
mixes, with and without returns): the FAST arithmetic forms (x * 0.001f for / 1000., S1 * (1 / H1), exp2(0.6 * log2 x) on
the SFU) stay within 6.1e-5 mm of the oracle in every tank -- tank 1 holds O(100 mm), whose fp32 ulp is 7.6e-6 mm, and the
other tanks are residuals of it -- so the floor is 8e-5 mm; the STRICT forms (`strict=True`: the Fortran's own
operations) stay within 1.5e-5 mm and 9.8e-6 relative (bit-identical in most tanks), floor 2e-5 mm.  Kinematic speeds
follow S**expo, hence 3e-4 relative.  `balance` is a catastrophic-cancellation quantity and is compared with an absolute
tolerance scaled by sum(StoOut)."""
import os

import numpy as np
import pytest
import torch

import helpers
import oracle
from conftest import golden_names
from test_oracle import _forc
from wmf_heavy_b200 import models, ops, wmfv2
from wmf_heavy_b200.models_py.core import ModelParams, ModelState, shia_v1

pytestmark = pytest.mark.gpu
ATOL_FAST, ATOL_STRICT = 8e-5, 2e-5          # [mm]; see the module docstring


def _mp(p, **kw):
    return ModelParams(dt=p.dt, h_coef=1, h_exp=1, v_coef=1, v_exp=1, max_capilar=p.max_capilar,
                       max_gravita=p.max_gravita, max_aquifer=p.max_aquifer, drena=p.drena, retorno_gr=p.retorno_gr,
                       storage_constant=0.0, k_perc_cap_to_grav=p.k_perc_cap_to_grav,
                       k_perc_grav_to_aqu=p.k_perc_grav_to_aqu, k_baseflow=p.k_baseflow,
                       max_infil_mm_h=p.max_infil_mm_h, eta_priority=p.eta_priority, **kw)


@pytest.mark.parametrize("name", golden_names("shia3"))
def test_shia3_golden(golden, name):
    f, grid, p = _forc(golden, name)
    s0 = golden[f"shia3/{name}/state0"]
    nt = f["rain_mm_h"].shape[0]
    st = ModelState(s0[0].copy(), s0[1].copy(), s0[2].copy())
    if int(golden[f"shia3/{name}/raises"]):
        with pytest.raises(RuntimeError, match="cumulative mass balance error >1%"):
            shia_v1(f, grid, _mp(p), st, nt)
        return
    st2, d = shia_v1(f, grid, _mp(p), st, nt)
    assert np.array_equal(st.storage_capilar, s0[0])                       # inputs are not mutated
    s1 = golden[f"shia3/{name}/state1"]
    for got, exp in zip((st2.storage_capilar, st2.storage_gravita, st2.storage_aquifer), s1):
        np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(st2.q_super_last, golden[f"shia3/{name}/q_last"][0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(st2.q_base_last, golden[f"shia3/{name}/q_last"][1], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(d["q_outlet"], golden[f"shia3/{name}/q_outlet"], rtol=1e-10)
    np.testing.assert_allclose(d["diag"]["mass_error_cum"], golden[f"shia3/{name}/mass_error_cum"], atol=1e-11)


def test_shia3_reference_tests():              # wmf_py/tests/test_models_core.py:57-192 through the drop-in API
    base = dict(dt=3600.0, h_coef=1, h_exp=1, v_coef=1, v_exp=1, drena=0.0, retorno_gr=0.0, storage_constant=0.0)
    z = lambda: ModelState(np.zeros((1, 1)), np.zeros((1, 1)), np.zeros((1, 1)))
    p = ModelParams(**base, max_capilar=100, max_gravita=100, max_aquifer=500, k_perc_cap_to_grav=1.0,
                    k_perc_grav_to_aqu=1.0, k_baseflow=0.2, max_infil_mm_h=100.0, separate_fluxes=True)
    _, out = shia_v1({"rain_mm_h": np.array([[[10.0]], [[0.0]]])}, {}, p, z(), nsteps=2)
    assert np.isclose(out["q_base_t"][0, 0, 0], 2.0) and np.isclose(out["q_base_t"][1, 0, 0], 1.6)
    p = ModelParams(**base, max_capilar=200, max_gravita=200, max_aquifer=200, max_infil_mm_h=100.0)
    s2, _ = shia_v1({"rain_mm_h": np.zeros((1, 1, 1)), "et_mm_h": np.array([[[60.0]]])}, {}, p,
                    ModelState(np.array([[50.0]]), np.array([[20.0]]), np.array([[10.0]])), nsteps=1)
    assert np.isclose(s2.storage_capilar[0, 0], 0.0) and np.isclose(s2.storage_gravita[0, 0], 10.0)
    assert np.isclose(s2.storage_aquifer[0, 0], 10.0)
    p = ModelParams(**base, max_capilar=100, max_gravita=100, max_aquifer=100, max_infil_mm_h=10.0,
                    separate_fluxes=True, save_storage=True)
    _, out = shia_v1({"rain_mm_h": np.array([[[50.0]]])}, {}, p, z(), nsteps=1)
    assert out["q_super_t"].shape == (1, 1, 1) and out["storage_aquifer_t"].shape == (1, 1, 1)
    assert np.isclose(out["q_super_t"][0, 0, 0], 40.0) and np.isclose(out["q_outlet"][0], 40.0)
    p2 = ModelParams(**base, max_capilar=100, max_gravita=100, max_aquifer=100, max_infil_mm_h=10.0)
    _, out2 = shia_v1({"rain_mm_h": np.array([[[50.0]]])}, {}, p2, z(), nsteps=1)
    assert "q_super_t" not in out2 and "storage_capilar_t" not in out2
    with pytest.raises(KeyError):
        shia_v1({}, {}, p2, z(), nsteps=1)
    # test_mass_balance_24_steps
    rain = np.full((24, 5, 5), 10.0)
    p = ModelParams(**base, max_capilar=100, max_gravita=200, max_aquifer=400, max_infil_mm_h=1e6)
    st = ModelState(np.zeros((5, 5)), np.zeros((5, 5)), np.zeros((5, 5)))
    st2, out = shia_v1({"rain_mm_h": rain}, {"dx": 30.0, "dy": 30.0}, p, st, nsteps=24)
    dS = st2.storage_capilar.sum() + st2.storage_gravita.sum() + st2.storage_aquifer.sum()
    q_mm = out["q_outlet"] * 3600.0 / (900.0 * 1e-3)
    assert abs(rain.sum() - dS - q_mm.sum()) / rain.sum() <= 0.005
    assert np.all(out["diag"]["mass_error_cum"] <= 0.005)


def test_shia3_vs_oracle_larger():
    rng = np.random.default_rng(50)
    ny, nx, nt = 70, 93, 60
    rain = np.where(rng.random((nt, ny, nx)) < 0.3, rng.gamma(2.0, 5.0, (nt, ny, nx)), 0.0)
    et = rng.random((nt, ny, nx)) * 0.4
    p = ModelParams(dt=600.0, h_coef=1, h_exp=1, v_coef=1, v_exp=1, max_capilar=80.0, max_gravita=60.0,
                    max_aquifer=400.0, drena=0.0, retorno_gr=0.02, storage_constant=0.0, k_perc_cap_to_grav=0.05,
                    k_perc_grav_to_aqu=0.04, k_baseflow=0.02, max_infil_mm_h=15.0, separate_fluxes=True, save_storage=True)
    s0 = [rng.random((ny, nx)) * m for m in (40, 30, 200)]
    for prio in ("capilar_first", "proportional"):
        p.eta_priority = prio
        f = {"rain_mm_h": rain, "et_mm_h": et}
        o = oracle.shia3(f, {"dx": 30.0, "dy": 30.0}, p, *s0, nt, separate_fluxes=True, save_storage=True)
        st2, d = shia_v1(f, {"dx": 30.0, "dy": 30.0}, p, ModelState(*[s.copy() for s in s0]), nt)
        for k, got in (("storage_capilar", st2.storage_capilar), ("storage_gravita", st2.storage_gravita),
                       ("storage_aquifer", st2.storage_aquifer)):
            np.testing.assert_allclose(got, o[k], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(d["q_outlet"], o["q_outlet"], rtol=1e-10)
        for k in ("q_super_t", "q_base_t", "q_total_t", "storage_capilar_t", "storage_gravita_t", "storage_aquifer_t"):
            np.testing.assert_allclose(d[k], o[k], rtol=1e-12, atol=1e-12)


def _run5(inp, calibs, sto0, show=True):
    cs = torch.from_numpy(helpers.cell_static(inp)).cuda()
    P = calibs.shape[0]
    N = sto0.shape[-1]
    sto = torch.from_numpy(np.broadcast_to(sto0, (P, 5, N)).copy()).cuda()
    return ops.shia5_run(cs, torch.from_numpy(inp.rain_records).cuda(), torch.from_numpy(inp.pos_evento).cuda(),
                         torch.from_numpy(inp.evp).cuda(), torch.from_numpy(calibs).cuda(), sto, None, dt=inp.dt,
                         retorno_gr=inp.retorno_gr, retorno_aq=inp.retorno_aq, calc_niter=inp.calc_niter,
                         speed_type=inp.speed_type, show_storage=show, show_mean_speed=show)


@pytest.mark.parametrize("speed_type", [(1, 1, 1), (2, 2, 2), (1, 2, 1)])
@pytest.mark.parametrize("ret", [(0.0, 0.0), (1.0, 1.0)])
def test_shia5_vs_oracle(speed_type, ret):
    N, T = 3001, 120
    inp = helpers.shia5_inputs(oracle, N, T, 60, speed_type=speed_type)
    inp.retorno_gr, inp.retorno_aq = ret
    rng = np.random.default_rng(61)
    calibs = np.vstack([np.ones(11), rng.uniform(0.5, 2.0, (2, 11))]).astype(np.float32)
    sto0 = (rng.random((5, N)) * np.array([[40], [10], [5], [100], [1]])).astype(np.float32) + 0.001
    out = _run5(inp, calibs, sto0)
    for p in range(calibs.shape[0]):
        o = oracle.f_shia_v1(inp, calibs[p], sto0)
        got = out["StoOut"][p].cpu().numpy()
        np.testing.assert_allclose(got, o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
        np.testing.assert_allclose(out["hspeed"][p].cpu().numpy(), o["hspeed"], rtol=3e-4, atol=1e-9)
        scale = float(o["StoOut"].sum())
        # tight against the fp64 sum of the same fp32 terms; loose (fp32 running-sum noise of the Fortran
        # formula itself, ~eps32 * sum(StoOut) * sqrt(5N)) against the reference-order fp32 balance
        np.testing.assert_allclose(out["balance"][p].cpu().numpy(), o["balance64"], rtol=1e-4, atol=1e-6 * scale)
        np.testing.assert_allclose(out["balance"][p].cpu().numpy(), o["balance"], atol=3e-5 * scale)
        np.testing.assert_allclose(out["mean_storage"][p].cpu().numpy(), o["mean_storage"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["mean_speed"][p].cpu().numpy(), o["mean_speed"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(out["mean_rain"].cpu().numpy(), o["mean_rain"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(out["acum_rain"].cpu().numpy(), o["acum_rain"], rtol=1e-6)   # RainInt * 0.001f vs / 1000.


def test_shia5_no_diag_and_resume():
    """Checkpoint/resume: running T steps == running T/2, feeding StoOut/hspeed back (StoIn/HspeedIn), T/2 more."""
    N, T = 1000, 64
    inp = helpers.shia5_inputs(oracle, N, T, 70, speed_type=(2, 2, 2))
    cal = np.ones((1, 11), np.float32)
    sto0 = np.full((5, N), 0.001, np.float32)
    full = _run5(inp, cal, sto0, show=False)
    assert full["mean_storage"] is None
    o = oracle.f_shia_v1(inp, cal[0], sto0)
    np.testing.assert_allclose(full["StoOut"][0].cpu().numpy(), o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    cs = torch.from_numpy(helpers.cell_static(inp)).cuda()
    rr, ev = torch.from_numpy(inp.rain_records).cuda(), torch.from_numpy(inp.evp).cuda()
    pe = torch.from_numpy(inp.pos_evento).cuda()
    kw = dict(dt=inp.dt, calc_niter=inp.calc_niter, speed_type=inp.speed_type)
    sto = torch.from_numpy(sto0[None].copy()).cuda()
    a = ops.shia5_run(cs, rr, pe[: T // 2].contiguous(), ev[: T // 2].contiguous(), torch.from_numpy(cal).cuda(), sto, None, **kw)
    b = ops.shia5_run(cs, rr, pe[T // 2:].contiguous(), ev[T // 2:].contiguous(), torch.from_numpy(cal).cuda(), a["StoOut"],
                      a["hspeed"], **kw)
    assert torch.equal(b["StoOut"], full["StoOut"]) and torch.equal(b["hspeed"], full["hspeed"])


def test_run_shia_object_api(tmp_path):
    """SimuBasin(...).run_shia(Calibracion, rain_path, N_intervals) with the reference's signature and
    returned dict keys (wmfv2.py:1924-2040), checked against the oracle run on the same files."""
    ny, nx = 120, 150
    fd = helpers.steepest(ny, nx, 80)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    key = helpers.to_keypad(fd)
    wmfv2.set_map_properties(ncols=nx, nrows=ny, xll=0.0, yll=0.0, dx=30.0, dxp=30.0)
    x, y = 30.0 * (c0 + 0.5), 30.0 * ((ny - (r0 + 1)) + 0.5)
    DIR = np.asfortranarray(key.T.astype(np.int32))
    DEM = np.asfortranarray((1000 - 0.3 * np.add.outer(np.arange(nx), np.arange(ny))).astype(np.float32))
    sb = wmfv2.SimuBasin(x, y, DEM, DIR, dt=300, threshold=30)
    N, T = sb.ncells, 48
    assert N == int(acc[r0, c0])
    inp = helpers.shia5_inputs(oracle, N, T, 81)
    inp.hill_long = models.hill_long[0].copy()
    inp.elem_area = models.elem_area[0].copy()
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None] for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.speed_type = np.array([1, 1, 1])
    models.show_storage = 1
    path = str(tmp_path / "lluvia")
    models.write_rain(path + ".bin", path + ".hdr", inp.rain_records, inp.pos_evento)
    cal = np.ones(11, np.float32)
    ret, qdf = sb.run_shia(cal, path, T, EvpSerie=inp.evp)
    assert set(ret) >= {"Qsim", "Balance", "Storage", "Rain_Acum", "Rain_hietogram"}
    o = oracle.f_shia_v1(inp, cal, np.full((5, N), 0.001, np.float32))
    np.testing.assert_allclose(ret["Storage"], o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    assert ret["Storage"].shape == (5, N) and ret["Qsim"].shape[1] == T and not ret["Qsim"].any()
    np.testing.assert_allclose(ret["Rain_hietogram"][0], o["mean_rain"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(ret["Rain_Acum"][0], o["acum_rain"], rtol=1e-6)
    assert len(qdf) == T
    pop = sb.run_shia_population([cal, cal * 1.5], path, T, EvpSerie=inp.evp)
    np.testing.assert_allclose(pop["StoOut"][0], o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    o2 = oracle.f_shia_v1(inp, cal * 1.5, np.full((5, N), 0.001, np.float32))
    np.testing.assert_allclose(pop["StoOut"][1], o2["StoOut"], rtol=1e-5, atol=ATOL_FAST)


def _prefix(inp, T):
    """The same inputs cut to the first T steps (the storages after step T - 1 of the long run)."""
    import copy
    q = copy.copy(inp)
    q.pos_evento, q.evp = inp.pos_evento[:T].copy(), inp.evp[:T].copy()
    return q


def test_shia5_storage_snapshots():
    """N3 / save_storage (Modelos_heavy.f90:496-500): the snapshot after step t equals the final storages of an
    oracle run over the first t + 1 steps; the run itself is unchanged by taking snapshots."""
    N, T = 2000, 40
    inp = helpers.shia5_inputs(oracle, N, T, 90, speed_type=(1, 2, 1))
    cal = np.vstack([np.ones(11), np.full(11, 1.3)]).astype(np.float32)
    sto0 = np.full((5, N), 0.001, np.float32)
    steps = [0, 7, 8, 39]
    slot = np.full(T, -1, np.int32)
    slot[steps] = np.arange(len(steps))
    cs = torch.from_numpy(helpers.cell_static(inp)).cuda()
    args = (cs, torch.from_numpy(inp.rain_records).cuda(), torch.from_numpy(inp.pos_evento).cuda(),
            torch.from_numpy(inp.evp).cuda(), torch.from_numpy(cal).cuda())
    kw = dict(dt=inp.dt, calc_niter=inp.calc_niter, speed_type=inp.speed_type)
    sto = torch.from_numpy(np.broadcast_to(sto0, (2, 5, N)).copy()).cuda()
    out = ops.shia5_run(*args, sto, None, snap_slot=torch.from_numpy(slot).cuda(), n_snap=len(steps), **kw)
    plain = ops.shia5_run(*args, torch.from_numpy(np.broadcast_to(sto0, (2, 5, N)).copy()).cuda(), None, **kw)
    assert plain["snapshots"] is None and torch.equal(plain["StoOut"], out["StoOut"])
    snaps = out["snapshots"].cpu().numpy()
    assert snaps.shape == (2, len(steps), 5, N)
    for p in range(2):
        for k, t in enumerate(steps):
            o = oracle.f_shia_v1(_prefix(inp, t + 1), cal[p], sto0)
            np.testing.assert_allclose(snaps[p, k], o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    assert np.array_equal(snaps[:, -1], out["StoOut"].cpu().numpy())         # the last step's snapshot is StoOut
    with pytest.raises(ValueError):
        ops.shia5_run(*args, sto, None, snap_slot=torch.from_numpy(slot[:5].copy()).cuda(), n_snap=2, **kw)


def test_shia_v1_writes_storage_records(tmp_path):
    """models.shia_v1 with save_storage = 1: records guarda_cond[t] of ruta_storage in write_float_basin's layout."""
    N, T = 500, 24
    inp = helpers.shia5_inputs(oracle, N, T, 91)
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None] for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.hill_long, models.elem_area = inp.hill_long[None], inp.elem_area[None]
    models.speed_type, models.evpserie, models.dt = np.array([1, 1, 1]), inp.evp, inp.dt
    models.storage, models.storage_constant, models.rain_first_point = None, 0.001, 1
    models.retorno_gr = models.retorno_aq = 0
    models.control, models.control_h = np.zeros((1, N), np.int32), np.zeros((1, N), np.int32)
    path = str(tmp_path / "rain")
    models.write_rain(path + ".bin", path + ".hdr", inp.rain_records, inp.pos_evento)
    sto_path = str(tmp_path / "states.StObin")
    models.save_storage = 1
    models.guarda_cond = np.zeros(T, np.int32)
    models.guarda_cond[[3, 10, 23]] = [1, 2, 3]
    try:
        res = models.shia_v1(path + ".bin", path + ".hdr", np.ones(11, np.float32), N, 1, 0, T, None, None, sto_path)
    finally:
        models.save_storage = 0
    import os
    assert os.path.getsize(sto_path) == 3 * 5 * N * 4
    for rec, t in ((1, 3), (2, 10), (3, 23)):
        v, rc = models.read_float_basin_ncol(sto_path, rec, N, 5)
        o = oracle.f_shia_v1(_prefix(inp, t + 1), np.ones(11, np.float32), np.full((5, N), 0.001, np.float32))
        assert rc == 0 and v.shape == (5, N)
        np.testing.assert_allclose(v, o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    np.testing.assert_array_equal(models.read_float_basin_ncol(sto_path, 3, N, 5)[0], res[9])   # record 3 == StoOut
    # a stored state restarts the model (StorageLoc of run_shia): steps 11.. from record 2 end where the full run ends
    v, _ = models.read_float_basin_ncol(sto_path, 2, N, 5)
    if v[0, 0] > 0:                                   # (StoIn is only honoured when StoIn(1,1) > 0, Modelos:229)
        models.write_rain(path + "_b.bin", path + "_b.hdr", inp.rain_records, inp.pos_evento[11:])
        models.evpserie = inp.evp[11:]
        res2 = models.shia_v1(path + "_b.bin", path + "_b.hdr", np.ones(11, np.float32), N, 1, 0, T - 11, v, None, None)
        np.testing.assert_allclose(res2[9], res[9], rtol=1e-6, atol=1e-6)  # (speed_type 1: hspeed carries no state)


def test_shia_v1_writes_step_records(tmp_path):
    """models.shia_v1 with save_speed / save_retorno / save_vfluxes + guarda_vfluxes (Modelos_heavy.f90:502-515): record t
    of ruta_speed = hspeed after step t, record t of ruta_retorno = the return flow of step t, record guarda_vfluxes(t)
    of ruta_vfluxes = the vertical fluxes of step t -- against the literal transliteration run with the same flags
    (kinematic speeds in two tanks so that hspeed changes from step to step, returns on)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fortran_transliteration",
                                                  os.path.join(os.path.dirname(__file__), "golden", "fortran_transliteration.py"))
    ft = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ft)
    N, T = 60, 14
    inp = helpers.shia5_inputs(oracle, N, T, 191, speed_type=(2, 1, 2))
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None] for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.hill_long, models.elem_area = inp.hill_long[None], inp.elem_area[None]
    models.speed_type, models.evpserie, models.dt = np.array([2, 1, 2]), inp.evp, inp.dt
    models.storage, models.storage_constant, models.rain_first_point = None, 0.001, 1
    models.retorno_gr, models.retorno_aq, models.calc_niter = 1, 1, 5
    models.control, models.control_h = np.zeros((1, N), np.int32), np.zeros((1, N), np.int32)
    path = str(tmp_path / "rain")
    models.write_rain(path + ".bin", path + ".hdr", inp.rain_records, inp.pos_evento)
    sp_path, re_path, vf_path = (str(tmp_path / nm) for nm in ("speed.bin", "retorno.bin", "vfluxes.bin"))
    gv = np.zeros(T, np.int32)
    gv[[2, 7, 13]] = [1, 2, 3]
    models.save_speed = models.save_retorno = models.save_vfluxes = 1
    models.guarda_vfluxes = gv
    models.strict_arithmetic = 1
    cal = np.linspace(0.8, 1.3, 11).astype(np.float32)
    rng = np.random.default_rng(192)                          # tanks 3 / 4 above capacity in many cells: return flow from step 1 on
    sto0 = np.stack([0.5 * inp.max_capilar, np.full(N, 5.0, np.float32), inp.max_gravita * rng.uniform(0.6, 1.6, N),
                     inp.max_aquifer * rng.uniform(0.8, 1.2, N), np.full(N, 0.001, np.float32)]).astype(np.float32)
    try:
        res = models.shia_v1(path + ".bin", path + ".hdr", cal, N, 1, 0, T, sto0, None, None, sp_path, vf_path,
                             None, None, None, None, re_path, None)
    finally:
        models.save_speed = models.save_retorno = models.save_vfluxes = 0
        models.guarda_vfluxes = None
        models.strict_arithmetic = 0
        models.retorno_gr = models.retorno_aq = 0
    recs = {}
    exp = ft.shia_v1(N, T, cal, inp.v_coef, inp.h_coef, inp.h_exp, inp.max_capilar, inp.max_gravita, inp.max_aquifer,
                     inp.hill_long, inp.elem_area, inp.rain_records, inp.pos_evento, inp.evp, inp.dt, (2, 1, 2, 1), 1, 1, 5,
                     StoIn=sto0, save_vfluxes=1, save_speed=1, save_retorno=1, step_records=recs)
    assert os.path.getsize(sp_path) == T * 4 * N * 4 and os.path.getsize(re_path) == T * N * 4
    assert os.path.getsize(vf_path) == 3 * 4 * N * 4
    changed = 0
    for t in range(T):
        v, rc = models.read_float_basin_ncol(sp_path, t + 1, N, 4)
        assert rc == 0
        np.testing.assert_allclose(v, recs["speed"][t], rtol=3e-4, atol=1e-9)
        changed += t > 0 and not np.array_equal(recs["speed"][t], recs["speed"][t - 1])
        r, rc = models.read_float_basin(re_path, t + 1, N)
        assert rc == 0
        np.testing.assert_allclose(r, recs["retorno"][t], rtol=1e-5, atol=ATOL_STRICT)
    assert changed > T // 2                                   # (the kinematic speeds do move)
    assert any(r.any() for r in recs["retorno"])              # (and there is return flow to record)
    for rec, t in ((1, 2), (2, 7), (3, 13)):
        v, rc = models.read_float_basin_ncol(vf_path, rec, N, 4)
        assert rc == 0
        np.testing.assert_allclose(v, recs["vfluxes"][t], rtol=1e-5, atol=ATOL_STRICT)
    np.testing.assert_allclose(res[9], exp["StoOut"], rtol=1e-5, atol=ATOL_STRICT)           # the run's ordinary outputs


def test_shia5_population_many_sets():
    """BASELINE config 5 in small: many parameter sets in one launch (grid.y), each equal to its own oracle run."""
    N, T, P = 3000, 37, 96
    inp = helpers.shia5_inputs(oracle, N, T, 95, speed_type=(1, 1, 1))
    rng = np.random.default_rng(96)
    calibs = rng.uniform(0.5, 2.0, (P, 11)).astype(np.float32)
    sto0 = np.full((5, N), 0.001, np.float32)
    out = _run5(inp, calibs, sto0, show=False)
    got = out["StoOut"].cpu().numpy()
    bal = out["balance"].cpu().numpy()
    for p in (0, 1, 17, 50, 95):
        o = oracle.f_shia_v1(inp, calibs[p], sto0)
        np.testing.assert_allclose(got[p], o["StoOut"], rtol=1e-5, atol=ATOL_FAST)
        np.testing.assert_allclose(bal[p], o["balance64"], rtol=1e-4, atol=1e-6 * float(o["StoOut"].sum()))
    assert np.isfinite(got).all()


def test_tables_edited_in_place_between_runs():
    """A calibration loop scales the module tables IN PLACE (same array object) and rebinds others between runs
    (the f2py idiom, wmfv2.py:1872-1916): every run must see the current values (no stale device copy)."""
    N, T = 1500, 24
    inp = helpers.shia5_inputs(oracle, N, T, 95)
    models.v_coef, models.h_coef, models.h_exp = inp.v_coef, inp.h_coef, inp.h_exp
    models.max_capilar, models.max_gravita, models.max_aquifer = (a[None].copy() for a in (inp.max_capilar, inp.max_gravita, inp.max_aquifer))
    models.hill_long, models.elem_area = inp.hill_long[None].copy(), inp.elem_area[None].copy()
    models.speed_type, models.evpserie, models.dt = np.array([1, 1, 1]), inp.evp, inp.dt
    models.storage, models.storage_constant = None, 0.001
    models.retorno_gr = models.retorno_aq = 0.0
    models.show_storage = models.show_mean_speed = models.save_storage = 0
    cal = np.ones(11, np.float32)
    sto0 = np.full((5, N), 0.001, np.float32)
    a = models.shia_v1_population(cal, inp.rain_records, inp.pos_evento, N, T)["StoOut"][0]
    np.testing.assert_allclose(a, oracle.f_shia_v1(inp, cal, sto0)["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    models.h_coef[0, :] *= 3.0                             # in place: the array object (and its id) stay the same
    models.max_aquifer = models.max_aquifer * 0.5          # rebound: a table the old cache key ignored
    inp.max_aquifer = models.max_aquifer[0]
    b = models.shia_v1_population(cal, inp.rain_records, inp.pos_evento, N, T)["StoOut"][0]
    np.testing.assert_allclose(b, oracle.f_shia_v1(inp, cal, sto0)["StoOut"], rtol=1e-5, atol=ATOL_FAST)
    assert not np.allclose(a[1], b[1], rtol=1e-3)


def test_rain_record_out_of_range_is_rejected():
    """posEvento naming a record beyond the rain file: the Fortran prints 'fuera del rango' and reads garbage
    (Modelos_heavy.f90:580-590); here the host layer raises before any launch."""
    N, T = 300, 10
    inp = helpers.shia5_inputs(oracle, N, T, 96)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    pos = inp.pos_evento.copy()
    for bad in (inp.rain_records.shape[0] + 1, 0, -3):
        pos[4] = bad
        with pytest.raises(ValueError, match="fuera del rango"):
            ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(pos), dev(inp.evp),
                          torch.ones((1, 11), device="cuda"), dev(np.full((1, 5, N), 0.001, np.float32)), None, dt=inp.dt)


@pytest.mark.parametrize("speed_type", [(1, 1, 1), (2, 2, 2), (1, 2, 1)])
def test_shia5_strict_arithmetic(speed_type):
    """`strict=True`: true division and powf, the Fortran's operation sequence -- the tight end of the error table."""
    N, T = 6000, 120
    inp = helpers.shia5_inputs(oracle, N, T, 97, speed_type=speed_type)
    inp.retorno_gr = inp.retorno_aq = 1.0
    sto0 = np.full((5, N), 0.001, np.float32)
    o = oracle.f_shia_v1(inp, np.ones(11, np.float32), sto0)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(inp.pos_evento), dev(inp.evp),
                        torch.ones((1, 11), device="cuda"), dev(sto0[None].copy()), None, dt=inp.dt, speed_type=speed_type,
                        retorno_gr=1.0, retorno_aq=1.0, calc_niter=inp.calc_niter, strict=True)
    np.testing.assert_allclose(out["StoOut"][0].cpu().numpy(), o["StoOut"], rtol=1e-5, atol=ATOL_STRICT)
    np.testing.assert_allclose(out["hspeed"][0].cpu().numpy(), o["hspeed"], rtol=2e-5, atol=1e-9)


FG = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fortran_golden.npz"))


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("name", sorted({k.split("/")[1] for k in FG.files if k.startswith("shia/")}))
def test_shia5_vs_fortran_transliteration(name, strict):
    """The CUDA path against the SECOND restatement of Modelos_heavy.f90 (tests/golden/fortran_transliteration.py, a
    literal 1-based np.float32 transliteration; vectors in fortran_golden.npz), not only against the C oracle."""
    from test_oracle import _fortran_shia_case
    inp, calib, sto, hs = _fortran_shia_case(name)
    N = sto.shape[1]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(inp.pos_evento), dev(inp.evp),
                        dev(calib[None].astype(np.float32)), dev(sto[None].astype(np.float32)),
                        None if hs is None else dev(hs[None].astype(np.float32)), dt=inp.dt, speed_type=inp.speed_type,
                        retorno_gr=inp.retorno_gr, retorno_aq=inp.retorno_aq, calc_niter=inp.calc_niter,
                        show_storage=True, show_mean_speed=True, strict=strict)
    q = f"shia/{name}/out/"
    np.testing.assert_allclose(out["StoOut"][0].cpu().numpy(), FG[q + "StoOut"], rtol=1e-5, atol=ATOL_STRICT if strict else ATOL_FAST)
    np.testing.assert_allclose(out["hspeed"][0].cpu().numpy(), FG[q + "hspeed"], rtol=2e-5 if strict else 3e-4, atol=1e-9)
    np.testing.assert_allclose(out["acum_rain"].cpu().numpy(), FG[q + "Acum_rain"], rtol=1e-6)
    np.testing.assert_allclose(out["mean_rain"].cpu().numpy(), FG[q + "Mean_Rain"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(out["mean_storage"][0].cpu().numpy(), FG[q + "mean_storage"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(out["mean_speed"][0].cpu().numpy(), FG[q + "mean_speed"], rtol=1e-4, atol=1e-9)


def test_shia5_optional_outputs_vs_transliteration():
    """The optional outputs of the shipped shia_v1 -- mean_retorno, mean_vfluxes, the last step's vfluxes, rc_coef
    (Modelos_heavy.f90:424-438, 502-507, 525-531, 543-544) -- against the literal transliteration's vectors (the C oracle
    does not carry them): the case runs with show_mean_retorno, save_vfluxes and save_rc set."""
    from test_oracle import _fortran_shia_case
    inp, calib, sto, hs = _fortran_shia_case("mixed_flags")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q = "shia/mixed_flags/out/"
    for strict, atol in ((True, ATOL_STRICT), (False, ATOL_FAST)):
        out = ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(inp.pos_evento), dev(inp.evp),
                            dev(calib[None].astype(np.float32)), dev(sto[None].astype(np.float32)), None, dt=inp.dt,
                            speed_type=inp.speed_type, retorno_gr=inp.retorno_gr, retorno_aq=inp.retorno_aq,
                            calc_niter=inp.calc_niter, strict=strict, show_mean_retorno=True, save_vfluxes=True, save_rc=True)
        np.testing.assert_allclose(out["StoOut"][0].cpu().numpy(), FG[q + "StoOut"], rtol=1e-5, atol=atol)
        np.testing.assert_allclose(out["mean_retorno"][0].cpu().numpy(), FG[q + "mean_retorno"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["mean_vfluxes"][0].cpu().numpy(), FG[q + "mean_vfluxes"], rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(out["vfluxes"][0].cpu().numpy(), FG[q + "vfluxes"], rtol=1e-5, atol=atol)
        np.testing.assert_allclose(out["rc_coef"][0].cpu().numpy(), FG[q + "rc_coef"], rtol=2e-5, atol=10 * atol)
        assert not out["retorned"].any()                     # cleared after every step when the mean is shown
    acc = ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(inp.pos_evento), dev(inp.evp),
                        dev(calib[None].astype(np.float32)), dev(sto[None].astype(np.float32)), None, dt=inp.dt,
                        speed_type=inp.speed_type, retorno_gr=inp.retorno_gr, retorno_aq=inp.retorno_aq,
                        calc_niter=inp.calc_niter, with_retorned=True)
    T = inp.pos_evento.shape[0]
    np.testing.assert_allclose(float(acc["retorned"].sum()) / sto.shape[1], float(FG[q + "mean_retorno"].astype(np.float64).sum()),
                               rtol=1e-4)                    # accumulated over the run = the per-step means added up
