"""Basin delineation parity: raster mask (wmf_py semantics), BFS-ordered structure, vector accumulation,
cell basics and stream/node marks (Fortran semantics) -- all bit-exact for integer outputs."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from conftest import golden_names
from wmf_heavy_b200 import _lib, cu, ops, wmfv2
from wmf_heavy_b200.cu_py import basics

pytestmark = pytest.mark.gpu


def test_cut_reference_counts():               # wmf_py/tests/test_cu_basics.py:61-90
    fd = np.zeros((7, 7), dtype=int); fd[:, -1] = 2; fd[-1, -1] = -1
    m = np.ones_like(fd, dtype=bool); dem = np.ones_like(fd, dtype=float)
    assert basics.basin_cut(dem, (6, 6), fd, m)["mask"].sum() == 49
    fd = np.zeros((7, 7), dtype=int); fd[:3, :] = 6; fd[3:, -1] = 2; fd[-1, -1] = -1
    assert basics.basin_cut(dem, (6, 6), fd, m)["mask"].sum() == 28
    fd = np.full((5, 5), 2, dtype=int); fd[4, :] = 0; fd[4, 2:] = 4; fd[4, 2] = -1
    assert basics.basin_cut(np.ones((5, 5)), (4, 2), fd, np.ones((5, 5), bool))["mask"].sum() == 25
    r = basics.basin_cut(np.ones((1, 5)), (0, 4), np.array([[0, 0, 8, 0, 0]]), np.ones((1, 5), bool))
    assert np.array_equal(r["mask"], [[False, False, False, True, True]])
    assert np.array_equal(r["flowdir"], np.zeros((1, 5))) and np.array_equal(r["dem"], [[0, 0, 0, 1, 1]])
    with pytest.raises(ValueError, match="outlet outside grid"):
        basics.basin_cut(dem, (7, 0), np.zeros((7, 7), int), m)
    m2 = m.copy(); m2[0, 0] = False
    with pytest.raises(ValueError, match="outlet not in mask"):
        basics.basin_cut(dem, (0, 0), np.zeros((7, 7), int), m2)


@pytest.mark.parametrize("name", golden_names("cut"))
def test_cut_golden(golden, name):
    fd, mask = golden[f"acum/{name}/flowdir"], golden[f"acum/{name}/mask"]
    dem = np.arange(fd.size, dtype=np.float64).reshape(fd.shape) + 1.0
    r = basics.basin_cut(dem, tuple(golden[f"cut/{name}/outlet"]), fd.astype(np.int64), mask)
    assert np.array_equal(r["mask"], golden[f"cut/{name}/mask"])
    assert np.array_equal(r["flowdir"], golden[f"cut/{name}/flowdir"])
    assert np.array_equal(r["dem"], golden[f"cut/{name}/dem"])


@pytest.mark.parametrize("case", ["steep", "random", "sch"])
def test_cut_vs_oracle(case):
    if case == "steep":
        fd, mask = helpers.steepest(600, 450, 21), np.random.default_rng(21).random((600, 450)) > 0.03
    elif case == "random":
        fd, mask = helpers.random_codes(200, 300, 22), np.random.default_rng(22).random((200, 300)) > 0.1
    else:
        fd, mask = ops.synth_scheidegger_host(800, 700, seed=23), np.ones((800, 700), bool)
    acc = oracle.basin_acum(fd, mask)
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    dem = np.random.default_rng(1).random(fd.shape)
    exp = oracle.basin_cut(dem, (r0, c0), fd.astype(np.int64), mask)
    got = basics.basin_cut(dem, (int(r0), int(c0)), fd, mask)
    for k in ("mask", "flowdir", "dem"):
        assert np.array_equal(got[k], exp[k]), k


@pytest.mark.parametrize("shape,kind,seed", [((64, 256), "steep", 41), ((65, 257), "random", 42), ((300, 517), "steep", 43),
                                             ((129, 1000), "random", 44), ((700, 300), "sch", 45), ((1, 700), "random", 46),
                                             ((513, 1), "random", 47)])
def test_cut_many_outlets(shape, kind, seed):
    """The tile contraction behind basin_cut on ragged rasters: outlets on tile edges and corners, on the raster's rim,
    inside cycles (random codes), at the largest accumulation and at random active cells -- mask identical to the oracle."""
    ny, nx = shape
    rng = np.random.default_rng(seed)
    if kind == "steep":
        fd = helpers.steepest(ny, nx, seed)
    elif kind == "random":
        fd = helpers.random_codes(ny, nx, seed)
    else:
        fd = ops.synth_scheidegger_host(ny, nx, seed=seed)
    mask = rng.random((ny, nx)) > 0.05
    acc = oracle.basin_acum(fd, mask)
    cand = [np.unravel_index(np.argmax(acc), acc.shape)]
    for r in (0, 63, 64, ny - 1):
        for c in (0, 255, 256, nx - 1):
            if 0 <= r < ny and 0 <= c < nx:
                cand.append((r, c))
    cand += [(int(rng.integers(ny)), int(rng.integers(nx))) for _ in range(6)]
    dem = rng.random(fd.shape)
    done = 0
    for r0, c0 in cand:
        if not mask[r0, c0]:
            continue
        exp = oracle.basin_cut(dem, (int(r0), int(c0)), fd.astype(np.int64), mask)
        got = basics.basin_cut(dem, (int(r0), int(c0)), fd, mask)
        assert np.array_equal(got["mask"], exp["mask"]), (shape, kind, r0, c0, int(exp["mask"].sum()), int(got["mask"].sum()))
        done += 1
    assert done >= 3


def test_cut_serpentine():
    """One path through every cell of a 1000 x 1030 raster (rows alternate east / west): the chain of tile exits is
    ~5000 long, the outlet is its last cell and the basin is the whole raster; from a cell half way, half of it."""
    ny, nx = 1000, 1030
    fd = helpers.serpentine(ny, nx)
    mask = np.ones((ny, nx), bool)
    dem = np.zeros((ny, nx))
    last = ny - 1
    out_rc = (last, nx - 1 if last % 2 == 0 else 0)
    got = basics.basin_cut(dem, out_rc, fd, mask)
    assert got["mask"].all()
    got = basics.basin_cut(dem, (500, 0), fd, mask)       # row 500 runs east: (500, 0) is its first cell
    assert int(got["mask"].sum()) == 500 * nx + 1 and got["mask"][:500].all() and not got["mask"][501:].any()


def _basin_case(ny, nx, seed, kind="steep"):
    fd = helpers.steepest(ny, nx, seed) if kind == "steep" else ops.synth_scheidegger_host(ny, nx, seed=seed)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    return helpers.to_keypad(fd), int(c0) + 1, int(r0) + 1


def test_structure_derived_example():          # SURVEY 8(a) D3/D4 [derived]
    d = torch.tensor([[3, 2, 1], [6, 3, 2], [6, 6, 0]], dtype=torch.int8, device="cuda")
    st = ops.basin_structure(d, 3, 3)
    assert st.cpu().tolist() == [[4, 2, 2, 2, 2, 1, 1, 1, 0], [1, 1, 3, 2, 1, 2, 3, 2, 3], [3, 2, 1, 1, 1, 3, 2, 2, 3]]
    assert ops.basin_acum_vec(st).cpu().tolist() == [1, 1, 1, 1, 1, 2, 1, 5, 9]


@pytest.mark.parametrize("shape,seed,kind", [((64, 80), 31, "steep"), ((500, 731), 32, "steep"),
                                             ((900, 900), 33, "sch"), ((1, 200), 34, "steep"), ((333, 2), 35, "steep")])
def test_structure_vs_oracle(shape, seed, kind):
    key, col, fil = _basin_case(*shape, seed, kind)
    exp = oracle.f_basin_find(key, col, fil)
    got = ops.basin_structure(torch.from_numpy(key).cuda(), col, fil).cpu().numpy()
    assert got.shape == exp.shape and np.array_equal(got, exp)
    n = exp.shape[1]
    st = torch.from_numpy(exp).cuda()
    # D4: vector accumulation + basics
    dem = (2000 - 0.5 * (exp[1] + exp[2]) + np.random.default_rng(seed).random(n)).astype(np.float32)
    dirv = key[exp[2] - 1, exp[1] - 1].astype(np.int32)
    e_ac, e_lon, e_pend, e_elev, e_sc = oracle.f_basin_basics(exp, dem, dirv, 30.0, 100.0, 200.0, 30.0, shape[0])
    g_ac, g_lon, g_pend, g_elev, g_sc = ops.basin_basics(st, torch.from_numpy(dem).cuda(), torch.from_numpy(dirv).cuda(),
                                                        30.0, 100.0, 200.0, 30.0, shape[0])
    assert np.array_equal(g_ac.cpu().numpy(), e_ac) and np.array_equal(ops.basin_acum_vec(st).cpu().numpy(), e_ac)
    assert np.array_equal(g_lon.cpu().numpy(), e_lon) and np.array_equal(g_elev.cpu().numpy(), e_elev)
    assert np.array_equal(g_pend.cpu().numpy(), e_pend)          # same fp32 operations, same rounding
    assert g_sc[0] == e_sc[0] and g_sc[3] == e_sc[3] and g_sc[4] == e_sc[4]
    # fp32 running sums in the Fortran vs an fp64 tree here: tolerance 1e-5 relative
    np.testing.assert_allclose(g_sc[1:3], e_sc[1:3], rtol=1e-5)
    X, Y = ops.basin_coordxy(st, 100.0, 200.0, 30.0, shape[0])
    eX, eY = oracle.f_basin_coordxy(exp, 100.0, 200.0, 30.0, shape[0])
    assert np.array_equal(X.cpu().numpy(), eX) and np.array_equal(Y.cpu().numpy(), eY)
    # D5: stream / node marks at two thresholds
    for umbral in (2, max(3, n // 50)):
        e = oracle.f_stream_nod(exp, e_ac, umbral)
        g = ops.stream_nod(st, g_ac, umbral)
        for a, b in zip(g[:3], e[:3]):
            assert np.array_equal(a.cpu().numpy(), b)
        assert g[3:] == e[3:]
    # map <-> basin
    raster = np.random.default_rng(2).random(shape).astype(np.float32)
    vec = ops.basin_map2basin(st, torch.from_numpy(raster).cuda())
    assert np.array_equal(vec.cpu().numpy(), raster[exp[2] - 1, exp[1] - 1])
    back = ops.basin_2map(st, vec, shape[0], shape[1], -9999.0).cpu().numpy()
    inb = np.zeros(shape, bool); inb[exp[2] - 1, exp[1] - 1] = True
    assert np.array_equal(back[inb], raster[inb]) and np.all(back[~inb] == -9999.0)


def test_structure_errors():
    d = torch.tensor([[6, 6, 5]], dtype=torch.int8, device="cuda")
    with pytest.raises(_lib.WmfError) as e:
        ops.basin_structure(d, 3, 1)
    assert e.value.code == _lib.E_CODE5
    d = torch.tensor([[6, 4, 4]], dtype=torch.int8, device="cuda")     # 2-cycle through the outlet
    with pytest.raises(_lib.WmfError) as e:
        ops.basin_structure(d, 2, 1)
    assert e.value.code == _lib.E_CYCLE
    with pytest.raises(_lib.WmfError) as e:
        ops.basin_structure(d, 9, 1)
    assert e.value.code == _lib.E_OUTLET_RANGE


def test_basin_object_api():
    """The reference's own call expressions (wmfv2.py:823-828, 976, 1868) through Basin / SimuBasin."""
    ny, nx = 300, 260
    key, col, fil = _basin_case(ny, nx, 41)
    wmfv2.set_map_properties(ncols=nx, nrows=ny, xll=1000.0, yll=5000.0, dx=30.0, dxp=30.0)
    x = 1000.0 + 30.0 * (col - 0.5)
    y = 5000.0 + 30.0 * ((ny - fil) + 0.5)
    assert ops.coord2fil_col(x, y, 1000.0, 5000.0, 30.0, nx, ny) == (col, fil)
    DIR = np.asfortranarray(key.T.astype(np.int32))            # (ncols, nrows) like read_map_raster
    DEM = np.asfortranarray(np.random.default_rng(3).random((nx, ny)).astype(np.float32) + 100)
    b = wmfv2.Basin(x, y, DEM, DIR, threshold=20)
    exp = oracle.f_basin_find(key, col, fil)
    assert b.ncells == exp.shape[1] and np.array_equal(b.structure, exp) and b.structure.flags["F_CONTIGUOUS"]
    assert np.array_equal(b.DIRvec, key[exp[2] - 1, exp[1] - 1])
    assert np.array_equal(b.DEMvec, DEM.T[exp[2] - 1, exp[1] - 1])
    b.GetGeo_Cell_Basics()
    assert np.array_equal(b.CellAcum, oracle.f_basin_acum(exp))
    assert b.CellCauce.sum() == (oracle.f_basin_acum(exp) > 20).sum()
    M, props = b.Transform_Basin2Map(b.CellAcum)
    assert M.shape == (nx, ny) and M[col - 1, fil - 1] == b.ncells


def test_dir_edited_in_place_between_calls():
    """cu.basin_find must read the caller's DIR as it is NOW: an in-place edit between two calls (same array
    object) changes the basin (a stale device copy keyed by id(DIR) would not)."""
    ny, nx = 90, 70
    key, col, fil = _basin_case(ny, nx, 43)
    wmfv2.set_map_properties(ncols=nx, nrows=ny, xll=0.0, yll=0.0, dx=30.0, dxp=30.0)
    x, y = 30.0 * (col - 0.5), 30.0 * ((ny - fil) + 0.5)
    DIR = np.asfortranarray(key.T.astype(np.int32))
    n1 = cu.basin_find(x, y, DIR, nx, ny)
    st1 = cu.basin_cut(n1)
    assert np.array_equal(st1, oracle.f_basin_find(key, col, fil))
    donors = np.nonzero(st1[0] == 1)[0]                        # cells draining straight into the outlet
    cut = donors[donors != n1 - 1][0]
    c, f = int(st1[1, cut]), int(st1[2, cut])
    DIR[c - 1, f - 1] = 5                                      # in place: that cell (and its subtree) leaves the basin
    key2 = key.copy()
    key2[f - 1, c - 1] = 5
    n2 = cu.basin_find(x, y, DIR, nx, ny)
    assert n2 < n1 and np.array_equal(cu.basin_cut(n2), oracle.f_basin_find(key2, col, fil))
