"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
host-only entry points agree with the oracle, and the host logic of the shims behaves like the
reference (error messages, formats)."""
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT
from wmf_heavy_b200 import _lib, models, ops
from wmf_heavy_b200.cu_py import basics


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wmf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(wmf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wmf_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.wmf_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        basics.basin_acum(np.zeros((3, 3), int), np.ones((3, 3), bool))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "wmf_heavy_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "wmforacle" not in src and "orc_" not in src, f


def test_coord2fil_col_matches_oracle():
    rng = np.random.default_rng(0)
    for _ in range(200):
        x, y = rng.uniform(-50, 400, 2)
        args = (x, y, 3.25, -7.5, 30.0, 11, 9)
        assert ops.coord2fil_col(*args) == oracle.f_coord2fil_col(*args)
    assert ops.coord2fil_col(75.0, 15.0, 0, 0, 30.0, 3, 3) == (3, 3)


def test_basin_find_py():                      # wmf_py/tests/test_cu_basics.py:14-16 + golden
    assert basics.basin_find(x=60, y=60, xll=0, yll=0, dx=30, dy=30, ncols=3, nrows=3) == (2, 2)
    g = np.load(os.path.join(ROOT, "tests", "golden", "wmf_py_golden.npz"))
    for a, rc in zip(g["find/args"], g["find/rc"]):
        assert basics.basin_find(*a[:6], int(a[6]), int(a[7])) == tuple(rc)
    with pytest.raises(ValueError, match="point outside raster bounds"):
        basics.basin_find(1e9, 0, 0, 0, 30, 30, 3, 3)


def test_scheidegger_host_is_deterministic_and_windowed():
    a = ops.synth_scheidegger_host(64, 96, seed=7)
    assert set(np.unique(a)) <= {0, 1, 2}
    frac = [(a == k).mean() for k in range(3)]
    assert abs(frac[1] - 0.5) < 0.05 and abs(frac[0] - 0.25) < 0.05
    w = ops.synth_scheidegger_host(10, 20, seed=7, row0=5, col0=30, nx_total=96)
    assert np.array_equal(w, a[5:15, 30:50])
    k = ops.synth_scheidegger_host(64, 96, seed=7, encoding=_lib.ENC_KEYPAD)
    assert np.array_equal(k, np.array([6, 3, 2], dtype=np.int8)[a])


def test_basin_record_files(tmp_path):        # N3: Modelos_heavy.f90:552-624 (direct access, RECL = 4 * N_col * N_cel)
    N = 11
    f = str(tmp_path / "sto.StObin")
    a = np.arange(5 * N, dtype=np.float32).reshape(5, N)
    with pytest.raises(FileNotFoundError):
        models.write_float_basin(f, a, 2, N, 5)              # status='old' for records other than 1
    models.write_float_basin(f, a, 0, N, 5)                  # record <= 0: nothing written (f90:603)
    assert not os.path.exists(f)
    models.write_float_basin(f, a, 1, N, 5)
    models.write_float_basin(f, 2 * a, 3, N, 5)              # leaves record 2 as a hole
    assert os.path.getsize(f) == 3 * 5 * N * 4
    raw = np.fromfile(f, dtype="<f4")
    assert np.array_equal(raw[:5], a[:, 0]) and np.array_equal(raw[5:10], a[:, 1])   # N_col values per cell
    v, res = models.read_float_basin_ncol(f, 3, N, 5)
    assert res == 0 and v.shape == (5, N) and np.array_equal(v, 2 * a)
    v, res = models.read_float_basin_ncol(f, 4, N, 5)
    assert res != 0                                          # iostat /= 0 past the end
    v1, res = models.read_float_basin(f, 2, N)               # same file read as N_col = 1 records
    assert res == 0 and np.array_equal(v1, raw[N:2 * N])
    models.write_float_basin(f, a, 1, N, 5)                  # record 1 replaces the file
    assert os.path.getsize(f) == 5 * N * 4
    g = str(tmp_path / "ints.bin")
    b = (np.arange(2 * N) * 7).astype(np.int32).reshape(2, N)
    models.write_int_basin(g, b, 1, N, 2)
    assert np.array_equal(np.fromfile(g, dtype="<i4").reshape(N, 2).T, b)
    assert np.array_equal(models.read_int_basin(g, 1, 2 * N), np.ascontiguousarray(b.T).reshape(-1))


def test_rain_file_roundtrip(tmp_path):        # Modelos_heavy.f90:580-590, 631-663
    N, T = 37, 12
    rng = np.random.default_rng(1)
    rec = np.vstack([np.zeros((1, N), np.int32), rng.integers(0, 5000, (4, N)).astype(np.int32)])
    pos = rng.integers(1, 6, T).astype(np.int32)
    b, h = str(tmp_path / "rain.bin"), str(tmp_path / "rain.hdr")
    models.write_rain(b, h, rec, pos)
    models.rain_first_point = 1
    ids, p2 = models.rain_read_ascii_table(h, T)
    assert np.array_equal(p2, pos) and np.array_equal(ids, np.arange(1, T + 1))
    assert np.array_equal(models.read_int_basin(b, 3, N), rec[2])
    assert np.array_equal(models.read_rain_records(b, N), rec)
    with pytest.raises(ValueError):
        models.rain_read_ascii_table(h, T + 5)


def test_route_topology_host():                # N4 extension, host part: depth / rank / donor lists from the drain ids
    from wmf_heavy_b200 import ops
    fd = ops.synth_scheidegger_host(60, 70, seed=12)
    acc = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    import helpers
    drain = oracle.f_basin_find(helpers.to_keypad(fd), int(c0) + 1, int(r0) + 1)[0].astype(np.int32)
    n = drain.size
    rank, dmax, ptr, idx = ops.route_topology(drain)
    recv = np.where(drain != 0, n - drain, -1)
    depth = np.zeros(n, int)
    for i in range(n - 2, -1, -1):             # receivers come later in the upstream-first order
        depth[i] = depth[recv[i]] + 1
    assert dmax == depth.max() and np.array_equal(rank, dmax - depth)
    assert ptr[0] == 0 and ptr[-1] == n - 1 and np.all(np.diff(ptr) >= 0)
    for j in range(n):
        assert list(idx[ptr[j]:ptr[j + 1]]) == list(np.nonzero(recv == j)[0])      # ascending donor indices
    with pytest.raises(ValueError):
        ops.route_topology(np.array([3, 0, 0], np.int32))


def test_nsga2_selection_helpers():
    """The numpy stand-ins for DEAP's selNSGA2 / selTournamentDCD behind Calib_NSGAII: non-dominated fronts, crowding."""
    from wmf_heavy_b200 import wmfv2
    w = (1.0, -1.0)                                      # maximise the first objective, minimise the second
    vals = [(0.9, 0.1), (0.8, 0.05), (0.7, 0.2), (0.5, 0.5), (0.95, 0.3), (0.2, 0.9)]
    pop = []
    for v in vals:
        ind = wmfv2.Individual([[np.array([0.0])] * 10])
        ind.fitness.values = v
        pop.append(ind)
    fronts = wmfv2._fronts(pop, w)
    assert sorted(i.fitness.values for i in fronts[0]) == [(0.8, 0.05), (0.9, 0.1), (0.95, 0.3)]
    assert [i.fitness.values for i in fronts[-1]] == [(0.2, 0.9)]
    sel = wmfv2._sel_nsga2(pop, 4, w)
    assert len(sel) == 4 and {(0.9, 0.1), (0.8, 0.05), (0.95, 0.3), (0.7, 0.2)} == {i.fitness.values for i in sel}
    np.random.seed(0)
    t = wmfv2._sel_tournament_dcd(pop[:4], 4, w)
    assert len(t) == 4 and all(i in pop[:4] for i in t)
