"""T2 (direction reclass) and N3 (basin <-> raster by linear index): the device kernels behind
wmf_heavy_b200.cu_py.basics.dir_reclass_* / basin_2map / basin_map2basin against fixtures generated from the
unmodified reference (tests/golden/make_golden_n3.py -> wmf_py_golden_n3.npz), bit-exact, plus the reference's own
tests (wmf_py/tests/test_cu_basics.py:19-32, test_transform.py:15-63) and the Fortran tables
(cuencas_heavy.f90:1086-1119) through the f2py-style `cu` namespace."""
import os

import numpy as np
import pytest

from conftest import ROOT
from wmf_heavy_b200 import cu
from wmf_heavy_b200.cu_py import basics

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "wmf_py_golden_n3.npz"))
names = lambda pre: sorted({k.split("/")[1] for k in G.files if k.startswith(pre + "/")})


@pytest.mark.parametrize("name", names("reclass"))
def test_reclass_golden(name):
    src = G[f"reclass/{name}/in"]
    for conv, fn in (("opentopo", basics.dir_reclass_opentopo), ("rwatershed", basics.dir_reclass_rwatershed)):
        keep = src.copy()
        res = fn(src)
        exp = G[f"reclass/{name}/{conv}"]
        assert res.dtype == exp.dtype and res.shape == exp.shape and np.array_equal(res, exp)
        assert np.array_equal(src, keep)                       # inputs are not mutated


@pytest.mark.parametrize("name", names("reclass_err"))
def test_reclass_unknown_codes(name):
    src = G[f"reclass_err/{name}/in"]
    with pytest.raises(ValueError) as e:
        basics.dir_reclass_opentopo(src)
    assert str(e.value) == str(G[f"reclass_err/{name}/msg"])
    with pytest.raises(ValueError):
        basics.dir_reclass_rwatershed(src)


def test_reclass_like_reference():             # wmf_py/tests/test_cu_basics.py:19-32
    src = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 1]])
    res = basics.dir_reclass_opentopo(src)
    assert set(np.unique(res)) <= set(range(8))
    assert np.array_equal(res, basics.dir_reclass_opentopo(res))
    with pytest.raises(ValueError):
        basics.dir_reclass_opentopo(np.array([[9]]))
    res = basics.dir_reclass_rwatershed(np.arange(1, 9).reshape(2, 4))
    assert np.array_equal(res, np.array([0, 7, 6, 5, 4, 3, 2, 1]).reshape(2, 4))
    assert basics.dir_reclass_opentopo(np.zeros((0, 3), dtype=np.int16)).shape == (0, 3)


def test_reclass_fortran_tables():             # cuencas_heavy.f90:1086-1119 (Mat_out = noData, eight where's)
    cu.nodata = -9999.0
    src = np.asfortranarray(np.array([[1, 2, 3, 4], [5, 6, 7, 8], [0, 9, -1, 300]], dtype=np.int32))
    ot = cu.dir_reclass_opentopo(src, 3, 4)
    rw = cu.dir_reclass_rwatershed(src, 3, 4)
    assert ot.dtype == np.int32 and ot.flags["F_CONTIGUOUS"] and ot.shape == src.shape
    assert np.array_equal(ot, [[6, 9, 8, 7], [4, 1, 2, 3], [-9999, -9999, -9999, -9999]])
    assert np.array_equal(rw, [[9, 8, 7, 4], [1, 2, 3, 6], [-9999, -9999, -9999, -9999]])
    # (the two reference flavours disagree on the sense of rotation of "opentopo": the Fortran table reads 2 as NE,
    # wmf_py reads it as SE -- each shim follows its own reference)
    assert cu.dir_reclass_rwatershed(np.arange(1, 9)).tolist() == [9, 8, 7, 4, 1, 2, 3, 6]


@pytest.mark.parametrize("name", names("map"))
def test_2map_golden(name):
    vals, idx = G[f"map/{name}/vals"], G[f"map/{name}/idx"].copy()
    shape, nodata = tuple(int(v) for v in G[f"map/{name}/shape"]), G[f"map/{name}/nodata"][()]
    m = basics.basin_2map(vals, idx, shape, nodata)
    exp = G[f"map/{name}/raster"]
    assert m.dtype == exp.dtype and m.shape == exp.shape and np.array_equal(m, exp)
    assert np.array_equal(idx, G[f"map/{name}/idx_after"])    # repaired indices are written back, like the reference
    back = basics.basin_map2basin(m, idx, vals.shape)
    assert back.dtype == G[f"map/{name}/back"].dtype and np.array_equal(back, G[f"map/{name}/back"])


def test_transform_like_reference():           # wmf_py/tests/test_transform.py:15-63
    idx = np.array([0, 10, 101, 555, 999, 1234, 2345, 3456, 4321, 4999])
    values = np.arange(idx.size)
    raster = basics.basin_2map(values, idx, (50, 50), -1.0)
    assert raster.shape == (50, 50)
    assert np.all(raster.ravel()[idx] == values) and (raster != -1.0).sum() == values.size
    assert np.array_equal(basics.basin_map2basin(raster, idx, (idx.size,)), values)
    idx = np.array([0, 2, 4, 6])
    values = np.array([[1, 2], [3, 4]])
    m = basics.basin_2map(values, idx, (3, 3), -9999.0)
    assert np.array_equal(basics.basin_map2basin(m, idx, values.shape), values)
    mask = np.indices((5, 5)).sum(axis=0) % 2 == 0
    idx = np.flatnonzero(mask.ravel(order="C"))
    out = basics.basin_2map(np.arange(idx.size, dtype=float), idx, (5, 5), -999.0)
    assert np.array_equal(out.ravel()[idx], np.arange(idx.size)) and np.all(out[~mask] == -999.0)
    # 4 is outside a 2x2 grid: the reference's repair reads it as row 0, col 4 % 2 = 0 -> a duplicate of index 0, and the
    # last value wins (measured on the reference: [[3, 2], [0, 0]], idx rewritten to [0, 1, 0])
    idx = np.array([0, 1, 4])
    assert np.array_equal(basics.basin_2map(np.array([1, 2, 3]), idx, (2, 2), 0.0), [[3, 2], [0, 0]]) and idx.tolist() == [0, 1, 0]
    with pytest.raises(ValueError, match="indices out of range"):
        basics.basin_2map(np.array([1, 2]), np.array([0, -450]), (2, 2), 0.0)
    with pytest.raises(ValueError, match="size of values and indices must match"):
        basics.basin_2map(np.array([1, 2, 3]), np.array([0, 1]), (2, 2), 0.0)
    with pytest.raises(ValueError, match="indices out of range"):
        basics.basin_map2basin(np.zeros((2, 2)), np.array([0, 7]), (2,))
