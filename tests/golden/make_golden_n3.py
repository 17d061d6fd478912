"""Golden vectors for the T2 (reclass) and N3 (basin <-> raster by linear index) rows from the UNMODIFIED Python
reference.  Run in the build container only:  python tests/golden/make_golden_n3.py -> tests/golden/wmf_py_golden_n3.npz
Reference functions: wmf_py/cu_py/basics.py:73-121 (_reclass, dir_reclass_opentopo / _rwatershed), :251-301
(basin_2map incl. its "row*100 + col" index repair and the write-back into the caller's index array, basin_map2basin)."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from wmf_py.cu_py.basics import basin_2map, basin_map2basin, dir_reclass_opentopo, dir_reclass_rwatershed  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(303)
out = {}

# ---- reclass: external 1..8 rasters, mixed with 0 and negatives (nodata), already-internal rasters, error lists
rc = {
    "ext_37x53": rng.integers(1, 9, size=(37, 53)),
    "ext_nodata_40x40": np.where(rng.random((40, 40)) < 0.1, -9999, rng.integers(1, 9, size=(40, 40))),
    "ext_with0_16x16": rng.integers(0, 9, size=(16, 16)),
    "internal_20x30": rng.integers(0, 8, size=(20, 30)),
    "int8_neg_9x9": np.where(rng.random((9, 9)) < 0.2, -1, rng.integers(1, 9, size=(9, 9))).astype(np.int8),
    "int32_25x7": rng.integers(1, 9, size=(25, 7)).astype(np.int32),
}
for name, a in rc.items():
    out[f"reclass/{name}/in"] = a
    out[f"reclass/{name}/opentopo"] = dir_reclass_opentopo(a.copy())
    out[f"reclass/{name}/rwatershed"] = dir_reclass_rwatershed(a.copy())
bad = {"bad_9": np.array([[1, 2], [9, 3]]), "bad_many": np.array([[1, 200, 17], [9, -5, 300]]),
       "bad_int8": np.array([[100, 2, 127]], dtype=np.int8)}
for name, a in bad.items():
    out[f"reclass_err/{name}/in"] = a
    try:
        dir_reclass_opentopo(a.copy())
        raise SystemExit("expected a ValueError")
    except ValueError as e:
        out[f"reclass_err/{name}/msg"] = np.array(str(e))

# ---- basin_2map / basin_map2basin
cases = {}
idx = np.sort(rng.choice(60 * 47, size=900, replace=False))
cases["f64_60x47"] = (rng.random(900), idx, (60, 47), -9999.0)
cases["i64_50x50"] = (np.arange(10), np.array([0, 10, 101, 555, 999, 1234, 2345, 3456, 4321, 4999]), (50, 50), -1.0)
cases["f32_dups_12x12"] = (rng.random(40).astype(np.float32), rng.integers(0, 144, size=40), (12, 12), -1.0)
cases["i32_2d_vals_3x3"] = (np.array([[1, 2], [3, 4]], dtype=np.int32), np.array([0, 2, 4, 6]), (3, 3), -9999.0)
# indices written as row*100 + col (some beyond the raster): the reference repairs them and writes them back
r, c = rng.integers(0, 30, size=25), rng.integers(0, 80, size=25)
cases["repair_30x80"] = (rng.random(25), r * 100 + c, (30, 80), -1.0)
cases["u8_8x8"] = (rng.integers(0, 255, size=20).astype(np.uint8), np.sort(rng.choice(64, size=20, replace=False)), (8, 8), 0)
for name, (vals, idx, shape, nodata) in cases.items():
    idx_work = idx.copy()
    m = basin_2map(vals, idx_work, shape, nodata)
    out[f"map/{name}/vals"] = vals
    out[f"map/{name}/idx"] = idx
    out[f"map/{name}/shape"] = np.array(shape)
    out[f"map/{name}/nodata"] = np.array(nodata)
    out[f"map/{name}/raster"] = m
    out[f"map/{name}/idx_after"] = idx_work
    out[f"map/{name}/back"] = basin_map2basin(m, idx_work, vals.shape)
np.savez_compressed(os.path.join(HERE, "wmf_py_golden_n3.npz"), **out)
print(len(out), "arrays")
