"""Golden vectors for the N1/N2 rows (streams / HAND / HDND) from the UNMODIFIED Python reference.
Run in the build container only:  python tests/golden/make_golden_n2.py  -> tests/golden/wmf_py_golden_n2.npz
Reference functions: wmf_py/cu_py/metrics.py:189-268 hydro_distance_and_receiver, :271-323 hand;
wmf_py/cu_py/streams.py:19-83 stream_find / stream_cut / stream_find_to_corr; basics.py:124-177 basin_acum."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from wmf_py.cu_py.basics import basin_acum  # noqa: E402
from wmf_py.cu_py.metrics import hand, hydro_distance_and_receiver  # noqa: E402
from wmf_py.cu_py.streams import stream_cut, stream_find, stream_find_to_corr  # noqa: E402
import helpers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
rng = np.random.default_rng(77)
cases = {
    "steep_40x31": (helpers.steepest(40, 31, 21, pit_frac=0.01), None, 30.0, 30.0),
    "steep_mask_33x45": (helpers.steepest(33, 45, 22, pit_frac=0.02), rng.random((33, 45)) > 0.06, 12.5, 20.0),
    "random_cycles_21x26": (helpers.random_codes(21, 26, 23), rng.random((21, 26)) > 0.1, 1.0, 2.0),
    "sch_30x30": (np.random.default_rng(24).choice(np.array([0, 1, 2], np.int8), size=(30, 30), p=[0.25, 0.5, 0.25]), None, 30.0, 30.0),
}
for name, (fd, mask, dx, dy) in cases.items():
    if mask is None:
        mask = np.ones(fd.shape, bool)
    acc = basin_acum(fd.astype(np.int64), mask)
    thr = max(3, int(np.sort(acc.ravel())[-max(4, acc.size // 12)]))
    stream = stream_find(acc, thr)
    if name.startswith("random"):                      # make sure some cycle cells are stream cells too
        stream = (stream.astype(bool) | (rng.random(fd.shape) < 0.08)).astype(np.uint8)
    stream = stream_cut(stream, mask)
    dem = 500.0 - 0.4 * np.add.outer(np.arange(fd.shape[0]), np.arange(fd.shape[1])) + rng.random(fd.shape)
    hd, aq = hydro_distance_and_receiver(fd.astype(np.int64), stream.astype(bool), mask, dx, dy)
    h = hand(dem, fd.astype(np.int64), stream.astype(bool), mask)
    r0, c0 = np.unravel_index(np.argmax(acc), acc.shape)
    corr = stream_find_to_corr(stream, fd.astype(np.int64), (int(r0), int(c0)))
    for k, v in (("flowdir", fd), ("mask", mask), ("stream", stream), ("dem", dem), ("dxdy", np.array([dx, dy])),
                 ("hdnd", hd), ("aquien", aq.astype(np.int64)), ("hand", h), ("thr", np.array(thr)),
                 ("outlet", np.array([r0, c0])), ("corr", corr), ("acc", acc.astype(np.int32))):
        out[f"{name}/{k}"] = v
np.savez_compressed(os.path.join(HERE, "wmf_py_golden_n2.npz"), **out)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "wmf_py_golden_n2.npz")), "bytes")
