"""Golden vectors for the N1/N2 rows (streams / HAND / HDND) from the UNMODIFIED Python reference.
Run in the build container only:  python tests/golden/make_golden_n2.py  -> tests/golden/wmf_py_golden_n2.npz
Reference functions: wmf_py/cu_py/metrics.py:189-268 hydro_distance_and_receiver, :271-323 hand;
wmf_py/cu_py/streams.py:19-83 stream_find / stream_cut / stream_find_to_corr, :282-384 basin_netxy_find / _cut;
metrics.py:337-490 basin_ppalstream_find / _cut, :528-599 basin_time_to_out; basics.py:124-177 basin_acum."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from wmf_py.cu_py.basics import basin_acum  # noqa: E402
from wmf_py.cu_py.metrics import (basin_ppalstream_cut, basin_ppalstream_find, basin_time_to_out, hand,  # noqa: E402
                                  hydro_distance_and_receiver)
from wmf_py.cu_py.streams import basin_netxy_cut, basin_netxy_find, stream_cut, stream_find, stream_find_to_corr  # noqa: E402
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
    # travel time: any raster (cycles give nan); principal stream / network segmentation: acyclic networks only (the
    # reference walks `while True` along the flow path)
    slope = 0.002 + 0.05 * rng.random(fd.shape)
    slope[rng.random(fd.shape) < 0.02] = 0.0            # no speed -> nan behind these cells
    manning = 0.03 + 0.02 * rng.random(fd.shape)
    out[f"{name}/slope"], out[f"{name}/manning"] = slope, manning
    out[f"{name}/time_to_out"] = basin_time_to_out(fd.astype(np.int64), dx, dy, slope, manning)
    if not name.startswith("random"):
        ncell, start = basin_ppalstream_find(corr.astype(bool), fd.astype(np.int64), dem, dx, dy, (int(r0), int(c0)))
        out[f"{name}/ppal_n"] = np.array([ncell, start[0], start[1]])
        out[f"{name}/ppal_profile"] = basin_ppalstream_cut()
        rows, cols, nodes = basin_netxy_find(corr, fd.astype(np.int64), (int(r0), int(c0)))
        out[f"{name}/net_nodes"] = np.array([[k, v[0], v[1]] for k, v in nodes.items()], dtype=np.int64).reshape(-1, 3)
        edges = basin_netxy_cut(corr, nodes, fd.astype(np.int64))
        out[f"{name}/net_edges"] = np.array([[e["id"], e["node_u"], e["node_v"], e["length"]] for e in edges], dtype=np.int64).reshape(-1, 4)
        out[f"{name}/net_cells"] = np.array([rc for e in edges for rc in e["cells"]], dtype=np.int64).reshape(-1, 2)
    for k, v in (("flowdir", fd), ("mask", mask), ("stream", stream), ("dem", dem), ("dxdy", np.array([dx, dy])),
                 ("hdnd", hd), ("aquien", aq.astype(np.int64)), ("hand", h), ("thr", np.array(thr)),
                 ("outlet", np.array([r0, c0])), ("corr", corr), ("acc", acc.astype(np.int32))):
        out[f"{name}/{k}"] = v
np.savez_compressed(os.path.join(HERE, "wmf_py_golden_n2.npz"), **out)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "wmf_py_golden_n2.npz")), "bytes")
