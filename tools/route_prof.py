"""Time the routing extension on the largest basin of a Scheidegger raster: python tools/route_prof.py [n] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench  # noqa: F401  (same synthetic forcing as bench_shia)
from wmf_heavy_b200 import _lib, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)
key = ops.synth_scheidegger(n, n, seed=20260101, encoding=_lib.ENC_KEYPAD)
acc = ops.d8_accum(key, None, encoding=_lib.ENC_KEYPAD)
flat = int(torch.argmax(acc))
r0, c0 = divmod(flat, n)
st = ops.basin_structure(key, c0 + 1, r0 + 1).cpu().numpy()
drain = st[0].astype(np.int32)
N = drain.size
rank, dmax, _, _ = ops.route_topology(drain)
g = torch.Generator(device=dev).manual_seed(7)
u = lambda lo, hi, *s: lo + (hi - lo) * torch.rand(*s, device=dev, generator=g)
cs = torch.empty((17, N), device=dev)
cs[0:4] = u(0.5, 1.5, 4, N) * torch.tensor([[1e-2], [3e-3], [1e-3], [1e-4]], device=dev)
cs[4:8] = u(0.5, 1.5, 4, N) * torch.tensor([[0.5], [0.05], [0.005], [1.0]], device=dev)
cs[8:12] = 1.0
cs[12], cs[13], cs[14] = u(50, 150, N), u(10, 50, N), u(200, 800, N)
cs[15], cs[16] = 30.0, 900.0
n_rec = 101
rec = (1000.0 * u(0.0, 4.0, n_rec, N)).to(torch.int32)
rec[0] = 0
pos = torch.where(torch.rand(T) < 0.7, torch.ones(T, dtype=torch.int64), torch.randint(2, n_rec + 1, (T,))).to(torch.int32).to(dev)
evp = torch.full((T,), 0.1, device=dev)
cal = torch.ones(11, device=dev)
ref = None
for persistent in (True, False):
    times = []
    for it in range(3):
        sto = torch.full((5, N), 0.001, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = ops.shia5_route_run(cs, rec, pos, evp, cal, sto, drain, None, 1, dt=300.0, persistent=persistent)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[1]
    same = "" if ref is None else f"; identical to the cooperative run: {bool(torch.equal(ref[0], out['Q']) and torch.equal(ref[1], out['StoOut']))}"
    ref = ref or (out["Q"].clone(), out["StoOut"].clone())
    print(f"routing extension ({'one cooperative launch' if persistent else str(T + dmax) + ' launches'}): basin of {N} cells "
          f"(depth {dmax}), {T} steps: {ms:.2f} ms = {N * T / ms / 1e6:.1f} G cell-steps/s; peak Q at the outlet "
          f"{float(out['Q'][0].max()):.1f} m3/s{same}")
