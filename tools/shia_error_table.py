"""Error of the SHIA 5-tank kernel against the CPU oracle, per tank, for the fast arithmetic forms (x * 0.001f,
S1 * (1 / H1), exp2(0.6 * log2 x) on the SFU) and for the strict ones (the Fortran's own operations):
    python tools/shia_error_table.py > profiles/r02_shia_error_table.md
Relative error = |gpu - oracle| / max(|oracle|, 1e-3 mm); also the largest absolute error [mm] and, for the strict
build, the largest distance in units in the last place."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers
import oracle
from wmf_heavy_b200 import ops


def ulps(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return int(np.abs(ia - ib).max())


print("# SHIA 5-tank kernel vs CPU oracle: per-tank error after the whole run (20 000 cells)\n")
print("| speed types | returns | steps | arithmetic | tank | max rel. error | max abs. error [mm] | max ulps | value range [mm] |")
print("|---|---|---|---|---|---|---|---|---|")
N = 20000
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
worst = {"fast": 0.0, "strict": 0.0}
for st in ((1, 1, 1), (2, 2, 2), (1, 2, 1)):
    for ret in (0.0, 1.0):
        for T in (200, 1000):
            inp = helpers.shia5_inputs(oracle, N, T, 700 + T, speed_type=st)
            inp.retorno_gr = inp.retorno_aq = ret
            sto0 = np.full((5, N), 0.001, np.float32)
            o = oracle.f_shia_v1(inp, np.ones(11, np.float32), sto0, show_storage=False, show_mean_speed=False)
            for mode in ("fast", "strict"):
                out = ops.shia5_run(dev(helpers.cell_static(inp)), dev(inp.rain_records), dev(inp.pos_evento), dev(inp.evp),
                                    torch.ones((1, 11), device="cuda"), dev(sto0[None].copy()), None, dt=inp.dt, speed_type=st,
                                    retorno_gr=ret, retorno_aq=ret, calc_niter=inp.calc_niter, strict=mode == "strict")
                g = out["StoOut"][0].cpu().numpy()
                for k in range(4):
                    ref = o["StoOut"][k]
                    rel = float((np.abs(g[k] - ref) / np.maximum(np.abs(ref), 1e-3)).max())
                    worst[mode] = max(worst[mode], rel)
                    print(f"| {st} | {int(ret)} | {T} | {mode} | {k + 1} | {rel:.2e} | {float(np.abs(g[k] - ref).max()):.2e} | "
                          f"{ulps(g[k], ref)} | {float(ref.min()):.3g} .. {float(ref.max()):.3g} |")
print(f"\nLargest relative error: fast {worst['fast']:.2e}, strict {worst['strict']:.2e} (north_star tolerance 1e-5).")
