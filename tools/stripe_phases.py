"""Time the phases of the striped accumulation on one device (CUDA events): python tools/stripe_phases.py [n] [G]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wmf_heavy_b200 import ops, stripes, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = ops.synth_scheidegger(n, n, seed=20260101, row0=0, nx_total=n)
ny_total = G * n


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


t_local, (rec, ws) = timed(lambda: ops.d8_stripe_local(d, 0, ny_total))
recs = torch.stack([rec] * G)
t_pack = 0.0
t_bound, (inflow, unres) = timed(lambda: stripes.stripe_inflow(recs, 0))
wide = G * n * n > 2**31 - 1
out = torch.empty((n, n), dtype=torch.int64 if wide else torch.int32, device="cuda")
def both():                                   # (a finish consumes the workspace of its local pass: time the pair)
    r, w = ops.d8_stripe_local(d, 0, ny_total)
    return ops.d8_stripe_finish(d, w, 0, ny_total, inflow, unres, acc_dtype=_lib.ACC_I64 if wide else _lib.ACC_I32, out=out)


t_pair, _ = timed(both)
t_fin = t_pair - t_local
out32 = torch.empty((n, n), dtype=torch.int32, device="cuda")
t_whole, _ = timed(lambda: ops.d8_accum(d, None, out=out32))
print(f"n={n} G={G}: local {t_local:.3f} ms, pack/unpack {t_pack:.3f}, boundary forest {t_bound:.3f}, finish {t_fin:.3f}; unstriped accumulate {t_whole:.3f}")
