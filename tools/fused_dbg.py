"""Timeline of the fused phase G per tile row (library built with EXTRA=-DWMF_FUSED_DEBUG)."""
import ctypes, sys, torch, numpy as np
sys.path.insert(0, '.')
from wmf_heavy_b200 import _lib, ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib = _lib.load()
d = ops.synth_scheidegger(n, n, seed=20260101)
acc = torch.empty((n, n), dtype=torch.int32, device='cuda')
for _ in range(3):
    ops.d8_accum(d, None, out=acc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.d8_accum(d, None, out=acc); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1))
ts = np.zeros(4 * 65536, dtype=np.uint64)
lib.wmf_dbg_ts(ts.ctypes.data_as(ctypes.c_void_p))
tx, ty = n // 256, n // 64
t = ts.reshape(4, 65536)[:, :tx * ty].astype(np.float64).reshape(4, ty, tx)
t0 = t[0].min()
t = (t - t0) / 1e3
names = ["A done", "built", "walk start", "walk end"]
print("row   " + "  ".join(f"{nm:>22s}" for nm in names) + "   (min / max us over the row's tiles)")
for r in list(range(0, ty, 16)) + [ty - 2, ty - 1]:
    print(f"{r:4d}  " + "  ".join(f"{t[k, r].min():10.1f} {t[k, r].max():10.1f} " for k in range(4)))
print("walk duration per tile: mean %.1f  p99 %.1f  max %.1f us" % ((t[3] - t[2]).mean(), np.percentile(t[3] - t[2], 99), (t[3] - t[2]).max()))
print("A done -> built: mean %.1f max %.1f; built -> walk start: mean %.1f max %.1f" % ((t[1] - t[0]).mean(), (t[1] - t[0]).max(), (t[2] - t[1]).mean(), (t[2] - t[1]).max()))
print("last A done %.1f, last walk end %.1f" % (t[0].max(), t[3].max()))
