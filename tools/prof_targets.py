"""One launch of every hot kernel at its bench size, inside a cudaProfilerStart/Stop range, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_full python tools/prof_targets.py
(warm-up runs stay outside the range).  D8: the 16384^2 Scheidegger raster (headline), the same with an all-true mask, the
terrain raster (jump solver); SHIA: 1 Mi cells x 1000 steps with type-1 and type-2 speeds, the fp64 3-tank model."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wmf_heavy_b200 import _lib, ops, synth

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = ops.synth_scheidegger(n, n, seed=bench.SEED)
ter = synth.terrain_dir(n, n, seed=5)
m1 = torch.ones((n, n), dtype=torch.uint8, device=dev)
acc = torch.empty((n, n), dtype=torch.int32, device=dev)
nc, nr = 1 << 20, 1000
cs, rec, pos, evp, _ = bench.shia_inputs(torch, dev, nc, nr, seed=7)
calib = torch.ones((1, 11), device=dev)
sto = lambda: torch.full((1, 5, nc), 0.001, device=dev)
rain3 = torch.rand((200, nc), device=dev, dtype=torch.float64) * 20.0
prm = _lib.Shia3Params(300.0, 100.0, 30.0, 500.0, 0.0, 0.01, 0.02, 0.01, 0.005, 1.0e6, 300.0 / 3600.0, 0, 0)
s3 = lambda: tuple(torch.full((nc,), v, device=dev, dtype=torch.float64) for v in (10.0, 5.0, 50.0))


def everything():
    ops.d8_accum(d, None, out=acc)
    ops.d8_accum(d, m1, out=acc)
    ops.d8_accum(ter, None, out=acc)
    ops.shia5_run(cs, rec, pos, evp, calib, sto(), None, dt=300.0, speed_type=(1, 1, 1), calc_niter=5)
    ops.shia5_run(cs, rec, pos, evp, calib, sto(), None, dt=300.0, speed_type=(2, 2, 2), calc_niter=5)
    a, b, c = s3()
    ops.shia3_run(rain3, None, a, b, c, prm, 200)


for _ in range(2):
    everything()
torch.cuda.synchronize()
torch.cuda.profiler.start()
everything()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
