"""One accumulation of the 16384^2 terrain raster (for an ncu launch list): python tools/prof_terrain.py [n] [mask]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wmf_heavy_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = synth.terrain_dir(n, n)
m = synth.nodata_mask(n, n, 0.05) if len(sys.argv) > 2 else None
acc = torch.empty((n, n), dtype=torch.int32, device="cuda")
for _ in range(3):
    ops.d8_accum(d, m, out=acc)
torch.cuda.synchronize()
