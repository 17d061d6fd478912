"""Timing of the basin path (mask, BFS structure, basics, stream nodes) at several raster sizes: python tools/basin_probe.py [n ...]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wmf_heavy_b200 import ops

dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
for n in [int(a) for a in sys.argv[1:]] or [2048, 8192]:
    print(json.dumps(bench.bench_basin(torch, ops, dev, peak, n)), flush=True)
