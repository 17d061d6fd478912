"""Run the SHIA 5-tank kernel once at the BASELINE config-3 size (for ncu): python tools/shia_prof.py [cells] [steps] [P]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wmf_heavy_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
t = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
r = bench.bench_shia(torch, ops, torch.device("cuda", 0), n, t, 2, 1)
print(r)
