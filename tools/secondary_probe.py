"""Timing of the accumulation on the secondary input classes: python tools/secondary_probe.py [n]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wmf_heavy_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)


def timed(d, m, reps=5):
    acc = torch.empty((n, n), dtype=torch.int32, device=dev)
    for _ in range(2):
        ops.d8_accum(d, m, out=acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.d8_accum(d, m, out=acc)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, acc


sch = ops.synth_scheidegger(n, n)
ter = synth.terrain_dir(n, n)
print("terrain code histogram", (torch.bincount(ter.view(-1).to(torch.int64) + 1, minlength=9).float() / ter.numel()).tolist())
m1 = torch.ones((n, n), dtype=torch.uint8, device=dev)
m95 = synth.nodata_mask(n, n, 0.05)
for name, d, m in (("scheidegger", sch, None), ("scheidegger+all-true mask", sch, m1), ("scheidegger+5% nodata", sch, m95),
                   ("terrain", ter, None), ("terrain+5% nodata", ter, m95)):
    ms, acc = timed(d, m)
    st = {}
    ops.d8_accum(d, m, out=acc, stats=st)
    print(json.dumps({"input": name, "tiles": st.get("tile_classes"), "rounds": st.get("jump_rounds"), "ms": ms, "mcells_s": n * n / ms / 1e3, "acc_sum": int(acc.sum(dtype=torch.int64))}), flush=True)
