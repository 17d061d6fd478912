"""2+ ranks, several stripes per rank (stripes.accumulate_striped_multi) against the one-device accumulation:
torchrun --nproc-per-node 2 tools/check_striped_multi.py [ny nx n_local]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from wmf_heavy_b200 import _lib, ops, stripes

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
ny, nx, k = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (6000, 4096, 3)
rows = stripes.stripe_rows(ny, world)
r0, n = rows[rank]
d_local = ops.synth_scheidegger(n, nx, seed=5, row0=r0, nx_total=nx)
got = stripes.accumulate_striped_multi(d_local, r0, ny, k, acc_dtype=_lib.ACC_I64)
ref = ops.d8_accum(ops.synth_scheidegger(ny, nx, seed=5), None, acc_dtype=_lib.ACC_I64)[r0:r0 + n]
ok = torch.equal(got, ref)
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("striped_multi", "OK" if int(flag) == 1 else "MISMATCH", f"world={world} n_local={k} raster={ny}x{nx}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
