"""SHIA sharding on real GPUs (NCCL): cells sharded and parameter sets sharded against the one-GPU run.
torchrun --nproc-per-node 2 tools/check_shia_sharded.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import helpers
import oracle
from wmf_heavy_b200 import dist as wdist
from wmf_heavy_b200 import ops

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
N, T, P = 5001, 48, 7
inp = helpers.shia5_inputs(oracle, N, T, 33)
cs = torch.from_numpy(helpers.cell_static(inp)).cuda()
rec, pos, evp = (torch.from_numpy(a).cuda() for a in (inp.rain_records, inp.pos_evento, inp.evp))
calibs = torch.from_numpy(np.random.default_rng(34).uniform(0.5, 2.0, (P, 11)).astype(np.float32)).cuda()
sto0 = torch.full((5, N), 0.001, device="cuda")
kw = dict(dt=inp.dt, speed_type=inp.speed_type)
ref = ops.shia5_run(cs, rec, pos, evp, calibs, sto0.expand(P, 5, N).contiguous(), None, **kw)
ok = True
# parameter sets sharded
bal, local = wdist.shia5_run_population_sharded(cs, rec, pos, evp, calibs, sto0, **kw)
lo, hi = wdist.shard_range(P, rank, world)
ok &= torch.equal(bal, ref["balance"]) and torch.equal(local["StoOut"], ref["StoOut"][lo:hi])
# cells sharded (one parameter set)
clo, chi = wdist.shard_range(N, rank, world)
loc, whole = wdist.shia5_run_cell_sharded(cs[:, clo:chi].contiguous(), rec[:, clo:chi].contiguous(), pos, evp, calibs[:1].contiguous(),
                                          sto0[None, :, clo:chi].contiguous(), None, **kw)
ok &= torch.equal(loc["StoOut"][0], ref["StoOut"][0][:, clo:chi])
ok &= torch.allclose(whole["balance"][0], ref["balance"][0], rtol=1e-4, atol=1e-4 * float(ref["StoOut"][0].sum()))
ok &= torch.allclose(whole["mean_rain"], ref["mean_rain"], rtol=1e-5, atol=1e-7)
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("shia_sharded", "OK" if int(flag) == 1 else "MISMATCH", f"world={world} cells={N} sets={P}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
