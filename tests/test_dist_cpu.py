"""Multi-rank host logic on CPU with gloo, world_size 2: the stripe exchange + boundary forest of the
row-striped accumulation, and the shard bookkeeping of the SHIA cell / population sharding.  The CUDA
phases are replaced by numpy restatements (tests/helpers.py) so that only the N>1 plumbing is under test."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
import oracle
from wmf_heavy_b200 import dist as wdist
from wmf_heavy_b200 import stripes


def _forest_cpu(parent, weight, blocked):
    acc, done = helpers.forest_numpy(parent.numpy(), weight.numpy(), blocked.numpy())
    return torch.from_numpy(acc), torch.from_numpy(done)


def _worker(rank, world, port, fd, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = stripes.stripe_rows(fd.shape[0], world)
        r0, n = rows[rank]
        acc_local = oracle.basin_acum(fd[r0:r0 + n], np.ones((n, fd.shape[1]), bool))
        rec = {k: torch.from_numpy(v) for k, v in helpers.stripe_records_numpy(fd, r0, n, acc_local).items()}
        recs_all = stripes.gather_records(stripes.pack_records(rec))
        inflow, unres = stripes.stripe_inflow(recs_all, rank, forest_fn=_forest_cpu)
        # SHIA shard bookkeeping: per-shard diagnostics recombine to the whole-basin ones
        N, T = 101, 7
        rng = np.random.default_rng(5)
        per_cell = torch.from_numpy(rng.random((T, N)))
        lo, hi = wdist.shard_range(N, rank, world)
        part = {"balance": per_cell[:, lo:hi].sum(1).float(), "mean_rain": per_cell[:, lo:hi].mean(1).float(),
                "mean_storage": None, "mean_speed": None}
        gathered = [dict() for _ in range(world)]
        for key in ("balance", "mean_rain"):
            for g, t in enumerate(wdist._all_gather(part[key])):
                gathered[g][key] = t
        for g in range(world):
            gathered[g]["mean_storage"] = gathered[g]["mean_speed"] = None
        comb = wdist.combine_cell_shards(gathered, [h - l for l, h in (wdist.shard_range(N, r, world) for r in range(world))])
        q.put((rank, inflow.numpy(), unres.numpy(), comb["balance"].numpy(), comb["mean_rain"].numpy(),
               per_cell.sum(1).numpy(), per_cell.mean(1).numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["sch", "steep"])
def test_stripe_exchange_gloo_world2(kind):
    from wmf_heavy_b200 import ops
    fd = ops.synth_scheidegger_host(48, 37, seed=3) if kind == "sch" else helpers.steepest(40, 33, 4, pit_frac=0.02)
    world = 2
    port = 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fd, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acc_full = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    exp = helpers.expected_stripe_inflow(fd, acc_full, stripes.stripe_rows(fd.shape[0], world))
    for rank, inflow, unres, bal, mr, bal_exp, mr_exp in res:
        assert np.array_equal(inflow, exp[rank]), f"rank {rank}"
        assert not unres.any()
        np.testing.assert_allclose(bal, bal_exp, rtol=1e-6)
        np.testing.assert_allclose(mr, mr_exp, rtol=1e-6)


def _worker_multi(rank, world, port, fd, k, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r0, n = stripes.stripe_rows(fd.shape[0], world)[rank]
        recs = []
        for s0, sn in stripes.stripe_rows(n, k):          # the rank's rows cut into k stripes (accumulate_striped_multi)
            a0 = r0 + s0
            acc_local = oracle.basin_acum(fd[a0:a0 + sn], np.ones((sn, fd.shape[1]), bool))
            rec = {kk: torch.from_numpy(v) for kk, v in helpers.stripe_records_numpy(fd, a0, sn, acc_local).items()}
            recs.append(stripes.pack_records(rec))
        gathered = stripes.gather_records(torch.stack(recs))               # [world, k, 2, 2, nx]
        recs_all = gathered.reshape((-1,) + tuple(gathered.shape[2:]))
        out = [stripes.stripe_inflow(recs_all, rank * k + i, forest_fn=_forest_cpu)[0].numpy() for i in range(k)]
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_several_stripes_per_rank_gloo_world2():
    """The exchange of stripes.accumulate_striped_multi: k stripes per rank behind one all-gather."""
    fd = helpers.steepest(66, 31, 14, pit_frac=0.02)
    world, k = 2, 3
    port = 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_multi, args=(r, world, port, fd, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = []
    for r0, n in stripes.stripe_rows(fd.shape[0], world):
        rows += [(r0 + s0, sn) for s0, sn in stripes.stripe_rows(n, k)]
    acc_full = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    exp = helpers.expected_stripe_inflow(fd, acc_full, rows)
    for rank, out in res:
        for i in range(k):
            assert np.array_equal(out[i], exp[rank * k + i]), f"rank {rank} stripe {i}"


def test_boundary_inflow_three_stripes_single_process():
    fd = helpers.steepest(61, 29, 9, pit_frac=0.01)
    rows = stripes.stripe_rows(fd.shape[0], 3)
    recs = []
    for r0, n in rows:
        acc_local = oracle.basin_acum(fd[r0:r0 + n], np.ones((n, fd.shape[1]), bool))
        recs.append({k: torch.from_numpy(v) for k, v in helpers.stripe_records_numpy(fd, r0, n, acc_local).items()})
    inflow, unres = stripes.boundary_inflow(recs, forest_fn=_forest_cpu)
    acc_full = oracle.basin_acum(fd, np.ones(fd.shape, bool))
    assert np.array_equal(inflow.numpy(), helpers.expected_stripe_inflow(fd, acc_full, rows))
    assert not unres.any()


def test_shard_ranges_cover():
    for n, w in ((10, 3), (512, 8), (7, 8), (250000, 8)):
        spans = [wdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert stripes.stripe_rows(65536, 8) == [(i * 8192, 8192) for i in range(8)]
