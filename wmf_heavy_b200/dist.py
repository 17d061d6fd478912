"""Sharding of the SHIA path across GPUs: cells of one basin, and parameter sets of a
NSGA-II population.  One process per GPU, torch.distributed for the plumbing; as shipped, cells do not
interact inside shia_v1, so neither shard needs a data-path collective -- only the per-step whole-basin
diagnostics (cells) or the results (population) are gathered at the end, in rank order, so the result
does not depend on the reduction schedule."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of n items owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _all_gather(t: torch.Tensor, group=None) -> List[torch.Tensor]:
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bufs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(bufs, t.contiguous(), group=group)
    return bufs


def combine_cell_shards(parts: List[Dict[str, torch.Tensor]], counts: List[int]) -> Dict[str, torch.Tensor]:
    """Whole-basin diagnostics from per-shard ones, summed in shard order in fp64: balance adds up, the
    per-step means are weighted by the shard's cell count (Modelos_heavy.f90:494, 518, 522, 533)."""
    n_tot = float(sum(counts))
    out: Dict[str, torch.Tensor] = {}
    out["balance"] = sum(p["balance"].double() for p in parts).float()
    for key in ("mean_rain", "mean_storage", "mean_speed"):
        if parts[0].get(key) is not None:
            out[key] = (sum(p[key].double() * c for p, c in zip(parts, counts)) / n_tot).float()
    return out


_DIAG_KEYS = ("balance", "mean_rain", "mean_storage", "mean_speed")


def shia5_run_cell_sharded(cell_static, rain_records, pos_evento, evp, calib, sto, hspeed=None, group=None, **kw):
    """Each rank holds a contiguous range of the basin's cells (its columns of cell_static / rain_records /
    sto): run the local shard, then gather the diagnostics.  Returns (local outputs, whole-basin diagnostics).
    ONE collective and no host round trip: the shard's cell count and every per-step series travel in one packed fp64
    buffer, and the ranks' rows are combined on the device in rank order (`combine_cell_shards`' arithmetic)."""
    import torch.distributed as dist
    local = ops.shia5_run(cell_static, rain_records, pos_evento, evp, calib, sto, hspeed, **kw)
    dev = sto.device
    keys = [k for k in _DIAG_KEYS if local.get(k) is not None]
    shapes = [tuple(local[k].shape) for k in keys]
    sizes = [int(local[k].numel()) for k in keys]
    pack = torch.cat([torch.full((1,), float(cell_static.shape[1]), dtype=torch.float64, device=dev)] +
                     [local[k].reshape(-1).double() for k in keys])
    world = dist.get_world_size(group)
    if pack.is_cuda:
        rows = torch.empty((world, pack.numel()), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(rows, pack, group=group)
    else:                                                 # gloo (the CPU tests)
        rows = torch.stack(_all_gather(pack, group))
    share = rows[:, 0] / rows[:, 0].sum()                 # cell-count weights of the means (fp64, like the Fortran's N_cel)
    acc = torch.zeros(pack.numel() - 1, dtype=torch.float64, device=dev)
    weight_one = torch.ones((), dtype=torch.float64, device=dev)
    off_bal = sizes[0] if keys and keys[0] == "balance" else 0
    for r in range(world):                                # rank order: the result does not depend on a reduction schedule
        row = rows[r, 1:]
        if off_bal:
            acc[:off_bal] += row[:off_bal] * weight_one
        acc[off_bal:] += row[off_bal:] * share[r]
    out: Dict[str, torch.Tensor] = {}
    o = 0
    for k, shp, n in zip(keys, shapes, sizes):
        out[k] = acc[o:o + n].reshape(shp).float()
        o += n
    return local, out


def shia5_run_population_sharded(cell_static, rain_records, pos_evento, evp, calibs, sto0, group=None, **kw):
    """NSGA-II population: every rank holds the whole basin and advances its slice of the parameter sets
    (calibs [P, 11]); returns the gathered balance [P, N_reg] and this rank's outputs."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    P = calibs.shape[0]
    if P < 1:                                             # the same test on every rank: all raise together
        raise ValueError("the population is empty")
    lo, hi = shard_range(P, rank, world)
    n = cell_static.shape[1]
    n_reg = pos_evento.shape[0]
    local = None
    if hi > lo:                                           # (P < world leaves some ranks without a set: they only join the gather)
        sto = sto0.expand(hi - lo, 5, n).contiguous() if sto0.dim() == 2 else sto0[lo:hi].contiguous()
        local = ops.shia5_run(cell_static, rain_records, pos_evento, evp, calibs[lo:hi].contiguous(), sto, None, **kw)
    sizes = [shard_range(P, r, world) for r in range(world)]
    pmax = max(h - l for l, h in sizes)
    pad = torch.zeros((pmax, n_reg), dtype=torch.float32, device=cell_static.device)
    if local is not None:
        pad[: hi - lo] = local["balance"]
    gathered = _all_gather(pad, group)
    balance = torch.cat([gathered[r][: h - l] for r, (l, h) in enumerate(sizes)])
    return balance, local
