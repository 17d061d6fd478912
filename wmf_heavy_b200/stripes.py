"""Row-stripe sharding of the full-raster D8 accumulation across GPUs (BASELINE config 4).

GPU g owns rows [row0_g, row0_g + ny_g) of a raster too large (or too slow) for one device.
    1. local   every rank solves its stripe in isolation and condenses it to 2*nx boundary records
               (ops.d8_stripe_local);
    2. exchange ONE all-gather of the records (NCCL over NVLink; ~15 bytes per boundary cell);
    3. boundary every rank redundantly solves the 2*G*nx-node boundary forest and sums, for each of its own
               boundary cells, what the neighbouring stripes send in (ops.stripe_boundary_inflow);
    4. finish  every rank injects that inflow and writes its final ACC rows (ops.d8_stripe_finish).
Integer adds make the result bit-identical to the single-device accumulation.  The boundary logic
(`boundary_inflow`) is plain tensor indexing and runs on any device, so it is covered on CPU with
gloo (tests/test_dist_cpu.py); phases 1 and 4 are CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import ACC_I64, ENC_INTERNAL

Records = Dict[str, torch.Tensor]     # k int8, exit_acc int64, link int32, unres uint8; each [2, nx]
_DC = (1, 1, 0, -1, -1, -1, 0, 1)     # column offset of direction k (basics.py:37-46)


def stripe_rows(ny_total: int, n_stripes: int) -> List[Tuple[int, int]]:
    """(row0, ny) of each stripe: contiguous, sizes differing by at most one row."""
    base, extra = divmod(ny_total, n_stripes)
    out, r = [], 0
    for g in range(n_stripes):
        n = base + (1 if g < extra else 0)
        out.append((r, n))
        r += n
    return out


def boundary_inflow(records: Sequence[Records], forest_fn: Optional[Callable] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Solve the boundary forest of all stripes -> (ext_inflow int64 [G, 2, nx], ext_unres uint8 [G, 2, nx]).

    Node (g, side, col) is a stripe exit when it drains across its stripe edge; its parent is the boundary
    cell where that flow leaves the neighbouring stripe again (the neighbour's `link` at the receiving cell)."""
    forest_fn = forest_fn or ops.forest_accum
    G = len(records)
    nx = records[0]["k"].shape[1]
    dev = records[0]["k"].device
    k = torch.stack([r["k"] for r in records]).to(torch.int64)                # [G, 2, nx]
    link = torch.stack([r["link"] for r in records]).to(torch.int64)
    w = torch.stack([r["exit_acc"] for r in records]).reshape(-1)
    blocked = torch.stack([r["unres"] for r in records]).reshape(-1)
    g_idx = torch.arange(G, device=dev).view(G, 1, 1).expand(G, 2, nx)
    side = torch.arange(2, device=dev).view(1, 2, 1).expand(G, 2, nx)
    col = torch.arange(nx, device=dev).view(1, 1, nx).expand(G, 2, nx)
    up = (side == 0) & (k >= 5) & (k <= 7) & (g_idx > 0)
    down = (side == 1) & (k >= 1) & (k <= 3) & (g_idx < G - 1)
    is_exit = up | down
    dc = torch.tensor(_DC + (0, 0), device=dev)[k.clamp(0, 9)]
    g2 = torch.where(up, g_idx - 1, g_idx + 1).clamp(0, G - 1)
    side2 = 1 - side
    col2 = (col + dc).clamp(0, nx - 1)
    recv = (g2 * 2 + side2) * nx + col2                                       # node id of the receiving cell
    recv_active = k.reshape(-1)[recv.reshape(-1)].view(G, 2, nx) != 9
    feeds = is_exit & recv_active
    recv_link = link.reshape(-1)[recv.reshape(-1)].view(G, 2, nx)
    parent = torch.where(feeds & (recv_link >= 0), g2 * 2 * nx + recv_link, torch.full_like(recv_link, -1))
    acc, resolved = forest_fn(parent.reshape(-1).to(torch.int32).contiguous(), w.contiguous(), blocked.contiguous())
    feeds_f, recv_f, ok = feeds.reshape(-1), recv.reshape(-1), resolved.reshape(-1) != 0
    # scatter-adds with zero contributions instead of boolean-mask indexing: no host synchronisation
    inflow = torch.zeros(G * 2 * nx, dtype=torch.int64, device=dev)
    inflow.index_add_(0, recv_f, torch.where(feeds_f & ok, acc, torch.zeros_like(acc)))
    bad = torch.zeros(G * 2 * nx, dtype=torch.int32, device=dev)
    bad.index_add_(0, recv_f, (feeds_f & ~ok).to(torch.int32))
    return inflow.view(G, 2, nx), (bad > 0).to(torch.uint8).view(G, 2, nx)


def pack_records(rec: Records) -> torch.Tensor:
    """[2, 2, nx] int64: plane 0 = exit_acc, plane 1 = link (low 32 bits) | k << 32 | unres << 40
    (the layout wmf_d8_stripe_local writes, include/wmf_b200.h)."""
    meta = (rec["link"].to(torch.int64) & 0xFFFFFFFF) | (rec["k"].to(torch.int64) << 32) | (rec["unres"].to(torch.int64) << 40)
    return torch.stack([rec["exit_acc"].to(torch.int64), meta])


def unpack_records(t: torch.Tensor) -> Records:
    meta = t[1]
    return {"exit_acc": t[0], "link": (meta & 0xFFFFFFFF).to(torch.int32), "k": ((meta >> 32) & 0xFF).to(torch.int8),
            "unres": ((meta >> 40) & 1).to(torch.uint8)}


def gather_records(rec_packed: torch.Tensor, group=None) -> torch.Tensor:
    """The one exchange step: a single all-gather of the packed boundary records -> [G, 2, 2, nx]."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if rec_packed.is_cuda:                               # NCCL: one buffer, no concatenation afterwards
        out = torch.empty((world,) + tuple(rec_packed.shape), dtype=rec_packed.dtype, device=rec_packed.device)
        dist.all_gather_into_tensor(out, rec_packed.contiguous(), group=group)
        return out
    bufs = [torch.empty_like(rec_packed) for _ in range(world)]      # gloo (the CPU tests)
    dist.all_gather(bufs, rec_packed.contiguous(), group=group)
    return torch.stack(bufs)


def stripe_inflow(recs_all: torch.Tensor, me: int, forest_fn: Optional[Callable] = None
                  ) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ext_inflow [2, nx], ext_unres [2, nx]) of stripe `me`.  CUDA tensors go through the library's boundary
    kernels; CPU tensors (tests, gloo) through the tensor-indexing restatement above."""
    if recs_all.is_cuda and forest_fn is None:
        return ops.stripe_boundary_inflow(recs_all.contiguous(), me)
    inflow, unres = boundary_inflow([unpack_records(recs_all[g]) for g in range(recs_all.shape[0])], forest_fn)
    return inflow[me].contiguous(), unres[me].contiguous()


def accumulate_striped(dir_stripe: torch.Tensor, row0: int, ny_total: int, group=None, encoding: int = ENC_INTERNAL,
                       mask_stripe: Optional[torch.Tensor] = None, acc_dtype: int = ACC_I64,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One rank's part of the striped accumulation (torch.distributed must be initialised, backend nccl).
    Ranks must own the stripes in rank order."""
    import torch.distributed as dist
    try:
        rec, ws = ops.d8_stripe_local(dir_stripe, row0, ny_total, mask_stripe, encoding)
    except Exception:
        # a stripe the library rejects (e.g. fewer than 2 rows) must not leave the other ranks waiting in the
        # collective: join it with an empty record, then fail here (the launcher takes the job down)
        gather_records(torch.zeros((2, 2, dir_stripe.shape[1]), dtype=torch.int64, device=dir_stripe.device), group)
        raise
    recs_all = gather_records(rec, group)
    inflow, unres = stripe_inflow(recs_all, dist.get_rank(group))
    return ops.d8_stripe_finish(dir_stripe, ws, row0, ny_total, inflow, unres, mask_stripe, encoding, acc_dtype, out)


def accumulate_striped_multi(dir_local: torch.Tensor, row0: int, ny_total: int, n_local: int, group=None,
                             encoding: int = ENC_INTERNAL, mask_local: Optional[torch.Tensor] = None,
                             acc_dtype: int = ACC_I64, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Like ``accumulate_striped`` with ``n_local`` stripes per rank (the same number on every rank): a rank's rows
    are cut into stripes that each hold fewer than 2^31 cells -- 65536^2 on two GPUs is 2 x 2^31 cells -- and still
    ONE all-gather carries every stripe's boundary records.  Stripe index = rank * n_local + i."""
    import torch.distributed as dist
    ny, nx = dir_local.shape
    rank = dist.get_rank(group)
    if n_local < 1:
        raise ValueError("n_local must be >= 1")
    rows = stripe_rows(ny, n_local)
    recs, wss = [], []
    try:
        for r0, n in rows:
            m = None if mask_local is None else mask_local[r0:r0 + n].contiguous()
            rec, ws = ops.d8_stripe_local(dir_local[r0:r0 + n].contiguous(), row0 + r0, ny_total, m, encoding)
            recs.append(rec)
            wss.append(ws)
    except Exception:
        # (see accumulate_striped) join the collective the other ranks are heading for, then fail
        gather_records(torch.zeros((n_local, 2, 2, nx), dtype=torch.int64, device=dir_local.device), group)
        raise
    gathered = gather_records(torch.stack(recs), group)                    # [world, n_local, 2, 2, nx]
    recs_all = gathered.reshape((-1,) + tuple(gathered.shape[2:]))
    if out is None:
        out = torch.empty((ny, nx), dtype=ops._ACC_TORCH[acc_dtype], device=dir_local.device)
    for i, (r0, n) in enumerate(rows):
        m = None if mask_local is None else mask_local[r0:r0 + n].contiguous()
        inflow, unres = stripe_inflow(recs_all, rank * n_local + i)
        ops.d8_stripe_finish(dir_local[r0:r0 + n].contiguous(), wss[i], row0 + r0, ny_total, inflow, unres, m, encoding,
                             acc_dtype, out[r0:r0 + n])
    return out


def accumulate_striped_single_device(dir_t: torch.Tensor, n_stripes: int, encoding: int = ENC_INTERNAL,
                                     mask_t: Optional[torch.Tensor] = None, acc_dtype: int = ACC_I64) -> torch.Tensor:
    """All stripes on ONE device, one after the other: same kernels and boundary logic as the multi-GPU
    path (used by the GPU parity tests and when fewer GPUs than stripes are available)."""
    ny, nx = dir_t.shape
    rows = stripe_rows(ny, n_stripes)
    recs, wss = [], []
    for r0, n in rows:
        m = None if mask_t is None else mask_t[r0:r0 + n].contiguous()
        rec, ws = ops.d8_stripe_local(dir_t[r0:r0 + n].contiguous(), r0, ny, m, encoding)
        recs.append(rec)
        wss.append(ws)
    recs_all = torch.stack(recs)
    out = torch.empty((ny, nx), dtype=ops._ACC_TORCH[acc_dtype], device=dir_t.device)
    for g, (r0, n) in enumerate(rows):
        m = None if mask_t is None else mask_t[r0:r0 + n].contiguous()
        inflow, unres = stripe_inflow(recs_all, g)
        ops.d8_stripe_finish(dir_t[r0:r0 + n].contiguous(), wss[g], r0, ny, inflow, unres, m, encoding, acc_dtype,
                             out[r0:r0 + n])
    return out
