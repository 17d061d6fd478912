"""Drop-in for the raster-sized functions of ``wmf_py/cu_py/streams.py``.

    stream_find(acum, threshold)                     <- streams.py:19-36   (acum >= threshold)
    stream_cut(stream, mask_basin)                   <- streams.py:39-42
    stream_find_to_corr(stream, flowdir, outlet_rc)  <- streams.py:45-83   (stream cells connected to the outlet)

The node / edge segmentation (basin_netxy_*) and the seed search produce small Python dictionaries and stay in
the reference's own Python (out of scope).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .. import ops
from .._dev import require_cuda, to_dev
from .._lib import ENC_INTERNAL
from .basics import _flowdir_to_dev, _mask_to_dev


def stream_find(acum, threshold: int) -> np.ndarray:
    """uint8 mask of the cells whose accumulation reaches ``threshold`` (works on host or device arrays)."""
    if isinstance(acum, torch.Tensor):
        return (acum >= threshold).to(torch.uint8)
    return (np.asarray(acum) >= threshold).astype(np.uint8)


def stream_cut(stream, mask_basin) -> np.ndarray:
    return (np.asarray(stream).astype(np.uint8) & np.asarray(mask_basin).astype(np.uint8)).astype(np.uint8)


def stream_find_to_corr(stream, flowdir, outlet_rc: Tuple[int, int]) -> np.ndarray:
    """Keep only the stream cells whose flow path reaches ``outlet_rc`` through stream cells."""
    require_cuda()
    ny, nx = np.asarray(stream).shape if not isinstance(stream, torch.Tensor) else tuple(stream.shape)
    r0, c0 = outlet_rc
    if not (0 <= r0 < ny and 0 <= c0 < nx):
        raise ValueError("outlet outside grid")
    s = _mask_to_dev(stream)
    active = s.clone()
    active[r0, c0] = 1                       # the search starts at the outlet whether or not it is a stream cell
    mb, _ = ops.basin_mask(_flowdir_to_dev(flowdir), active, (int(r0), int(c0)), ENC_INTERNAL)
    return (mb & s).cpu().numpy().astype(np.uint8)


# ---- network nodes and segmentation (streams.py:282-384) -----------------------------------------------------
def basin_netxy_find(stream, flowdir, outlet_rc: Tuple[int, int]):
    """(rows, cols, nodes): coordinates of all stream cells and ``{id: (row, col)}`` of the network nodes -- stream
    cells whose degree (stream donors + stream receiver) is not 2, and the outlet -- numbered from 1 in row-major order."""
    require_cuda()
    s_host = np.asarray(stream)
    ny, nx = s_host.shape
    r0, c0 = outlet_rc
    outlet = int(r0) * nx + int(c0) if (0 <= r0 < ny and 0 <= c0 < nx) else -1
    node = ops.stream_nodes(_flowdir_to_dev(flowdir), _mask_to_dev(s_host != 0), outlet, ENC_INTERNAL)[0]
    rows, cols = np.nonzero(s_host)
    nr, nc = np.nonzero(node.cpu().numpy())
    nodes = {i + 1: (int(r), int(c)) for i, (r, c) in enumerate(zip(nr, nc))}
    return rows, cols, nodes


def basin_netxy_cut(stream, nodes, flowdir):
    """Edges between nodes: ``id``, ``node_u`` (upstream), ``node_v`` (downstream), ``length`` (cells, both ends
    included) and ``cells``.  One edge per node with a stream receiver whose path meets another node, in the order of
    ``nodes``; a path that returns to its own node or meets none gives no edge (the reference would not terminate)."""
    require_cuda()
    s_host = np.asarray(stream)
    ny, nx = s_host.shape
    d = _flowdir_to_dev(flowdir)
    # the node set is whatever the caller passes: rebuild its raster and let the kernels use it through `outlet`-free
    # degrees only when it matches basin_netxy_find's; otherwise walk with the given set
    node_map = np.zeros((ny, nx), dtype=np.int64)
    for nid, (r, c) in nodes.items():
        node_map[r, c] = nid
    node_t, dn, ds, un, us = ops.stream_nodes(d, _mask_to_dev(s_host != 0), -1, ENC_INTERNAL)
    gpu_nodes = node_t.cpu().numpy().astype(bool)
    extra = (node_map > 0) & ~gpu_nodes                  # e.g. the outlet, a node by decree
    if ((node_map > 0) != gpu_nodes).any() and not (extra.sum() <= 1 and not (gpu_nodes & ~(node_map > 0)).any()):
        return _netxy_cut_host(s_host, nodes, np.asarray(flowdir), node_map)
    if extra.any():                                      # re-run with that cell as the outlet
        r, c = np.argwhere(extra)[0]
        node_t, dn, ds, un, us = ops.stream_nodes(d, _mask_to_dev(s_host != 0), int(r) * nx + int(c), ENC_INTERNAL)
    dn, ds, un, us = (t.cpu().numpy().reshape(-1) for t in (dn, ds, un, us))
    flat_nodes = node_map.reshape(-1)
    # cells strictly inside an edge know their upstream node and their distance from it: one sort orders every edge
    inner = np.nonzero((np.asarray(s_host).reshape(-1) != 0) & (flat_nodes == 0) & (un >= 0))[0]
    order = inner[np.lexsort((us[inner], un[inner]))]
    starts = {}
    if order.size:
        key = un[order]
        cut = np.nonzero(np.diff(key))[0] + 1
        for a, b in zip(np.concatenate([[0], cut]), np.concatenate([cut, [order.size]])):
            starts[int(key[a])] = order[a:b]
    edges = []
    edge_id = 1
    for nid, (r, c) in nodes.items():
        u = r * nx + c
        v = int(dn[u])
        if v < 0 or flat_nodes[v] == nid:
            continue
        mid = starts.get(u, np.empty(0, dtype=np.int64))
        mid = mid[: int(ds[u]) - 1]
        path = [(int(r), int(c))] + [(int(i // nx), int(i % nx)) for i in mid] + [(int(v // nx), int(v % nx))]
        edges.append({"id": edge_id, "node_u": nid, "node_v": int(flat_nodes[v]), "length": len(path), "cells": path})
        edge_id += 1
    return edges


def _netxy_cut_host(s_host, nodes, fd, node_map):
    """A node set that is not basin_netxy_find's: follow each path on the host (streams.py:336-384)."""
    off = {0: (0, 1), 1: (1, 1), 2: (1, 0), 3: (1, -1), 4: (0, -1), 5: (-1, -1), 6: (-1, 0), 7: (-1, 1)}
    ny, nx = s_host.shape
    edges, edge_id = [], 1
    for nid, (r, c) in nodes.items():
        o = off.get(int(fd[r, c]))
        if not o:
            continue
        rr, cc = r + o[0], c + o[1]
        if not (0 <= rr < ny and 0 <= cc < nx) or not s_host[rr, cc]:
            continue
        path = [(r, c)]
        for _ in range(ny * nx):
            path.append((rr, cc))
            if node_map[rr, cc] and node_map[rr, cc] != nid:
                edges.append({"id": edge_id, "node_u": nid, "node_v": int(node_map[rr, cc]), "length": len(path), "cells": path})
                edge_id += 1
                break
            o = off.get(int(fd[rr, cc]))
            if not o:
                break
            rr, cc = rr + o[0], cc + o[1]
            if not (0 <= rr < ny and 0 <= cc < nx):
                break
    return edges
