"""View-sharded data parallelism: the host-side logic, independent of the device.

Views of one step are partitioned round-robin over the ranks; every rank accumulates the gradients of its views
in a flat planar arena; one all-reduce(sum) makes the arenas identical; every rank then applies the same Adam
update, so parameters stay bit-identical without a broadcast.  Backend: NCCL on the GPUs (NVLink 5 / NVSwitch),
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def shard_views(n_views: int, world_size: int, rank: int, costs: Sequence[float] | None = None) -> List[int]:
    """Views handled by `rank`.  Without costs: rank, rank + world_size, ...  (view v -> rank v mod world_size).
    With per-view cost estimates (e.g. the tile-instance counts of the cameras from the previous epoch -- the blend
    kernels, 60 % of a view, scale with them): longest-processing-time-first assignment with equal view counts per rank,
    so that the slowest rank of a step, which every other rank waits for at the exchange barrier, is as fast as possible.
    Deterministic: every rank computes the same partition from the same costs."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if costs is None:
        return list(range(rank, n_views, world_size))
    if len(costs) != n_views:
        raise ValueError(f"{len(costs)} costs for {n_views} views")
    quota = [(n_views + world_size - 1 - r) // world_size for r in range(world_size)]  # same counts as round-robin
    load = [0.0] * world_size
    mine: List[List[int]] = [[] for _ in range(world_size)]
    for v in sorted(range(n_views), key=lambda i: (-float(costs[i]), i)):
        r = min((r for r in range(world_size) if len(mine[r]) < quota[r]), key=lambda r: (load[r], r))
        mine[r].append(v)
        load[r] += float(costs[v])
    return sorted(mine[rank])


def plane_table(K: int) -> dict:
    """First plane index of every parameter group of the planar arena (group order of the reference,
    strategies/strategy_utils.cpp:35-40): means 3 | sh0 3 | shN 3(K-1) | scaling 3 | rotation 4 | opacity 1."""
    t, o = {}, 0
    for name, n in (("means", 3), ("sh0", 3), ("shN", 3 * (K - 1)), ("scaling", 3), ("rotation", 4), ("opacity", 1)):
        t[name] = (o, n)
        o += n
    t["_total"] = o
    return t


def pack_planar(groups: dict, n: int, K: int) -> np.ndarray:
    """AoS tensors in the reference's SplatData layout -> planar arena [(11+3K) * N_pad] (what lfs_trainer_pack
    does on the device; used by the CPU tests and for checkpoint import)."""
    n_pad = (n + 3) // 4 * 4
    tab = plane_table(K)
    out = np.zeros((tab["_total"], n_pad), dtype=np.float32)
    for name in ("means", "sh0", "shN", "scaling", "rotation", "opacity"):
        first, cnt = tab[name]
        if cnt:
            out[first:first + cnt, :n] = np.asarray(groups[name], np.float32).reshape(n, cnt).T
    return out.reshape(-1)


def unpack_planar(arena: np.ndarray, n: int, K: int) -> dict:
    n_pad = (n + 3) // 4 * 4
    tab = plane_table(K)
    a = np.asarray(arena).reshape(tab["_total"], n_pad)
    shapes = {"means": (n, 3), "sh0": (n, 1, 3), "shN": (n, K - 1, 3), "scaling": (n, 3), "rotation": (n, 4),
              "opacity": (n, 1)}
    return {name: a[tab[name][0]:tab[name][0] + tab[name][1], :n].T.reshape(shapes[name]).copy()
            for name in shapes}


def allreduce_sum_(tensor) -> None:
    """In-place sum over all ranks of the default process group (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
