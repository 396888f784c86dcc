"""R-sharding of the ensemble across the GPUs of one box (one process per GPU, torch.distributed
over NCCL / NVLink).  SURVEY.md section 8(e): windows are independent, so rank g scans its own
rows [row_offset, row_offset + R_local) with replicated queries, and ONE small exchange -- an
all-gather of the per-rank (distance, [trajectory, offset]) records, B*k*12 bytes per rank --
lets every rank merge G*k -> k by (distance bits, global flat index).  The merge of exact
per-shard top-k's is exactly the global top-k, so results are bit-identical to one GPU holding
the whole ensemble (replaces the cross-split merge of path_shadowing.py:170-173).  Winning
paths are gathered by their owner and assembled with an all-reduce (x + 0 == x exactly).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib

_INF = float("inf")
_PAD_ROW = 2 ** 31 - 1


def shard_bounds(R: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of R rows: the first R % world ranks hold one extra row."""
    base, extra = divmod(R, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_scan(ps, rows: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int):
    """Local exact top-k on this rank's rows, all-gather, merge.  Every rank returns the global
    (dist (B,k), idx (B,k,2)) with GLOBAL trajectory indices."""
    pg = ps._pg
    world = dist.get_world_size(pg)
    B, W = q.shape
    Tp = T - W - H + 1
    if Tp <= 0:
        raise RuntimeError(f"context ({W}) + horizon ({H}) longer than the trajectories ({T})")
    n_local = rows.shape[0] * Tp
    tot = torch.tensor([n_local], dtype=torch.int64, device=rows.device)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=pg)
    if k > int(tot.item()):
        raise RuntimeError(f"selected index k out of range: k={k} > {int(tot.item())} windows")
    k_loc = min(k, n_local)
    d_loc = torch.full((B, k), _INF, dtype=torch.float32, device=rows.device)
    i_loc = torch.empty((B, k, 2), dtype=torch.int32, device=rows.device)
    i_loc[..., 0] = _PAD_ROW
    i_loc[..., 1] = 0
    if k_loc > 0:
        mode, aux = ps._mode_and_aux(rows, T, W, H)
        d, i, ps._workspace = _lib.scan_topk(rows, T, q, H, k_loc, ps._row_offset, mode, ps._workspace, aux)
        d_loc[:, :k_loc] = d
        i_loc[:, :k_loc] = i
    d_all = torch.empty((world, B, k), dtype=torch.float32, device=rows.device)
    i_all = torch.empty((world, B, k, 2), dtype=torch.int32, device=rows.device)
    # all_gather on views of one buffer: NCCL coalesces it into a single ncclAllGather; gloo
    # (CPU tests) has no all_gather_into_tensor
    dist.all_gather(list(d_all.unbind(0)), d_loc, group=pg)
    dist.all_gather(list(i_all.unbind(0)), i_loc, group=pg)
    return _lib.merge_topk(d_all, i_all, Tp)


def sharded_gather(ps, rows: torch.Tensor, T: int, idx: torch.Tensor, L: int) -> torch.Tensor:
    """Each rank copies the winners it owns (zeros elsewhere); the sum over ranks is the result."""
    paths = _lib.gather_paths(rows, T, idx, L, ps._row_offset)
    dist.all_reduce(paths, op=dist.ReduceOp.SUM, group=ps._pg)
    return paths
