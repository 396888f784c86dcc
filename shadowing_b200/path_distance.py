"""Distance plugins: the surface of the reference's shadowing/path_shadowing/path_distance.py.

`RelativeMSE` marks the distance the CUDA scan evaluates (||x-y||_2 / ||x||_2,
path_distance.py:62-65).  `forward` works on any torch tensors (user-side checks such as
testing.ipynb:62-78); `forward_topk` (path_distance.py:10-49) is kept for API compatibility.
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Tuple

import torch
import torch.nn as nn


class PathDistance(nn.Module):

    def forward_topk(self, x: torch.Tensor, y: torch.Tensor, k: int,
                     n_splits: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """k smallest distances between x (B1, d) and every y[i1, ..., :] of y (B2, ..., d).

        Returns (B1, k) distances ascending and (B1, k, y.ndim-1) indices, invariant to `n_splits` and
        prefix-consistent in k (the property testing.ipynb:43-53 asserts).  The indices are int64 as the
        live reference returns them (its int32 buffer of path_distance.py:29 is promoted by the torch.cat
        of :40 with the int64 product indices; pinned by tests/golden/forward_topk_B8_d34.npz); slots that
        no distance ever filled (k > number of entries) keep the reference's fill value 2147483647."""
        lead = y.shape[:-1]
        n_tot = 1
        for s in lead:
            n_tot *= s
        best_d = x.new_full((x.shape[0], k), float("inf"))
        best_i = torch.full((x.shape[0], k), torch.iinfo(torch.int64).max, dtype=torch.int64, device=x.device)
        xq = x.view((x.shape[0],) + (1,) * (y.ndim - 1) + (x.shape[-1],))
        step = max(y.shape[0] // n_splits, 1)
        per_row = n_tot // y.shape[0]
        for r0 in range(0, y.shape[0], step):
            blk = y[None, r0:r0 + step]
            d = self(xq, blk).reshape(x.shape[0], -1)
            flat = torch.arange(r0 * per_row, r0 * per_row + d.shape[1], device=x.device).expand_as(d)
            cd = torch.cat([best_d, d], 1)
            ci = torch.cat([best_i, flat], 1)
            # total order (distance, flat index): deterministic under ties and split-invariant
            order = torch.argsort(ci, dim=1, stable=True)
            cd, ci = cd.gather(1, order), ci.gather(1, order)
            order = torch.argsort(cd, dim=1, stable=True)[:, :k]
            best_d, best_i = cd.gather(1, order), ci.gather(1, order)
        coords = []
        rem = best_i
        for s in reversed(lead):
            coords.append(rem % s)
            rem = rem // s
        idces = torch.stack(coords[::-1], dim=-1)
        idces[best_i == torch.iinfo(torch.int64).max] = 2147483647
        return best_d, idces

    @abstractmethod
    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Distance between x (..., d) and y (..., d) over the last axis."""


class RelativeMSE(PathDistance):

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return (x - y).norm(dim=-1) / x.norm(dim=-1)
