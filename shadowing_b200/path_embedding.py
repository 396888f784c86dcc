"""Embeddings and context managers: the plugin surface of the reference's
shadowing/path_shadowing/path_embedding.py, kept name- and signature-compatible.

On the B200 path the embedding is never *applied* to the dataset: `Identity(W)` tells the scan
that the embedded window IS the raw window (the reference's conv1d with eye(W),
path_embedding.py:129-139); any other linear kernel (`Foveal`, `PathEmbedding(kernel)`) is
decomposed into runs of equal taps (`kernel_runs`) that the embedded scan evaluates from prefix
sums in shared memory; and the context manager tells it how many trailing samples of each
window are out-of-context (pad_context, path_embedding.py:48-51).  `forward` is still provided
(the scan embeds the few query windows with it, exactly as the reference does, and user code
embeds small tensors for plots and checks, e.g. tutorial.ipynb cell 8).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ArrayType = np.ndarray | torch.Tensor


class ContextManagerBase:
    """Splits a path into in-context (shadowed) and out-context (predicted) parts.
    Mirrors path_embedding.py:13-30."""

    def select_in_context(self, x: ArrayType) -> ArrayType:
        raise NotImplementedError

    def select_out_context(self, x: ArrayType) -> ArrayType:
        raise NotImplementedError

    def pad_context(self, x_in_context: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def get_out_times(self):
        raise NotImplementedError


class PredictionContext(ContextManagerBase):
    """in-context = the past, out-context = the next `horizon` steps (path_embedding.py:33-56)."""

    def __init__(self, horizon: int | None = None):
        self.horizon = horizon

    def select_in_context(self, x: ArrayType) -> ArrayType:
        return x if self.horizon is None else x[..., :-self.horizon]

    def select_out_context(self, x: ArrayType) -> ArrayType:
        return x if self.horizon is None else x[..., -self.horizon:]

    def pad_context(self, x_in_context: torch.Tensor) -> torch.Tensor:
        return x_in_context if self.horizon is None else F.pad(x_in_context, (0, self.horizon))

    def get_out_times(self):
        return 0 if self.horizon is None else self.horizon


class ImputationContext(ContextManagerBase):
    """in-context = the first `l` and the last `r` samples of a window, out-context = the `c` samples in
    between (`portion = (l, c, r)`; path_embedding.py:59-87).  The scan compares the l + r context
    samples only: the embedding kernel is padded with `c` zero taps in the middle (`pad_context`)."""

    def __init__(self, portion: tuple | None = None):
        self.portion = portion

    def select_in_context(self, x: ArrayType) -> ArrayType:
        if self.portion is None:
            return x
        l, _, r = self.portion
        if isinstance(x, torch.Tensor):
            return torch.cat([x[..., :l], x[..., -r:]], dim=-1)
        return np.concatenate([x[..., :l], x[..., -r:]], axis=-1)

    def select_out_context(self, x: ArrayType) -> ArrayType:
        if self.portion is None:
            return x
        l, _, r = self.portion
        return x[..., l:-r]

    # the reference spells it `slect_out_context` (path_embedding.py:71); both names work here
    slect_out_context = select_out_context

    def pad_context(self, x_in_context: torch.Tensor) -> torch.Tensor:
        if self.portion is None:
            return x_in_context
        l, c, r = self.portion
        zeros_middle = x_in_context.new_zeros(x_in_context.shape[:-1] + (c,))
        return torch.cat([x_in_context[..., :l], zeros_middle, x_in_context[..., -r:]], dim=-1)

    def get_out_times(self):
        return 0 if self.portion is None else self.portion[1]


class CrossChannelContext(ContextManagerBase):
    """in-context = the first channels of a (.., C, T) path, out-context = its last `out_context_channels`
    channels (path_embedding.py:90-114): the scan compares the in-context channel, the shadowing paths
    come back with all channels."""

    def __init__(self, out_context_channels: int):
        self.out_context_channels = out_context_channels

    def select_in_context(self, x: ArrayType) -> ArrayType:
        return x[..., :x.shape[-2] - self.out_context_channels, :]

    def select_out_context(self, x: ArrayType) -> ArrayType:
        if self.out_context_channels is None:
            return x
        return x[..., -self.out_context_channels:, :]

    def pad_context(self, x_in_context: torch.Tensor) -> torch.Tensor:
        if self.out_context_channels is None:
            return x_in_context
        shape = list(x_in_context.shape)
        shape[-2] = self.out_context_channels
        return torch.cat([x_in_context, x_in_context.new_zeros(shape)], dim=-2)

    def get_out_times(self):
        return 0


class PathEmbedding(nn.Module):
    """Linear embedding given by a (d, 1, W) kernel buffer (path_embedding.py:117-132)."""

    def __init__(self, kernel: torch.Tensor):
        super().__init__()
        self.register_buffer("kernel", kernel)

    def adjust_to_context(self, context: ContextManagerBase) -> "PathEmbedding":
        return PathEmbedding(context.pad_context(self.kernel))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        # (b, 1, t) -> (b, t', d): every length-W window projected on the d kernel rows
        return F.conv1d(x, self.kernel).transpose(1, 2)


class Identity(PathEmbedding):
    """The embedded window is the window itself (path_embedding.py:135-139)."""

    def __init__(self, dimension: int):
        self.d = dimension
        super().__init__(torch.eye(dimension)[:, None, :])


class Foveal(PathEmbedding):
    """Foveal embedding: the context seen at a resolution that coarsens with the distance to the
    present -- `dim = floor(log(max_context) / log(alpha))` trailing box sums of lengths
    `int(alpha ** n)`, n = 1..dim, each weighted by `length ** -beta`.  Same constructor, attributes
    (`alpha`, `beta`, `max_context`, `dim`, `slices`) and kernel as path_embedding.py:142-172."""

    def __init__(self, alpha: float, beta: float, max_context: int, device: str = "cpu"):
        self.alpha, self.beta, self.max_context = alpha, beta, max_context
        self.dim = int(np.floor(np.log(max_context) / np.log(alpha)))
        lengths = np.array([int(alpha ** n) for n in range(1, self.dim + 1)], dtype=np.int64)
        self.slices = [slice(-int(le), None) for le in lengths]
        # row n: weight length**-beta on the last `length` taps, zero before (one trailing box per row)
        weights = np.array([int(le) ** (-beta) for le in lengths], dtype=np.float64)
        trailing = np.arange(max_context)[None, :] >= (max_context - lengths)[:, None]
        kernel = torch.tensor(np.where(trailing, weights[:, None], 0.0)[:, None, :], dtype=torch.float32, device=device)
        super().__init__(kernel)


def kernel_runs(kernel: torch.Tensor) -> np.ndarray:
    """Decompose a (d, 1, W) embedding kernel into runs of equal non-zero taps: a structured array
    of (row, a, b, c) with kernel[row, 0, a:b] == c, rows ascending -- what the embedded scan
    evaluates as c * (P[t+b] - P[t+a]) on a prefix sum P.  Foveal: one run per row."""
    K = kernel.detach().cpu().numpy()
    if K.ndim != 3 or K.shape[1] != 1:
        raise RuntimeError(f"expected a (d, 1, W) embedding kernel, got {tuple(K.shape)}")
    K = K[:, 0, :].astype(np.float32)
    runs = []
    for n in range(K.shape[0]):
        row = K[n]
        cuts = np.flatnonzero(np.diff(row.view(np.uint32)) != 0) + 1   # bit-pattern changes (NaN-safe)
        starts = np.concatenate(([0], cuts))
        ends = np.concatenate((cuts, [row.shape[0]]))
        for a, b in zip(starts, ends):
            if row[a] != 0.0:
                runs.append((n, int(a), int(b), float(row[a])))
    dt = np.dtype([("row", np.int32), ("a", np.int32), ("b", np.int32), ("c", np.float32)])
    return np.array(runs, dtype=dt)
