"""Embeddings and context managers: the plugin surface of the reference's
shadowing/path_shadowing/path_embedding.py, kept name- and signature-compatible.

On the B200 path the embedding is never *applied* to the dataset: `Identity(W)` tells the scan
that the embedded window IS the raw window (the reference's conv1d with eye(W),
path_embedding.py:129-139), and the context manager tells it how many trailing samples of each
window are out-of-context (pad_context, path_embedding.py:48-51).  `forward` is still provided
(user code embeds small tensors for plots and checks, e.g. tutorial.ipynb cell 8).
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
