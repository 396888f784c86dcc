"""Statistics of out-context paths (reference: shadowing/statistics.py)."""
from __future__ import annotations

from typing import Iterable

import numpy as np


def realized_variance(x: np.ndarray, Ts: Iterable, vol: bool):
    """Annualised realised variance of log-returns x (..., T) at maturities Ts -> (..., len(Ts)).
    Same contract as statistics.py:5-16 (mean of squares over the first T steps, times 252;
    square root if `vol`)."""
    sq = np.square(x)
    out = np.stack([sq[..., :T].mean(-1) for T in Ts], axis=-1) * 252
    return out ** 0.5 if vol else out


class RealizedVariance:
    """A `to_predict` callable equal to
    `lambda x: realized_variance(x, Ts, vol)[:, :, 0, :]` (README.md:77-80), which
    `PathShadowing.predict_from_paths` / `predict` recognise and evaluate with the fused
    realised-variance + aggregation CUDA kernel instead of on the host."""

    def __init__(self, Ts: Iterable, vol: bool = False):
        self.Ts = [int(t) for t in Ts]
        self.vol = bool(vol)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        return realized_variance(x, self.Ts, self.vol)[:, :, 0, :]
