"""Dataset ingestion: a stand-in for scatspectra's `TimeSeriesDataset` on the one call site of the
path -- `PathShadowing.__init__` (path_shadowing.py:84-87: `TimeSeriesDataset(dpath=..., R=None)
.load()`; README.md:41-42: `TimeSeriesDataset(dpath, R=32768)`).

scatspectra (RudyMorel/scattering_spectra v2.0.2) is not vendored in the reference tree, so its
loader cannot be pinned; this class reads the on-disk format the reference's own scripts write --
a directory of `.npy` files, one trajectory or one batch of trajectories per file
(scripts/snp_generation.py:39-50 via scatspectra.generate's cache, scripts/batch_generations.py:28-40
`batchNNNN.npy` = np.concatenate of 256 files) -- in file-name order, into ONE (R, C, T) float32
array (a single allocation filled from memory-mapped files, no list-and-concatenate copy).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def _as_rct(a: np.ndarray) -> np.ndarray:
    """(T,) -> (1,1,T); (n,T) -> (n,1,T); (n,C,T) unchanged (path_shadowing.py:16-26 `_dim_array`)."""
    if a.ndim == 1:
        return a[None, None, :]
    if a.ndim == 2:
        return a[:, None, :]
    if a.ndim == 3:
        return a
    raise ValueError(f"cannot read an array of shape {a.shape} as (R, C, T) trajectories")


class TimeSeriesDataset:
    """A directory of `.npy` trajectory files, loaded lazily.

    :param dpath: directory holding the `.npy` files
    :param R: number of trajectories to load (None: all of them)
    """

    def __init__(self, dpath: str | Path, R: int | None = None):
        self.dpath = Path(dpath)
        self.R = R
        self._files = None

    def files(self) -> list[Path]:
        if self._files is None:
            self._files = sorted(self.dpath.glob("*.npy"))
            if not self._files:
                raise FileNotFoundError(f"no .npy files under {self.dpath}")
        return self._files

    def load(self) -> np.ndarray:
        """(R, C, T) float32 array of the first R trajectories in file-name order."""
        maps = [_as_rct(np.load(f, mmap_mode="r")) for f in self.files()]
        C, T = maps[0].shape[1:]
        for f, m in zip(self.files(), maps):
            if m.shape[1:] != (C, T):
                raise ValueError(f"{f.name}: trajectories of shape {m.shape[1:]}, expected {(C, T)}")
        total = sum(m.shape[0] for m in maps)
        R = total if self.R is None else int(self.R)
        if R > total:
            raise ValueError(f"{self.dpath} holds {total} trajectories, R={R} requested")
        out = np.empty((R, C, T), np.float32)
        r = 0
        for m in maps:
            n = min(m.shape[0], R - r)
            if n <= 0:
                break
            out[r:r + n] = m[:n]
            r += n
        return out

    def __len__(self) -> int:
        total = sum(_as_rct(np.load(f, mmap_mode="r")).shape[0] for f in self.files())
        return total if self.R is None else min(int(self.R), total)
