"""Dataset ingestion: a stand-in for scatspectra's `TimeSeriesDataset` on the one call site of the
path -- `PathShadowing.__init__` (path_shadowing.py:84-87: `TimeSeriesDataset(dpath=..., R=None)
.load()`; README.md:41-42: `TimeSeriesDataset(dpath, R=32768)`).

scatspectra (RudyMorel/scattering_spectra v2.0.2) is not vendored in the reference tree, so its
loader cannot be pinned; this class reads the on-disk format the reference's own scripts write --
a directory of `.npy` files, one trajectory or one batch of trajectories per file
(scripts/snp_generation.py:39-50 via scatspectra.generate's cache, scripts/batch_generations.py:28-40
`batchNNNN.npy` = np.concatenate of 256 files) -- in file-name order, into ONE (R, C, T) float32
array (a single allocation filled from memory-mapped files, no list-and-concatenate copy).

`to_device` is the B200-side loader: the files are streamed through two pinned host buffers straight
into the resident (R * C, row_stride) device rows -- the disk read of chunk i + 1 overlaps the
host-to-device copy of chunk i, and no (R, C, T) host array is ever built (8 GiB at BASELINE configs[3]).
`PathShadowing(..., dataset=TimeSeriesDataset(...), stream_dataset=True)` uses it.
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

    # ------------------------------------------------------------------ array-likeness (lazy)
    @property
    def shape(self) -> tuple[int, int, int]:
        m = _as_rct(np.load(self.files()[0], mmap_mode="r"))
        return (len(self), m.shape[1], m.shape[2])

    def __array__(self, dtype=None, copy=None):
        a = self.load()
        return a if dtype is None else a.astype(dtype, copy=False)

    # ------------------------------------------------------------------ pinned streaming upload
    def chunks(self, chunk_rows: int):
        """Yield (first row, (n, C, T) float32 array-like) over the first R trajectories in file-name order,
        at most `chunk_rows` trajectories at a time (memory-mapped: nothing is read before it is copied)."""
        R = len(self)
        r = 0
        for f in self.files():
            m = _as_rct(np.load(f, mmap_mode="r"))
            for a in range(0, m.shape[0], chunk_rows):
                n = min(chunk_rows, m.shape[0] - a, R - r)
                if n <= 0:
                    return
                yield r, m[a:a + n]
                r += n

    def to_device(self, device, chunk_bytes: int = 64 << 20):
        """Stream the trajectories into device memory: returns ((R * C, row_stride) float32 rows with
        row_stride % 4 == 0 -- channel c of trajectory r is row r * C + c --, T, C).  Two pinned staging
        buffers and a copy stream: reading chunk i + 1 from disk overlaps the upload of chunk i."""
        import torch
        R, C, T = self.shape
        stride = (T + 3) // 4 * 4
        dev = torch.device(device)
        rows = torch.zeros((R * C, stride), dtype=torch.float32, device=dev)
        chunk_rows = max(1, chunk_bytes // (C * T * 4))
        pin = dev.type == "cuda"
        stage = [torch.empty((chunk_rows, C, T), dtype=torch.float32, pin_memory=pin) for _ in range(2)]
        if not pin:   # (CPU tensors: tests without a device)
            for r0, a in self.chunks(chunk_rows):
                rows.view(R, C, stride)[r0:r0 + a.shape[0], :, :T] = torch.from_numpy(np.array(a, dtype=np.float32))
            return rows, T, C
        copy_stream = torch.cuda.Stream(device=dev)
        copy_stream.wait_stream(torch.cuda.current_stream(dev))   # (the zero fill of `rows`)
        done = [None, None]
        for i, (r0, a) in enumerate(self.chunks(chunk_rows)):
            buf = stage[i & 1]
            if done[i & 1] is not None:
                done[i & 1].synchronize()                 # the upload that last used this buffer
            n = a.shape[0]
            np.copyto(buf[:n].numpy(), a, casting="same_kind")   # disk (page cache) -> pinned
            with torch.cuda.stream(copy_stream):
                rows.view(R, C, stride)[r0:r0 + n, :, :T].copy_(buf[:n], non_blocking=True)
                done[i & 1] = torch.cuda.Event()
                done[i & 1].record(copy_stream)
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
        return rows, T, C
