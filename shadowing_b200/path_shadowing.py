"""PathShadowing on B200: drop-in for the reference's shadowing/path_shadowing/path_shadowing.py.

Same class, method names, argument order, defaults, return types and error behaviour as the
reference (`PathShadowing.shadow`, `batched_distance`, `predict_from_paths`, `predict`,
`init_averaging_proba`; path_shadowing.py:61-301), but the scan runs as hand-written sm_100a
CUDA through libpshadow.so (include/pshadow.h):

* the ensemble is uploaded ONCE (first use) and stays resident in HBM as (R, row_stride) fp32
  rows -- the reference re-copies it on every call (path_shadowing.py:204-205,152-155);
* windows are never materialised (the reference's conv1d blows the data up W times), so
  `n_splits` / `n_dataset_splits` are accepted and ignored;
* `cuda` is accepted and ignored: this implementation has no CPU path.

Linear embeddings (Identity: exact/filter/fft scans; Foveal and any PathEmbedding(kernel): the
embedded scan) + RelativeMSE + PredictionContext run on the device; any other plugin combination
raises NotImplementedError (no silent fallback).
"""
from __future__ import annotations

import os
import weakref
from pathlib import Path
from typing import Callable

import numpy as np
import torch

from . import _lib
from .averaging import DiscreteProba, Softmax, Uniform
from .dataset import TimeSeriesDataset
from .path_distance import PathDistance, RelativeMSE
from .path_embedding import (ArrayType, ContextManagerBase, CrossChannelContext, Foveal, Identity, ImputationContext,
                             PathEmbedding, PredictionContext, kernel_runs)
from .statistics import RealizedVariance


def _dim_array(x: ArrayType) -> ArrayType:
    """Bring x to (B, C, T): 1-D is a single series, 2-D is (batch, time) (path_shadowing.py:16-26)."""
    if x is None:
        return x
    if x.ndim == 1:
        return x[None, None, :]
    if x.ndim == 2:
        return x[:, None, :]
    if x.ndim == 3:
        return x
    raise Exception("Array cannot be formatted to (B, C, T) shape.")


def _torch(x: ArrayType) -> torch.Tensor:
    """numpy (any float dtype) -> float32 tensor; tensors pass through (path_shadowing.py:29-33)."""
    if isinstance(x, torch.Tensor):
        return x
    return torch.tensor(x, dtype=torch.float32)


def _numpy(x: ArrayType) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return x
    return x.cpu().numpy()


def select_cartesian_product(indices: torch.Tensor, tensors: list[torch.Tensor]) -> torch.Tensor:
    """Rows `indices` of the cartesian product of `tensors` without building it
    (path_shadowing.py:43-58): flat index -> mixed-radix coordinates -> per-axis lookup."""
    sizes = [int(t.shape[0]) for t in tensors]
    coords = []
    rem = indices
    for n in reversed(sizes):
        coords.append(rem % n)
        rem = torch.div(rem, n, rounding_mode="floor")
    coords.reverse()
    return torch.stack([t[c] for t, c in zip(tensors, coords)], dim=-1)


def _recycle(pool: list, outstanding: list, host_buf: torch.Tensor) -> None:
    """Finaliser of a result buffer handed out by `shadow`: every numpy view of it is gone."""
    outstanding[0] -= 1
    pool.append(host_buf)


class PathShadowing:
    """Path shadowing: scan a generated dataset for the paths closest to an observed context.

    Attributes (as in the reference, path_shadowing.py:61-95):
        embedding, distance, dataset, context.
    Extra keyword-only arguments (all optional) select the device and, for an ensemble sharded
    over the GPUs of one box, describe this rank's shard (see shadowing_b200.distributed).
    """

    def __init__(
        self,
        embedding: PathEmbedding,
        distance: PathDistance,
        dataset,
        context: ContextManagerBase | None = None,
        *,
        device: torch.device | str | None = None,
        row_offset: int = 0,
        process_group=None,
        scan_mode: str = "auto",
        stream_dataset: bool = False,
    ):
        if isinstance(dataset, (str, Path)):
            dataset = TimeSeriesDataset(dpath=dataset, R=None)              # path_shadowing.py:84-85
        if hasattr(dataset, "load") and not isinstance(dataset, (np.ndarray, torch.Tensor)):
            # a TimeSeriesDataset(-like) object (path_shadowing.py:86-87): loaded into a host array as the
            # reference does, or -- stream_dataset=True -- kept as it is and streamed file by file through
            # pinned buffers into the resident device rows (no host copy of the ensemble)
            if not (stream_dataset and hasattr(dataset, "to_device")):
                dataset = dataset.load()
        self.dataset = dataset
        self.embedding = embedding
        self.distance = distance
        self.context = context or PredictionContext(horizon=None)

        self._device = torch.device(device) if device is not None else None
        self._row_offset = int(row_offset)
        self._pg = process_group
        if scan_mode not in ("auto", "fft", "filter", "exact"):
            raise ValueError("scan_mode must be 'auto', 'fft', 'filter' or 'exact'")
        # every mode returns bit-identical results; they differ in how candidates are filtered:
        #   exact  -- every window with the reference's sub/mul/add sequence
        #   filter -- 1 FMA per element lower bound, exact re-rank of survivors
        #   fft    -- lower bound through one inverse 4096-point FFT per pair of trajectory pieces
        #   auto   -- fft when the context is long enough to pay for it, else filter
        self._scan_mode = scan_mode
        self._resident = None  # (key, device rows (R, row_stride), T)
        self._workspace = None
        self._fft_aux = None   # (key, aux buffer) for (resident rows, W, H)
        self._staging = None   # pinned host buffers for the results of shadow()
        self._runs = None      # (key, device run table) of a non-Identity embedding kernel
        self._pipeline_B = 0   # queries of the last enqueue-only scan on the main workspace
        self._side = None      # side streams [stream, workspace, B] of pipelined enqueue-only scans
        self._pipe_streams = max(1, int(os.environ.get("PSH_STREAMS", "1")))
        self._lanes = None     # lanes (stream + per-stream state) of pipelined enqueue-only SHARDED scans
        self._padded_kernel = None   # (key, kernel padded by an ImputationContext)

    # ------------------------------------------------------------------ device residency
    def _dev(self) -> torch.device:
        _lib.require_cuda()
        if self._device is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    def _check_plugins(self) -> None:
        if not isinstance(self.embedding, PathEmbedding) or self.embedding.kernel.dim() != 3:
            raise NotImplementedError(
                f"the B200 scan implements linear embeddings (PathEmbedding with a (d,1,W) kernel: Identity, "
                f"Foveal, ...); got {type(self.embedding).__name__}")
        if type(self.distance) is not RelativeMSE:
            raise NotImplementedError(
                f"the B200 scan implements the RelativeMSE distance; got {type(self.distance).__name__}")
        if type(self.context) not in (PredictionContext, ImputationContext, CrossChannelContext):
            raise NotImplementedError(
                f"the B200 scan implements PredictionContext, ImputationContext and CrossChannelContext; "
                f"got {type(self.context).__name__}")

    def _imputation(self) -> bool:
        return type(self.context) is ImputationContext and self.context.portion is not None

    def _scan_kernel(self) -> torch.Tensor:
        """The kernel the DATASET is embedded with (`embedding.adjust_to_context(context)`,
        path_shadowing.py:140): ImputationContext pads the kernel with zero taps in the middle -- the
        embedded scan skips them (runs of equal non-zero taps) -- any other context leaves the taps that
        matter where they are (PredictionContext pads zeros behind them: the horizon H of the scan)."""
        kernel = self.embedding.kernel
        if not self._imputation():
            return kernel
        key = (kernel.data_ptr(), kernel._version, tuple(self.context.portion))
        if self._padded_kernel is None or self._padded_kernel[0] != key:
            self._padded_kernel = (key, self.context.pad_context(kernel.detach()).contiguous())
        return self._padded_kernel[1]

    def _scan_horizon(self) -> int:
        """Trailing out-of-context samples of a window (0 for the contexts whose out-context is elsewhere)."""
        return self.context.get_out_times() if type(self.context) is PredictionContext else 0

    def _identity_scan(self) -> bool:
        """Raw windows compared sample by sample: Identity without zero taps in the middle."""
        return type(self.embedding) is Identity and not self._imputation()

    def _rows_to_device(self, y: ArrayType, dev: torch.device) -> tuple[torch.Tensor, int]:
        """(R, C, T) host/device array -> resident fp32 rows, row stride % 4 == 0 so every row starts
        16-byte aligned for the TMA bulk copies.  Returns the rows the scan reads -- channel 0, a
        (R, C * row_stride)-strided view -- and T; all channels stay reachable through `._channels`."""
        want = 1 + self.context.out_context_channels if type(self.context) is CrossChannelContext else 1
        if hasattr(y, "to_device") and not isinstance(y, (np.ndarray, torch.Tensor)):   # streamed from its files
            allc, T, C = y.to_device(dev)
            if C != want:
                raise RuntimeError(f"expected a dataset with {want} channel(s), the files hold {C}")
            rows = allc[0::C] if C > 1 else allc
            rows._channels = (allc, C)
            return rows, T
        y = _dim_array(y)
        C = y.shape[1]
        if C != want:
            raise RuntimeError(
                f"expected a dataset with {want} channel(s) (R, {want}, T), got {tuple(y.shape)}: the embedding is a "
                "conv1d over the in-context channel (path_embedding.py:130, CrossChannelContext.pad_context)")
        y = _torch(y)
        if y.dtype != torch.float32:
            raise RuntimeError(f"expected a float32 dataset tensor, got {y.dtype}")
        R, _, T = y.shape
        stride = (T + 3) // 4 * 4
        allc = torch.zeros((R * C, stride), dtype=torch.float32, device=dev)
        allc.view(R, C, stride)[:, :, :T].copy_(y, non_blocking=True)
        rows = allc[0::C] if C > 1 else allc
        rows._channels = (allc, C)
        return rows, T

    def _resident_rows(self) -> tuple[torch.Tensor, int]:
        ds = self.dataset
        key = (id(ds), tuple(ds.shape), type(self.context), getattr(self.context, "out_context_channels", None))
        if self._resident is None or self._resident[0] != key:
            rows, T = self._rows_to_device(ds, self._dev())
            self._resident = (key, rows, T)
            self._workspace = None
            self._fft_aux = None
        return self._resident[1], self._resident[2]

    def invalidate(self) -> None:
        """Forget the resident copy of the dataset and everything derived from it (spectra, energies): call
        it after editing `self.dataset` IN PLACE.  The ensemble is uploaded once and kept resident (the
        reference re-reads it on every call); replacing `self.dataset` by another object is noticed."""
        self._resident = None
        self._workspace = None
        self._fft_aux = None

    def _gather(self, rows: torch.Tensor, T: int, idx: torch.Tensor, L: int, out=None) -> torch.Tensor:
        """paths (B, k, C, L) of the winners: dataset[r, :, t:t+L] (path_shadowing.py:210-216)."""
        allc, C = getattr(rows, "_channels", (rows, 1))
        if C == 1:
            return _lib.gather_paths(rows, T, idx, L, self._row_offset, out)
        return torch.cat([_lib.gather_paths(allc[c::C], T, idx, L, self._row_offset) for c in range(C)], dim=2)

    def _mode_and_aux(self, rows: torch.Tensor, T: int, W: int, H: int):
        mode = self._scan_mode
        if mode == "auto":
            # the FFT flavour pays off for long contexts on trajectories that fill a 4096-point
            # transform reasonably (longer ones are cut into overlapping pieces)
            mode = "fft" if (64 <= W <= 1024 and T >= 1024) else "filter"
        if mode == "exact":
            return _lib.PSH_MODE_EXACT, None
        if mode == "filter" or W > _lib.FFT_MAX_W:
            return _lib.PSH_MODE_FILTER, None
        # the spectra / window energies are cached for the RESIDENT rows only, and the cache entry holds
        # the rows tensor itself (compared with `is`): a foreign `y` of batched_distance gets a throwaway
        # aux -- a data_ptr() key could be recycled by the caching allocator for the next same-shaped y
        if self._resident is None or rows is not self._resident[1]:
            return _lib.PSH_MODE_FFT, _lib.fft_prepare(rows, T, W, H)
        key = (T, W, H, os.environ.get("PSH_FFT_N", ""))   # (the library picks the transform length from W and this knob)
        if self._fft_aux is None or self._fft_aux[0] != key or self._fft_aux[2] is not rows:
            self._fft_aux = (key, _lib.fft_prepare(rows, T, W, H), rows)
        return _lib.PSH_MODE_FFT, self._fft_aux[1]

    # ------------------------------------------------------------------ scan
    def _scan_device(self, x: torch.Tensor, rows: torch.Tensor, T: int, k: int, out=None, nosync: bool = False):
        """x (B, 1, W) -> device (dist (B,k), idx (B,k,2)); all-reduced across the process group
        when the ensemble is sharded.  `nosync`: enqueue only (a pipeline of scans on one stream);
        the caller must call `_check_pipeline()` before trusting any of the results."""
        self._check_plugins()
        dev = rows.device
        if x.dtype != torch.float32:
            raise RuntimeError(f"expected a float32 context, got {x.dtype}")  # reference: conv1d dtype error
        if x.shape[1] != 1:
            raise RuntimeError(f"expected a single-channel context (B, 1, W), got {tuple(x.shape)}")
        H = self._scan_horizon()
        if not self._identity_scan():
            return self._scan_embedded(x, rows, T, k, H, out, nosync)
        q = x[:, 0, :].to(dev, non_blocking=True).contiguous()
        W = q.shape[1]
        if self._pg is None:
            n_windows = rows.shape[0] * (T - W - H + 1)
            if T - W - H + 1 <= 0:
                raise RuntimeError(f"context ({W}) + horizon ({H}) longer than the trajectories ({T})")
            if k > n_windows:
                raise RuntimeError(f"selected index k out of range: k={k} > {n_windows} windows")
            mode, aux = self._mode_and_aux(rows, T, W, H)
            if nosync and out is None and self._pipe_streams > 1:   # (shadow() brings its own `out`: one stream)
                return self._scan_on_side_stream(rows, T, q, H, k, mode, aux)
            if (self._pipeline_B and self._workspace is not None and self._workspace.numel() <
                    _lib.lib().psh_scan_workspace_bytes(rows.shape[0], T, q.shape[0], W, H, k)):
                self._check_pipeline()   # the workspace is about to be replaced: settle the scans still pending on it
            dist, idx, self._workspace = _lib.scan_topk(
                rows, T, q, H, k, self._row_offset, mode | (_lib.PSH_FLAG_NOSYNC if nosync else 0),
                self._workspace, aux, out)
            self._pipeline_B = max(self._pipeline_B, q.shape[0]) if nosync else q.shape[0]
            return dist, idx
        from .distributed import finish_sharded, sharded_scan
        if nosync and out is None and self._pipe_streams > 1:
            return self._sharded_scan_on_lane(rows, T, q, H, k)
        res = sharded_scan(self, rows, T, q, H, k, defer=nosync)
        return res if nosync else finish_sharded(self, rows, T, q, H, k, res)

    def _sharded_scan_on_lane(self, rows, T, q, H, k):
        """Sharded pipelines of enqueue-only scans on `pipeline_streams` > 1 lanes: the same overlap as
        `_scan_on_side_stream` (query i+1's preparation and scan run while query i's re-rank, select and
        exchange finish), every lane with its own workspace, record buffers and flag.  Every rank
        alternates lanes identically, so the exchange epochs stay aligned."""
        from .distributed import Lane, sharded_scan, sharded_scan_fast
        dev = rows.device
        if self._pipe_streams > 4:
            raise ValueError("sharded pipelines alternate between at most 4 streams (the exchange buffers hold 8 epochs)")
        if self._lanes is None or len(self._lanes) != self._pipe_streams:
            self._lanes = [Lane(self, torch.cuda.Stream(device=dev)) for _ in range(self._pipe_streams)]
            self._lane_i = 0
        lane = self._lanes[self._lane_i % self._pipe_streams]
        self._lane_i += 1
        cur = torch.cuda.current_stream(dev)
        lane.stream.wait_stream(cur)
        q.record_stream(lane.stream)
        out = sharded_scan_fast(lane, rows, T, q, H, k)      # steady state: raw stream handles, no stream switch
        if out is not None:
            dist, idx = out
            dist.record_stream(lane.stream)                  # allocated on `cur`, written on the lane's stream
            idx.record_stream(lane.stream)
            return dist, idx
        with torch.cuda.stream(lane.stream):
            dist, idx = sharded_scan(lane, rows, T, q, H, k, defer=True)
        dist.record_stream(cur)
        idx.record_stream(cur)
        return dist, idx

    def _scan_on_side_stream(self, rows, T, q, H, k, mode, aux):
        """Pipelines of enqueue-only scans, `pipeline_streams` > 1: consecutive queries alternate
        between side streams, each with its own workspace, so that query i+1's prologue and main
        launch overlap query i's re-rank and select (two tiny grids that leave the GPU idle).
        `_check_pipeline` joins the side streams back into the caller's stream."""
        dev = rows.device
        if self._side is None or len(self._side) != self._pipe_streams:
            self._side = [[torch.cuda.Stream(device=dev), None, 0] for _ in range(self._pipe_streams)]
            self._side_i = 0
        slot = self._side[self._side_i % self._pipe_streams]
        self._side_i += 1
        side, cur = slot[0], torch.cuda.current_stream(dev)
        side.wait_stream(cur)                      # the query comes from `cur`
        q.record_stream(side)
        # the scan is enqueued on the side stream through its raw handle (no stream switch on the host:
        # `with torch.cuda.stream(...)` costs more than the four launches); outputs and a first workspace are
        # allocated on `cur` and handed to the side stream
        ws_before = slot[1]
        dist, idx, slot[1] = _lib.scan_topk(rows, T, q, H, k, self._row_offset,
                                            mode | _lib.PSH_FLAG_NOSYNC | _lib.PSH_FLAG_SHARE_SMS, slot[1], aux,
                                            stream=side.cuda_stream)
        if slot[1] is not ws_before:               # a new (first or larger) workspace
            slot[1].record_stream(side)
        slot[2] = q.shape[0]
        dist.record_stream(side)                   # written on `side`, consumed on `cur` after the join
        idx.record_stream(side)
        return dist, idx

    def _run_table(self, device: torch.device):
        """Device run table of a non-Identity embedding kernel (None for Identity)."""
        if self._identity_scan():
            return None
        kernel = self._scan_kernel()
        key = (kernel.data_ptr(), tuple(kernel.shape), kernel._version, str(device))
        if self._runs is None or self._runs[0] != key:
            runs = kernel_runs(kernel)
            if runs.shape[0] == 0:
                raise RuntimeError("the embedding kernel is identically zero")
            words = torch.from_numpy(runs.view(np.int32).reshape(-1, 4).copy())
            self._runs = (key, words.to(device))
        return self._runs[1]

    def _embed_flavour(self, rows: torch.Tensor, T: int, W: int, H: int, ex: torch.Tensor) -> dict:
        """Extra arguments of the embedded scan's fft flavour (empty: exact flavour).  The squared
        embedded distance is ||ex||^2 - 2 g.y_t + ||K y_t||^2 with g = K^T ex: one correlation per
        trajectory (the Identity flavour's spectra) plus a per-(dataset, kernel) energy table."""
        mode = self._scan_mode
        if mode == "auto":
            mode = "fft" if (64 <= W <= 1024 and T >= 1024) else "exact"
        if mode != "fft" or W > _lib.FFT_MAX_W:
            return {}
        kernel = self._scan_kernel()
        runs = self._run_table(rows.device)
        key = (T, W, H, kernel.data_ptr(), kernel._version, os.environ.get("PSH_FFT_N", ""))
        resident = self._resident is not None and rows is self._resident[1]
        if not resident:   # a foreign `y`: throwaway aux (see _mode_and_aux)
            K = kernel.detach().cpu()[:, 0, :].double()
            aux = _lib.fft_prepare_embed(rows, T, W, H, runs)
        else:
            if (self._fft_aux is None or self._fft_aux[0] != key or len(self._fft_aux) != 4
                    or self._fft_aux[2] is not rows):
                K = kernel.detach().cpu()[:, 0, :].double()
                self._fft_aux = (key, _lib.fft_prepare_embed(rows, T, W, H, runs), rows, K)
            _, aux, _, K = self._fft_aux
        g = (ex.detach().cpu().double() @ K).float().to(rows.device, non_blocking=True).contiguous()
        return {"g": g, "aux": aux}

    def _scan_embedded(self, x: torch.Tensor, rows: torch.Tensor, T: int, k: int, H: int, out, nosync: bool):
        """Foveal / PathEmbedding(kernel): the few query windows are embedded on the host with the
        embedding's own forward -- the reference's `embedding(x)[:, 0, :]`, path_shadowing.py:138 --
        and the ensemble is scanned in embedded space from prefix sums (never materialised)."""
        W = int(self._scan_kernel().shape[-1])   # window length in the dataset (l + c + r under an ImputationContext)
        if x.shape[-1] != int(self.embedding.kernel.shape[-1]):
            raise RuntimeError(f"context length {x.shape[-1]} does not match the embedding kernel "
                               f"({int(self.embedding.kernel.shape[-1])})")
        Tp = T - W - H + 1
        if Tp <= 0:
            raise RuntimeError(f"context ({W}) + horizon ({H}) longer than the trajectories ({T})")
        with torch.no_grad():
            ex_host = self.embedding.to(x.device)(x)[:, 0, :]     # (B, d), as the reference embeds the context
        ex = ex_host.to(rows.device, non_blocking=True).contiguous()
        if self._pg is not None:
            from .distributed import finish_sharded, sharded_scan
            self._ex_host = ex_host
            res = sharded_scan(self, rows, T, ex, H, k, W, defer=nosync)
            return res if nosync else finish_sharded(self, rows, T, ex, H, k, res, W)
        if k > rows.shape[0] * Tp:
            raise RuntimeError(f"selected index k out of range: k={k} > {rows.shape[0] * Tp} windows")
        dist, idx, self._workspace = _lib.scan_topk_embed(rows, T, ex, W, H, k, self._run_table(rows.device),
                                                          self._row_offset, nosync, self._workspace, out,
                                                          **self._embed_flavour(rows, T, W, H, ex_host))
        self._pipeline_B = ex.shape[0]
        return dist, idx

    def _check_pipeline(self) -> None:
        """Synchronise behind a pipeline of `nosync` scans; raises if any of them overflowed a
        candidate buffer (adversarially ordered data: those scans must be repeated synchronously)."""
        if self._pg is None:
            bad = False
            if self._side is not None:
                cur = torch.cuda.current_stream(self._dev())
                for slot in self._side:
                    if slot[1] is not None and slot[2] > 0:
                        cur.wait_stream(slot[0])
                        bad = _lib.scan_overflowed(slot[1], slot[2]) or bad
                        slot[2] = 0
            if self._pipeline_B:
                bad = _lib.scan_overflowed(self._workspace, self._pipeline_B) or bad
                self._pipeline_B = 0
        else:
            from .distributed import flush_deferred_merge
            bad = False
            cur = torch.cuda.current_stream(self._dev()) if torch.cuda.is_available() else None
            for holder in [self] + list(self._lanes or []):
                if holder is not self:
                    with torch.cuda.stream(holder.stream):
                        flush_deferred_merge(holder)
                    cur.wait_stream(holder.stream)
                else:
                    flush_deferred_merge(holder)
                flag = getattr(holder, "_pending_flag", None)
                if flag is not None and int(flag.item()) != 0:
                    bad = True
                    flag.zero_()
                holder._pending_flag = None
        if bad:
            raise _lib.PshadowError(_lib.PSH_E_OVERFLOW, "nosync scan pipeline")

    def batched_distance(self, x: torch.Tensor, y: torch.Tensor, k: int, n_splits: int,
                         cuda: bool) -> tuple[torch.Tensor, torch.Tensor]:
        """k smallest distances between the context(s) x (B, C, T) and every window of y
        (S, C, T): CPU tensors (B, k) ascending and (B, k, 2) int32 [trajectory, offset].
        Same contract as path_shadowing.py:97-179; `n_splits` and `cuda` are ignored."""
        del n_splits, cuda
        if y is self.dataset or (self._resident is not None and y is self._resident[1]):
            rows, T = self._resident_rows()
        else:
            rows, T = self._rows_to_device(y, self._dev())
        dist, idx = self._scan_device(_torch(_dim_array(x)), rows, T, k)
        return dist.cpu(), idx.cpu()

    def shadow_device(self, x_context: ArrayType, k: int = 1, _packed: torch.Tensor | None = None,
                      _nosync: bool = False):
        """`shadow` without the device->host copy: (dist (B,k), paths (B,k,1,W+H), idx (B,k,2))
        as CUDA tensors (used by `predict` to keep the whole pipeline on the GPU)."""
        if self.embedding.kernel.shape[-1] != 0 and self.embedding.kernel.shape[-1] != x_context.shape[-1]:
            raise Exception("The embedding kernel should be of the same size as the context.")
        x = _torch(_dim_array(x_context))
        rows, T = self._resident_rows()
        L = x.shape[-1] + self.context.get_out_times()
        out = out_paths = None
        if _packed is not None and self._pg is None:  # views of one device buffer: [dist | idx | paths]
            B = x.shape[0]
            nd, ni = B * k, B * k * 2
            out = (_packed[:nd].view(torch.float32).view(B, k), _packed[nd:nd + ni].view(B, k, 2))
            out_paths = _packed[nd + ni:nd + ni + B * k * L].view(torch.float32).view(B, k, 1, L)
        dist, idx = self._scan_device(x, rows, T, k, out, nosync=_nosync and self._pg is None)
        if self._pg is None:
            paths = self._gather(rows, T, idx, L, out_paths)
        else:
            from .distributed import sharded_gather
            paths = sharded_gather(self, rows, T, idx, L)
        return dist, paths, idx

    def shadow(self, x_context: ArrayType, k: int = 1, n_splits: int = 1,
               cuda: bool = False) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Scan the dataset for the k paths closest to each context (path_shadowing.py:181-218).

        :param x_context: (B, C, T) / (B, T) / (T,) array, the B paths to shadow
        :param k: number of closest paths to keep
        :param n_splits: accepted for compatibility; the scan needs no memory splits
        :param cuda: accepted for compatibility; the scan always runs on the GPU
        :return: numpy (distances (B,k) f32 ascending, paths (B,k,C,W+H) f32, indices (B,k,2) i32)
        """
        del n_splits, cuda
        if self._pg is not None and torch.cuda.is_available() and type(self.context) is not CrossChannelContext:
            return self._shadow_sharded(x_context, k)
        if self._pg is not None or not torch.cuda.is_available() or type(self.context) is CrossChannelContext:
            dist, paths, idx = self.shadow_device(x_context, k)
            return _numpy(dist), _numpy(paths), _numpy(idx)
        # results land in ONE device buffer [dist | idx | paths] (4-byte words) that is copied with a
        # single async copy into a pinned host buffer; the returned numpy arrays are views of that
        # buffer (no second copy), which goes back to a small pool when the caller drops them
        shp = _dim_array(x_context).shape
        B, L = shp[0], shp[-1] + self.context.get_out_times()
        words = B * k * (3 + L)
        if self._staging is None or self._staging[0].numel() != words:
            # (two pinned buffers up front: a caller that keeps the previous result while asking for the next
            # one would otherwise pay a 2 ms cudaHostAlloc inside its second call)
            self._staging = (torch.empty(words, dtype=torch.int32, device=self._dev()),
                             [torch.empty(words, dtype=torch.int32, pin_memory=True) for _ in range(2)], [0])
        dev_buf, pool, outstanding = self._staging
        if pool:
            host_buf = pool.pop()
        elif outstanding[0] < self._POOL_MAX:
            host_buf = torch.empty(words, dtype=torch.int32, pin_memory=True)
        else:
            host_buf = None   # the caller keeps many results alive: fall back to pageable copies
        stage = host_buf if host_buf is not None else self._spill_buffer(words)
        # scan, gather and the copy back are enqueued together; the single synchronisation is the
        # status read, which also tells whether a candidate buffer overflowed (adversarially
        # ordered data) -- then the synchronous scan repeats it in its safe schedule
        self.shadow_device(x_context, k, _packed=dev_buf, _nosync=True)
        stage.copy_(dev_buf, non_blocking=True)
        if _lib.scan_overflowed(self._workspace, B):
            self.shadow_device(x_context, k, _packed=dev_buf)
            stage.copy_(dev_buf, non_blocking=True)
            torch.cuda.current_stream(dev_buf.device).synchronize()
        if host_buf is not None:
            flat = host_buf.numpy()
            outstanding[0] += 1
            weakref.finalize(flat, _recycle, pool, outstanding, host_buf)
        else:
            flat = stage.numpy().copy()
        nd, ni = B * k, B * k * 2
        return (flat[:nd].view(np.float32).reshape(B, k),
                flat[nd + ni:].view(np.float32).reshape(B, k, 1, L),
                flat[nd:nd + ni].reshape(B, k, 2))

    def _shadow_sharded(self, x_context: ArrayType, k: int):
        """`shadow` on an ensemble sharded over the process group: local scan (enqueue only), fused
        exchange + merge, owner gather + all-reduce of the paths, then ONE device buffer
        [dist | idx | paths | overflow flag] copied with ONE asynchronous copy into pinned memory and ONE
        synchronisation (round 1: three pageable copies, an `.item()` on the flag, 1.0 ms at 8 GPUs)."""
        from .distributed import flush_deferred_merge, sharded_gather
        if self.embedding.kernel.shape[-1] != 0 and self.embedding.kernel.shape[-1] != _dim_array(x_context).shape[-1]:
            raise Exception("The embedding kernel should be of the same size as the context.")
        x = _torch(_dim_array(x_context))
        rows, T = self._resident_rows()
        B, L = x.shape[0], x.shape[-1] + self.context.get_out_times()
        nd, ni, npth = B * k, B * k * 2, B * k * L
        words = nd + ni + npth + 1
        if self._staging is None or self._staging[0].numel() != words:
            self._staging = (torch.empty(words, dtype=torch.int32, device=self._dev()),
                             [torch.empty(words, dtype=torch.int32, pin_memory=True) for _ in range(2)], [0])
        dev_buf, pool, outstanding = self._staging
        host_buf = pool.pop() if pool else torch.empty(words, dtype=torch.int32, pin_memory=True)
        streams, self._pipe_streams = self._pipe_streams, 1               # this call's own stream, no lanes
        try:
            dist, idx = self._scan_device(x, rows, T, k, nosync=True)
        finally:
            self._pipe_streams = streams
        flush_deferred_merge(self)                                       # (PSH_DEFER=1: the split exchange)
        flag = getattr(self, "_pending_flag", None)
        self._pending_flag = None
        paths = sharded_gather(self, rows, T, idx, L)
        dev_buf[:nd].view(torch.float32).copy_(dist.reshape(-1))
        dev_buf[nd:nd + ni].copy_(idx.reshape(-1))
        dev_buf[nd + ni:nd + ni + npth].view(torch.float32).copy_(paths.reshape(-1))
        if flag is not None:
            dev_buf[-1:].copy_(flag)
        else:
            dev_buf[-1:].zero_()
        host_buf.copy_(dev_buf, non_blocking=True)
        torch.cuda.current_stream(dev_buf.device).synchronize()
        flat = host_buf.numpy()
        if int(flat[-1]) != 0:      # an overflow somewhere (every rank sees the same flag) or a peer timed out
            pool.append(host_buf)
            if flag is not None:
                flag.zero_()
            if int(flat[-1]) & 2:
                raise RuntimeError("peer-memory all-gather: a rank did not deliver its records within the timeout")
            dist, paths, idx = self.shadow_device(x_context, k)           # synchronous scans: the safe schedule
            return _numpy(dist), _numpy(paths), _numpy(idx)
        outstanding[0] += 1
        weakref.finalize(flat, _recycle, pool, outstanding, host_buf)
        return (flat[:nd].view(np.float32).reshape(B, k),
                flat[nd + ni:nd + ni + npth].view(np.float32).reshape(B, k, 1, L),
                flat[nd:nd + ni].reshape(B, k, 2))

    _POOL_MAX = 8   # pinned result buffers handed out at once before falling back to copies

    def _spill_buffer(self, words: int) -> torch.Tensor:
        sp = getattr(self, "_spill", None)
        if sp is None or sp.numel() != words:
            sp = torch.empty(words, dtype=torch.int32, pin_memory=True)
            self._spill = sp
        return sp

    # ------------------------------------------------------------------ aggregation
    @staticmethod
    def init_averaging_proba(proba_name: str, distances: np.ndarray, eta: float | None) -> DiscreteProba:
        """The probability used to average out-context predictions (path_shadowing.py:220-232)."""
        if proba_name == "uniform":
            return Uniform()
        elif proba_name == "softmax":
            return Softmax(distances, eta)
        else:
            raise ValueError("Unrecognized averaging proba")

    def _predict_device(self, dist: torch.Tensor, paths: torch.Tensor, rv: RealizedVariance, proba_name: str,
                        eta: float | None):
        if proba_name not in ("uniform", "softmax"):
            raise ValueError("Unrecognized averaging proba")
        if proba_name == "softmax" and not (eta is not None and eta > 0):
            raise ValueError("softmax averaging needs a positive eta (the width of the Gaussian weights)")
        H = self.context.get_out_times() or paths.shape[-1]
        order = sorted(range(len(rv.Ts)), key=lambda i: rv.Ts[i])
        Ts = torch.tensor([rv.Ts[i] for i in order], dtype=torch.int32, device=paths.device)
        mean, std = _lib.rv_aggregate(paths, dist, H, Ts, eta, 1 if proba_name == "softmax" else 0, rv.vol)
        inv = torch.empty(len(order), dtype=torch.long)
        inv[torch.tensor(order)] = torch.arange(len(order))
        inv = inv.to(paths.device)
        return mean[:, inv], std[:, inv]

    def predict_from_paths(self, distances: np.ndarray, paths: np.ndarray, to_predict: Callable,
                           proba_name: str, eta: float | None) -> tuple[np.ndarray, np.ndarray]:
        """Aggregate predictions over shadowing paths (path_shadowing.py:234-254).

        A `RealizedVariance` callable is evaluated by the fused CUDA kernel; any other callable
        is the user's own host function and is applied to the out-context on the host, exactly
        as the reference does (called twice there; once here)."""
        if (isinstance(to_predict, RealizedVariance) and paths.shape[2] == 1 and type(self.context) is PredictionContext
                and len(to_predict.Ts) <= _lib.AGG_MAX_T):
            dev = self._dev()
            d = torch.as_tensor(distances, dtype=torch.float32).to(dev)
            p = torch.as_tensor(paths, dtype=torch.float32).to(dev)
            mean, std = self._predict_device(d, p, to_predict, proba_name, eta)
            return _numpy(mean), _numpy(std)
        out = self.context.select_out_context(_numpy(paths) if isinstance(paths, torch.Tensor) else paths)
        proba = self.init_averaging_proba(proba_name, np.asarray(distances)[:, :, None], eta)
        values = to_predict(out)
        return proba.avg(values, axis=1), proba.std(values, axis=1)

    def predict(self, x_context: ArrayType, k: int, to_predict: Callable, eta: float | None = None,
                proba_name: str = "softmax", n_dataset_splits: int = 1, n_context_splits: int = 1,
                cuda: bool = False) -> tuple[np.ndarray, np.ndarray]:
        """Shadow then aggregate, chunking the contexts (path_shadowing.py:256-301).
        With a `RealizedVariance` target nothing but the (B, nT) results leaves the GPU."""
        del n_dataset_splits, cuda
        x = _torch(_dim_array(x_context))
        B = x.shape[0]
        preds, stds = [], []
        for bs in torch.arange(B).split(max(B // n_context_splits, 1)):
            xb = x[bs, ...]
            if (isinstance(to_predict, RealizedVariance) and type(self.context) is PredictionContext
                    and len(to_predict.Ts) <= _lib.AGG_MAX_T):
                dist, paths, _ = self.shadow_device(xb, k)
                mean, std = self._predict_device(dist, paths, to_predict, proba_name, eta)
                preds.append(_numpy(mean))
                stds.append(_numpy(std))
            else:
                dist, paths, _ = self.shadow(xb, k)
                res = self.predict_from_paths(dist, paths, to_predict, proba_name, eta)
                preds.append(res[0])
                stds.append(res[1])
        return np.concatenate(preds), np.concatenate(stds)
