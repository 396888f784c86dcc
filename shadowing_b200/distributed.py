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

import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib

_INF = float("inf")
_PAD_ROW = 2 ** 31 - 1


class Lane:
    """One stream of a pipeline of enqueue-only sharded scans: its own workspace, record buffers, overflow
    flag and deferred merge; everything else (the plugin objects, the resident rows, the exchange buffers
    and their epoch counter) is the PathShadowing object's.  `sharded_scan(lane, ...)` then runs unchanged."""
    _OWN = frozenset({"stream", "_workspace", "_shard_bufs", "_pending_flag", "_pending_merge"})

    def __init__(self, ps, stream):
        object.__setattr__(self, "_ps", ps)
        self.stream = stream
        self._workspace = None
        self._shard_bufs = None
        self._pending_flag = None
        self._pending_merge = None

    def __getattr__(self, name):           # (only reached for names the lane does not hold itself)
        return getattr(object.__getattribute__(self, "_ps"), name)

    def __setattr__(self, name, value):
        if name in Lane._OWN:
            object.__setattr__(self, name, value)
        else:
            setattr(object.__getattribute__(self, "_ps"), name, value)


def shard_bounds(R: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition of R rows: the first R % world ranks hold one extra row."""
    base, extra = divmod(R, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _local_records(ps, rows, T, q, H, k, rec, nosync: bool, W: int):
    """This rank's k best windows as packed (B,k,3) int32 records [distance bits, r, t] in `rec`.
    `q` holds the contexts (B, W) -- or, for a non-Identity embedding, the EMBEDDED contexts (B, d)."""
    n_local = rows.shape[0] * (T - W - H + 1)
    k_loc = min(k, n_local)
    runs = ps._run_table(rows.device) if rows.is_cuda else None
    if runs is not None:   # Foveal / PathEmbedding(kernel): the embedded scan
        rec[..., 0] = 0x7F800000
        rec[..., 1] = _PAD_ROW
        rec[..., 2] = 0
        flavour = ps._embed_flavour(rows, T, W, H, ps._ex_host) if rows.shape[0] > 0 else {}
        if k_loc == k:
            _, _, ps._workspace = _lib.scan_topk_embed(rows, T, q, W, H, k, runs, ps._row_offset, nosync,
                                                       ps._workspace, rec=rec, **flavour)
        elif k_loc > 0:
            d, i, ps._workspace = _lib.scan_topk_embed(rows, T, q, W, H, k_loc, runs, ps._row_offset, False,
                                                       ps._workspace, **flavour)
            rec[:, :k_loc, 0] = d.view(torch.int32)
            rec[:, :k_loc, 1:] = i
        return
    if rows.is_cuda and k_loc == k:
        mode, aux = ps._mode_and_aux(rows, T, W, H)
        # (a pipeline alternating between lanes: leave two SMs to the other lane's re-rank / select / exchange)
        # (ten: next to re-rank and select, the G exchange CTAs of every lane wait there for the peers;
        # measured at 8 GPUs: 0.187 ms per step with six spare SMs, 0.169 with ten)
        share = _lib.share_sms(10) if (nosync and getattr(ps, "_pipe_streams", 1) > 1) else 0
        ps._workspace = _lib.scan_topk_packed(rows, T, q, H, k, ps._row_offset,
                                              mode | (_lib.PSH_FLAG_NOSYNC if nosync else 0) | share, ps._workspace, aux, rec)
        return
    # a shard with fewer than k windows pads with +inf records that sort last
    rec[..., 0] = 0x7F800000
    rec[..., 1] = _PAD_ROW
    rec[..., 2] = 0
    if k_loc > 0:
        mode, aux = ps._mode_and_aux(rows, T, W, H)
        d, i, ps._workspace = _lib.scan_topk(rows, T, q, H, k_loc, ps._row_offset, mode, ps._workspace, aux)
        rec[:, :k_loc, 0] = d.view(torch.int32)
        rec[:, :k_loc, 1:] = i


class _PeerExchange:
    """Exchange buffers of all ranks mapped into this process (CUDA IPC over NVLink peer access).
    Creation is collective: every rank allocates its buffer, the 64-byte IPC handles travel through
    one all_gather_object, and the ranks agree (all-reduce MIN) on whether every mapping worked."""

    def __init__(self, pg, device: torch.device, B: int, k: int):
        self.world, self.rank = dist.get_world_size(pg), dist.get_rank(pg)
        self.device, self.shape, self.epoch = device, (B, k), 0
        self.bufs: list[int] = []
        self.own = 0
        ok = 1
        handle = b""
        try:
            self.own, handle = _lib.xchg_create(_lib.xchg_bytes(self.world, B, k), device)
        except Exception:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=pg)
        if ok and all(handles):
            try:
                for g, h in enumerate(handles):
                    self.bufs.append(self.own if g == self.rank else _lib.xchg_open(h, device))
            except Exception:
                ok = 0
        else:
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=pg)
        self.ok = bool(agree.item())
        if not self.ok:
            self.close()
        # the host array of peer pointers the exchange launches take, built once
        self.arr = (ctypes.c_void_p * len(self.bufs))(*self.bufs) if self.ok else None

    def close(self):
        for g, p in enumerate(self.bufs):
            if g != self.rank and p:
                _lib.xchg_close(p, self.device)
        self.bufs = []
        if self.own:
            _lib.xchg_destroy(self.own, self.device)
            self.own = 0


def _peer_exchange(ps, device, B, k):
    """The peer-memory exchange for (B, k), or None when disabled (PSH_P2P=0) / unavailable --
    then the step uses ncclAllGather + merge_kernel."""
    if os.environ.get("PSH_P2P", "1") == "0":
        return None
    ex = getattr(ps, "_xchg", None)
    if ex is None or ex.shape != (B, k):
        if ex is not None:
            torch.cuda.synchronize(device)
            dist.barrier(group=ps._pg)     # nobody is still writing into the buffers being replaced
            ex.close()
        ex = _PeerExchange(ps._pg, device, B, k)
        ps._xchg = ex
    return ex if ex.ok else None


def flush_deferred_merge(ps) -> None:
    """Enqueue the wait + merge of the step whose records were sent last (pipelined scans)."""
    pend = getattr(ps, "_pending_merge", None)
    if pend is None:
        return
    ps._pending_merge = None
    ex, epoch, B, k, Tp, dist_, idx_, flag = pend
    _lib.xchg_merge(ex.bufs, ex.rank, B, k, Tp, epoch, dist_, idx_, flag)


def _exchange_and_merge(ps, rows, rec, rec_all, Tp, flag, defer: bool = False):
    """All-gather of the per-rank records + merge.  `defer` (pipelines of enqueue-only scans): only
    the send half is enqueued now; the wait + merge follows the NEXT step's scan and send (or the
    pipeline's final check), so no rank idles while a slower peer finishes the same step.  The
    returned tensors are filled by then -- the caller must not read them before `_check_pipeline`."""
    pg = ps._pg
    if rows.is_cuda:
        B, k = rec.shape[0], rec.shape[1]
        pend = getattr(ps, "_pending_merge", None)
        if pend is not None and (not defer or pend[2:4] != (B, k)):
            flush_deferred_merge(ps)
        ex = _peer_exchange(ps, rows.device, B, k)
        if ex is not None:
            ex.epoch += 1
            # (measured at N = 2 and 4: the split form is no faster -- rank skew is not what the exchange
            # costs -- so the single fused launch stays the default; PSH_DEFER=1 selects the split form)
            if not defer or os.environ.get("PSH_DEFER", "0") != "1":
                return _lib.allgather_merge_packed(rec, ex.bufs, ex.rank, Tp, ex.epoch, flag)
            _lib.xchg_send(rec, ex.bufs, ex.rank, ex.epoch)
            flush_deferred_merge(ps)     # the previous step's merge, behind this step's scan and send
            dist_ = torch.empty((B, k), dtype=torch.float32, device=rows.device)
            idx_ = torch.empty((B, k, 2), dtype=torch.int32, device=rows.device)
            ps._pending_merge = (ex, ex.epoch, B, k, Tp, dist_, idx_, flag)
            return dist_, idx_
        dist.all_gather_into_tensor(rec_all, rec, group=pg)          # one ncclAllGather, B*k*12 bytes per rank
        return _lib.merge_topk_packed(rec_all, Tp, flag)
    dist.all_gather(list(rec_all.unbind(0)), rec, group=pg)          # gloo (CPU tests) has no *_into_tensor
    return _lib.merge_topk(rec_all[..., 0].contiguous().view(torch.float32), rec_all[..., 1:].contiguous(), Tp)


def sharded_scan(ps, rows: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int, W: int | None = None,
                 defer: bool = False):
    """Local exact top-k on this rank's rows, all-gather, merge.  Every rank returns the global
    (dist (B,k), idx (B,k,2)) with GLOBAL trajectory indices.  On CUDA the scan, the all-gather
    and the merge are enqueued back to back without a host synchronisation; the candidate-buffer
    overflow flag travels inside the records and is read once per step by `finish_sharded`."""
    pg = ps._pg
    world = dist.get_world_size(pg)
    B = q.shape[0]
    W = q.shape[1] if W is None else W
    Tp = T - W - H + 1
    if Tp <= 0:
        raise RuntimeError(f"context ({W}) + horizon ({H}) longer than the trajectories ({T})")
    n_local = rows.shape[0] * Tp
    # total number of windows over all ranks: one collective per (shard, W, H), then cached
    key = (rows.data_ptr(), rows.shape[0], Tp)
    cache = getattr(ps, "_total_windows", None)
    if cache is None or cache[0] != key:
        tot = torch.tensor([n_local], dtype=torch.int64, device=rows.device)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=pg)
        cache = (key, int(tot.item()))
        ps._total_windows = cache
    if k > cache[1]:
        raise RuntimeError(f"selected index k out of range: k={k} > {cache[1]} windows")
    bufs = getattr(ps, "_shard_bufs", None)
    if bufs is None or bufs[0].shape != (B, k, 3) or bufs[0].device != rows.device:
        bufs = (torch.empty((B, k, 3), dtype=torch.int32, device=rows.device),
                torch.empty((world, B, k, 3), dtype=torch.int32, device=rows.device),
                torch.zeros(1, dtype=torch.int32, device=rows.device))
        ps._shard_bufs = bufs
    rec, rec_all, flag = bufs
    _local_records(ps, rows, T, q, H, k, rec, True, W)
    out = _exchange_and_merge(ps, rows, rec, rec_all, Tp, flag, defer)
    ps._pending_flag = flag if rows.is_cuda else None
    return out


def sharded_scan_fast(lane, rows: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int):
    """The steady state of a pipelined sharded step -- Identity scan of a shard that holds at least k windows,
    buffers, window total and peer exchange already set up, fused exchange -- enqueued on the lane's stream
    through RAW stream handles: no stream switch on the host, no proxy attribute look-ups.  With 4 or 8 ranks
    on one box the pipeline was bound by the 0.13-0.16 ms of host time a step took through `sharded_scan`.
    Returns None when anything else is needed (first step of a lane, padding, NCCL fallback, split form):
    the caller then takes `sharded_scan` inside the lane's stream."""
    ps = object.__getattribute__(lane, "_ps")
    if os.environ.get("PSH_P2P", "1") == "0" or os.environ.get("PSH_DEFER", "0") == "1":
        return None
    B, W = q.shape
    Tp = T - W - H + 1
    cache = getattr(ps, "_total_windows", None)
    bufs = lane._shard_bufs
    ex = getattr(ps, "_xchg", None)
    if (Tp <= 0 or rows.shape[0] * Tp < k or cache is None or cache[0] != (rows.data_ptr(), rows.shape[0], Tp)
            or k > cache[1] or bufs is None or bufs[0].shape != (B, k, 3) or lane._workspace is None
            or ex is None or ex.shape != (B, k) or not ex.ok or lane._pending_merge is not None):
        return None
    rec, _, flag = bufs
    mode, aux = ps._mode_and_aux(rows, T, W, H)
    s = lane.stream.cuda_stream
    lane._workspace = _lib.scan_topk_packed(rows, T, q, H, k, ps._row_offset,
                                            mode | _lib.PSH_FLAG_NOSYNC | _lib.share_sms(10), lane._workspace, aux, rec,
                                            stream=s)
    ex.epoch += 1
    out = _lib.allgather_merge_packed(rec, ex.bufs, ex.rank, Tp, ex.epoch, flag, stream=s, arr=ex.arr)
    lane._pending_flag = flag
    return out


def finish_sharded(ps, rows: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int, out, W: int | None = None):
    """Second half of the sharded scan: the single host synchronisation of a step.  If a candidate
    buffer overflowed on ANY rank (adversarially ordered data; every rank reads the same flag from
    the gathered records) all ranks repeat the step together with synchronous scans, which re-run
    the overflowing queries in the safe schedule."""
    flag = getattr(ps, "_pending_flag", None)
    if flag is None:
        return out
    ps._pending_flag = None
    status = int(flag.item())
    if status == 0:
        return out
    flag.zero_()
    if status & 2:
        raise RuntimeError("peer-memory all-gather: a rank did not deliver its records within the timeout")
    W = q.shape[1] if W is None else W
    rec, rec_all, _ = ps._shard_bufs
    _local_records(ps, rows, T, q, H, k, rec, False, W)
    return _exchange_and_merge(ps, rows, rec, rec_all, T - W - H + 1, None)


def sharded_gather(ps, rows: torch.Tensor, T: int, idx: torch.Tensor, L: int) -> torch.Tensor:
    """Each rank copies the winners it owns (zeros elsewhere); the sum over ranks is the result."""
    paths = _lib.gather_paths(rows, T, idx, L, ps._row_offset)
    dist.all_reduce(paths, op=dist.ReduceOp.SUM, group=ps._pg)
    return paths
