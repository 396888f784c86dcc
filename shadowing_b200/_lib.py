"""ctypes binding of libpshadow.so (include/pshadow.h).  No fallback: if the CUDA library is
missing or no CUDA device is present, every entry point raises."""
from __future__ import annotations

import ctypes
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
import os as _os
LIB_PATH = _PKG / _os.environ.get("PSH_LIB", "libpshadow.so")   # (A/B builds: PSH_LIB=libpshadow_<variant>.so)

PSH_MODE_EXACT = 0
PSH_MODE_FILTER = 1
PSH_MODE_FFT = 2
PSH_FLAG_NOSYNC = 0x100
PSH_FLAG_SHARE_SMS = 0x200


def share_sms(n: int) -> int:
    """PSH_SHARE_SMS(n): the pipelined scan leaves n SMs to the other streams' small kernels."""
    return PSH_FLAG_SHARE_SMS | ((n & 0x3F) << 12)

PSH_E_OVERFLOW = -6
FFT_MAX_W = 2048   # psh_fft_prepare: context length at most half a 4096-point transform
AGG_MAX_T = 16     # psh_rv_aggregate: maturities per launch (more: the host aggregation)

_lib = None


class PshadowError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().psh_error_string(code).decode()
        super().__init__(f"{where}: {msg} (code {code})")


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is not built. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). shadowing_b200 has no CPU fallback.")
        L = ctypes.CDLL(str(LIB_PATH))
        vp, i64, i32, ci, sz, f32 = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int,
                                     ctypes.c_size_t, ctypes.c_float)
        L.psh_version.restype = ci
        L.psh_version.argtypes = []
        L.psh_error_string.restype = ctypes.c_char_p
        L.psh_error_string.argtypes = [ci]
        L.psh_launch_count.restype = ctypes.c_uint64
        L.psh_launch_count.argtypes = []
        L.psh_scan_workspace_bytes.restype = sz
        L.psh_scan_workspace_bytes.argtypes = [i64, i64, ci, ci, ci, i64]
        L.psh_scan_topk_f32.restype = ci
        L.psh_scan_topk_f32.argtypes = [vp, i64, i64, i64, vp, ci, ci, ci, i64, i32, ci, vp, vp, vp, sz, vp, sz, vp]
        L.psh_scan_topk_embed_f32.restype = ci
        L.psh_scan_topk_embed_f32.argtypes = [vp, i64, i64, i64, vp, ci, ci, ci, ci, i64, i32, ci, vp, ci,
                                              vp, vp, ctypes.c_size_t, vp, vp, vp, ctypes.c_size_t, vp]
        L.psh_fft_prepare_embed.restype = ci
        L.psh_fft_prepare_embed.argtypes = [vp, i64, i64, i64, ci, ci, vp, ci, vp, ctypes.c_size_t, vp]
        L.psh_scan_overflowed.restype = ci
        L.psh_scan_overflowed.argtypes = [vp, ci, vp]
        L.psh_fft_aux_bytes.restype = sz
        L.psh_fft_aux_bytes.argtypes = [i64, i64, ci, ci]
        L.psh_fft_prepare.restype = ci
        L.psh_fft_prepare.argtypes = [vp, i64, i64, i64, ci, ci, vp, sz, vp]
        L.psh_debug_fft4096.restype = ci
        L.psh_debug_fft4096.argtypes = [vp, vp, ci, ci, vp, vp]
        L.psh_debug_fft1024.restype = ci
        L.psh_debug_fft1024.argtypes = [vp, vp, ci, ci, vp, vp]
        L.psh_merge_topk.restype = ci
        L.psh_merge_topk.argtypes = [vp, vp, ci, ci, i64, i64, vp, vp, vp]
        L.psh_merge_topk_packed.restype = ci
        L.psh_merge_topk_packed.argtypes = [vp, ci, ci, i64, i64, vp, vp, vp, vp]
        L.psh_xchg_bytes.restype = ctypes.c_size_t
        L.psh_xchg_bytes.argtypes = [ci, ci, i64]
        L.psh_xchg_create.restype = ci
        L.psh_xchg_create.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp), ctypes.c_char_p]
        L.psh_xchg_open.restype = ci
        L.psh_xchg_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(vp)]
        L.psh_xchg_close.restype = ci
        L.psh_xchg_close.argtypes = [vp]
        L.psh_xchg_destroy.restype = ci
        L.psh_xchg_destroy.argtypes = [vp]
        L.psh_allgather_merge_packed.restype = ci
        L.psh_allgather_merge_packed.argtypes = [vp, ctypes.POINTER(vp), ci, ci, ci, i64, i64, ctypes.c_uint32,
                                                 vp, vp, vp, vp]
        L.psh_xchg_send.restype = ci
        L.psh_xchg_send.argtypes = [vp, ctypes.POINTER(vp), ci, ci, ci, i64, ctypes.c_uint32, vp]
        L.psh_xchg_merge.restype = ci
        L.psh_xchg_merge.argtypes = [ctypes.POINTER(vp), ci, ci, ci, i64, i64, ctypes.c_uint32, vp, vp, vp, vp]
        L.psh_gather_paths.restype = ci
        L.psh_gather_paths.argtypes = [vp, i64, i64, i64, vp, i64, i32, ci, vp, vp]
        L.psh_rv_aggregate.restype = ci
        L.psh_rv_aggregate.argtypes = [vp, vp, ci, i64, ci, ci, vp, ci, f32, ci, ci, vp, vp, vp]
        _lib = L
    return _lib


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("shadowing_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _check(rc: int, where: str) -> None:
    if rc != 0:
        raise PshadowError(rc, where)


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class _on_device:
    """`with torch.cuda.device(dev)` that costs nothing when `dev` is already current (the common case; the
    context manager is ~8 us of host time per call, a tenth of what it takes to enqueue a whole scan)."""
    __slots__ = ("_cm",)

    def __init__(self, dev: torch.device):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self._cm = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self._cm is not None:
            self._cm.__enter__()

    def __exit__(self, *exc):
        if self._cm is not None:
            return self._cm.__exit__(*exc)
        return False


_ws_bytes_cache: dict = {}


def _workspace_bytes(R: int, T: int, B: int, W: int, H: int, k: int) -> int:
    key = (R, T, B, W, H, k)
    n = _ws_bytes_cache.get(key)
    if n is None:
        n = int(lib().psh_scan_workspace_bytes(R, T, B, W, H, k))
        if len(_ws_bytes_cache) < 4096:
            _ws_bytes_cache[key] = n
    return n


def launch_count() -> int:
    return int(lib().psh_launch_count())


def scan_topk_packed(ds: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int, row_offset: int, mode: int,
                     workspace: torch.Tensor | None, aux: torch.Tensor | None, rec: torch.Tensor,
                     stream: int | None = None):
    """Same scan, results as packed (B,k,3) int32 records [distance bits, r, t] written to `rec`.
    `stream`: raw CUDA stream to enqueue on (default: torch's current stream of ds.device)."""
    L = lib()
    R, row_stride = ds.shape[0], ds.stride(0)
    B, W = q.shape
    need = _workspace_bytes(R, T, B, W, H, k) or 256
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=ds.device)
    with _on_device(ds.device):
        rc = L.psh_scan_topk_f32(ds.data_ptr(), R, T, row_stride, q.data_ptr(), B, W, H, k, row_offset, mode,
                                 rec.data_ptr(), None, workspace.data_ptr(), workspace.numel(),
                                 aux.data_ptr() if aux is not None else None, aux.numel() if aux is not None else 0,
                                 _stream(ds) if stream is None else stream)
    _check(rc, "psh_scan_topk_f32")
    return workspace


def scan_overflowed(workspace: torch.Tensor, B: int) -> bool:
    """Synchronise and report whether a PSH_FLAG_NOSYNC scan overflowed a candidate buffer."""
    L = lib()
    with torch.cuda.device(workspace.device):
        rc = L.psh_scan_overflowed(workspace.data_ptr(), B, _stream(workspace))
    if rc == PSH_E_OVERFLOW:
        return True
    _check(rc, "psh_scan_overflowed")
    return False


def fft_prepare(ds: torch.Tensor, T: int, W: int, H: int) -> torch.Tensor:
    """Dataset-side precomputation for PSH_MODE_FFT: spectra of row pairs, window energies, norms."""
    L = lib()
    need = L.psh_fft_aux_bytes(ds.shape[0], T, W, H)
    if need == 0:
        raise PshadowError(-5, "psh_fft_aux_bytes")
    aux = torch.empty(need, dtype=torch.uint8, device=ds.device)
    with torch.cuda.device(ds.device):
        rc = L.psh_fft_prepare(ds.data_ptr(), ds.shape[0], T, ds.stride(0), W, H, aux.data_ptr(), aux.numel(),
                               _stream(ds))
    _check(rc, "psh_fft_prepare")
    return aux


def fft_prepare_embed(ds: torch.Tensor, T: int, W: int, H: int, runs: torch.Tensor) -> torch.Tensor:
    """Dataset- and kernel-side precomputation for the fft flavour of the embedded scan: spectra of
    row pairs, EMBEDDED window energies sum_n e_n(t)^2, pair norms."""
    L = lib()
    need = L.psh_fft_aux_bytes(ds.shape[0], T, W, H)
    if need == 0:
        raise PshadowError(-5, "psh_fft_aux_bytes")
    aux = torch.empty(need, dtype=torch.uint8, device=ds.device)
    with torch.cuda.device(ds.device):
        rc = L.psh_fft_prepare_embed(ds.data_ptr(), ds.shape[0], T, ds.stride(0), W, H, runs.data_ptr(),
                                     runs.shape[0], aux.data_ptr(), aux.numel(), _stream(ds))
    _check(rc, "psh_fft_prepare_embed")
    return aux


def debug_fft4096(x: torch.Tensor, direction: int, aux: torch.Tensor) -> torch.Tensor:
    """x (n, 4096) complex64 cuda -> unnormalised DFT (direction -1) / inverse (+1), test hook."""
    L = lib()
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = L.psh_debug_fft4096(x.data_ptr(), out.data_ptr(), x.shape[0], direction, aux.data_ptr(), _stream(x))
    _check(rc, "psh_debug_fft4096")
    return out


def debug_fft1024(x: torch.Tensor, direction: int, aux: torch.Tensor) -> torch.Tensor:
    """x (n, 1024) complex64 cuda -> unnormalised DFT (direction -1) / inverse (+1) with the warp-level
    transform of the 1024-point fft flavour, test hook."""
    L = lib()
    x = x.contiguous()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = L.psh_debug_fft1024(x.data_ptr(), out.data_ptr(), x.shape[0], direction, aux.data_ptr(), _stream(x))
    _check(rc, "psh_debug_fft1024")
    return out


def scan_topk(ds: torch.Tensor, T: int, q: torch.Tensor, H: int, k: int, row_offset: int = 0,
              mode: int = PSH_MODE_FILTER, workspace: torch.Tensor | None = None, aux: torch.Tensor | None = None,
              out: tuple[torch.Tensor, torch.Tensor] | None = None, stream: int | None = None):
    """ds (R, row_stride) f32 cuda, q (B, W) f32 cuda -> (dist (B,k) f32, idx (B,k,2) i32) cuda.
    `stream`: raw CUDA stream to enqueue on (default: torch's current stream of ds.device)."""
    L = lib()
    assert ds.is_cuda and q.is_cuda and ds.dtype == torch.float32 and q.dtype == torch.float32
    assert ds.dim() == 2 and ds.stride(1) == 1 and q.is_contiguous()
    R, row_stride = ds.shape[0], ds.stride(0)
    B, W = q.shape
    need = _workspace_bytes(R, T, B, W, H, k)
    if need == 0:
        # invalid sizes: let the library name the error
        need = 256
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=ds.device)
    if out is not None:
        dist, idx = out
    else:
        dist = torch.empty((B, k), dtype=torch.float32, device=ds.device)
        idx = torch.empty((B, k, 2), dtype=torch.int32, device=ds.device)
    with _on_device(ds.device):
        rc = L.psh_scan_topk_f32(ds.data_ptr(), R, T, row_stride, q.data_ptr(), B, W, H, k, row_offset, mode,
                                 dist.data_ptr(), idx.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                 aux.data_ptr() if aux is not None else None, aux.numel() if aux is not None else 0,
                                 _stream(ds) if stream is None else stream)
    _check(rc, "psh_scan_topk_f32")
    return dist, idx, workspace


def scan_topk_embed(ds: torch.Tensor, T: int, ex: torch.Tensor, W: int, H: int, k: int, runs: torch.Tensor,
                    row_offset: int = 0, nosync: bool = False, workspace: torch.Tensor | None = None,
                    out: tuple[torch.Tensor, torch.Tensor] | None = None, rec: torch.Tensor | None = None,
                    g: torch.Tensor | None = None, aux: torch.Tensor | None = None):
    """Scan in embedded space: ex (B, d) f32 cuda embedded queries, runs (nruns, 4) 32-bit words cuda
    [row, a, b, c] (path_embedding.kernel_runs) -> (dist (B,k) f32, idx (B,k,2) i32) cuda; with
    `rec` (B,k,3) i32 the results are written there as packed [distance bits, r, t] records.
    fft flavour: `aux` from fft_prepare_embed, `g` (B, W) = ex @ K."""
    L = lib()
    assert ds.is_cuda and ex.is_cuda and runs.is_cuda and ex.dtype == torch.float32 and ex.is_contiguous()
    R, row_stride = ds.shape[0], ds.stride(0)
    B, d = ex.shape
    need = L.psh_scan_workspace_bytes(R, T, B, W, H, k) or 256
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=ds.device)
    if rec is not None:
        dist, idx = rec, None
    elif out is not None:
        dist, idx = out
    else:
        dist = torch.empty((B, k), dtype=torch.float32, device=ds.device)
        idx = torch.empty((B, k, 2), dtype=torch.int32, device=ds.device)
    with torch.cuda.device(ds.device):
        rc = L.psh_scan_topk_embed_f32(ds.data_ptr(), R, T, row_stride, ex.data_ptr(), B, d, W, H, k, row_offset,
                                       PSH_FLAG_NOSYNC if nosync else 0, runs.data_ptr(), runs.shape[0],
                                       g.data_ptr() if aux is not None else None,
                                       aux.data_ptr() if aux is not None else None,
                                       aux.numel() if aux is not None else 0,
                                       dist.data_ptr(), idx.data_ptr() if idx is not None else None,
                                       workspace.data_ptr(), workspace.numel(), _stream(ds))
    _check(rc, "psh_scan_topk_embed_f32")
    return dist, idx, workspace


def merge_topk(dist_parts: torch.Tensor, idx_parts: torch.Tensor, Tp: int):
    """(G,B,k) f32 + (G,B,k,2) i32 -> merged (B,k), (B,k,2)."""
    L = lib()
    G, B, k = dist_parts.shape
    dist_parts = dist_parts.contiguous()
    idx_parts = idx_parts.contiguous()
    dist = torch.empty((B, k), dtype=torch.float32, device=dist_parts.device)
    idx = torch.empty((B, k, 2), dtype=torch.int32, device=dist_parts.device)
    with torch.cuda.device(dist_parts.device):
        rc = L.psh_merge_topk(dist_parts.data_ptr(), idx_parts.data_ptr(), G, B, k, Tp, dist.data_ptr(),
                              idx.data_ptr(), _stream(dist_parts))
    _check(rc, "psh_merge_topk")
    return dist, idx


def merge_topk_packed(rec_parts: torch.Tensor, Tp: int, flag: torch.Tensor | None = None):
    """(G,B,k,3) i32 [distance bits, r, t] -> merged (B,k) f32, (B,k,2) i32.  `flag` (int32[1],
    zeroed by the caller) is set if any shard's NOSYNC scan overflowed."""
    L = lib()
    G, B, k, _ = rec_parts.shape
    rec_parts = rec_parts.contiguous()
    dist = torch.empty((B, k), dtype=torch.float32, device=rec_parts.device)
    idx = torch.empty((B, k, 2), dtype=torch.int32, device=rec_parts.device)
    with torch.cuda.device(rec_parts.device):
        rc = L.psh_merge_topk_packed(rec_parts.data_ptr(), G, B, k, Tp, dist.data_ptr(), idx.data_ptr(),
                                     flag.data_ptr() if flag is not None else None, _stream(rec_parts))
    _check(rc, "psh_merge_topk_packed")
    return dist, idx


def gather_paths(ds: torch.Tensor, T: int, idx: torch.Tensor, L_out: int, row_offset: int = 0,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    """idx (B,k,2) i32 -> paths (B,k,1,L) f32 (rows outside this shard are zeros)."""
    L = lib()
    B, k, _ = idx.shape
    idx = idx.contiguous()
    if out is None:
        out = torch.empty((B, k, 1, L_out), dtype=torch.float32, device=ds.device)
    with torch.cuda.device(ds.device):
        rc = L.psh_gather_paths(ds.data_ptr(), ds.shape[0], T, ds.stride(0), idx.data_ptr(), B * k, row_offset,
                                L_out, out.data_ptr(), _stream(ds))
    _check(rc, "psh_gather_paths")
    return out


def rv_aggregate(paths: torch.Tensor, dist: torch.Tensor, H: int, Ts: torch.Tensor, eta: float, proba: int,
                 vol: bool):
    """paths (B,k,1,L), dist (B,k), Ts (nT) i32 ascending -> mean, std (B,nT) f32."""
    L = lib()
    B, k = dist.shape
    Lp = paths.shape[-1]
    paths = paths.contiguous()
    dist = dist.contiguous()
    nT = Ts.numel()
    mean = torch.empty((B, nT), dtype=torch.float32, device=paths.device)
    std = torch.empty((B, nT), dtype=torch.float32, device=paths.device)
    with torch.cuda.device(paths.device):
        rc = L.psh_rv_aggregate(paths.data_ptr(), dist.data_ptr(), B, k, Lp, H, Ts.data_ptr(), nT,
                                float(eta if eta is not None else 0.0), proba, 1 if vol else 0, mean.data_ptr(),
                                std.data_ptr(), _stream(paths))
    _check(rc, "psh_rv_aggregate")
    return mean, std


# ---------------------------------------------------------------------------------------------
# peer-memory exchange (multi-GPU): buffers every rank maps through CUDA IPC
# ---------------------------------------------------------------------------------------------
def xchg_bytes(G: int, B: int, k: int) -> int:
    return int(lib().psh_xchg_bytes(G, B, k))


def xchg_create(nbytes: int, device: torch.device) -> tuple[int, bytes]:
    """cudaMalloc + zero an exchange buffer on `device`; returns (device pointer, 64-byte IPC handle)."""
    L = lib()
    ptr = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    with torch.cuda.device(device):
        _check(L.psh_xchg_create(nbytes, ctypes.byref(ptr), handle), "psh_xchg_create")
    return int(ptr.value), handle.raw


def xchg_open(handle: bytes, device: torch.device) -> int:
    L = lib()
    ptr = ctypes.c_void_p()
    with torch.cuda.device(device):
        _check(L.psh_xchg_open(handle, ctypes.byref(ptr)), "psh_xchg_open")
    return int(ptr.value)


def xchg_close(ptr: int, device: torch.device) -> None:
    with torch.cuda.device(device):
        lib().psh_xchg_close(ptr)


def xchg_destroy(ptr: int, device: torch.device) -> None:
    with torch.cuda.device(device):
        lib().psh_xchg_destroy(ptr)


def allgather_merge_packed(rec: torch.Tensor, bufs: list[int], rank: int, Tp: int, epoch: int,
                           flag: torch.Tensor | None = None, stream: int | None = None, arr=None):
    """rec (B,k,3) i32 [distance bits, r, t] of this rank -> merged (B,k) f32, (B,k,2) i32 over all
    ranks: ONE kernel stores the records into every rank's exchange buffer over NVLink, waits for
    the peers' records and merges.  `flag` bit 0: a shard overflowed; bit 1: a peer timed out."""
    L = lib()
    B, k, _ = rec.shape
    G = len(bufs)
    if arr is None:                      # (callers on the hot path keep the pointer array: _PeerExchange.arr)
        arr = (ctypes.c_void_p * G)(*bufs)
    dist = torch.empty((B, k), dtype=torch.float32, device=rec.device)
    idx = torch.empty((B, k, 2), dtype=torch.int32, device=rec.device)
    with _on_device(rec.device):
        rc = L.psh_allgather_merge_packed(rec.data_ptr(), arr, G, rank, B, k, Tp, epoch, dist.data_ptr(),
                                          idx.data_ptr(), flag.data_ptr() if flag is not None else None,
                                          _stream(rec) if stream is None else stream)
    _check(rc, "psh_allgather_merge_packed")
    return dist, idx


def xchg_send(rec: torch.Tensor, bufs: list[int], rank: int, epoch: int) -> None:
    """First half of allgather_merge_packed: store this rank's records into every rank's buffer and
    raise the flags; never waits."""
    L = lib()
    B, k, _ = rec.shape
    G = len(bufs)
    arr = (ctypes.c_void_p * G)(*bufs)
    with torch.cuda.device(rec.device):
        rc = L.psh_xchg_send(rec.data_ptr(), arr, G, rank, B, k, epoch, _stream(rec))
    _check(rc, "psh_xchg_send")


def xchg_merge(bufs: list[int], rank: int, B: int, k: int, Tp: int, epoch: int, dist: torch.Tensor,
               idx: torch.Tensor, flag: torch.Tensor | None = None) -> None:
    """Second half: wait for the epoch's records of all ranks, merge into dist (B,k) / idx (B,k,2)."""
    L = lib()
    G = len(bufs)
    arr = (ctypes.c_void_p * G)(*bufs)
    with torch.cuda.device(dist.device):
        rc = L.psh_xchg_merge(arr, G, rank, B, k, Tp, epoch, dist.data_ptr(), idx.data_ptr(),
                              flag.data_ptr() if flag is not None else None, _stream(dist))
    _check(rc, "psh_xchg_merge")
