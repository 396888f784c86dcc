"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Two restatements of the reference path (RudyMorel/shadowing @ 751a800):

* `shadow_topk` / `distances` / `gather_paths`  -> ctypes calls into `liboracle.so`
  (oracle/shadow_oracle.c, multi-threaded, used at every size the tests and the bench
  `cpu_baseline` leg need);
* `np_distances` / `np_shadow_topk`            -> an independent numpy restatement of the same
  6-line arithmetic spec, for small cases, so the C code is itself cross-checked.

plus `predict_from_paths` -- numpy restatement of `path_shadowing.py:234-254` with
`statistics.py:5-16` and the (parity-UNPINNED, scatspectra-resident) Softmax/Uniform weights.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

Parity pin: tests/golden/*.npz were produced by the live reference (tests/gen_golden.py);
tests/test_oracle.py checks both restatements against them bit-for-bit.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "shadow_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "liboracle.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "liboracle.so"
        if not so.exists():
            build()
        L = ctypes.CDLL(str(so))
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        i64 = ctypes.c_int64
        L.orc_qnorm.restype = ctypes.c_float
        L.orc_qnorm.argtypes = [fp, ctypes.c_int]
        L.orc_distances.restype = None
        L.orc_distances.argtypes = [fp, i64, i64, i64, i64, fp, ctypes.c_int, ctypes.c_int, fp]
        L.orc_shadow_topk.restype = ctypes.c_int
        L.orc_shadow_topk.argtypes = [fp, i64, i64, i64, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      i64, ctypes.c_int32, fp, ip, ctypes.c_int]
        L.orc_embed_topk.restype = ctypes.c_int
        L.orc_embed_topk.argtypes = [fp, i64, i64, i64, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp,
                                     ctypes.c_int, i64, ctypes.c_int32, fp, ip, ctypes.c_int]
        L.orc_embed_row.restype = None
        L.orc_embed_row.argtypes = [fp, i64, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        L.orc_gather_paths.restype = None
        L.orc_gather_paths.argtypes = [fp, i64, ip, i64, ctypes.c_int, fp]
        L.orc_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def num_threads() -> int:
    return int(lib().orc_num_threads())


def _rows(dataset) -> np.ndarray:
    """(R,1,T) / (R,T) / (T,) -> (R,T) float32, as `_dim_array` (path_shadowing.py:16-26), C=1."""
    ds = np.asarray(dataset)
    if ds.ndim == 1:
        ds = ds[None, :]
    if ds.ndim == 3:
        assert ds.shape[1] == 1, "the Identity/conv1d path is single-channel (path_embedding.py:130)"
        ds = ds[:, 0, :]
    return _f32(ds)


def _queries(x) -> np.ndarray:
    q = np.asarray(x)
    if q.ndim == 1:
        q = q[None, :]
    if q.ndim == 3:
        assert q.shape[1] == 1
        q = q[:, 0, :]
    return _f32(q)


def qnorm(q) -> np.float32:
    q = _f32(q)
    return np.float32(lib().orc_qnorm(_fp(q), int(q.shape[-1])))


def distances(dataset, q, H: int) -> np.ndarray:
    """All distances of ONE query: (R, T') float32 (C oracle)."""
    ds = _rows(dataset)
    q = _f32(q).reshape(-1)
    R, T = ds.shape
    W = q.shape[0]
    Tp = T - W - H + 1
    out = np.empty((R, Tp), np.float32)
    lib().orc_distances(_fp(ds), T, T, 0, R, _fp(q), W, H, _fp(out))
    return out


def shadow_topk(dataset, x_context, k: int, H: int, row_offset: int = 0, nthreads: int = 0):
    """(d (B,k) f32 ascending, idx (B,k,2) i32 [r,t]); ties by (d, r*T'+t)."""
    ds = _rows(dataset)
    q = _queries(x_context)
    R, T = ds.shape
    B, W = q.shape
    d = np.empty((B, k), np.float32)
    idx = np.empty((B, k, 2), np.int32)
    rc = lib().orc_shadow_topk(_fp(ds), R, T, T, _fp(q), B, W, H, k, row_offset, _fp(d),
                               idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle: invalid arguments (rc={rc}): k={k} windows={R * (T - W - H + 1)}")
    return d, idx


def gather_paths(dataset, idx: np.ndarray, L: int) -> np.ndarray:
    """paths (B,k,1,L) = dataset[r, t:t+L]  (path_shadowing.py:210-216)."""
    ds = _rows(dataset)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, k, _ = idx.shape
    out = np.empty((B, k, 1, L), np.float32)
    lib().orc_gather_paths(_fp(ds), ds.shape[1], idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                           B * k, L, _fp(out))
    return out


def shadow(dataset, x_context, k: int, H: int):
    """Restatement of PathShadowing.shadow (path_shadowing.py:181-218): (d, paths, idx)."""
    q = _queries(x_context)
    d, idx = shadow_topk(dataset, q, k, H)
    return d, gather_paths(dataset, idx, q.shape[1] + H), idx


# ----------------------------------------------------------------------------------------
# general linear embedding (PathEmbedding(kernel), Foveal): path_embedding.py:117-132,142-172
# ----------------------------------------------------------------------------------------
def foveal_kernel(alpha: float, beta: float, max_context: int) -> np.ndarray:
    """(d, W) kernel of Foveal(alpha, beta, max_context) restated from path_embedding.py:152-170:
    d = floor(log W / log alpha) trailing boxes of lengths int(alpha**n), weight length**-beta."""
    dim = int(np.floor(np.log(max_context) / np.log(alpha)))
    K = np.zeros((dim, max_context), np.float32)
    for n in range(1, dim + 1):
        le = int(alpha ** n)
        K[n - 1, max_context - le:] = np.float32(le ** (-beta))
    return K


def embed_queries(kernel, x_context) -> np.ndarray:
    """ex (B, d): the context embedded (one window); fp64 dot products rounded once."""
    K = _f32(kernel).reshape(kernel.shape[0], -1)
    q = _queries(x_context)
    return (K.astype(np.float64) @ q.astype(np.float64).T).T.astype(np.float32)


def embed_row(y, kernel, H: int) -> np.ndarray:
    """(T', d) embedded windows of one trajectory."""
    K = _f32(kernel).reshape(kernel.shape[0], -1)
    y = _f32(y).reshape(-1)
    d, W = K.shape
    out = np.empty((y.shape[0] - W - H + 1, d), np.float32)
    lib().orc_embed_row(_fp(y), y.shape[0], _fp(K), d, W, H, _fp(out))
    return out


def embed_topk(dataset, kernel, ex, k: int, H: int, row_offset: int = 0, nthreads: int = 0):
    """k closest windows in EMBEDDED space: ex (B, d) embedded queries, kernel (d, W) or (d,1,W).
    (d (B,k) f32 ascending, idx (B,k,2) i32 [r,t]); ties by (d, r*T'+t)."""
    ds = _rows(dataset)
    K = _f32(kernel).reshape(kernel.shape[0], -1)
    ex = _f32(ex)
    R, T = ds.shape
    d_, W = K.shape
    B = ex.shape[0]
    assert ex.shape[1] == d_
    d = np.empty((B, k), np.float32)
    idx = np.empty((B, k, 2), np.int32)
    rc = lib().orc_embed_topk(_fp(ds), R, T, T, _fp(K), d_, W, H, _fp(ex), B, k, row_offset, _fp(d),
                              idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle: invalid arguments (rc={rc})")
    return d, idx


# ----------------------------------------------------------------------------------------
# independent numpy restatement (small cases)
# ----------------------------------------------------------------------------------------
def np_qnorm(q) -> np.float32:
    q = _f32(q).reshape(-1)
    W = q.shape[0]
    n8 = (W // 8) * 8
    acc = np.zeros(8, np.float32)
    sq = (q * q).astype(np.float32)
    for j in range(n8):
        acc[j % 8] = np.float32(acc[j % 8] + sq[j])
    s = np.float32(0.0)
    for lane in range(8):
        s = np.float32(s + acc[lane])
    for j in range(n8, W):
        s = np.float32(s + sq[j])
    return np.sqrt(s, dtype=np.float32)


def np_distances(dataset, q, H: int) -> np.ndarray:
    ds = _rows(dataset)
    q = _f32(q).reshape(-1)
    R, T = ds.shape
    W = q.shape[0]
    Tp = T - W - H + 1
    s = np.zeros((R, Tp), np.float32)
    for j in range(W):  # sequential in j, non-fused float32: the reference's reduction order
        df = (q[j] - ds[:, j:j + Tp]).astype(np.float32)
        s = (s + (df * df).astype(np.float32)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (np.sqrt(s, dtype=np.float32) / np_qnorm(q)).astype(np.float32)


def np_shadow_topk(dataset, x_context, k: int, H: int):
    ds = _rows(dataset)
    q = _queries(x_context)
    R, T = ds.shape
    B, W = q.shape
    Tp = T - W - H + 1
    d = np.empty((B, k), np.float32)
    idx = np.empty((B, k, 2), np.int32)
    for b in range(B):
        flat = np_distances(ds, q[b], H).ravel()
        order = np.lexsort((np.arange(flat.size), flat.view(np.uint32)))[:k]
        d[b] = flat[order]
        idx[b, :, 0] = order // Tp
        idx[b, :, 1] = order % Tp
    return d, idx


# ----------------------------------------------------------------------------------------
# predict_from_paths (path_shadowing.py:234-254) -- Softmax/Uniform live in scatspectra v2.0.2,
# which is NOT under /root/reference: the weight formula below is PARITY-UNPINNED.
# ----------------------------------------------------------------------------------------
def realized_variance(x: np.ndarray, Ts, vol: bool) -> np.ndarray:
    """statistics.py:5-16 restated."""
    x2 = x ** 2
    rv = np.stack([x2[..., :T].mean(-1) for T in Ts], -1) * 252
    return rv ** 0.5 if vol else rv


def softmax_weights(distances: np.ndarray, eta: float, axis: int, dtype=np.float64) -> np.ndarray:
    """w ∝ exp(-d²/(2η²)) normalised along `axis` ("the width of a Gaussian in the Gaussian
    average", plot_utils.py:59-65).  Shifted by min d² for stability (cancels in the ratio)."""
    d = np.asarray(distances, dtype=dtype)
    e = -(d ** 2) / (2.0 * eta ** 2)
    e = e - e.max(axis=axis, keepdims=True)
    w = np.exp(e)
    return w / w.sum(axis=axis, keepdims=True)


def predict_from_paths(distances: np.ndarray, paths: np.ndarray, H: int, Ts, vol: bool,
                       proba_name: str, eta, dtype=np.float64):
    """(pred, std) over axis 1 of realized_variance(out-context)[:, :, 0, :]  -> (B, nT) each.
    avg = Σ w x ;  std = sqrt(Σ w x² − (Σ w x)²)."""
    out = paths[..., -H:] if H else paths
    x = realized_variance(out.astype(dtype), Ts, vol)[:, :, 0, :]
    if proba_name == "uniform":
        w = np.full(distances.shape + (1,), 1.0 / distances.shape[1], dtype)
    elif proba_name == "softmax":
        w = softmax_weights(distances[:, :, None], eta, axis=1, dtype=dtype)
    else:
        raise ValueError("Unrecognized averaging proba")
    avg = (w * x).sum(1)
    var = (w * x * x).sum(1) - avg ** 2
    return avg, np.sqrt(np.maximum(var, 0.0))
