"""GPU parity: the CUDA path (through the reference-shaped Python API and the raw C ABI)
against the live-reference golden fixtures and the CPU oracle.  Integer outputs bit-exact,
distances bit-exact (the scan reproduces the reference's fp32 sequence), predictions within
1e-6 relative (north_star's stated tolerance)."""
import math

import numpy as np
import pytest
import torch

from conftest import assert_topk_equal, load_golden, make_inputs
from oracle import oracle

pytestmark = pytest.mark.gpu

import shadowing_b200 as sb  # noqa: E402
from shadowing_b200 import _lib  # noqa: E402


def _obj(ds, W, H, **kw):
    return sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H), **kw)


@pytest.mark.parametrize("mode", ["exact", "filter", "fft"])
def test_golden_shadow_bit_exact(golden, mode):
    obj = _obj(golden["dataset"], golden["W"], golden["H"], scan_mode=mode)
    d, paths, idx = obj.shadow(golden["x_context"], k=golden["k"], n_splits=golden["n_splits"], cuda=True)
    assert d.dtype == np.float32 and paths.dtype == np.float32 and idx.dtype == np.int32
    assert_topk_equal(d, idx, golden["distances"], golden["indices"])
    assert np.array_equal(idx, golden["indices"])
    assert paths.shape == golden["paths"].shape
    assert np.array_equal(paths, golden["paths"])


@pytest.mark.parametrize("mode", ["exact", "filter", "fft"])
@pytest.mark.parametrize("R,T,W,H,k,B", [
    (512, 4096, 252, 20, 1024, 2),    # north-star window/k on a subsample of rows
    (300, 1000, 100, 0, 77, 3),       # no horizon, odd sizes
    (64, 4099, 17, 3, 5000, 1),       # T % 4 != 0 (padded row stride), k > SEG
    (2048, 512, 20, 20, 16, 40),      # many queries: several query groups
    (5, 300, 252, 20, 1, 2),          # k = 1, fewer rows than seed rows
    (16, 2000, 1, 0, 9, 2),           # W = 1
    (24, 8192, 252, 20, 500, 2),      # T > 4096: the fft flavour cuts rows into overlapping pieces
    (7, 12001, 100, 0, 300, 1),       # odd number of virtual rows, ragged last piece
])
def test_oracle_parity_seeded(R, T, W, H, k, B, mode):
    ds, q = make_inputs(R, T, W, B, seed=100 + R)
    obj = _obj(ds, W, H or None, scan_mode=mode)
    d, paths, idx = obj.shadow(q, k=k)
    do, po, io = oracle.shadow(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)
    assert np.array_equal(paths, po)


def test_k_equals_all_windows():
    ds, q = make_inputs(7, 90, 10, 2, seed=5)
    n = 7 * (90 - 10 - 5 + 1)
    d, _, idx = _obj(ds, 10, 5).shadow(q, k=n)
    do, io = oracle.shadow_topk(ds, q, n, 5)
    assert_topk_equal(d, idx, do, io)


def test_large_k_global_sort_path():
    ds, q = make_inputs(40, 1500, 30, 1, seed=9)
    k = 20000  # > 16384: the finalise sort runs in global memory
    d, _, idx = _obj(ds, 30, 10).shadow(q, k=k)
    do, io = oracle.shadow_topk(ds, q, k, 10)
    assert_topk_equal(d, idx, do, io)


def test_unaligned_rows_take_the_plain_load_path():
    """row_stride % 4 != 0: no TMA bulk copy possible, the warp loads its segment itself."""
    ds, q = make_inputs(33, 1001, 50, 2, seed=21)
    rows = torch.tensor(ds[:, 0, :]).cuda()  # stride 1001
    qd = torch.tensor(q[:, 0, :]).cuda()
    for mode in (_lib.PSH_MODE_EXACT, _lib.PSH_MODE_FILTER, _lib.PSH_MODE_FFT):
        aux = _lib.fft_prepare(rows, 1001, 50, 7) if mode == _lib.PSH_MODE_FFT else None
        d, idx, _ = _lib.scan_topk(rows, 1001, qd, 7, 300, 0, mode, None, aux)
        do, io = oracle.shadow_topk(ds, q, 300, 7)
        assert_topk_equal(d.cpu().numpy(), idx.cpu().numpy(), do, io)


def test_fft4096_matches_numpy_and_error_budget():
    """The library's 4096-point FFT against numpy (fp64), both directions, and the constant the
    lower bound relies on: |c^_t - c_t| <= CF u Qmax ynorm with CF = 512 (observed: a few units)."""
    rng = np.random.default_rng(0)
    ds = (rng.standard_normal((6, 4096)) * 0.01).astype(np.float32)
    rows = torch.tensor(ds).cuda()
    aux = _lib.fft_prepare(rows, 4096, 252, 20)
    z = (rng.standard_normal((5, 4096)) + 1j * rng.standard_normal((5, 4096))).astype(np.complex64)
    zd = torch.tensor(z).cuda()
    # direction -1: forward, table twiddles (dataset spectra); +1: inverse, table twiddles;
    # 3 / 4: the scan's inverse (packed fp32 arithmetic; pass-B twiddles rebuilt from register-resident
    # seeds / read from the shared table)
    inv = np.fft.ifft(z.astype(np.complex128), axis=1) * 4096
    for direction, ref in ((-1, np.fft.fft(z.astype(np.complex128), axis=1)), (1, inv), (3, inv)):
        out = _lib.debug_fft4096(zd, direction, aux).cpu().numpy()
        err = np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert err.max() < 40 * 2.0 ** -24, (direction, err)   # theory: ~80-120 u worst case, few u typical
    # full pipeline: forward (library), pointwise conj(Q)/N, inverse (library) vs fp64 correlation
    q = (rng.standard_normal(252) * 0.01).astype(np.float32)
    pair = (ds[0] + 1j * ds[1]).astype(np.complex64)[None]
    Z = _lib.debug_fft4096(torch.tensor(pair).cuda(), -1, aux)
    Q = np.fft.fft(np.pad(q.astype(np.float64), (0, 4096 - 252)))
    Qc = torch.tensor((np.conj(Q) / 4096).astype(np.complex64)).cuda()
    c = _lib.debug_fft4096(Z * Qc[None], 3, aux).cpu().numpy()[0]
    Tp = 4096 - 252 + 1
    ca = np.array([np.dot(q.astype(np.float64), ds[0, t:t + 252].astype(np.float64)) for t in range(Tp)])
    cb = np.array([np.dot(q.astype(np.float64), ds[1, t:t + 252].astype(np.float64)) for t in range(Tp)])
    err = max(np.abs(c.real[:Tp] - ca).max(), np.abs(c.imag[:Tp] - cb).max())
    bound_unit = 2.0 ** -24 * np.abs(Q).max() * np.sqrt((ds[0].astype(np.float64) ** 2).sum() + (ds[1].astype(np.float64) ** 2).sum())
    assert err < 16 * bound_unit, (err / bound_unit)   # CF = 512 leaves >= 32x head-room


def test_fft1024_matches_numpy_and_error_budget():
    """The warp-level 1024-point transform (pshadow_fft3.cuh) against numpy (fp64), both directions, and the
    constant the lower bound relies on, through the full pipeline forward -> conj(Q)/N -> inverse."""
    rng = np.random.default_rng(1)
    ds = (rng.standard_normal((6, 1024)) * 0.01).astype(np.float32)
    aux = _lib.fft_prepare(torch.tensor(ds).cuda(), 1024, 252, 20)
    z = (rng.standard_normal((7, 1024)) + 1j * rng.standard_normal((7, 1024))).astype(np.complex64)
    zd = torch.tensor(z).cuda()
    for direction, ref in ((-1, np.fft.fft(z.astype(np.complex128), axis=1)),
                           (1, np.fft.ifft(z.astype(np.complex128), axis=1) * 1024)):
        out = _lib.debug_fft1024(zd, direction, aux).cpu().numpy()
        err = np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)
        assert err.max() < 40 * 2.0 ** -24, (direction, err)
    q = (rng.standard_normal(252) * 0.01).astype(np.float32)
    pair = (ds[0] + 1j * ds[1]).astype(np.complex64)[None]
    Z = _lib.debug_fft1024(torch.tensor(pair).cuda(), -1, aux)
    Q = np.fft.fft(np.pad(q.astype(np.float64), (0, 1024 - 252)))
    Qc = torch.tensor((np.conj(Q) / 1024).astype(np.complex64)).cuda()
    c = _lib.debug_fft1024(Z * Qc[None], 1, aux).cpu().numpy()[0]
    Tp = 1024 - 252 + 1
    ca = np.array([np.dot(q.astype(np.float64), ds[0, t:t + 252].astype(np.float64)) for t in range(Tp)])
    cb = np.array([np.dot(q.astype(np.float64), ds[1, t:t + 252].astype(np.float64)) for t in range(Tp)])
    err = max(np.abs(c.real[:Tp] - ca).max(), np.abs(c.imag[:Tp] - cb).max())
    bound_unit = 2.0 ** -24 * np.abs(Q).max() * np.sqrt((ds[0].astype(np.float64) ** 2).sum() + (ds[1].astype(np.float64) ** 2).sum())
    assert err < 16 * bound_unit, (err / bound_unit)   # CF = 512 leaves >= 32x head-room


@pytest.mark.parametrize("nfft", ["1024", "4096"])
@pytest.mark.parametrize("R,T,W,H,k,B", [
    (2048, 4096, 252, 20, 1024, 1),   # north-star window/k: seedless schedule, one query (spectrum staged in the tile)
    (2048, 4096, 252, 20, 1024, 3),   # a group of queries (separate spectrum buffer, the pair transformed twice while seeding)
    (600, 3000, 500, 10, 200, 2),     # long context: 1024-point pieces yield 525 windows each
    (37, 1024, 64, 0, 50, 1),         # T = one piece exactly
    (9, 1030, 300, 3, 40, 2),         # a second piece holding a handful of windows
    (3, 5000, 380, 0, 4000, 1),       # odd number of rows, k close to the number of windows
])
def test_fft_flavours_bit_exact(R, T, W, H, k, B, nfft, monkeypatch):
    """Both transform lengths of the fft flavour (warp-level 1024-point pieces / CTA-level 4096-point rows),
    forced through PSH_FFT_N, against the oracle."""
    monkeypatch.setenv("PSH_FFT_N", nfft)
    ds, q = make_inputs(R, T, W, B, seed=4000 + R)
    obj = _obj(ds, W, H or None, scan_mode="fft")
    d, paths, idx = obj.shadow(q, k=k)
    do, po, io = oracle.shadow(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)
    assert np.array_equal(paths, po)


@pytest.mark.parametrize("prepared_with", ["4096", "1024"])
def test_scan_uses_the_length_its_aux_was_prepared_with(prepared_with, monkeypatch):
    """The PSH_FFT_N knob may change between psh_fft_prepare and the scan: the library remembers the transform
    length every aux buffer was laid out with."""
    ds, q = make_inputs(64, 2048, 100, 2, seed=11)
    rows = torch.tensor(ds[:, 0, :]).cuda()
    qd = torch.tensor(q[:, 0, :]).cuda()
    monkeypatch.setenv("PSH_FFT_N", prepared_with)
    aux = _lib.fft_prepare(rows, 2048, 100, 5)
    monkeypatch.setenv("PSH_FFT_N", "1024" if prepared_with == "4096" else "4096")
    d, idx, _ = _lib.scan_topk(rows, 2048, qd, 5, 50, 0, _lib.PSH_MODE_FFT, None, aux)
    do, io = oracle.shadow_topk(ds, q, 50, 5)
    assert_topk_equal(d.cpu().numpy(), idx.cpu().numpy(), do, io)


def _perm_stride(R):
    if R <= 2:
        return 1
    p = max(int(0.6180339887498949 * R), 1)
    while math.gcd(p, R) != 1:
        p += 1
    return p % R or 1


@pytest.mark.parametrize("mode", ["exact", "filter", "fft"])
def test_adversarial_order_overflows_then_recovers(mode):
    """Seed rows far, every other row near: the candidate buffer overflows and the scan must
    re-run in its safe schedule and still return the exact answer."""
    R, T, W, H, k = 512, 1024, 32, 0, 64
    ds, q = make_inputs(R, T, W, 1, seed=33)
    Tp = T - W + 1
    n0 = min(max(-(-16 * k // Tp), 1), R)
    P = _perm_stride(R)
    seed_rows = {(i * P) % R for i in range(n0)}
    ds = ds * 1e-3
    for r in seed_rows:
        ds[r] *= 1e5
    d, _, idx = _obj(ds, W, None, scan_mode=mode).shadow(q, k=k)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)


def test_nosync_packed_records_and_overflow_flag():
    """The multi-GPU building blocks on one device: packed [distance bits, r, t] records from a
    PSH_FLAG_NOSYNC scan, the deferred overflow status, and the poisoned record that makes
    psh_merge_topk_packed raise its flag when a shard overflowed."""
    ds, q = make_inputs(200, 1500, 64, 3, seed=41)
    rows = torch.tensor(ds[:, 0, :]).cuda()
    qd = torch.tensor(q[:, 0, :]).cuda()
    k, H = 100, 6
    do, io = oracle.shadow_topk(ds, q, k, H, row_offset=7)
    for mode in (_lib.PSH_MODE_FILTER, _lib.PSH_MODE_FFT):
        aux = _lib.fft_prepare(rows, 1500, 64, H) if mode == _lib.PSH_MODE_FFT else None
        rec = torch.empty((3, k, 3), dtype=torch.int32, device="cuda")
        ws = _lib.scan_topk_packed(rows, 1500, qd, H, k, 7, mode | _lib.PSH_FLAG_NOSYNC, None, aux, rec)
        assert not _lib.scan_overflowed(ws, 3)
        r = rec.cpu().numpy()
        assert np.array_equal(r[..., 0].view(np.uint32), do.view(np.uint32)) and np.array_equal(r[..., 1:], io)
    # adversarial order (as in test_adversarial_order_overflows_then_recovers): the NOSYNC scan must report it
    R, T, W, k = 512, 1024, 32, 64
    ds, q = make_inputs(R, T, W, 1, seed=33)
    Tp = T - W + 1
    n0 = min(max(-(-16 * k // Tp), 1), R)
    P = _perm_stride(R)
    ds = ds * 1e-3
    for i in range(n0):
        ds[(i * P) % R] *= 1e5
    rows = torch.tensor(ds[:, 0, :]).cuda()
    qd = torch.tensor(q[:, 0, :]).cuda()
    rec = torch.empty((1, k, 3), dtype=torch.int32, device="cuda")
    ws = _lib.scan_topk_packed(rows, T, qd, 0, k, 0, _lib.PSH_MODE_FILTER | _lib.PSH_FLAG_NOSYNC, None, None, rec)
    assert _lib.scan_overflowed(ws, 1)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.merge_topk_packed(rec[None], Tp, flag)
    assert int(flag.item()) == 1
    ws = _lib.scan_topk_packed(rows, T, qd, 0, k, 0, _lib.PSH_MODE_FILTER, ws, None, rec)   # synchronous: safe re-run
    do, io = oracle.shadow_topk(ds, q, k, 0)
    r = rec.cpu().numpy()
    assert np.array_equal(r[..., 0].view(np.uint32), do.view(np.uint32)) and np.array_equal(r[..., 1:], io)
    flag.zero_()
    _lib.merge_topk_packed(rec[None], Tp, flag)
    assert int(flag.item()) == 0


def test_ties_are_ordered_by_flat_index():
    """Identical rows: every distance value appears R times; order must be (d, r*T'+t)."""
    base, q = make_inputs(1, 600, 40, 1, seed=3)
    ds = np.repeat(base, 50, axis=0)
    d, _, idx = _obj(ds, 40, 10).shadow(q, k=333)
    do, io = oracle.shadow_topk(ds, q, 333, 10)
    assert np.array_equal(d, do) and np.array_equal(idx, io)


def test_zero_query_all_inf():
    ds, _ = make_inputs(4, 200, 16, 1)
    d, _, idx = _obj(ds, 16, None).shadow(np.zeros((1, 1, 16), np.float32), k=10)
    assert np.isinf(d).all()
    assert np.array_equal(idx[0, :, 0], np.zeros(10)) and np.array_equal(idx[0, :, 1], np.arange(10))


def test_input_conventions_and_errors():
    ds, q = make_inputs(8, 256, 20, 2, seed=1)
    obj = _obj(ds[:, 0, :], 20, 5)  # 2-D dataset
    d3, p3, i3 = obj.shadow(q, k=4)
    d2, p2, i2 = obj.shadow(q[:, 0, :], k=4)            # 2-D contexts
    d1, p1, i1 = obj.shadow(q[0, 0, :], k=4)            # 1-D context
    dt, _, _ = obj.shadow(torch.tensor(q), k=4)         # torch input
    d64, _, _ = obj.shadow(q.astype(np.float64), k=4)   # numpy f64 is cast (path_shadowing.py:33)
    assert np.array_equal(d3, d2) and np.array_equal(d3[:1], d1) and np.array_equal(d3, dt)
    assert np.array_equal(d3, d64) and p3.shape == (2, 4, 1, 25) and p1.shape == (1, 4, 1, 25)
    with pytest.raises(Exception, match="same size as the context"):
        obj.shadow(q[:, :, :10], k=4)
    with pytest.raises(RuntimeError):
        obj.shadow(torch.tensor(q, dtype=torch.float64), k=4)   # torch f64 context (reference raises too)
    with pytest.raises(RuntimeError):
        obj.shadow(q, k=8 * 256)                                 # k > number of windows
    with pytest.raises(ValueError):
        obj.predict_from_paths(d3, p3, sb.RealizedVariance([2]), "gaussian", 0.1)
    class OtherContext(sb.PredictionContext):   # only PredictionContext itself runs on the device
        pass
    bad = sb.PathShadowing(sb.Identity(20), sb.RelativeMSE(), ds, OtherContext(5))
    with pytest.raises(NotImplementedError):
        bad.shadow(q, k=4)


def test_batched_distance_contract():
    ds, q = make_inputs(16, 300, 24, 3, seed=2)
    obj = _obj(ds, 24, 6)
    d, i = obj.batched_distance(torch.tensor(q), torch.tensor(ds), k=12, n_splits=4, cuda=True)
    assert isinstance(d, torch.Tensor) and d.device.type == "cpu" and i.dtype == torch.int32
    do, io = oracle.shadow_topk(ds, q, 12, 6)
    assert_topk_equal(d.numpy(), i.numpy(), do, io)


def test_batched_distance_foreign_y_never_reuses_stale_spectra():
    """batched_distance(x, y) with a `y` that is not the resident dataset, twice, same shapes, fft
    flavour: the second call must not filter with the first y's spectra / window energies (the
    caching allocator hands the second upload the first one's address)."""
    R, T, W, H, k = 64, 2048, 64, 4, 50
    obj = _obj(make_inputs(R, T, W, 1, seed=41)[0], W, H, scan_mode="fft")
    for seed in (42, 43, 44):
        y, q = make_inputs(R, T, W, 2, seed=seed)
        d, idx = obj.batched_distance(torch.tensor(q), torch.tensor(y), k, 1, False)
        do, io = oracle.shadow_topk(y, q, k, H)
        assert_topk_equal(d.numpy(), idx.numpy(), do, io)
    # and the resident dataset's own cache is still valid afterwards
    q = make_inputs(R, T, W, 2, seed=41)[1]
    d, _, idx = obj.shadow(q, k=k)
    do, io = oracle.shadow_topk(obj.dataset, q, k, H)
    assert_topk_equal(d, idx, do, io)


def test_testing_ipynb_self_consistency():
    """testing.ipynb:62-78 re-expressed for Identity: re-embedding the returned paths and
    re-computing the distance reproduces the returned distances (their rtol 1e-2; here 1e-6)."""
    ds, q = make_inputs(32, 4096, 126, 8, seed=4)
    obj = _obj(ds, 126, 252)
    d, paths, _ = obj.shadow(q, k=1024)
    pin = torch.tensor(obj.context.select_in_context(paths))[:, :, 0, :]
    ds_test = obj.distance(torch.tensor(q), pin)
    assert torch.allclose(ds_test, torch.tensor(d), rtol=1e-6)


@pytest.mark.parametrize("proba,eta,vol", [("softmax", 0.1, False), ("softmax", 0.02, True), ("uniform", None, False)])
def test_predict_from_paths_matches_oracle(proba, eta, vol):
    ds, q = make_inputs(256, 2048, 60, 5, seed=8)
    obj = _obj(ds, 60, 20)
    d, paths, _ = obj.shadow(q, k=512)
    Ts = [10, 5, 20]  # unsorted on purpose
    mean, std = obj.predict_from_paths(d, paths, sb.RealizedVariance(Ts, vol), proba, eta)
    mo, so = oracle.predict_from_paths(d, paths, 20, Ts, vol, proba, eta)
    assert mean.shape == (5, 3) and mean.dtype == np.float32
    assert np.allclose(mean, mo, rtol=1e-6, atol=0)
    assert np.allclose(std, so, rtol=1e-5, atol=0)
    # the host plugin path (arbitrary callable, as README.md:77-80) agrees with the fused kernel
    lam = lambda x: sb.realized_variance(x, Ts=Ts, vol=vol)[:, :, 0, :]  # noqa: E731
    mh, sh = obj.predict_from_paths(d, paths, lam, proba, eta)
    assert np.allclose(mh, mo, rtol=2e-6) and np.allclose(sh, so, rtol=1e-4)


def test_predict_end_to_end_chunks():
    ds, q = make_inputs(128, 1024, 40, 6, seed=12)
    obj = _obj(ds, 40, 20)
    rv = sb.RealizedVariance([5, 10, 20])
    p1, s1 = obj.predict(q, k=200, to_predict=rv, eta=0.1, n_context_splits=1)
    p3, s3 = obj.predict(q, k=200, to_predict=rv, eta=0.1, n_context_splits=3, n_dataset_splits=8, cuda=True)
    assert p1.shape == (6, 3) and np.array_equal(p1, p3) and np.array_equal(s1, s3)
    d, paths, _ = obj.shadow(q, k=200)
    mo, so = oracle.predict_from_paths(d, paths, 20, [5, 10, 20], False, "softmax", 0.1)
    assert np.allclose(p1, mo, rtol=1e-6) and np.allclose(s1, so, rtol=1e-5)


@pytest.mark.parametrize("force_sort", ["0", "1"])   # merge by rank (default) / bitonic-sort merge
def test_merge_topk_equals_global(force_sort, monkeypatch):
    monkeypatch.setenv("PSH_MERGE_SORT", force_sort)
    ds, q = make_inputs(64, 700, 30, 3, seed=14)
    k, H = 128, 10
    rows = torch.tensor(ds[:, 0, :]).cuda()
    qd = torch.tensor(q[:, 0, :]).cuda()
    parts_d, parts_i = [], []
    for s in range(0, 64, 16):
        d, i, _ = _lib.scan_topk(rows[s:s + 16].contiguous(), 700, qd, H, k, s)
        parts_d.append(d)
        parts_i.append(i)
    d, i = _lib.merge_topk(torch.stack(parts_d), torch.stack(parts_i), 700 - 30 - H + 1)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert_topk_equal(d.cpu().numpy(), i.cpu().numpy(), do, io)
    # packed records [distance bits, r, t], the layout of the single all-gather
    rec = torch.cat([torch.stack(parts_d).view(torch.int32).unsqueeze(-1), torch.stack(parts_i)], dim=-1)
    d2, i2 = _lib.merge_topk_packed(rec, 700 - 30 - H + 1)
    assert torch.equal(d, d2) and torch.equal(i, i2)


def test_full_size_properties_cfg2():
    """BASELINE configs[1] (R=32768 x T=4096, W=252, k=1024) through size-independent
    properties: ascending distances, every returned (r,t) re-evaluates to its distance on the
    oracle, planted near-copies of the query are found at rank 0.., and the k-th distance
    bounds an oracle-evaluated random sample of windows."""
    R, T, W, H, k = 32768, 4096, 252, 20, 1024
    g = torch.Generator().manual_seed(0)
    ds = torch.randn(R, 1, T, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(1)
    q = torch.randn(1, 1, W, generator=g, dtype=torch.float32) * 0.01
    planted = [(31000, 77), (5, 3800), (16384, 0)]
    for n, (r, t) in enumerate(planted):
        ds[r, 0, t:t + W] = q[0, 0] * (1.0 + 0.01 * (n + 1))
    obj = _obj(ds, W, H)
    d, paths, idx = obj.shadow(q, k=k)
    assert (np.diff(d[0]) >= 0).all()
    assert [tuple(v) for v in idx[0, :3]] == planted
    dsn = ds.numpy()
    qn = q.numpy()[0, 0]
    # every winner's distance re-evaluated by the oracle on its own window
    for j in range(0, k, 7):
        r, t = idx[0, j]
        dd = oracle.distances(dsn[r:r + 1, :, t:t + W + H], qn, H)
        assert dd.shape == (1, 1) and dd[0, 0] == d[0, j]
        assert np.array_equal(paths[0, j, 0], dsn[r, 0, t:t + W + H])
    # no window of an oracle-scanned row sample beats the k-th distance without being returned
    rows = np.arange(0, R, 257)
    do, io = oracle.shadow_topk(dsn[rows], qn, 64, H)
    got = {tuple(v) for v in idx[0]}
    for dist_, (rr, tt) in zip(do[0], io[0]):
        if dist_ < d[0, -1]:
            assert (int(rows[rr]), int(tt)) in got


# ---------------------------------------------------------------------------------------------
# seedless fft schedule (large ensembles): seed launch -> one launch over all pairs -> re-rank ->
# select.  Results must be bit-identical to the oracle and to the exact-seeded chunk schedule.
# ---------------------------------------------------------------------------------------------
def _shadow_both_schedules(ds, q, W, H, k, monkeypatch):
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PSH_SEEDLESS", flag)
        obj = _obj(ds, W, H or None, scan_mode="fft")
        obj.shadow(q[:1], k=k)  # residency + fft aux
        n0 = _lib.launch_count()
        out[flag] = obj.shadow(q, k=k) + (_lib.launch_count() - n0,)
    return out


@pytest.mark.parametrize("R,T,W,H,k,B", [
    (2048, 4096, 252, 20, 1024, 1),   # north-star window/k, 1024 row pairs
    (2048, 4096, 252, 20, 1024, 3),   # several queries share the seed and the main launch
    (1100, 8192, 100, 0, 300, 2),     # virtual rows (T > 4096), no horizon
    (4096, 1024, 64, 8, 4096, 1),     # k at the limit of the fused final sort
    (4099, 1500, 80, 5, 5000, 1),     # k > SEL_LIST: separate finalise kernel, odd number of rows
])
def test_seedless_schedule_bit_exact(R, T, W, H, k, B, monkeypatch):
    ds, q = make_inputs(R, T, W, B, seed=700 + R)
    out = _shadow_both_schedules(ds, q, W, H, k, monkeypatch)
    do, po, io = oracle.shadow(ds, q, k, H)
    for flag in ("1", "0"):
        d, paths, idx, _ = out[flag]
        assert_topk_equal(d, idx, do, io)
        assert np.array_equal(paths, po)
    # the seedless schedule really ran: qprep, qfft, seed, scan, rerank, select (+finalise) + gather
    assert out["1"][3] < out["0"][3] and out["1"][3] <= 8


def test_seedless_scale_mismatch_recovers(monkeypatch):
    """Dataset 1000x the query's scale: every upper bound lies beyond the seed histogram's 16
    binades above Q2, the seed threshold stays +inf, the candidate buffer overflows and the safe
    schedule must still return the exact answer."""
    monkeypatch.setenv("PSH_SEEDLESS", "1")
    R, T, W, H, k = 2048, 2048, 64, 4, 128
    ds, q = make_inputs(R, T, W, 1, seed=77)
    ds = ds * 1000.0
    d, _, idx = _obj(ds, W, H, scan_mode="fft").shadow(q, k=k)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)


def test_seedless_many_near_copies(monkeypatch):
    """More than k windows closer than 2^-8 ||q|| (below the seed histogram's lowest bin): the seed
    threshold is the lowest bin's edge and the main launch tightens it by itself."""
    monkeypatch.setenv("PSH_SEEDLESS", "1")
    R, T, W, H, k = 2048, 2048, 64, 0, 512
    ds, q = make_inputs(R, T, W, 1, seed=78)
    rng = np.random.default_rng(5)
    reps = np.tile(q[0, 0], T // W)
    ds[::2, 0, :] = reps[None, :] * (1.0 + 1e-4 * rng.standard_normal((R // 2, T)).astype(np.float32))
    d, _, idx = _obj(ds, W, None, scan_mode="fft").shadow(q, k=k)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)


def test_seedless_adversarial_seed_pairs(monkeypatch):
    """The seed wave only sees far rows, everything else is near: thresholds start loose, the
    in-launch tightening (or, failing that, the overflow re-run) must keep the answer exact."""
    monkeypatch.setenv("PSH_SEEDLESS", "1")
    R, T, W, H, k = 4096, 1024, 32, 0, 64
    ds, q = make_inputs(R, T, W, 1, seed=79)
    npairs = R // 2
    P = _perm_stride(npairs)
    nseed = min(3 * torch.cuda.get_device_properties(0).multi_processor_count, npairs)   # every CTA's first pair
    ds = ds * 1e-3
    for i in range(nseed):
        pr = (i * P) % npairs
        ds[2 * pr] *= 1e5
        ds[2 * pr + 1] *= 1e5
    d, _, idx = _obj(ds, W, None, scan_mode="fft").shadow(q, k=k)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert_topk_equal(d, idx, do, io)


def test_nosync_pipeline_sticky_overflow():
    """A pipeline of enqueue-only scans on one workspace (bench.py's device loop): results equal
    the synchronous ones, and an overflow in ANY scan of the pipeline is still reported at the
    single check at its end (the per-call state is reset by every scan, the sticky flag is not)."""
    R, T, W, H, k = 2048, 2048, 64, 4, 128
    ds, q = make_inputs(R, T, W, 4, seed=91)
    obj = _obj(ds, W, H, scan_mode="fft")
    rows, T_ = obj._resident_rows()
    qd = torch.tensor(q).cuda()
    outs = []
    for i in range(4):
        out = (torch.empty((1, k), dtype=torch.float32, device="cuda"),
               torch.empty((1, k, 2), dtype=torch.int32, device="cuda"))
        outs.append(obj._scan_device(qd[i:i + 1], rows, T_, k, out=out, nosync=True))
    obj._check_pipeline()
    do, io = oracle.shadow_topk(ds, q, k, H)
    for i, (d, idx) in enumerate(outs):
        assert_topk_equal(d.cpu().numpy(), idx.cpu().numpy(), do[i:i + 1], io[i:i + 1])
    # scan 1 of 3 overflows (query at 1e-3 of the data's scale: seed threshold +inf), scans 2-3 are clean
    obj._scan_device(qd[:1] * 1e-3, rows, T_, k, nosync=True)
    obj._scan_device(qd[1:2], rows, T_, k, nosync=True)
    obj._scan_device(qd[2:3], rows, T_, k, nosync=True)
    with pytest.raises(_lib.PshadowError):
        obj._check_pipeline()
    # reported once; the next pipeline starts clean, and the synchronous call repairs by itself
    obj._scan_device(qd[1:2], rows, T_, k, nosync=True)
    obj._check_pipeline()
    d, idx = obj._scan_device(qd[:1] * 1e-3, rows, T_, k)
    do, io = oracle.shadow_topk(ds, q[:1] * np.float32(1e-3), k, H)
    assert_topk_equal(d.cpu().numpy(), idx.cpu().numpy(), do, io)
    obj._scan_device(qd[1:2], rows, T_, k, nosync=True)
    obj._check_pipeline()


@pytest.mark.parametrize("force_sort", ["0", "1"])
def test_merge_ties_and_padding(force_sort, monkeypatch):
    """Merge of shards that tie in distance (identical rows live on different shards) and of a short
    shard padded with +inf records: order must be (distance, global flat index), padding last."""
    monkeypatch.setenv("PSH_MERGE_SORT", force_sort)
    base, q = make_inputs(1, 600, 40, 2, seed=3)
    ds = np.repeat(base, 12, axis=0)
    k, H, Tp = 200, 10, 600 - 40 - 10 + 1
    rows = torch.tensor(ds[:, 0, :]).cuda()
    qd = torch.tensor(q[:, 0, :]).cuda()
    recs = []
    for s in range(0, 12, 4):
        d, i, _ = _lib.scan_topk(rows[s:s + 4].contiguous(), 600, qd, H, k, s)
        recs.append(torch.cat([d.view(torch.int32).unsqueeze(-1), i], dim=-1))
    pad = torch.empty_like(recs[0])
    pad[..., 0] = 0x7F800000
    pad[..., 1] = 2 ** 31 - 1
    pad[..., 2] = 0
    d, i = _lib.merge_topk_packed(torch.stack(recs + [pad, pad]), Tp)
    do, io = oracle.shadow_topk(ds, q, k, H)
    assert np.array_equal(d.cpu().numpy(), do) and np.array_equal(i.cpu().numpy(), io)


# ---------------------------------------------------------------------------------------------
# embedded scans: Foveal and generic PathEmbedding(kernel)  
# ---------------------------------------------------------------------------------------------
def _gather_ref(ds, idx, L):
    ds = np.asarray(ds).reshape(ds.shape[0], -1)
    return np.stack([np.stack([ds[r, t:t + L] for r, t in row]) for row in idx])[:, :, None, :]


@pytest.mark.parametrize("mode", ["exact", "fft"])
@pytest.mark.parametrize("name", ["foveal_R32_T4096_W126", "dense_kernel_R16_T300_W16"])
def test_embedded_scan_matches_reference_fixture(name, mode):
    """Live-reference fixtures (testing.ipynb:62-78's Foveal configuration; a dense random kernel):
    distances within 1e-6 relative, indices equal up to near-ties, paths = the indexed windows."""
    from conftest import assert_topk_close
    g = load_golden(name)
    if "foveal" in g:
        a, b, w = g["foveal"]
        emb = sb.Foveal(float(a), float(b), int(w))
        assert torch.equal(emb.kernel, torch.tensor(g["kernel"]))
    else:
        emb = sb.PathEmbedding(torch.tensor(g["kernel"]))
    obj = sb.PathShadowing(emb, sb.RelativeMSE(), g["dataset"], sb.PredictionContext(g["H"]), scan_mode=mode)
    d, paths, idx = obj.shadow(g["x_context"], k=g["k"], n_splits=g["n_splits"], cuda=True)
    assert d.dtype == np.float32 and idx.dtype == np.int32 and paths.shape == (g["B"], g["k"], 1, g["W"] + g["H"])
    assert_topk_close(d, idx, g["distances"], g["indices"])
    assert np.array_equal(paths, _gather_ref(g["dataset"], idx, g["W"] + g["H"]))
    if "paths" in g and np.array_equal(idx, g["indices"]):
        assert np.array_equal(paths, g["paths"])


@pytest.mark.parametrize("mode", ["exact", "fft"])
@pytest.mark.parametrize("R,T,W,H,k,B,alpha,beta", [
    (512, 4096, 126, 252, 1024, 5, 1.15, 0.9),   # the reference benchmark's embedding, 3 + 2 query groups
    (700, 1000, 64, 0, 300, 1, 1.3, 0.5),        # no horizon, one query
    (33, 5000, 252, 20, 2000, 4, 1.15, 0.9),     # long rows (fft: two overlapping pieces per row), k > SEG
    (1024, 4096, 126, 252, 1024, 2, 1.15, 0.9),  # 512 row pairs: the fft flavour runs its seedless schedule
])
def test_foveal_matches_oracle(R, T, W, H, k, B, alpha, beta, mode):
    from conftest import assert_topk_close
    ds, q = make_inputs(R, T, W, B, seed=900 + R)
    emb = sb.Foveal(alpha, beta, W)
    obj = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(H or None), scan_mode=mode)
    n0 = _lib.launch_count()
    d, paths, idx = obj.shadow(q, k=k)
    ex = emb(torch.tensor(q))[:, 0, :].numpy()
    do, io = oracle.embed_topk(ds, emb.kernel.numpy()[:, 0, :], ex, k, H)
    assert_topk_close(d, idx, do, io)
    assert np.array_equal(paths, _gather_ref(ds, idx, W + H))
    assert (np.diff(d, axis=1) >= 0).all()
    if mode == "fft" and R == 1024:
        n1 = _lib.launch_count()
        obj.shadow(q, k=k)   # second call: aux is cached; qfft, scan (seeds itself), rerank, select + gather
        assert _lib.launch_count() - n1 == 5, _lib.launch_count() - n1


def test_foveal_fft_flavour_equals_exact_flavour_and_recovers():
    """Both embedded flavours return the same windows; a query far off the data's scale leaves the
    seed threshold at +inf, the candidate buffer overflows and the safe schedule still answers."""
    from conftest import assert_topk_close
    ds, q = make_inputs(1024, 2048, 100, 3, seed=71)
    emb = sb.Foveal(1.2, 0.7, 100)
    outs = {}
    for mode in ("exact", "fft"):
        obj = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(30), scan_mode=mode)
        outs[mode] = obj.shadow(q, k=500)
    assert_topk_close(outs["fft"][0], outs["fft"][2], outs["exact"][0], outs["exact"][2])
    qs = q * np.float32(1e-3)
    d, _, idx = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(30), scan_mode="fft").shadow(qs, k=500)
    ex = emb(torch.tensor(qs))[:, 0, :].numpy()
    do, io = oracle.embed_topk(ds, emb.kernel.numpy()[:, 0, :], ex, 500, 30)
    assert_topk_close(d, idx, do, io)


def test_generic_kernels_through_the_embedded_scan():
    """Identity as a generic kernel (W one-tap runs) agrees with the exact Identity scan; a
    piecewise-constant kernel with several runs per row and an all-zero row agrees with the oracle."""
    from conftest import assert_topk_close
    ds, q = make_inputs(40, 900, 24, 3, seed=61)
    d0, _, i0 = _obj(ds, 24, 6).shadow(q, k=200)
    gen = sb.PathEmbedding(torch.eye(24)[:, None, :].clone())
    d1, _, i1 = sb.PathShadowing(gen, sb.RelativeMSE(), ds, sb.PredictionContext(6)).shadow(q, k=200)
    assert_topk_close(d1, i1, d0, i0)
    K = torch.zeros(4, 1, 24)
    K[0, 0, 2:9] = 0.5; K[0, 0, 9:15] = -1.25; K[0, 0, 20:] = 2.0
    K[2, 0, :] = 0.1
    K[3, 0, ::2] = 1.0
    emb = sb.PathEmbedding(K)
    d2, _, i2 = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(6)).shadow(q, k=200)
    ex = emb(torch.tensor(q))[:, 0, :].numpy()
    do, io = oracle.embed_topk(ds, K.numpy()[:, 0, :], ex, 200, 6)
    assert_topk_close(d2, i2, do, io)


def test_testing_ipynb_foveal_self_consistency():
    """testing.ipynb:62-78 'shadowing correctly selected paths': the returned paths, re-embedded and
    re-distanced with the plugins' own forward, reproduce the returned distances (rtol 1e-2 there)."""
    g = torch.Generator().manual_seed(5)
    x_context = torch.randn(8, 1, 126, generator=g)
    x_dataset = torch.randn(32, 1, 4096, generator=g)
    embedding = sb.Foveal(1.15, 0.9, 126)
    context = sb.PredictionContext(252)
    obj = sb.PathShadowing(embedding, sb.RelativeMSE(), x_dataset, context)
    distances, paths, _ = obj.shadow(x_context, k=1024, cuda=True)
    x_emb = embedding(x_context)[:, 0, :]
    p_in = context.select_in_context(torch.tensor(paths))
    p_emb = torch.stack([embedding(p_in[b])[:, 0, :] for b in range(8)])
    d_check = sb.RelativeMSE()(x_emb[:, None, :], p_emb).numpy()
    assert np.allclose(d_check, distances, rtol=1e-5)


def test_predict_with_foveal_end_to_end():
    ds, q = make_inputs(300, 1500, 64, 6, seed=15)
    emb = sb.Foveal(1.2, 0.8, 64)
    obj = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(20))
    rv = sb.RealizedVariance([5, 10, 20])
    pred, std = obj.predict(q, k=256, to_predict=rv, eta=0.1, n_context_splits=2)
    d, paths, _ = obj.shadow(q, k=256)
    mo, so = oracle.predict_from_paths(d, paths, 20, [5, 10, 20], False, "softmax", 0.1)
    assert np.allclose(pred, mo, rtol=1e-6) and np.allclose(std, so, rtol=1e-5)


@pytest.mark.parametrize("mode", ["exact", "fft"])
def test_imputation_context_matches_reference_fixture(mode):
    """ImputationContext((l, c, r)) (path_embedding.py:59-87): the windows are compared on their first l and
    last r samples, the c samples in between come back as out-context.  Identity and Foveal embeddings
    against live-reference fixtures (the kernel padded with zero taps runs through the embedded scan:
    distances within 1e-6, same windows up to near-ties)."""
    from conftest import GOLDEN, assert_topk_close
    g = np.load(GOLDEN / "imputation_R24_T700.npz")
    l, c, r = (int(v) for v in g["portion"])
    ds, x = g["dataset"], g["x_context"]
    obj = sb.PathShadowing(sb.Identity(l + r), sb.RelativeMSE(), ds, sb.ImputationContext((l, c, r)), scan_mode=mode)
    d, paths, idx = obj.shadow(x, k=40)
    assert paths.shape == g["paths"].shape == (3, 40, 1, l + c + r) and idx.dtype == np.int32
    assert_topk_close(d, idx, g["distances"], g["indices"])
    assert np.array_equal(paths, _gather_ref(ds, idx, l + c + r))
    out = obj.context.select_out_context(paths)
    assert out.shape[-1] == c and np.array_equal(out, paths[..., l:l + c])
    a, b, _ = (float(v) for v in g["foveal"])
    fobj = sb.PathShadowing(sb.Foveal(a, b, l + r), sb.RelativeMSE(), ds, sb.ImputationContext((l, c, r)), scan_mode=mode)
    fd, _, fidx = fobj.shadow(x, k=40)
    assert_topk_close(fd, fidx, g["foveal_distances"], g["foveal_indices"])
    # portion=None: the whole window is in-context (== PredictionContext(None))
    d0, _, i0 = sb.PathShadowing(sb.Identity(l + r), sb.RelativeMSE(), ds, sb.ImputationContext(None)).shadow(x, k=40)
    d1, _, i1 = sb.PathShadowing(sb.Identity(l + r), sb.RelativeMSE(), ds, sb.PredictionContext(None)).shadow(x, k=40)
    assert np.array_equal(d0, d1) and np.array_equal(i0, i1)


@pytest.mark.parametrize("mode", ["exact", "filter", "fft"])
def test_cross_channel_context_bit_exact(mode):
    """CrossChannelContext(2) (path_embedding.py:90-114): a (R, 3, T) dataset is scanned on its in-context
    channel, the shadowing paths come back with all three channels -- bit-exact against the live-reference
    fixture (Identity windows are exact)."""
    from conftest import GOLDEN
    g = np.load(GOLDEN / "crosschannel_R16_C3_T500.npz")
    obj = sb.PathShadowing(sb.Identity(25), sb.RelativeMSE(), g["dataset"], sb.CrossChannelContext(2), scan_mode=mode)
    d, paths, idx = obj.shadow(g["x_context"], k=30)
    assert_topk_equal(d, idx, g["distances"], g["indices"])
    assert np.array_equal(idx, g["indices"])
    assert paths.shape == (2, 30, 3, 25) and np.array_equal(paths, g["paths"])
    assert np.array_equal(obj.context.select_out_context(paths), paths[:, :, 1:, :])
    with pytest.raises(RuntimeError):   # a single-channel dataset does not fit a 1 + 2 channel context
        sb.PathShadowing(sb.Identity(25), sb.RelativeMSE(), g["dataset"][:, :1], sb.CrossChannelContext(2)).shadow(
            g["x_context"], k=3)


def test_streamed_dataset_equals_host_dataset(tmp_path):
    """`stream_dataset=True`: `.npy` batch files (scripts/batch_generations.py:28-40) go through two pinned
    staging buffers straight into the resident rows; results equal those of the host-array path."""
    ds, q = make_inputs(48, 1030, 64, 2, seed=61)
    for i in range(3):
        np.save(tmp_path / f"batch{i + 1:04}.npy", ds[16 * i:16 * (i + 1)])
    tsd = sb.TimeSeriesDataset(tmp_path)
    rows, T, C = tsd.to_device("cuda", chunk_bytes=5 * 1030 * 4)    # chunks that do not divide a file
    assert (T, C) == (1030, 1) and rows.shape == (48, 1032)
    assert np.array_equal(rows[:, :1030].cpu().numpy(), ds[:, 0, :])
    a = sb.PathShadowing(sb.Identity(64), sb.RelativeMSE(), tsd, sb.PredictionContext(5), stream_dataset=True)
    b = _obj(ds, 64, 5)
    for x, y in zip(a.shadow(q, k=70), b.shadow(q, k=70)):
        assert np.array_equal(x, y)


def test_unsupported_plugins_raise():
    class Cosine(sb.PathDistance):
        def forward(self, x, y):
            return 1 - (x * y).sum(-1)
    ds, q = make_inputs(4, 100, 8, 1)
    with pytest.raises(NotImplementedError):
        sb.PathShadowing(sb.Identity(8), Cosine(), ds, sb.PredictionContext(2)).shadow(q, k=3)


@pytest.mark.parametrize("streams", [2, 3])
def test_two_stream_pipeline_matches_single_stream(streams):
    """`_pipe_streams = 2 | 3`: consecutive enqueue-only scans alternate between side streams with
    their own workspaces (the scans carry PSH_SHARE_SMS: their persistent kernel leaves a few SMs to the
    neighbours' small kernels); results are those of the one-stream pipeline, overflow is still caught."""
    R, T, W, H, k = 2048, 2048, 64, 4, 128
    ds, q = make_inputs(R, T, W, 6, seed=93)
    obj = _obj(ds, W, H, scan_mode="fft")
    rows, T_ = obj._resident_rows()
    qd = torch.tensor(q).cuda()
    do, io = oracle.shadow_topk(ds, q, k, H)
    obj._pipe_streams = streams
    outs = [obj._scan_device(qd[i:i + 1], rows, T_, k, nosync=True) for i in range(6)]
    obj._check_pipeline()
    for i, (d, idx) in enumerate(outs):
        assert_topk_equal(d.cpu().numpy(), idx.cpu().numpy(), do[i:i + 1], io[i:i + 1])
    obj._scan_device(qd[:1] * 1e-3, rows, T_, k, nosync=True)     # overflows (seed threshold +inf)
    obj._scan_device(qd[1:2], rows, T_, k, nosync=True)
    with pytest.raises(_lib.PshadowError):
        obj._check_pipeline()
    obj._scan_device(qd[1:2], rows, T_, k, nosync=True)
    obj._check_pipeline()
    d, _, idx = obj.shadow(q[:2], k=k)                              # shadow() keeps to one stream
    assert_topk_equal(d, idx, do[:2], io[:2])
