"""CPU: the C-ABI library loads and exports every symbol include/pshadow.h declares; argument
validation that needs no device; host-side logic of the Python plugin surface."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import load_golden

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "pshadow.h").read_text()
    names = set(re.findall(r"\b(psh_[a-z0-9_]+)\s*\(", hdr))
    assert {"psh_scan_topk_f32", "psh_merge_topk", "psh_merge_topk_packed", "psh_gather_paths", "psh_rv_aggregate",
            "psh_scan_workspace_bytes", "psh_scan_overflowed", "psh_fft_aux_bytes", "psh_fft_prepare",
            "psh_debug_fft4096", "psh_profile_begin", "psh_profile_end",
            "psh_version", "psh_error_string", "psh_launch_count"} <= names
    so = ROOT / "shadowing_b200" / "libpshadow.so"
    assert so.exists(), "build with __graft_entry__.build()"
    L = ctypes.CDLL(str(so))
    for n in names:
        assert hasattr(L, n), n


def test_version_errors_and_workspace_sizing():
    from shadowing_b200 import _lib
    L = _lib.lib()
    assert L.psh_version() == 100
    assert b"k exceeds" in L.psh_error_string(-2)
    assert L.psh_scan_workspace_bytes(32768, 4096, 1, 252, 20, 1024) > 0
    assert L.psh_scan_workspace_bytes(32768, 4096, 256, 252, 20, 1024) > L.psh_scan_workspace_bytes(32768, 4096, 1, 252, 20, 1024)
    assert L.psh_scan_workspace_bytes(10, 100, 1, 90, 20, 4) == 0   # W + H > T
    assert L.psh_scan_workspace_bytes(0, 100, 1, 10, 0, 4) == 0
    # null pointers are rejected before any device work
    assert L.psh_scan_topk_f32(None, 1, 100, 100, None, 1, 10, 0, 1, 0, 0, None, None, None, 0, None, 0, None) == -1
    assert L.psh_fft_aux_bytes(32768, 4096, 252, 20) > 2 * 32768 * 4096 * 4 // 2
    assert L.psh_fft_aux_bytes(8, 8192, 252, 20) > 3 * 8 * 4096 * 4   # T > 4096: three overlapping pieces per row
    assert L.psh_fft_aux_bytes(8, 8192, 3000, 20) == 0                # context longer than half a transform
    assert L.psh_gather_paths(None, 1, 1, 1, None, 1, 0, 1, None, None) == -1
    assert L.psh_merge_topk(None, None, 1, 1, 1, 1, None, None, None) == -1
    assert L.psh_scan_overflowed(None, 1, None) == -1
    assert b"overflow" in L.psh_error_string(-6)


def test_plugin_surface_matches_reference_names():
    import shadowing_b200 as sb
    for n in ["PathShadowing", "PathEmbedding", "Identity", "PathDistance", "RelativeMSE", "ContextManagerBase",
              "PredictionContext", "ArrayType", "realized_variance", "Softmax", "Uniform", "DiscreteProba"]:
        assert hasattr(sb, n)
    import inspect
    sig = inspect.signature(sb.PathShadowing.shadow)
    assert list(sig.parameters)[1:] == ["x_context", "k", "n_splits", "cuda"]
    sig = inspect.signature(sb.PathShadowing.predict)
    assert list(sig.parameters)[1:] == ["x_context", "k", "to_predict", "eta", "proba_name", "n_dataset_splits",
                                        "n_context_splits", "cuda"]
    assert sig.parameters["proba_name"].default == "softmax"
    sig = inspect.signature(sb.PathShadowing.__init__)
    assert list(sig.parameters)[1:5] == ["embedding", "distance", "dataset", "context"]


def test_prediction_context_and_identity():
    import shadowing_b200 as sb
    c = sb.PredictionContext(5)
    x = np.arange(20.0).reshape(2, 10)
    assert np.array_equal(c.select_in_context(x), x[:, :5]) and np.array_equal(c.select_out_context(x), x[:, 5:])
    assert c.get_out_times() == 5 and sb.PredictionContext().get_out_times() == 0
    assert sb.PredictionContext().select_out_context(x) is x
    e = sb.Identity(4)
    assert e.d == 4 and tuple(e.kernel.shape) == (4, 1, 4)
    adj = e.adjust_to_context(c)
    assert tuple(adj.kernel.shape) == (4, 1, 9) and float(adj.kernel[:, :, 4:].abs().sum()) == 0.0
    y = torch.arange(12.0).reshape(1, 1, 12)
    emb = adj(y)  # (1, t', 4): exact sliding windows
    assert tuple(emb.shape) == (1, 4, 4) and torch.equal(emb[0, 2], y[0, 0, 2:6])


def test_imputation_and_cross_channel_contexts():
    """ImputationContext / CrossChannelContext mirror path_embedding.py:59-114 (checked against what the live
    reference returned for the committed fixtures)."""
    import numpy as np
    import torch
    import shadowing_b200 as sb
    from conftest import GOLDEN
    g = np.load(GOLDEN / "imputation_R24_T700.npz")
    l, c, r = (int(v) for v in g["portion"])
    ctx = sb.ImputationContext((l, c, r))
    assert ctx.get_out_times() == c and sb.ImputationContext(None).get_out_times() == 0
    k = torch.arange(float(2 * (l + r))).reshape(2, 1, l + r)
    kp = ctx.pad_context(k)
    assert kp.shape == (2, 1, l + c + r) and torch.equal(kp[..., :l], k[..., :l]) and torch.equal(kp[..., -r:], k[..., -r:])
    assert float(kp[..., l:l + c].abs().sum()) == 0.0
    paths = g["paths"]
    assert np.array_equal(ctx.select_in_context(paths), np.concatenate([paths[..., :l], paths[..., -r:]], -1))
    assert np.array_equal(ctx.select_out_context(paths), paths[..., l:-r])
    assert np.array_equal(ctx.slect_out_context(paths), paths[..., l:-r])   # the reference's spelling
    # the fixture's distances are the relative error on the in-context samples of the returned paths
    x = g["x_context"]
    err = np.linalg.norm(ctx.select_in_context(paths)[:, :, 0, :] - x, axis=-1) / np.linalg.norm(x, axis=-1)
    assert np.allclose(err, g["distances"], rtol=1e-5)
    h = np.load(GOLDEN / "crosschannel_R16_C3_T500.npz")
    cc = sb.CrossChannelContext(2)
    assert cc.get_out_times() == 0
    assert cc.pad_context(torch.ones(4, 1, 25)).shape == (4, 3, 25)
    assert float(cc.pad_context(torch.ones(4, 1, 25))[:, 1:].abs().sum()) == 0.0
    assert np.array_equal(cc.select_in_context(h["paths"]), h["paths"][:, :, :1, :])
    assert np.array_equal(cc.select_out_context(h["paths"]), h["paths"][:, :, 1:, :])


def test_relative_mse_and_forward_topk_match_reference_fixture():
    import shadowing_b200 as sb
    g = load_golden("forward_topk_B8_d34")
    dist = sb.RelativeMSE()
    x, y = torch.tensor(g["x"]), torch.tensor(g["y"])
    ds1, id1 = dist.forward_topk(x, y, k=32, n_splits=4)
    ds2, id2 = dist.forward_topk(x, y, k=64, n_splits=8)
    assert torch.equal(ds1, ds2[:, :32]) and torch.equal(id1, id2[:, :32])   # testing.ipynb:43-53
    assert np.array_equal(ds2.numpy(), g["ds"]) and np.array_equal(id2.numpy(), g["idces"])


def test_realized_variance_matches_reference_fixture():
    import shadowing_b200 as sb
    g = load_golden("cfg1_R128_T512_W20")
    out = g["paths"][..., -20:]
    Ts = [int(t) for t in g["Ts"]]
    assert np.array_equal(sb.realized_variance(out, Ts, False), g["rv"])
    assert np.array_equal(sb.realized_variance(out, Ts, True), g["rvol"])
    assert np.array_equal(sb.RealizedVariance(Ts)(out), g["rv"][:, :, 0, :])


def test_softmax_uniform_contracts():
    import shadowing_b200 as sb
    from oracle import oracle
    rng = np.random.default_rng(1)
    d = np.sort(rng.uniform(0.9, 1.3, (4, 64)).astype(np.float32), 1)
    x = rng.normal(size=(4, 64, 3)).astype(np.float32)
    p = sb.PathShadowing.init_averaging_proba("softmax", d[:, :, None], 0.1)
    w = oracle.softmax_weights(d[:, :, None], 0.1, axis=1)
    assert np.allclose(p.avg(x, axis=1), (w * x).sum(1), rtol=1e-5, atol=1e-7)
    var = (w * x * x).sum(1) - (w * x).sum(1) ** 2
    assert np.allclose(p.std(x, axis=1), np.sqrt(var), rtol=1e-5)
    u = sb.PathShadowing.init_averaging_proba("uniform", d[:, :, None], None)
    assert np.allclose(u.avg(x, axis=1), x.mean(1, dtype=np.float64), rtol=1e-5, atol=1e-7)
    assert np.allclose(u.std(x, axis=1), x.std(1, dtype=np.float64), rtol=1e-5)
    with pytest.raises(ValueError):
        sb.PathShadowing.init_averaging_proba("nope", d, 0.1)
    # plot_utils.py:74-76 pattern: (k,) distances against (k, 1, T) paths over axis 0
    paths = rng.normal(size=(64, 1, 30)).astype(np.float32)
    s = sb.Softmax(distances=d[0], eta=0.09)
    w0 = oracle.softmax_weights(d[0], 0.09, axis=0)
    assert np.allclose(s.avg(paths, axis=0), (w0[:, None, None] * paths).sum(0), rtol=1e-5, atol=1e-7)
    assert s.avg(paths, axis=0)[0, :].shape == (30,)


def test_select_cartesian_product():
    import shadowing_b200 as sb
    a, b = torch.tensor([10, 20, 30]), torch.arange(4)
    idx = torch.tensor([[0, 5, 11]])
    out = sb.select_cartesian_product(idx, [a, b])
    assert torch.equal(out, torch.cartesian_prod(a, b)[idx])


def test_no_cpu_fallback_without_device():
    import shadowing_b200 as sb
    if torch.cuda.is_available():
        pytest.skip("device present")
    obj = sb.PathShadowing(sb.Identity(8), sb.RelativeMSE(), np.zeros((2, 1, 64), np.float32), sb.PredictionContext(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        obj.shadow(np.ones((1, 1, 8), np.float32), k=2)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under shadowing_b200/ may import, load or call it."""
    pkg = ROOT / "shadowing_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("Makefile")):
        text = f.read_text()
        assert "oracle" not in text.lower(), f"{f} mentions the oracle"
        assert "/root/reference" not in text, f"{f} reads the reference tree"


def test_new_entry_points_validate_arguments():
    from shadowing_b200 import _lib
    L = _lib.lib()
    # embedded scan: null run table / non-positive sizes / unknown flags are rejected before device work
    assert L.psh_scan_topk_embed_f32(None, 1, 100, 100, None, 1, 4, 10, 0, 1, 0, 0, None, 0, None, None, 0,
                                     None, None, None, 0, None) == -1
    assert L.psh_fft_prepare_embed(None, 1, 100, 100, 10, 0, None, 0, None, 0, None) == -1
    # exchange buffers: sizes scale with ranks, queries and k; more than 16 peers is not supported
    a = L.psh_xchg_bytes(2, 1, 1024)
    assert a >= 2 * (2 * 1024 * 12 + 2 * 4) and L.psh_xchg_bytes(8, 4, 1024) > 4 * a
    assert L.psh_xchg_bytes(17, 1, 1024) == 0 and L.psh_xchg_bytes(0, 1, 1) == 0
    assert L.psh_allgather_merge_packed(None, None, 2, 0, 1, 16, 100, 1, None, None, None, None) == -1
    assert L.psh_xchg_open(None, None) == -1 and L.psh_xchg_close(None) == -1 and L.psh_xchg_destroy(None) == -1


def test_foveal_matches_reference_fixture_and_kernel_runs():
    """Foveal's kernel equals the live reference's (fixture) bit for bit; kernel_runs turns any
    (d,1,W) kernel into runs that reconstruct it exactly."""
    import shadowing_b200 as sb
    from shadowing_b200.path_embedding import kernel_runs
    g = load_golden("foveal_R32_T4096_W126")
    a, b, w = g["foveal"]
    f = sb.Foveal(float(a), float(b), int(w))
    assert f.dim == 34 and f.alpha == float(a) and f.beta == float(b) and f.max_context == 126
    assert torch.equal(f.kernel, torch.tensor(g["kernel"]))
    assert [s.start for s in f.slices][:6] == [-1, -1, -1, -1, -2, -2] and f.slices[-1].start == -115
    assert np.array_equal(f(torch.tensor(g["x_context"]))[:, 0, :].numpy(), g["ex"])   # same conv1d as the reference
    runs = kernel_runs(f.kernel)
    assert len(runs) == 34 and (runs["b"] == 126).all() and (np.diff(runs["row"]) == 1).all()
    assert tuple(f.adjust_to_context(sb.PredictionContext(252)).kernel.shape) == (34, 1, 378)
    rng = np.random.default_rng(0)
    for K in (torch.eye(7)[:, None, :], torch.tensor(rng.integers(-2, 3, size=(5, 1, 40)).astype(np.float32)),
              torch.tensor(g["kernel"]), torch.zeros(2, 1, 9)):
        runs = kernel_runs(K)
        rebuilt = np.zeros(tuple(K.shape), np.float32)
        for r in runs:
            assert 0 <= r["a"] < r["b"] <= K.shape[-1] and r["c"] != 0
            assert (rebuilt[r["row"], 0, r["a"]:r["b"]] == 0).all()
            rebuilt[r["row"], 0, r["a"]:r["b"]] = r["c"]
        assert np.array_equal(rebuilt, K.numpy()) and (np.diff(runs["row"]) >= 0).all()
    with pytest.raises(RuntimeError):
        kernel_runs(torch.zeros(3, 2, 5))


def test_time_series_dataset_reads_npy_batches(tmp_path):
    """The on-disk format of scripts/batch_generations.py:28-40 (batchNNNN.npy) and of single
    trajectory files, in file-name order, with the README's `R=` limit (README.md:41-42)."""
    import shadowing_b200 as sb
    rng = np.random.default_rng(1)
    parts = [rng.standard_normal((5, 1, 64)).astype(np.float32), rng.standard_normal((3, 64)).astype(np.float32),
             rng.standard_normal(64)]
    for i, a in enumerate(parts):
        np.save(tmp_path / f"batch{i + 1:04}.npy", a)
    full = np.concatenate([parts[0], parts[1][:, None, :], parts[2][None, None, :].astype(np.float32)])
    ds = sb.TimeSeriesDataset(tmp_path)
    assert len(ds) == 9 and np.array_equal(ds.load(), full) and ds.load().dtype == np.float32
    assert np.array_equal(sb.TimeSeriesDataset(tmp_path, R=6).load(), full[:6])
    with pytest.raises(ValueError):
        sb.TimeSeriesDataset(tmp_path, R=10).load()
    with pytest.raises(FileNotFoundError):
        sb.TimeSeriesDataset(tmp_path / "nothing").load()
    # PathShadowing accepts the directory or the dataset object, as path_shadowing.py:84-88
    obj = sb.PathShadowing(sb.Identity(8), sb.RelativeMSE(), tmp_path, sb.PredictionContext(2))
    assert np.array_equal(obj.dataset, full)
    obj = sb.PathShadowing(sb.Identity(8), sb.RelativeMSE(), sb.TimeSeriesDataset(tmp_path, R=4))
    assert obj.dataset.shape == (4, 1, 64) and obj.context.get_out_times() == 0
    # the streaming loader (chunks smaller than a file, R limit, row stride padded to 16 bytes)
    for R in (None, 7):
        tsd = sb.TimeSeriesDataset(tmp_path, R=R)
        rows, T, C = tsd.to_device("cpu", chunk_bytes=2 * 64 * 4)
        want = full if R is None else full[:R]
        assert (T, C) == (64, 1) and rows.shape == (want.shape[0], 64)
        assert np.array_equal(rows.numpy(), want[:, 0, :]) and tsd.shape == want.shape
        assert np.array_equal(np.asarray(tsd), want)
    keep = sb.PathShadowing(sb.Identity(8), sb.RelativeMSE(), sb.TimeSeriesDataset(tmp_path), stream_dataset=True)
    assert isinstance(keep.dataset, sb.TimeSeriesDataset) and keep.dataset.shape == (9, 1, 64)


def test_share_sms_flag_encoding():
    """PSH_SHARE_SMS(n) of include/pshadow.h and its Python twin: flag bit 0x200 + n in bits 12..17."""
    from shadowing_b200 import _lib
    hdr = (ROOT / "include" / "pshadow.h").read_text()
    assert "#define PSH_FLAG_SHARE_SMS 0x200" in hdr and "#define PSH_SHARE_SMS(n)" in hdr
    assert _lib.PSH_FLAG_SHARE_SMS == 0x200 and _lib.PSH_FLAG_NOSYNC == 0x100
    for n in (1, 6, 10, 63):
        v = _lib.share_sms(n)
        assert v & _lib.PSH_FLAG_SHARE_SMS and (v >> 12) & 0x3F == n
        assert v & 0xFF == 0          # never collides with the scan mode


def test_fft_aux_sizing_follows_the_transform_length(monkeypatch):
    """psh_fft_aux_bytes (host arithmetic only): 1024-point pieces for W <= 384 -- ceil(T'/hop) pieces per row,
    hop = (1025 - W) & ~3, 8 KiB per pair of pieces -- and 4096-point row pairs beyond; PSH_FFT_N forces one."""
    L = ctypes.CDLL(str(ROOT / "shadowing_b200" / "libpshadow.so"))
    L.psh_fft_aux_bytes.restype = ctypes.c_size_t
    L.psh_fft_aux_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    tables = 4096 * 8 + 4096 * 16 + 512 * 16

    def expect(R, T, W, H, nfft):
        Tp = T - W - H + 1
        if T <= nfft:
            nseg = 1
        else:
            hop = (nfft - W + 1) & ~3
            nseg = -(-Tp // hop)
        npairs = (R * nseg + 1) // 2
        return tables + (npairs * 16 + 255) // 256 * 256 + 2 * 4 * nfft * npairs

    monkeypatch.delenv("PSH_FFT_N", raising=False)
    for (R, T, W, H), nfft in (((32768, 4096, 252, 20), 1024), ((7, 12001, 100, 0), 1024), ((37, 1024, 64, 0), 1024),
                              ((600, 3000, 500, 10), 4096), ((24, 8192, 1024, 20), 4096)):
        assert L.psh_fft_aux_bytes(R, T, W, H) == expect(R, T, W, H, nfft), (R, T, W, H)
    monkeypatch.setenv("PSH_FFT_N", "4096")
    assert L.psh_fft_aux_bytes(32768, 4096, 252, 20) == expect(32768, 4096, 252, 20, 4096)
    monkeypatch.setenv("PSH_FFT_N", "1024")
    assert L.psh_fft_aux_bytes(600, 3000, 500, 10) == expect(600, 3000, 500, 10, 1024)
    assert L.psh_fft_aux_bytes(24, 8192, 1024, 20) == expect(24, 8192, 1024, 20, 4096)   # W > 768: never 1024-point
    assert L.psh_fft_aux_bytes(10, 100, 2049, 0) == 0 and L.psh_fft_aux_bytes(10, 100, 90, 20) == 0
