"""CPU: the oracle (C and numpy restatements) against the live-reference golden fixtures."""
import numpy as np
import pytest

from conftest import assert_topk_equal, load_golden, make_inputs
from oracle import oracle, ref_loader


def test_oracle_builds_and_loads():
    oracle.build()
    assert oracle.num_threads() >= 1


def test_qnorm_matches_reference(golden):
    q = golden["x_context"].reshape(golden["B"], golden["W"])
    for b in range(golden["B"]):
        assert oracle.qnorm(q[b]) == golden["qnorm"][b]
        assert oracle.np_qnorm(q[b]) == golden["qnorm"][b]


def test_c_oracle_topk_bit_exact(golden):
    H = golden["H"] or 0
    d, idx = oracle.shadow_topk(golden["dataset"], golden["x_context"], golden["k"], H)
    assert_topk_equal(d, idx, golden["distances"], golden["indices"])
    assert np.array_equal(idx, golden["indices"])  # fixtures are tie-free


def test_numpy_oracle_topk_bit_exact(golden):
    H = golden["H"] or 0
    d, idx = oracle.np_shadow_topk(golden["dataset"], golden["x_context"], golden["k"], H)
    assert_topk_equal(d, idx, golden["distances"], golden["indices"])


def test_oracle_paths_bit_exact(golden):
    H = golden["H"] or 0
    d, paths, idx = oracle.shadow(golden["dataset"], golden["x_context"], golden["k"], H)
    assert paths.shape == golden["paths"].shape
    assert np.array_equal(paths, golden["paths"])


def test_c_and_numpy_distances_agree_everywhere():
    ds, q = make_inputs(9, 300, 33, 2, seed=7)
    for b in range(2):
        a = oracle.distances(ds, q[b], 5)
        n = oracle.np_distances(ds, q[b], 5)
        assert np.array_equal(a.view(np.uint32), n.view(np.uint32))


def test_threads_do_not_change_result():
    ds, q = make_inputs(64, 700, 40, 3, seed=3)
    a = oracle.shadow_topk(ds, q, 200, 10, nthreads=1)
    b = oracle.shadow_topk(ds, q, 200, 10, nthreads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_row_offset_and_shard_merge_equals_global():
    """Merge of per-shard exact top-k == global top-k."""
    ds, q = make_inputs(48, 400, 24, 2, seed=5)
    k, H = 64, 8
    dg, ig = oracle.shadow_topk(ds, q, k, H)
    parts = [oracle.shadow_topk(ds[s:s + 12], q, k, H, row_offset=s) for s in range(0, 48, 12)]
    d_all = np.concatenate([p[0] for p in parts], 1)
    i_all = np.concatenate([p[1] for p in parts], 1)
    Tp = 400 - 24 - 8 + 1
    for b in range(2):
        flat = i_all[b, :, 0].astype(np.int64) * Tp + i_all[b, :, 1]
        order = np.lexsort((flat, d_all[b].view(np.uint32)))[:k]
        assert np.array_equal(d_all[b][order], dg[b])
        assert np.array_equal(i_all[b][order], ig[b])


def test_k_larger_than_windows_raises():
    ds, q = make_inputs(2, 40, 8, 1)
    with pytest.raises(RuntimeError):
        oracle.shadow_topk(ds, q, 2 * 40, 4)


def test_zero_query_gives_inf():
    ds, _ = make_inputs(2, 64, 8, 1)
    d, idx = oracle.shadow_topk(ds, np.zeros((1, 1, 8), np.float32), 4, 0)
    assert np.isinf(d).all()
    assert np.array_equal(idx[0, :, 1], np.arange(4))  # index order among ties


def test_realized_variance_matches_reference(golden):
    H = golden["H"] or 0
    out = golden["paths"][..., -H:] if H else golden["paths"]
    Ts = [int(t) for t in golden["Ts"]]
    assert np.array_equal(oracle.realized_variance(out, Ts, False), golden["rv"])
    assert np.array_equal(oracle.realized_variance(out, Ts, True), golden["rvol"])


def test_softmax_weight_properties():
    rng = np.random.default_rng(0)
    d = np.sort(rng.uniform(0.8, 1.4, (3, 50)).astype(np.float32), 1)
    w = oracle.softmax_weights(d[:, :, None], 0.1, axis=1)
    assert np.allclose(w.sum(1), 1.0)
    assert (np.diff(w[:, :, 0], axis=1) <= 0).all()          # closer paths weigh more
    wu = oracle.softmax_weights(d[:, :, None], 1e6, axis=1)   # eta -> inf == Uniform
    assert np.allclose(wu, 1.0 / 50, rtol=1e-9)
    wa = oracle.softmax_weights(d[:, :, None], 1e-3, axis=1)  # eta -> 0 == argmin
    assert np.allclose(wa[:, 0, 0], 1.0)
    # the second call-site broadcast pattern: (k,) weights against (k,1,T) over axis 0 (plot_utils.py:74-76)
    w0 = oracle.softmax_weights(d[0], 0.1, axis=0)
    assert np.allclose(w0, w[0, :, 0])


@pytest.mark.skipif(not ref_loader.available(), reason="live reference only exists in the build container")
def test_live_reference_matches_oracle_on_fresh_inputs():
    import torch
    ref = ref_loader.load()
    ds, q = make_inputs(40, 600, 50, 2, seed=11)
    obj = ref.path_shadowing.PathShadowing(ref.path_embedding.Identity(50), ref.path_distance.RelativeMSE(),
                                           torch.tensor(ds), ref.path_embedding.PredictionContext(12))
    d, paths, idx = obj.shadow(torch.tensor(q), k=300, n_splits=5, cuda=False)
    do, po, io = oracle.shadow(ds, q, 300, 12)
    assert_topk_equal(do, io, d, idx)
    assert np.array_equal(po, paths)


# ---------------------------------------------------------------------------------------------
# embedded scans (Foveal / PathEmbedding(kernel)): the oracle's fp64-dot-product definition
# against live-reference fixtures, within north_star's 1e-6 relative tolerance
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["foveal_R32_T4096_W126", "dense_kernel_R16_T300_W16"])
def test_embed_oracle_matches_reference_fixture(name):
    from conftest import assert_topk_close
    g = load_golden(name)
    K = g["kernel"][:, 0, :]
    if "foveal" in g:
        a, b, w = g["foveal"]
        assert np.array_equal(oracle.foveal_kernel(float(a), float(b), int(w)), K)
    ex = oracle.embed_queries(K, g["x_context"])
    assert np.abs(ex - g["ex"]).max() <= 1e-6 * np.abs(g["ex"]).max()   # fp32 conv vs fp64 dot product
    d, idx = oracle.embed_topk(g["dataset"], K, g["ex"], g["k"], g["H"])
    assert_topk_close(d, idx, g["distances"], g["indices"])
    assert (np.diff(d, axis=1) >= 0).all()


def test_embed_oracle_identity_kernel_equals_identity_scan():
    """PathEmbedding(eye(W)) through the embedded oracle == the Identity oracle, bit for bit
    (an fp64 'dot product' with one unit tap is the sample itself)."""
    from conftest import make_inputs
    ds, q = make_inputs(9, 200, 12, 2, seed=4)
    d1, i1 = oracle.shadow_topk(ds, q, 40, 3)
    d2, i2 = oracle.embed_topk(ds, np.eye(12, dtype=np.float32), q[:, 0, :], 40, 3)
    assert np.array_equal(d1, d2) and np.array_equal(i1, i2)
