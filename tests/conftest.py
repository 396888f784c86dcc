import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
SHADOW_CASES = ["cfg1_R128_T512_W20", "w252_R96_T2048", "nohorizon_R64_T1024",
                "ragged_R37_T513_W21", "single_T4096_W64", "k_all_R3_T40_W8"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    g = {k: z[k] for k in z.files}
    if "meta" in g:
        R, T, W, H, k, B, ns = (int(v) for v in g["meta"])
        g.update(R=R, T=T, W=W, H=None if H < 0 else H, k=k, B=B, n_splits=ns)
    return g


@pytest.fixture(params=SHADOW_CASES)
def golden(request):
    return load_golden(request.param)


def make_inputs(R, T, W, B, seed=0, scale=0.01):
    """Seeded synthetic inputs of SURVEY.md section 8(d): identical bits for oracle and GPU."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ds = (torch.randn(R, 1, T, generator=g, dtype=torch.float32) * scale).numpy()
    g = torch.Generator().manual_seed(seed + 1)
    q = (torch.randn(B, 1, W, generator=g, dtype=torch.float32) * scale).numpy()
    return ds, q


def assert_topk_equal(d, idx, d_ref, idx_ref):
    """Bit-exact distances; indices equal up to permutation inside groups of tied distances
    (the reference's tie order is unspecified, SURVEY.md section 7 'Ties')."""
    d = np.asarray(d); d_ref = np.asarray(d_ref)
    assert d.dtype == np.float32 and d.shape == d_ref.shape
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32)), "distances differ"
    idx = np.asarray(idx); idx_ref = np.asarray(idx_ref)
    assert idx.shape == idx_ref.shape
    if np.array_equal(idx, idx_ref):
        return
    for b in range(d.shape[0]):
        k = d.shape[1]
        i = 0
        while i < k:
            j = i
            while j + 1 < k and d[b, j + 1] == d[b, i]:
                j += 1
            a = {tuple(v) for v in idx[b, i:j + 1]}
            r = {tuple(v) for v in idx_ref[b, i:j + 1]}
            if j + 1 < k or i == j:
                assert a == r, f"indices differ at query {b}, ranks {i}..{j}"
            # a tie group cut by the k boundary may legitimately hold different members
            i = j + 1


def assert_topk_close(d, idx, d_ref, idx_ref, rtol=1e-6):
    """Tolerance parity for embedded scans (Foveal / PathEmbedding(kernel)): the reference's conv1d
    accumulates in an order that cannot be replayed, so the embedded values e agree to ~1e-7
    relative and the distance d = ||ex - e|| / ||ex|| to `rtol` (north_star: 1e-6) of the embedded
    scale, |d - d_ref| <= rtol (1 + d_ref) (a distance much smaller than 1 is a cancellation and
    carries the error of e, not of d).  Indices agree up to (i) swaps among windows whose distances
    lie within the tolerance of each other and (ii) exchanges at the k-th boundary."""
    d = np.asarray(d); d_ref = np.asarray(d_ref); idx = np.asarray(idx); idx_ref = np.asarray(idx_ref)
    assert d.shape == d_ref.shape and idx.shape == idx_ref.shape
    tol = rtol * (1.0 + np.abs(d_ref.astype(np.float64)))
    assert (np.abs(d.astype(np.float64) - d_ref) <= tol).all(), float(np.max(np.abs(d - d_ref) / tol)) * rtol
    for b in range(d.shape[0]):
        ref_pos = {tuple(v): j for j, v in enumerate(idx_ref[b])}
        for j, v in enumerate(idx[b]):
            jr = ref_pos.get(tuple(v))
            if jr is None:   # only a window as far as the boundary may be exchanged for another one
                assert d[b, j] >= d_ref[b, -1] - 4 * tol[b, -1], (b, j)
            else:            # same window: same distance up to the tolerance, wherever it was ranked
                assert abs(float(d[b, j]) - float(d_ref[b, jr])) <= 2 * tol[b, jr], (b, j, jr)
