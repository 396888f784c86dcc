"""CPU: the rigorous bounds of the FFT filter flavours, restated in numpy and checked against fp64.

The CUDA filter (csrc/pshadow.cu `fft_scan_kernel`, csrc/pshadow_embed_fft.cuh) keeps a window iff a
LOWER bound of its squared distance passes the threshold and tightens thresholds from UPPER bounds;
results are exact only if  LB <= S_true <= UB  holds for every window.  These tests restate the two
bound formulas (same constants) with a single-precision FFT pipeline standing in for the kernel's
(numpy's pocketfft in complex64: a different but comparably accurate fp32 transform) and verify the
inequalities window by window on benign and on badly conditioned data, plus the seed histogram's
threshold rule.  They guard the constants: a slack that is tuned down too far fails here first.
"""
import numpy as np
import pytest

U = 2.0 ** -24
CF_U = np.float32(512.0 * U)


def _pair_correlation_f32(ya, yb, g):
    """D^_a[t], D^_b[t] = sum_j g_j y[t+j] through a 4096-point complex64 FFT of ya + i yb."""
    N = 4096
    z = np.zeros(N, np.complex64)
    z[:ya.size] = ya.astype(np.float32) + 1j * yb.astype(np.float32)
    Z = np.fft.fft(z).astype(np.complex64)
    G = np.fft.fft(np.pad(g.astype(np.float64), (0, N - g.size)))
    Qc = (np.conj(G) / N).astype(np.complex64)
    c = np.fft.ifft((Z * Qc).astype(np.complex64) * np.complex64(N)).astype(np.complex64)   # unnormalised inverse
    qmax = np.float32(np.abs(G).max() * (1 + 1e-7))
    return c.real.astype(np.float32), c.imag.astype(np.float32), qmax


def _rows(rng, kind, T):
    if kind == "gauss":
        return rng.standard_normal((2, T)) * 0.01
    if kind == "heavy":          # heavy tails + a level shift: large pair norm, small windows
        y = rng.standard_t(2.5, size=(2, T)) * 0.01
        y[:, T // 2:] += 0.5
        return y
    if kind == "near_copy":      # windows that almost equal the query (cancellation)
        return None
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["gauss", "heavy", "near_copy"])
def test_identity_fft_bounds_hold(kind):
    rng = np.random.default_rng(7)
    T, W = 4096, 252
    q = (rng.standard_normal(W) * 0.01).astype(np.float32)
    y = _rows(rng, kind, T)
    if y is None:
        y = np.tile(q.astype(np.float64), (2, T // W + 1))[:, :T] * (1 + 1e-4 * rng.standard_normal((2, T)))
    y = y.astype(np.float32)
    Tp = T - W + 1
    Da, Db, qmax = _pair_correlation_f32(y[0], y[1], q)
    yn = np.float32(np.sqrt((y.astype(np.float64) ** 2).sum()) * (1 + 1e-7))
    q2 = np.float32((q.astype(np.float64) ** 2).sum())
    slack = np.float32((2 * CF_U * qmax * yn + np.float32(8 * U) * (q2 + yn * yn)) * np.float32(1.0001))
    for row, D in ((0, Da), (1, Db)):
        y64 = y[row].astype(np.float64)
        c2 = np.concatenate(([0.0], np.cumsum(y64 ** 2)))
        Y2 = (c2[W:W + Tp] - c2[:Tp]).astype(np.float32)
        S = np.array([((q.astype(np.float64) - y64[t:t + W]) ** 2).sum() for t in range(0, Tp, 3)])
        lb = (np.float32(-2) * D[:Tp] + Y2 + (q2 - slack))[::3].astype(np.float64)
        ub = lb + 2 * float(slack)
        assert (lb <= S).all(), float((lb - S).max())
        assert (S <= ub).all(), float((S - ub).max())
        # the bound is not vacuous: the slack is a small fraction of a typical squared distance
        if kind == "gauss":
            assert slack < 0.01 * np.median(S)


@pytest.mark.parametrize("kind", ["gauss", "heavy"])
def test_embedded_fft_bounds_hold(kind):
    """S_t = ||ex - K y_t||^2 = ||ex||^2 - 2 g.y_t + E2_t with the stored energies scaled by (1 - 16u),
    slack = 2 cf_u max|FFT(g)| ||y_pair|| + 16u ||ex||^2 + 2u ||g|| ||y_pair||, UB - LB = 2 slack + 32u E2."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle import oracle
    rng = np.random.default_rng(11)
    T, W = 4096, 126
    K = oracle.foveal_kernel(1.15, 0.9, W).astype(np.float64)
    x = rng.standard_normal(W) * (0.01 if kind == "gauss" else 1.0)
    y = _rows(rng, kind, T).astype(np.float32)
    ex = (K @ x).astype(np.float32)
    g = (ex.astype(np.float64) @ K).astype(np.float32)
    Tp = T - W + 1
    Da, Db, qmax = _pair_correlation_f32(y[0], y[1], g)
    yn = np.float32(np.sqrt((y.astype(np.float64) ** 2).sum()) * (1 + 1e-7))
    gn = np.float32(np.linalg.norm(g.astype(np.float64)) * (1 + 1e-7))
    q2 = np.float32((ex.astype(np.float64) ** 2).sum())
    slack = np.float32((2 * CF_U * qmax * yn + np.float32(16 * U) * q2 + np.float32(2 * U) * gn * yn) * np.float32(1.0001))
    for row, D in ((0, Da), (1, Db)):
        y64 = y[row].astype(np.float64)
        ts = np.arange(0, Tp, 5)
        E = np.stack([K @ y64[t:t + W] for t in ts])                    # (n, d) embedded windows, fp64
        E2 = (E ** 2).sum(1)
        S = ((ex.astype(np.float64)[None, :] - E) ** 2).sum(1)
        e2s = np.nextafter((E2 * (1 - 16 * U)).astype(np.float32), np.float32(0)) # stored: scaled, rounded down
        lb = (np.float32(-2) * D[ts] + e2s + (q2 - slack)).astype(np.float64)
        ub = lb + 2 * float(slack) + 2 * 16 * U * 1.001 * e2s.astype(np.float64)
        assert (lb <= S).all(), float((lb - S).max())
        assert (S <= ub).all(), float((S - ub).max())
        # the property the rigour rests on: |2 D| <= ||ex||^2 + E2 (Cauchy-Schwarz in embedded space)
        Dtrue = np.array([g.astype(np.float64) @ y64[t:t + W] for t in ts])
        assert (2 * np.abs(Dtrue) <= float(q2) * (1 + 1e-6) + E2 * (1 + 1e-6) + 1e-12).all()


def test_seed_histogram_threshold_rule():
    """fft_scan_kernel<SEED>: per-thread minima of UB in logarithmic bins (upper 16 bits of the float,
    4096 bins centred on Q2); the published threshold -- the upper edge of the bin holding the k-th
    minimum -- is >= the k-th smallest minimum, hence >= the k-th smallest UB of the ensemble."""
    rng = np.random.default_rng(3)
    q2 = np.float32(0.0254)
    for scale, k in ((1.0, 1024), (1e-3, 64), (30.0, 5000)):
        ub = (np.abs(rng.standard_normal(75776)) * 0.2 + 1.0).astype(np.float32) * np.float32(q2 * scale)
        base = (int(q2.view(np.uint32)) >> 16) - 2048
        bins = np.clip((ub.view(np.uint32) >> 16).astype(np.int64) - base, 0, 4095)
        hist = np.bincount(bins, minlength=4096)
        b = int(np.searchsorted(np.cumsum(hist), k))          # first bin with cum >= k
        eb = base + b + 1
        thr = np.inf if (b >= 4095 or eb <= 0 or eb >= 0x7F80) else np.array([eb << 16], np.uint32).view(np.float32)[0]
        kth = np.partition(ub, k - 1)[k - 1]
        assert thr >= kth
        if np.isfinite(thr) and b > 0:
            assert thr <= kth * (1 + 2.0 ** -6)               # ... and within one bin (0.8 %) of it
