"""CPU: the rigorous bounds of the FFT filter flavours, restated in numpy and checked against fp64.

The CUDA filter (csrc/pshadow_fftscan.cuh `fft_scan_kernel`, csrc/pshadow_embed_fft.cuh) keeps a window iff
a LOWER bound of its squared distance passes the threshold and tightens thresholds from UPPER bounds;
results are exact only if  LB <= S_true <= UB  holds for every window.  These tests restate the bound
formulas (same constants, same data formats: spectra quantised to fp16 pairs with a per-pair power-of-two
scale and their MEASURED error, window energies scaled and rounded DOWN to fp16) with a single-precision
FFT pipeline standing in for the kernel's (numpy's pocketfft in complex64: a different but comparably
accurate fp32 transform) and verify the inequalities window by window on benign and on badly conditioned
data, plus the threshold histogram's rule.  They guard the constants: a slack that is tuned down too far
fails here first.
"""
import numpy as np
import pytest

U = 2.0 ** -24
CF_U = np.float32(512.0 * U)


def _pow2_scale(m, top):
    """power of two s with m * s in [2^(top-1), 2^top) (pshadow_fftscan.cuh: pow2_scale)"""
    if not (m > 0 and np.isfinite(m)):
        return np.float32(1.0)
    _, e = np.frexp(np.float32(m))
    return np.float32(np.ldexp(1.0, int(np.clip(top - e, -100, 100))))


def _floor_f16(x):
    """fp32 -> fp16 rounded toward -inf (__float2half_rd), +inf kept"""
    h = x.astype(np.float16)
    up = h.astype(np.float32) > x
    h[up] = np.nextafter(h[up], np.float16(-np.inf))
    return h


def _floor_bf16(x):
    """fp32 >= 0 -> bf16 rounded toward -inf = the upper 16 bits (fft_prep_energy_kernel<., 1024>), +inf kept"""
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)


def _pair_correlation_f32(ya, yb, g, quantise=True, N=4096):
    """v[t] (the kernel's transform output, in the pair's scaled units when `quantise`), and the per-pair
    statistics psh_fft_prepare stores: zs, zqerr.  D^[t] = v[t] / zs.  N: the transform length of the
    flavour (4096: one CTA per transform; 1024: one warp per transform, pieces of 1024 samples)."""
    z = np.zeros(N, np.complex64)
    z[:ya.size] = ya.astype(np.float32) + 1j * yb.astype(np.float32)
    Z = np.fft.fft(z).astype(np.complex64)
    zs, zqerr = np.float32(1.0), np.float32(0.0)
    if quantise:
        zs = _pow2_scale(max(np.abs(Z.real).max(), np.abs(Z.imag).max()), 14)
        zh = (Z.real * zs).astype(np.float16).astype(np.float32) + 1j * (Z.imag * zs).astype(np.float16).astype(np.float32)
        err = np.abs(zh.astype(np.complex128) / float(zs) - Z.astype(np.complex128))
        zqerr = np.float32(np.sqrt((err ** 2).sum() / N) * (1 + 1e-6))
        Z = zh.astype(np.complex64)
    G = np.fft.fft(np.pad(g.astype(np.float64), (0, N - g.size)))
    Qc = (np.conj(G) / N).astype(np.complex64)
    c = np.fft.ifft((Z * Qc).astype(np.complex64) * np.complex64(N)).astype(np.complex64)   # unnormalised inverse
    qmax = np.float32(np.abs(G).max() * (1 + 1e-7))
    return c.real.astype(np.float32), c.imag.astype(np.float32), qmax, zs, zqerr


def _rows(rng, kind, T):
    if kind == "gauss":
        return rng.standard_normal((2, T)) * 0.01
    if kind == "heavy":          # heavy tails + a level shift: large pair norm, small windows
        y = rng.standard_t(2.5, size=(2, T)) * 0.01
        y[:, T // 2:] += 0.5
        return y
    if kind == "mixed_scale":    # one row 1000 x the other: the pair shares ONE energy / spectrum scale
        y = rng.standard_normal((2, T)) * 0.01
        y[1] *= 1e-3
        return y
    if kind == "near_copy":      # windows that almost equal the query (cancellation)
        return None
    raise ValueError(kind)


def _staged_energies(y, W, Tp, embed=None, bf16=False):
    """(yf (2, Tp) fp32 = what the kernel reads: energies * es floored to fp16 -- bf16 in the 1024-point
    flavour), es"""
    e = []
    for row in range(2):
        y64 = y[row].astype(np.float64)
        if embed is None:
            c2 = np.concatenate(([0.0], np.cumsum(y64 ** 2)))
            e.append(np.nextafter((c2[W:W + Tp] - c2[:Tp]).astype(np.float32), np.float32(-np.inf)).clip(min=0))
        else:
            E = np.stack([embed @ y64[t:t + W] for t in range(Tp)])
            e.append(np.nextafter(((E ** 2).sum(1) * (1 - 16 * U)).astype(np.float32), np.float32(-np.inf)).clip(min=0))
    e = np.stack(e)
    es = _pow2_scale(e.max(), 15)
    if bf16:
        return _floor_bf16(np.minimum(e * es, np.float32(65504.0))), es
    return _floor_f16(np.minimum(e * es, np.float32(65504.0))).astype(np.float32), es


@pytest.mark.parametrize("N", [4096, 1024])
@pytest.mark.parametrize("kind", ["gauss", "heavy", "near_copy", "mixed_scale"])
def test_identity_fft_bounds_hold(kind, N):
    """LB = Q2 + Y2^ - 2 D^ - slack <= S <= UB = LB + 2 slack + 2^-10 Y2^ (+ 2^-24 / es), with
    slack = 2 cf_u Qmax ynorm + 2 zqerr ||q|| + 12u (Q2 + ynorm^2), evaluated as the kernel does:
    fma(m2, v, yf) in the pair's scaled units.  N = 1024: the warp-level flavour's piece of 1024 samples,
    energies floored to bf16 (UB - LB grows to 2^-7 Y2^)."""
    rng = np.random.default_rng(7)
    T, W = N, 252
    q = (rng.standard_normal(W) * 0.01).astype(np.float32)
    y = _rows(rng, kind, T)
    if y is None:
        y = np.tile(q.astype(np.float64), (2, T // W + 1))[:, :T] * (1 + 1e-4 * rng.standard_normal((2, T)))
    y = y.astype(np.float32)
    Tp = T - W + 1
    va, vb, qmax, zs, zqerr = _pair_correlation_f32(y[0], y[1], q, N=N)
    yn = np.float32(np.sqrt((y.astype(np.float64) ** 2).sum()) * (1 + 1e-7))
    gn = np.float32(np.linalg.norm(q.astype(np.float64)) * (1 + 1e-7))
    q2 = np.float32((q.astype(np.float64) ** 2).sum())
    slack = np.float32((2 * CF_U * qmax * yn + 2 * zqerr * gn + np.float32(12 * U) * (q2 + yn * yn)) * np.float32(1.0001))
    yf, es = _staged_energies(y, W, Tp, bf16=(N == 1024))
    m2 = np.float32(-2.0) * es / zs
    cu = np.float32((2.0 ** -7 if N == 1024 else 2.0 ** -10) * 1.01)
    loosest = 0.0
    for row, v in ((0, va), (1, vb)):
        y64 = y[row].astype(np.float64)
        ts = np.arange(0, Tp, 3)
        S = np.array([((q.astype(np.float64) - y64[t:t + W]) ** 2).sum() for t in ts])
        val = (m2 * v[:Tp] + yf[row])[ts]                                   # scaled units (fp32, as the FFMA2)
        lb = val.astype(np.float64) / float(es) + float(q2 - slack)
        ub = (((val + cu * yf[row][ts]).astype(np.float64) + 2.0 ** -24) / float(es) + float(q2 - slack) + 2 * float(slack)) * 1.000001
        assert (lb <= S).all(), float((lb - S).max())
        assert (S <= ub).all(), float((S - ub).max())
        loosest = max(loosest, float(np.median(ub - lb) / np.median(S)))
    # the bound is not vacuous: UB - LB is a small fraction of a typical squared distance
    if kind == "gauss":
        assert loosest < 0.01, loosest


@pytest.mark.parametrize("kind", ["gauss", "heavy"])
def test_embedded_fft_bounds_hold(kind):
    """S_t = ||ex - K y_t||^2 = ||ex||^2 - 2 g.y_t + E2_t with the stored energies scaled by (1 - 16u) and
    floored to fp16, slack = 2 cf_u max|FFT(g)| ||y_pair|| + 2 zqerr ||g|| + 20u ||ex||^2 + 2u ||g|| ||y_pair||,
    UB - LB = 2 slack + (2^-10 + 32u) E2^."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle import oracle
    rng = np.random.default_rng(11)
    T, W = 2048, 126
    K = oracle.foveal_kernel(1.15, 0.9, W).astype(np.float64)
    x = rng.standard_normal(W) * (0.01 if kind == "gauss" else 1.0)
    y = _rows(rng, kind, T).astype(np.float32)
    ex = (K @ x).astype(np.float32)
    g = (ex.astype(np.float64) @ K).astype(np.float32)
    Tp = T - W + 1
    va, vb, qmax, zs, zqerr = _pair_correlation_f32(y[0], y[1], g)
    yn = np.float32(np.sqrt((y.astype(np.float64) ** 2).sum()) * (1 + 1e-7))
    gn = np.float32(np.linalg.norm(g.astype(np.float64)) * (1 + 1e-7))
    q2 = np.float32((ex.astype(np.float64) ** 2).sum())
    slack = np.float32((2 * CF_U * qmax * yn + 2 * zqerr * gn + np.float32(20 * U) * q2 + np.float32(2 * U) * gn * yn) * np.float32(1.0001))
    yf, es = _staged_energies(y, W, Tp, embed=K)
    m2 = np.float32(-2.0) * es / zs
    cu = np.float32(2.0 ** -10 * 1.01 + 2 * 16 * U * 1.001)
    for row, v in ((0, va), (1, vb)):
        y64 = y[row].astype(np.float64)
        ts = np.arange(0, Tp, 5)
        E = np.stack([K @ y64[t:t + W] for t in ts])                    # (n, d) embedded windows, fp64
        E2 = (E ** 2).sum(1)
        S = ((ex.astype(np.float64)[None, :] - E) ** 2).sum(1)
        val = (m2 * v[:Tp] + yf[row])[ts]
        lb = val.astype(np.float64) / float(es) + float(q2 - slack)
        ub = (((val + cu * yf[row][ts]).astype(np.float64) + 2.0 ** -24) / float(es) + float(q2 - slack) + 2 * float(slack)) * 1.000001
        assert (lb <= S).all(), float((lb - S).max())
        assert (S <= ub).all(), float((S - ub).max())
        # the property the rigour rests on: |2 D| <= ||ex||^2 + E2 (Cauchy-Schwarz in embedded space)
        Dtrue = np.array([g.astype(np.float64) @ y64[t:t + W] for t in ts])
        assert (2 * np.abs(Dtrue) <= float(q2) * (1 + 1e-6) + E2 * (1 + 1e-6) + 1e-12).all()


def test_spectrum_quantisation_error_bound():
    """|D^_t - D_t| <= zqerr ||q||_2 for every t (Cauchy-Schwarz on the quantisation error of the spectrum),
    with zqerr = ||Z^ - Z||_2 / sqrt(N) measured by psh_fft_prepare -- checked in fp64 so that only the
    quantisation is tested."""
    rng = np.random.default_rng(5)
    N, W = 4096, 252
    for scale in (0.01, 37.0):
        z = (rng.standard_normal(N) + 1j * rng.standard_normal(N)) * scale
        q = rng.standard_normal(W) * 0.01
        Z = np.fft.fft(z)
        zs = float(_pow2_scale(max(np.abs(Z.real).max(), np.abs(Z.imag).max()), 14))
        Zh = ((Z.real * zs).astype(np.float16).astype(np.float64) + 1j * (Z.imag * zs).astype(np.float16).astype(np.float64)) / zs
        zqerr = np.sqrt((np.abs(Zh - Z) ** 2).sum() / N)
        Qc = np.conj(np.fft.fft(np.pad(q, (0, N - W))))
        d_true = np.fft.ifft(Z * Qc)
        d_q = np.fft.ifft(Zh * Qc)
        assert np.abs(d_q - d_true).max() <= zqerr * np.linalg.norm(q) * (1 + 1e-9)
        # ... and the bound is tight enough to matter: ~2^-12 of ||z|| ||q||
        assert zqerr < 2.0 ** -11 * np.linalg.norm(z)


def test_threshold_histogram_rule():
    """fft_scan_kernel: upper bounds in logarithmic bins on the float's bit pattern (bits >> 13: 0.1 % wide;
    8192 fine bins = 8 binades below 8 Q2, 64 coarse bins of 128); the threshold is the upper edge of the
    bin holding the k-th entry -- >= the k-th smallest UB -- found through the coarse counts first."""
    rng = np.random.default_rng(3)
    q2 = np.float32(0.0254)
    HB, HC = 8192, 64
    hbase = (int(np.float32(8.0 * q2).view(np.uint32)) >> 13) - HB
    for scale, k in ((1.0, 1024), (1e-3, 64), (3.0, 5000), (40.0, 100)):
        ub = (np.abs(rng.standard_normal(75776)) * 0.2 + 1.0).astype(np.float32) * np.float32(q2 * scale)
        bins = (ub.view(np.uint32) >> 13).astype(np.int64) - hbase
        keep = bins < HB                                   # beyond 8 Q2: not counted
        bins = np.clip(bins[keep], 0, None)
        fine = np.bincount(bins, minlength=HB)
        coarse = fine.reshape(HC, HB // HC).sum(1)
        if coarse.sum() < k:                               # the histogram certifies nothing: thresholds stay +inf
            assert scale >= 8.0
            continue
        cb = int(np.searchsorted(np.cumsum(coarse), k))
        below = int(coarse[:cb].sum())
        e = cb * (HB // HC) + int(np.searchsorted(np.cumsum(fine[cb * (HB // HC):(cb + 1) * (HB // HC)]) + below, k))
        thr = np.array([(hbase + e + 1) << 13], np.uint32).view(np.float32)[0]
        kth = np.partition(ub, k - 1)[k - 1]
        assert thr >= kth
        if e > 0:
            assert thr <= kth * (1 + 2.0 ** -9)            # ... and within one bin (0.1 - 0.2 %) of it
