"""CPU: the index algebra of the warp-level 1024-point transform (csrc/pshadow_fft3.cuh), restated in numpy.

The kernel's decomposition -- pass A on lane n1 over n2, twiddle w1024^(n1 k2) from the table
{w^(lane 2j), w^(lane (2j+1))} at [j * 32 + lane], 32 x 32 transpose, pass B on lane k2 over n1, output register
k1 of lane k2 = X[32 k1 + k2] -- and the storage orders psh_fft_prepare writes for it (4-byte elements: perm4,
8-byte elements: perm2) are checked against numpy's FFT and against each other.  These are the maps the GPU
parity tests exercise end to end; here they are pinned without a GPU."""
import numpy as np

N = 1024


def perm4(e):   # fx3::perm4
    lane, i = e & 31, e >> 5
    return ((i >> 2) * 32 + lane) * 4 + (i & 3)


def perm2(e):   # fx3::perm2
    lane, i = e & 31, e >> 5
    return ((i >> 1) * 32 + lane) * 2 + (i & 1)


def unperm4(pos):   # fx3::unperm4 / fft_window_of_pos<1024>
    j, lane, g = pos & 3, (pos >> 2) & 31, pos >> 7
    return lane + 32 * (4 * g + j)


def test_permutations_are_bijections_and_inverse():
    e = np.arange(N)
    assert sorted(perm4(e)) == list(range(N)) and sorted(perm2(e)) == list(range(N))
    assert np.array_equal(unperm4(perm4(e)), e)
    # a lane's eight 16-byte loads: group g of lane l holds elements l + 32 (4g .. 4g+3), contiguous in memory
    for lane in (0, 7, 31):
        for g in range(8):
            pos = perm4(lane + 32 * (4 * g + np.arange(4)))
            assert np.array_equal(pos, pos[0] + np.arange(4)) and pos[0] % 4 == 0 and pos[0] == (g * 32 + lane) * 4
    # the query spectrum: float4 index (i >> 1) * 32 + lane holds elements i even (.xy) and i odd (.zw)
    for lane in (0, 19):
        for i in range(0, 32, 2):
            p0, p1 = perm2(lane + 32 * i), perm2(lane + 32 * (i + 1))
            assert p1 == p0 + 1 and p0 // 2 == (i >> 1) * 32 + lane


def test_two_pass_decomposition_equals_inverse_dft():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    w = np.exp(2j * np.pi * np.arange(N) / N)
    # twiddle table as psh_fft_prepare builds it: tw2[j * 32 + lane] = (w^(lane 2j), w^(lane (2j+1)))
    tw2 = np.empty((16 * 32, 2), complex)
    for j in range(16):
        for lane in range(32):
            tw2[j * 32 + lane] = (w[(lane * 2 * j) & 1023], w[(lane * (2 * j + 1)) & 1023])
    v = np.stack([x[lane + 32 * np.arange(32)] for lane in range(32)])        # v[lane][i] = x[lane + 32 i]
    A = np.fft.ifft(v, axis=1) * 32                                            # pass A: inverse 32-point DFT over i
    for j in range(16):
        for lane in range(32):
            A[lane, 2 * j] *= tw2[j * 32 + lane, 0]
            A[lane, 2 * j + 1] *= tw2[j * 32 + lane, 1]
    B = A.T.copy()                                                             # transpose: lane k2 holds A[n1][k2]
    X = np.fft.ifft(B, axis=1) * 32                                            # pass B over n1: X[lane k2][k1]
    out = np.empty(N, complex)
    for lane in range(32):
        out[lane + 32 * np.arange(32)] = X[lane]                               # register c of a lane = X[lane + 32 c]
    ref = np.fft.ifft(x) * N
    assert np.abs(out - ref).max() < 1e-9 * np.abs(ref).max()


def test_radix32_split_used_by_ifft32():
    """ifft32 = one radix-2 stage (decimation in frequency) + two 16-point transforms:
    X[2m] = IDFT16(a + b)[m], X[2m+1] = IDFT16((a - b) w32^n)[m], with w32^8 = i folded into a butterfly."""
    rng = np.random.default_rng(4)
    v = rng.standard_normal(32) + 1j * rng.standard_normal(32)
    w32 = np.exp(2j * np.pi * np.arange(16) / 32)
    assert abs(w32[8] - 1j) < 1e-15
    e = v[:16] + v[16:]
    o = (v[:16] - v[16:]) * w32
    X = np.empty(32, complex)
    X[0::2] = np.fft.ifft(e) * 16
    X[1::2] = np.fft.ifft(o) * 16
    assert np.abs(X - np.fft.ifft(v) * 32).max() < 1e-12


def test_piece_geometry():
    """Overlap-save pieces (fft_aux_layout): hop = (1025 - W) & ~3 windows per 1024-sample piece, every window
    of a row owned by exactly one piece, none of them wrapping around the transform."""
    for T, W, H in ((4096, 252, 20), (8192, 252, 20), (3000, 380, 0), (1030, 300, 3), (5000, 64, 8)):
        Tp = T - W - H + 1
        hop = (1025 - W) & ~3
        nseg = (Tp + hop - 1) // hop
        owned = np.zeros(Tp, int)
        for piece in range(nseg):
            t = np.arange(hop)
            glob = piece * hop + t
            ok = glob < Tp
            assert (t[ok] + W <= 1024).all()                 # no circular wrap inside the piece
            assert (glob[ok] + W <= T).all()                 # the window's samples exist
            owned[glob[ok]] += 1
        assert (owned == 1).all()
        assert (hop + 127) // 128 <= 8                       # energy groups staged per piece
