"""GPU parity of the sharded data plane and of the full-size BASELINE configs, driver-visible
(`pytest -m gpu`):

* the peer-memory exchange kernels (`psh_allgather_merge_packed`, `psh_xchg_send` + `psh_xchg_merge`;
  self-validating-word and fence+flag forms, merge by rank and bitonic-sort merge) with G = 2, 4, 8
  VIRTUAL ranks inside one process: G exchange buffers on one device, one stream per rank, no IPC --
  exactly the kernels the multi-process path launches, with same-device pointers standing in for the
  NVLink-mapped ones.  Every rank's merged result must equal the CPU oracle's global top-k
  (path_shadowing.py:170-173, the cross-split merge), over >= 7 steps so that every rotating buffer
  is reused, including the overflow and timeout flags;
* BASELINE configs[1] at full size, bit-exact against a full oracle scan;
* BASELINE configs[2] (256 query dates + predict_from_paths realised variance, softmax eta = 0.1),
  sampled queries bit-exact, predictions within 1e-6;
* one GPU's shard of BASELINE configs[3] (R = 32768 x T = 8192), bit-exact against a full oracle scan.
"""
import numpy as np
import pytest
import torch

from conftest import assert_topk_equal, make_inputs
from oracle import oracle

pytestmark = pytest.mark.gpu

import shadowing_b200 as sb  # noqa: E402
from shadowing_b200 import _lib  # noqa: E402
from shadowing_b200.distributed import shard_bounds  # noqa: E402

PAD_ROW = 2 ** 31 - 1


class VirtualRanks:
    """G ranks of the sharded scan inside one process: shard g lives on the same device, has its own
    stream and its own exchange buffer; `bufs` is what every rank of the multi-process path holds
    after the IPC handshake (shadowing_b200/distributed.py:_PeerExchange).

    All virtual ranks share ONE CUDA context, which real ranks do not: an exchange kernel that waits for a
    peer must never be resident while that peer still has anything to do that needs the whole context
    (lazy loading of a kernel's module on its first launch, an allocation) -- so the tests compute every
    rank's records first, synchronise, and only then launch the exchange kernels concurrently."""

    def __init__(self, ds, G, B, k, T, W, H, mode=_lib.PSH_MODE_FILTER):
        self.G, self.B, self.k, self.T, self.W, self.H, self.mode = G, B, k, T, W, H, mode
        self.Tp = T - W - H + 1
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.bounds = [shard_bounds(ds.shape[0], G, g) for g in range(G)]
        self.rows = [torch.tensor(ds[lo:hi, 0, :]).to(self.dev).contiguous() for lo, hi in self.bounds]
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(G)]
        self.ws = [None] * G
        self.bufs = [_lib.xchg_create(_lib.xchg_bytes(G, B, k), self.dev)[0] for _ in range(G)]
        self.flags = [torch.zeros(1, dtype=torch.int32, device=self.dev) for _ in range(G)]
        self.rec = [torch.empty((B, k, 3), dtype=torch.int32, device=self.dev) for _ in range(G)]

    def close(self):
        torch.cuda.synchronize()
        for p in self.bufs:
            _lib.xchg_destroy(p, self.dev)

    def local_records(self, g, q):
        """Rank g's k best windows as packed records (padding a shard shorter than k with +inf records,
        as distributed._local_records does)."""
        rows, (lo, _) = self.rows[g], self.bounds[g]
        rec = self.rec[g]
        n_local = rows.shape[0] * self.Tp
        if n_local >= self.k:
            self.ws[g] = _lib.scan_topk_packed(rows, self.T, q, self.H, self.k, lo, self.mode, self.ws[g], None, rec)
            return rec
        rec[..., 0] = 0x7F800000
        rec[..., 1] = PAD_ROW
        rec[..., 2] = 0
        if n_local > 0:
            d, i, self.ws[g] = _lib.scan_topk(rows, self.T, q, self.H, n_local, lo, self.mode, self.ws[g])
            rec[:, :n_local, 0] = d.view(torch.int32)
            rec[:, :n_local, 1:] = i
        return rec


def _oracle_steps(ds, qs, k, H):
    return [oracle.shadow_topk(ds, q, k, H) for q in qs]


@pytest.mark.parametrize("form", ["ll", "flags", "sort"])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_virtual_rank_exchange_fused(G, form, monkeypatch):
    """One launch per rank and step (`psh_allgather_merge_packed`), 7 steps (buffers rotate over the
    epochs), fresh queries every step so that a stale record from an earlier epoch cannot pass."""
    monkeypatch.setenv("PSH_XCHG_LL", "1" if form == "ll" else "0")
    monkeypatch.setenv("PSH_MERGE_SORT", "1" if form == "sort" else "0")
    R, T, W, H, k, B, steps = 8 * G + 3, 700, 30, 10, 128, 3, 7
    ds, _ = make_inputs(R, T, W, B, seed=900 + G)
    qs = [make_inputs(1, 8, W, B, seed=950 + s)[1] for s in range(steps)]
    vr = VirtualRanks(ds, G, B, k, T, W, H)
    try:
        outs = []
        for s in range(steps):
            qd = torch.tensor(qs[s][:, 0, :]).to(vr.dev)
            step = []
            for g in range(G):
                with torch.cuda.stream(vr.streams[g]):
                    vr.streams[g].wait_stream(torch.cuda.current_stream())
                    vr.local_records(g, qd)
            torch.cuda.synchronize()
            for g in range(G):
                with torch.cuda.stream(vr.streams[g]):
                    step.append(_lib.allgather_merge_packed(vr.rec[g], vr.bufs, g, vr.Tp, s + 1, vr.flags[g]))
            outs.append(step)
        torch.cuda.synchronize()
        for s, (do, io) in enumerate(_oracle_steps(ds, qs, k, H)):
            for g in range(G):
                d, i = outs[s][g]
                assert_topk_equal(d.cpu().numpy(), i.cpu().numpy(), do, io)
                assert np.array_equal(i.cpu().numpy(), io), (s, g)
        assert all(int(f.item()) == 0 for f in vr.flags)
    finally:
        vr.close()


@pytest.mark.parametrize("form", ["ll", "flags"])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_virtual_rank_exchange_split_pipeline(G, form, monkeypatch):
    """The split form as the pipelined device loop enqueues it: scan(s+1), send(s+1), merge(s) -- a rank
    writes step s+1 while its peers may still merge step s (rotating buffers); 8 steps."""
    monkeypatch.setenv("PSH_XCHG_LL", "1" if form == "ll" else "0")
    R, T, W, H, k, B, steps = 5 * G + 1, 512, 24, 6, 96, 2, 8
    ds, _ = make_inputs(R, T, W, B, seed=700 + G)
    qs = [make_inputs(1, 8, W, B, seed=750 + s)[1] for s in range(steps)]
    vr = VirtualRanks(ds, G, B, k, T, W, H)
    try:
        outs = [[(torch.empty((B, k), dtype=torch.float32, device=vr.dev),
                  torch.empty((B, k, 2), dtype=torch.int32, device=vr.dev)) for _ in range(G)] for _ in range(steps)]
        for s in range(steps + 1):
            qd = torch.tensor(qs[s][:, 0, :]).to(vr.dev) if s < steps else None
            if s < steps:
                for g in range(G):
                    with torch.cuda.stream(vr.streams[g]):
                        vr.streams[g].wait_stream(torch.cuda.current_stream())
                        vr.local_records(g, qd)
                torch.cuda.synchronize()
            for g in range(G):
                with torch.cuda.stream(vr.streams[g]):
                    if s < steps:
                        _lib.xchg_send(vr.rec[g], vr.bufs, g, s + 1)
                    if s > 0:
                        d, i = outs[s - 1][g]
                        _lib.xchg_merge(vr.bufs, g, B, k, vr.Tp, s, d, i, vr.flags[g])
        torch.cuda.synchronize()
        for s, (do, io) in enumerate(_oracle_steps(ds, qs, k, H)):
            for g in range(G):
                d, i = outs[s][g]
                assert np.array_equal(d.cpu().numpy().view(np.uint32), do.view(np.uint32)), (s, g)
                assert np.array_equal(i.cpu().numpy(), io), (s, g)
        assert all(int(f.item()) == 0 for f in vr.flags)
    finally:
        vr.close()


def test_virtual_rank_exchange_short_shards_and_ties():
    """World larger than the row count (empty and short shards padded with +inf records) and rows
    duplicated across shards (ties in distance are ordered by the global flat index)."""
    G, T, W, H, k, B = 8, 600, 40, 10, 200, 2
    base, q = make_inputs(1, T, W, B, seed=3)
    ds = np.repeat(base, 5, axis=0)   # 5 identical rows over 8 ranks: 3 empty shards, every distance 5-fold
    vr = VirtualRanks(ds, G, B, k, T, W, H)
    try:
        qd = torch.tensor(q[:, 0, :]).to(vr.dev)
        outs = []
        for g in range(G):
            with torch.cuda.stream(vr.streams[g]):
                vr.streams[g].wait_stream(torch.cuda.current_stream())
                vr.local_records(g, qd)
        torch.cuda.synchronize()
        for g in range(G):
            with torch.cuda.stream(vr.streams[g]):
                outs.append(_lib.allgather_merge_packed(vr.rec[g], vr.bufs, g, vr.Tp, 1, vr.flags[g]))
        torch.cuda.synchronize()
        do, io = oracle.shadow_topk(ds, q, k, H)
        for d, i in outs:
            assert np.array_equal(d.cpu().numpy().view(np.uint32), do.view(np.uint32))
            assert np.array_equal(i.cpu().numpy(), io)
    finally:
        vr.close()


@pytest.mark.parametrize("form", ["ll", "flags"])
def test_virtual_rank_exchange_overflow_and_timeout_flags(form, monkeypatch):
    """A shard whose enqueue-only scan overflowed poisons its first record: every rank must raise bit 0
    of its flag.  A peer that never sends raises bit 1 after the timeout instead of hanging the GPU."""
    monkeypatch.setenv("PSH_XCHG_LL", "1" if form == "ll" else "0")
    monkeypatch.setenv("PSH_XCHG_TIMEOUT_MS", "200")
    G, T, W, H, k, B = 4, 400, 16, 4, 64, 2
    ds, q = make_inputs(40, T, W, B, seed=31)
    vr = VirtualRanks(ds, G, B, k, T, W, H)
    try:
        qd = torch.tensor(q[:, 0, :]).to(vr.dev)
        for g in range(G):
            with torch.cuda.stream(vr.streams[g]):
                vr.streams[g].wait_stream(torch.cuda.current_stream())
                rec = vr.local_records(g, qd)
                if g == 2:
                    rec[1, 0, 0] = -1   # 0xffffffff: the overflow poison of select_kernel / finalize_kernel
        torch.cuda.synchronize()
        for g in range(G):
            with torch.cuda.stream(vr.streams[g]):
                _lib.allgather_merge_packed(vr.rec[g], vr.bufs, g, vr.Tp, 1, vr.flags[g])
        torch.cuda.synchronize()
        assert all(int(f.item()) == 1 for f in vr.flags)
        # epoch 2: rank 3 never sends -> the others time out (bit 1), nobody hangs
        for f in vr.flags:
            f.zero_()
        for g in range(G - 1):
            with torch.cuda.stream(vr.streams[g]):
                _lib.allgather_merge_packed(vr.rec[g], vr.bufs, g, vr.Tp, 2, vr.flags[g])
        torch.cuda.synchronize()
        assert all(int(f.item()) & 2 for f in vr.flags[:G - 1])
    finally:
        vr.close()


# ---------------------------------------------------------------------------------------------
# full-size BASELINE configs
# ---------------------------------------------------------------------------------------------
def _cfg2_inputs(B=1):
    g = torch.Generator().manual_seed(0)
    ds = torch.randn(32768, 1, 4096, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, 1, 252, generator=g, dtype=torch.float32) * 0.01
    return ds, q


@pytest.fixture(scope="module")
def cfg2():
    """The cfg2 ensemble, resident once for the tests of this module (512 MiB + spectra)."""
    ds, q = _cfg2_inputs(256)
    obj = sb.PathShadowing(sb.Identity(252), sb.RelativeMSE(), ds, sb.PredictionContext(20))
    yield ds.numpy(), q, obj
    del obj
    torch.cuda.empty_cache()


def test_full_cfg2_bit_exact(cfg2):
    """BASELINE configs[1]: R=32768 x T=4096, W=252, H=20, k=1024, one query date -- distances, indices
    and paths bit-exact against a FULL scan by the C oracle (125 337 600 windows)."""
    dsn, q, obj = cfg2
    for b in (0, 1):
        d, paths, idx = obj.shadow(q[b:b + 1], k=1024)
        do, po, io = oracle.shadow(dsn, q[b:b + 1].numpy(), 1024, 20)
        assert np.array_equal(d.view(np.uint32), do.view(np.uint32))
        assert np.array_equal(idx, io)
        assert np.array_equal(paths, po)


def test_cfg3_batched_predict_sampled(cfg2):
    """BASELINE configs[2]: 256 query dates in ONE call, predict_from_paths realised variance
    Ts=[5,10,20], softmax eta=0.1.  8 sampled queries are checked against full oracle scans (indices
    bit-exact), all 256 predictions against the numpy restatement on the returned paths (1e-6)."""
    dsn, q, obj = cfg2
    k, H, Ts = 1024, 20, [5, 10, 20]
    d, paths, idx = obj.shadow(q, k=k)
    assert d.shape == (256, k) and paths.shape == (256, k, 1, 272) and idx.shape == (256, k, 2)
    for b in (0, 17, 31, 32, 100, 128, 200, 255):
        do, io = oracle.shadow_topk(dsn, q[b:b + 1].numpy(), k, H)
        assert np.array_equal(d[b:b + 1].view(np.uint32), do.view(np.uint32)), b
        assert np.array_equal(idx[b:b + 1], io), b
        assert np.array_equal(paths[b:b + 1], oracle.gather_paths(dsn, io, 272)), b
    rv = sb.RealizedVariance(Ts, vol=False)
    pred, pstd = obj.predict(q, k=k, to_predict=rv, eta=0.1, proba_name="softmax")
    mo, so = oracle.predict_from_paths(d, paths, H, Ts, False, "softmax", 0.1)
    assert pred.shape == (256, 3)
    assert np.allclose(pred, mo, rtol=1e-6, atol=0) and np.allclose(pstd, so, rtol=1e-5, atol=0)
    # the same through chunked contexts (path_shadowing.py:283-299)
    pred8, pstd8 = obj.predict(q, k=k, to_predict=rv, eta=0.1, n_context_splits=8)
    assert np.array_equal(pred8, pred) and np.array_equal(pstd8, pstd)


def test_cfg4_shard_bit_exact():
    """One GPU's share of BASELINE configs[3]: 32768 rows x T=8192 (rank 3's seed, rows 98304..131071 of
    the 262144-row ensemble), W=252, H=20, k=1024 -- bit-exact against a full oracle scan, with GLOBAL
    trajectory indices (row_offset)."""
    R, T, W, H, k, rank = 32768, 8192, 252, 20, 1024, 3
    g = torch.Generator().manual_seed(0 + rank)
    ds = torch.randn(R, 1, T, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(1)
    q = torch.randn(1, 1, W, generator=g, dtype=torch.float32) * 0.01
    obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H))
    rows, T_ = obj._resident_rows()
    mode, aux = obj._mode_and_aux(rows, T_, W, H)
    assert mode == _lib.PSH_MODE_FFT   # T > 4096: overlapping 4096-sample pieces (virtual rows)
    d, idx, _ = _lib.scan_topk(rows, T_, q[:, 0, :].cuda().contiguous(), H, k, rank * R, mode, None, aux)
    do, io = oracle.shadow_topk(ds.numpy(), q.numpy(), k, H, row_offset=rank * R)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), do.view(np.uint32))
    assert np.array_equal(idx.cpu().numpy(), io)
    del obj
    torch.cuda.empty_cache()
