"""CPU, world_size=2, gloo: the host-side logic of the R-sharded path (shadowing_b200/distributed.py):
shard bounds, global row offsets, padding of short shards, all-gather + merge, owner-gather +
all-reduce of paths.  The three device entry points are replaced by oracle-backed stand-ins
(the oracle is the checker here; the CUDA merge/gather themselves are covered by -m gpu tests)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _standins():
    from oracle import oracle

    def scan_topk(rows, T, q, H, k, row_offset=0, mode=1, workspace=None, aux=None):
        d, i = oracle.shadow_topk(rows[:, :T].numpy(), q.numpy(), k, H, row_offset=row_offset)
        return torch.from_numpy(d), torch.from_numpy(i), workspace

    def merge_topk(d_parts, i_parts, Tp):
        G, B, k = d_parts.shape
        d = d_parts.permute(1, 0, 2).reshape(B, G * k).numpy()
        i = i_parts.permute(1, 0, 2, 3).reshape(B, G * k, 2).numpy()
        od = np.empty((B, k), np.float32)
        oi = np.empty((B, k, 2), np.int32)
        for b in range(B):
            flat = i[b, :, 0].astype(np.int64) * Tp + i[b, :, 1]
            order = np.lexsort((flat, d[b].view(np.uint32)))[:k]
            od[b], oi[b] = d[b][order], i[b][order]
        return torch.from_numpy(od), torch.from_numpy(oi)

    def gather_paths(rows, T, idx, L, row_offset=0):
        B, k, _ = idx.shape
        out = torch.zeros((B, k, 1, L), dtype=torch.float32)
        r = idx[..., 0].long() - row_offset
        own = (r >= 0) & (r < rows.shape[0])
        for b in range(B):
            for j in range(k):
                if own[b, j]:
                    t = int(idx[b, j, 1])
                    out[b, j, 0] = rows[r[b, j], t:t + L]
        return out

    return scan_topk, merge_topk, gather_paths


def _worker(rank, world, port, R, T, W, H, k, B, tmp):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import make_inputs
        import shadowing_b200 as sb
        from shadowing_b200 import _lib, distributed

        _lib.scan_topk, _lib.merge_topk, _lib.gather_paths = _standins()
        _lib.require_cuda = lambda: None
        ds, q = make_inputs(R, T, W, B, seed=77)
        lo, hi = distributed.shard_bounds(R, world, rank)
        obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds[lo:hi], sb.PredictionContext(H),
                               device="cpu", row_offset=lo, process_group=dist.group.WORLD, scan_mode="filter")
        d, paths, idx = obj.shadow(q, k=k)
        np.savez(Path(tmp) / f"rank{rank}.npz", d=d, paths=paths, idx=idx)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R,T,W,H,k,B", [(37, 300, 20, 5, 64, 3), (3, 200, 16, 4, 300, 2)])
def test_sharded_shadow_equals_single_process(tmp_path, R, T, W, H, k, B):
    from conftest import make_inputs
    from oracle import oracle
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, R, T, W, H, k, B, str(tmp_path)), nprocs=2, join=True)
    ds, q = make_inputs(R, T, W, B, seed=77)
    do, po, io = oracle.shadow(ds, q, k, H)
    for rank in range(2):
        z = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(z["d"], do) and np.array_equal(z["idx"], io) and np.array_equal(z["paths"], po)


def test_shard_bounds_cover_rows():
    from shadowing_b200.distributed import shard_bounds
    for R in (1, 7, 8, 32768, 262144 + 3):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(R, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == R
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1
