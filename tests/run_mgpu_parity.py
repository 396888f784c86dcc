"""Multi-GPU parity (run under torchrun on a GPU box, not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/run_mgpu_parity.py
The ensemble is sharded over N ranks (NCCL); every rank must return exactly the CPU oracle's
single-process answer (distances, indices, paths) -- SURVEY.md section 8(e)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    from conftest import make_inputs
    import shadowing_b200 as sb
    from shadowing_b200.distributed import shard_bounds
    from oracle import oracle

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [(1024, 4096, 252, 20, 1024, 2, False), (37, 700, 20, 5, 300, 3, False), (3, 600, 16, 4, 900, 1, False),
             (1, 600, 16, 4, 100, 1, False),          # world > rows: empty shards
             (1024, 1024, 32, 0, 64, 1, True)]        # adversarial order: candidate buffers overflow on every rank
    for (R, T, W, H, k, B, adversarial) in cases:
        ds, q = make_inputs(R, T, W, B, seed=500 + R)
        if adversarial:
            # every rank's seed rows (its first permuted slots) are far, all its other rows near:
            # the candidate buffers overflow and all ranks must repeat the step together
            import math
            ds = ds * 1e-3
            Tp = T - W - H + 1
            for rk in range(world):
                lo_, hi_ = shard_bounds(R, world, rk)
                n = hi_ - lo_
                p = max(int(0.6180339887498949 * n), 1) if n > 2 else 1
                while n > 2 and math.gcd(p, n) != 1:
                    p += 1
                p = (p % n or 1) if n > 2 else 1
                n0 = min(max(-(-16 * k // Tp), 1), n)
                for i in range(n0):
                    ds[lo_ + (i * p) % n] *= 1e5
        lo, hi = shard_bounds(R, world, rank)
        obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds[lo:hi], sb.PredictionContext(H or None), device=dev,
                               row_offset=lo, process_group=dist.group.WORLD,
                               scan_mode="filter" if adversarial else "auto")
        d, paths, idx = obj.shadow(q, k=k)
        do, po, io = oracle.shadow(ds, q, k, H)
        if H:
            pred, pstd = obj.predict(q, k=k, to_predict=sb.RealizedVariance([2, 4]), eta=0.1)
            mo, so = oracle.predict_from_paths(do, po, H, [2, 4], False, "softmax", 0.1)
        else:
            pred = pstd = mo = so = np.zeros(1)
        good = (np.array_equal(d.view(np.uint32), do.view(np.uint32)) and np.array_equal(idx, io)
                and np.array_equal(paths, po) and np.allclose(pred, mo, rtol=1e-6) and np.allclose(pstd, so, rtol=1e-5))
        print(f"rank {rank}/{world} R={R} T={T} W={W} k={k} B={B} adv={adversarial}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    # pipelined sharded scans (bench.py's device loop): scan(i+1), send(i+1), merge(i) -- every query's
    # result must equal the synchronous sharded result once the pipeline has been checked
    R, T, W, H, k = 700, 2048, 100, 10, 300
    ds, q = make_inputs(R, T, W, 6, seed=640)
    lo, hi = shard_bounds(R, world, rank)
    obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds[lo:hi], sb.PredictionContext(H), device=dev,
                           row_offset=lo, process_group=dist.group.WORLD)
    rows, T_ = obj._resident_rows()
    qd = torch.tensor(q)
    do, io = oracle.shadow_topk(ds, q, k, H)
    for defer in ("0", "1"):   # fused exchange launch / split send + deferred merge (PSH_DEFER)
        os.environ["PSH_DEFER"] = defer
        outs = [obj._scan_device(qd[i:i + 1], rows, T_, k, nosync=True) for i in range(6)]
        obj._check_pipeline()
        good = all(np.array_equal(d_.cpu().numpy().view(np.uint32), do[i:i + 1].view(np.uint32))
                   and np.array_equal(i_.cpu().numpy(), io[i:i + 1]) for i, (d_, i_) in enumerate(outs))
        print(f"rank {rank}/{world} pipelined sharded scans (PSH_DEFER={defer}): {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    os.environ.pop("PSH_DEFER", None)
    # Foveal embedding, sharded: every rank must return what ONE GPU holding all rows returns
    # (bit-identical: same kernel, same per-window arithmetic, exact merge), and that agrees with the
    # CPU oracle within the embedded scan's tolerance
    from conftest import assert_topk_close
    for (R, T, W, H, k, B) in [(301, 2000, 126, 50, 700, 4), (2, 900, 64, 0, 1200, 1)]:
        ds, q = make_inputs(R, T, W, B, seed=600 + R)
        emb = sb.Foveal(1.15, 0.9, W)
        lo, hi = shard_bounds(R, world, rank)
        obj = sb.PathShadowing(emb, sb.RelativeMSE(), ds[lo:hi], sb.PredictionContext(H or None), device=dev,
                               row_offset=lo, process_group=dist.group.WORLD)
        d, paths, idx = obj.shadow(q, k=k)
        one = sb.PathShadowing(emb, sb.RelativeMSE(), ds, sb.PredictionContext(H or None), device=dev)
        d1, p1, i1 = one.shadow(q, k=k)
        good = np.array_equal(d.view(np.uint32), d1.view(np.uint32)) and np.array_equal(idx, i1) and np.array_equal(paths, p1)
        ex = emb(torch.tensor(q))[:, 0, :].numpy()
        do, io = oracle.embed_topk(ds, emb.kernel.numpy()[:, 0, :], ex, k, H)
        try:
            assert_topk_close(d, idx, do, io)
        except AssertionError:
            good = False
        print(f"rank {rank}/{world} Foveal R={R} T={T} W={W} k={k} B={B}: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if int(flag.item()) != 0:
        sys.exit(1)
    if rank == 0:
        print("MGPU PARITY OK")


if __name__ == "__main__":
    main()
