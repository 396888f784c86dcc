"""BASELINE configs[2] on a GPU box (not collected by pytest):
256 query dates, R=32768 x T=4096, W=252, k=1024, predict_from_paths realised variance
Ts=[5,10,20], softmax eta=0.1.  Times PathShadowing.predict end to end and checks a sample of
the queries against the CPU oracle (indices bit-exact, predictions <= 1e-6 relative)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import shadowing_b200 as sb
    from oracle import oracle
    R, T, W, H, k, B = 32768, 4096, 252, 20, 1024, 256
    g = torch.Generator().manual_seed(0)
    ds = torch.randn(R, 1, T, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, 1, W, generator=g, dtype=torch.float32) * 0.01
    mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
    obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H), scan_mode=mode)
    rv = sb.RealizedVariance([5, 10, 20], vol=False)
    obj.predict(q[:4], k=k, to_predict=rv, eta=0.1)  # warm-up: upload, spectra
    torch.cuda.synchronize()
    for splits in (1, 8):
        t0 = time.perf_counter()
        pred, pstd = obj.predict(q, k=k, to_predict=rv, eta=0.1, proba_name="softmax", n_context_splits=splits)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"mode={mode} n_context_splits={splits}: predict(256 queries) {dt * 1e3:.1f} ms -> "
              f"{B * R * (T - W - H + 1) / dt:.3e} query-windows/s, pred[0]={pred[0]}")
    # parity on a sample of queries
    dsn = ds.numpy()
    for b in (0, 17, 255):
        d, paths, idx = obj.shadow(q[b:b + 1], k=k)
        do, io = oracle.shadow_topk(dsn, q[b:b + 1].numpy(), k, H)
        assert np.array_equal(d.view(np.uint32), do.view(np.uint32)) and np.array_equal(idx, io), b
        mo, so = oracle.predict_from_paths(do, oracle.gather_paths(dsn, io, W + H), H, [5, 10, 20], False, "softmax", 0.1)
        assert np.allclose(pred[b], mo[0], rtol=1e-6) and np.allclose(pstd[b], so[0], rtol=1e-5), b
    print("CFG3 PARITY OK")


if __name__ == "__main__":
    main()
