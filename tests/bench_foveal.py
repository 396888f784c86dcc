"""Reference benchmark configuration (testing.ipynb:95-103): predict() with Foveal(1.15, 0.9, 126),
PredictionContext(252), k=10000, realized_variance Ts=[2,7,252], eta=0.1 on a randn ensemble
(reference: R=131072 x T=4096, 2.65 s/it on an unstated GPU).  Run on a GPU box:
    python tests/bench_foveal.py [R]"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import shadowing_b200 as sb

R = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T, W, H, k = 4096, 126, 252, 10000
g = torch.Generator().manual_seed(0)
ds = torch.randn(R, 1, T, generator=g)
x = torch.randn(4, 1, W, generator=g)
obj = sb.PathShadowing(sb.Foveal(1.15, 0.9, W), sb.RelativeMSE(), ds, sb.PredictionContext(H))
rv = sb.RealizedVariance([2, 7, 252], vol=False)
obj.predict(x[:1], k=k, to_predict=rv, eta=0.1)   # residency + warm-up
torch.cuda.synchronize()
for B in (1, 4):
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        pred, std = obj.predict(x[:B], k=k, to_predict=rv, eta=0.1, n_dataset_splits=8, cuda=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"foveal predict R={R} B={B}: {dt * 1e3:.2f} ms/it  ({B * R * (T - W - H + 1) / dt:.3e} windows/s); "
          f"reference published 2650 ms/it at R=131072, B=1; pred[0]={pred[0]}")
