"""Builder tool: per-step wall time and launch count of shadow() (host numpy in / out) -- finds re-runs and host stalls."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench, shadowing_b200 as sb
from shadowing_b200 import _lib
g = torch.Generator().manual_seed(1234); ds = torch.randn(bench.R_FULL, 1, bench.T, generator=g) * 0.01
qs = bench.make_queries(128); qp = qs.clone().pin_memory()
obj = sb.PathShadowing(sb.Identity(bench.W), sb.RelativeMSE(), ds, sb.PredictionContext(bench.H), device="cuda:0")
for i in range(3): obj.shadow(qp[i:i+1], k=1024)
torch.cuda.synchronize()
ts, ls = [], []
for i in range(120):
    n0 = _lib.launch_count(); t = time.perf_counter()
    out = obj.shadow(qp[i % 128:i % 128 + 1], k=1024)
    ts.append((time.perf_counter() - t) * 1e3); ls.append(_lib.launch_count() - n0)
ts = np.array(ts); ls = np.array(ls)
print("median %.3f mean %.3f max %.3f ms; launches/step: %s" % (np.median(ts), ts.mean(), ts.max(), np.unique(ls, return_counts=True)))
bad = np.argsort(-ts)[:10]
print("slowest steps:", [(int(i), round(float(ts[i]), 3), int(ls[i])) for i in bad])
