"""Where does predict() at B=256 spend its time?  (GPU box; not collected by pytest)"""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import shadowing_b200 as sb

R, T, W, H, k, B = 32768, 4096, 252, 20, 1024, 256
g = torch.Generator().manual_seed(0)
ds = torch.randn(R, 1, T, generator=g) * 0.01
g = torch.Generator().manual_seed(1)
q = torch.randn(B, 1, W, generator=g) * 0.01
obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H))
rv = sb.RealizedVariance([5, 10, 20])
obj.predict(q[:4], k=k, to_predict=rv, eta=0.1)
torch.cuda.synchronize()
rows, T_ = obj._resident_rows()


def timed(label, f, n=3):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        print(f"{label} run {i}: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
    return out


for nb in (32, 64, 256):
    timed(f"_scan_device B={nb}", lambda: obj._scan_device(q[:nb], rows, T_, k))
d, paths, idx = timed("shadow_device B=256", lambda: obj.shadow_device(q, k))
timed("_predict_device B=256", lambda: obj._predict_device(d, paths, rv, "softmax", 0.1))
timed("predict B=256 splits=1", lambda: obj.predict(q, k=k, to_predict=rv, eta=0.1))
timed("predict B=256 splits=8", lambda: obj.predict(q, k=k, to_predict=rv, eta=0.1, n_context_splits=8))
