"""Embedded (Foveal) scan timing by batch size / k (GPU box; not collected by pytest)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import shadowing_b200 as sb
from shadowing_b200 import _lib

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
T, W, H = 4096, 126, 252
g = torch.Generator().manual_seed(0)
ds = torch.randn(R, 1, T, generator=g)
x = torch.randn(8, 1, W, generator=g)
for mode in ("fft", "exact"):
    obj = sb.PathShadowing(sb.Foveal(1.15, 0.9, W), sb.RelativeMSE(), ds, sb.PredictionContext(H), scan_mode=mode)
    rows, T_ = obj._resident_rows()
    for k in (1024, 10000):
        for B in (1, 2, 4, 8):
            obj._scan_device(x[:B], rows, T_, k)
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            t0 = time.perf_counter()
            for _ in range(3):
                obj._scan_device(x[:B], rows, T_, k)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            print(f"mode={mode} R={R} k={k} B={B}: {dt * 1e3:.2f} ms per call, {(_lib.launch_count() - n0) / 3:.1f} launches", flush=True)
