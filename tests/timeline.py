"""Per-CTA (4096-point flavour) / per-warp (1024-point flavour) timeline of the fft scan kernel
(debug: PSH_FFT_DBG = device pointer of a (units, 8) uint64 buffer)."""
import os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import shadowing_b200 as sb
R, T, W, H, k = 32768, 4096, 252, 20, 1024
g = torch.Generator().manual_seed(0); ds = torch.randn(R, 1, T, generator=g) * 0.01
g = torch.Generator().manual_seed(1); q = torch.randn(8, 1, W, generator=g) * 0.01
obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H))
print("env", {k: v for k, v in os.environ.items() if k.startswith("PSH_")})
for i in range(3): obj.shadow(q[i:i+1], k=k)
dbg = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
os.environ["PSH_FFT_DBG"] = str(dbg.data_ptr())
obj.shadow(q[4:5], k=k); torch.cuda.synchronize()
os.environ.pop("PSH_FFT_DBG")
t = dbg.cpu().numpy().reshape(-1, 8); t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
rel = (t - t0) / 1e3
names = (["start", "seed_epi_done", "seed_wait_done", "iter0_end", "iter1_end", "iter8_end", "iter24_end", "end"] if len(t) > 1024 else
         ["start", "seed_epi_done", "seed_wait_done", "seed_refresh_done", "iter0_end", "iter1_end", "iter8_end", "end"])
print("units (CTAs / warps):", len(t))
for j, n in enumerate(names):
    c = rel[:, j]
    c = c[c >= 0]
    print(f"{n:18s} min {c.min():8.2f} median {np.median(c):8.2f} max {c.max():8.2f} us  (n={len(c)})")
