"""One GPU's share of BASELINE configs[3] (R=262144 x T=8192 over 8 GPUs -> 32768 x 8192 per GPU,
1 GiB): times PathShadowing.shadow per scan flavour and checks the result against the CPU oracle
(not collected by pytest; run on a GPU box)."""
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
    R, T, W, H, k = 32768, 8192, 252, 20, 1024
    g = torch.Generator().manual_seed(3)
    ds = torch.randn(R, 1, T, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(4)
    q = torch.randn(8, 1, W, generator=g, dtype=torch.float32) * 0.01
    do, io = oracle.shadow_topk(ds.numpy(), q[:1].numpy(), k, H)
    for mode in ("fft", "filter"):
        obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds, sb.PredictionContext(H), scan_mode=mode)
        d, paths, idx = obj.shadow(q[:1], k=k)
        ok = np.array_equal(d.view(np.uint32), do.view(np.uint32)) and np.array_equal(idx, io)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(8):
            obj.shadow(q[i:i + 1], k=k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 8
        print(f"mode={mode}: {dt * 1e3:.3f} ms per query e2e -> {R * (T - W - H + 1) / dt:.3e} windows/s, parity {'OK' if ok else 'MISMATCH'}")
        del obj
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
