"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Runs the unmodified reference `PathShadowing.shadow(..., cuda=False)` (path_shadowing.py:181)
and `realized_variance` (statistics.py:5) through oracle/ref_loader.py on seeded inputs and
stores inputs + outputs.  The fixtures pin oracle/ (tests/test_oracle.py) and, through it and
directly, the CUDA path (tests/test_gpu_parity.py).  Re-run: `python tests/gen_golden.py`.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

OUT = ROOT / "tests" / "golden"

# name: (R, T, W, H, k, B, n_splits, dataset_ndim, context_ndim)
CASES = {
    "cfg1_R128_T512_W20": (128, 512, 20, 20, 16, 3, 1, 3, 3),      # BASELINE.json configs[0]
    "w252_R96_T2048": (96, 2048, 252, 20, 256, 2, 3, 3, 3),        # north-star W, multi-split
    "nohorizon_R64_T1024": (64, 1024, 252, None, 128, 2, 4, 3, 3),  # PredictionContext(None)
    "ragged_R37_T513_W21": (37, 513, 21, 7, 50, 4, 1, 2, 2),       # odd sizes, 2-D inputs
    "single_T4096_W64": (1, 4096, 64, 32, 32, 1, 1, 1, 1),         # 1-D dataset and context
    "k_all_R3_T40_W8": (3, 40, 8, 4, 87, 2, 1, 3, 3),              # k == number of windows
}


def make_inputs(R, T, W, B, ds_ndim, q_ndim):
    g = torch.Generator().manual_seed(0)
    ds = torch.randn(R, 1, T, generator=g, dtype=torch.float32) * 0.01
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, 1, W, generator=g, dtype=torch.float32) * 0.01
    if ds_ndim == 2:
        ds = ds[:, 0, :]
    if ds_ndim == 1:
        ds = ds[0, 0, :]
    if q_ndim == 2:
        q = q[:, 0, :]
    if q_ndim == 1:
        q = q[0, 0, :]
    return ds.contiguous(), q.contiguous()


def main():
    torch.set_num_threads(8)
    ref = ref_loader.load()
    PS = ref.path_shadowing.PathShadowing
    OUT.mkdir(parents=True, exist_ok=True)
    for name, (R, T, W, H, k, B, ns, dnd, qnd) in CASES.items():
        ds, q = make_inputs(R, T, W, B, dnd, qnd)
        obj = PS(ref.path_embedding.Identity(W), ref.path_distance.RelativeMSE(), ds,
                 ref.path_embedding.PredictionContext(H))
        d, paths, idx = obj.shadow(q, k=k, n_splits=ns, cuda=False)
        # the reference's norm of each query (path_distance.py:65), to pin the 8-lane restatement
        qn = torch.as_tensor(q).reshape(B, W).norm(dim=-1).numpy()
        Ts = [T_ for T_ in (2, 5, 10, 20) if T_ <= (H or W)]
        out_ctx = obj.context.select_out_context(paths)
        rv = ref.statistics.realized_variance(out_ctx, Ts, vol=False)
        rvol = ref.statistics.realized_variance(out_ctx, Ts, vol=True)
        np.savez(OUT / f"{name}.npz", dataset=ds.numpy(), x_context=q.numpy(),
                 meta=np.array([R, T, W, -1 if H is None else H, k, B, ns], np.int64),
                 distances=d, paths=paths, indices=idx, qnorm=qn,
                 Ts=np.array(Ts, np.int64), rv=rv, rvol=rvol)
        print(name, d.shape, paths.shape, idx.shape, d.dtype, paths.dtype, idx.dtype,
              "ties:", int((np.diff(d, axis=1) == 0).sum()))

    # testing.ipynb:43-53 -- RelativeMSE.forward_topk prefix/split invariance (its own assert)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(8, 34, generator=g)
    y = torch.randn(16, 64, 34, generator=g)
    dist = ref.path_distance.RelativeMSE()
    ds1, id1 = dist.forward_topk(x, y, k=32, n_splits=4)
    ds2, id2 = dist.forward_topk(x, y, k=64, n_splits=8)
    assert torch.equal(ds1, ds2[:, :32])
    np.savez(OUT / "forward_topk_B8_d34.npz", x=x.numpy(), y=y.numpy(), ds=ds2.numpy(), idces=id2.numpy())
    print("forward_topk", ds2.shape, id2.shape, id2.dtype)


if __name__ == "__main__":
    main()
