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

    # Foveal embedding (path_embedding.py:142-172) on the configuration of the reference's own
    # self-consistency cell, testing.ipynb:62-78: x_context randn(8,1,126), dataset randn(32,1,4096),
    # Foveal(1.15, 0.9, 126), PredictionContext(252), k=1024.  Paths are not stored (12 MB): they
    # are dataset[r, t:t+W+H] of the stored indices.
    g = torch.Generator().manual_seed(10)
    ds = torch.randn(32, 1, 4096, generator=g)
    x = torch.randn(8, 1, 126, generator=g)
    emb = ref.path_embedding.Foveal(1.15, 0.9, 126)
    obj = PS(emb, ref.path_distance.RelativeMSE(), ds, ref.path_embedding.PredictionContext(252))
    d, paths, idx = obj.shadow(x, k=1024, n_splits=4, cuda=False)
    assert np.array_equal(paths[3, 7, 0], ds.numpy()[idx[3, 7, 0], 0, idx[3, 7, 1]:idx[3, 7, 1] + 378])
    np.savez_compressed(OUT / "foveal_R32_T4096_W126.npz", dataset=ds.numpy(), x_context=x.numpy(),
                        kernel=emb.kernel.numpy(), ex=emb(x)[:, 0, :].numpy(), distances=d, indices=idx,
                        meta=np.array([32, 4096, 126, 252, 1024, 8, 4], np.int64),
                        foveal=np.array([1.15, 0.9, 126.0]))
    print("foveal", d.shape, idx.shape, emb.kernel.shape)

    # a generic dense PathEmbedding(kernel) (path_embedding.py:117-132): 5 random rows of 16 taps
    g = torch.Generator().manual_seed(11)
    ds = torch.randn(16, 1, 300, generator=g) * 0.01
    x = torch.randn(3, 1, 16, generator=g) * 0.01
    emb = ref.path_embedding.PathEmbedding(torch.randn(5, 1, 16, generator=g))
    obj = PS(emb, ref.path_distance.RelativeMSE(), ds, ref.path_embedding.PredictionContext(4))
    d, paths, idx = obj.shadow(x, k=64, n_splits=2, cuda=False)
    np.savez_compressed(OUT / "dense_kernel_R16_T300_W16.npz", dataset=ds.numpy(), x_context=x.numpy(),
                        kernel=emb.kernel.numpy(), ex=emb(x)[:, 0, :].numpy(), distances=d, indices=idx,
                        paths=paths, meta=np.array([16, 300, 16, 4, 64, 3, 2], np.int64))
    print("dense kernel", d.shape, idx.shape)

    # ImputationContext (path_embedding.py:59-87) and CrossChannelContext (:90-114): `shadow` only -- the
    # reference's `slect_out_context` typo makes `predict` unusable with ImputationContext upstream
    g = torch.Generator().manual_seed(20)
    ds = torch.randn(24, 1, 700, generator=g) * 0.01
    l, c, r = 30, 12, 20
    x = torch.randn(3, 1, l + r, generator=g) * 0.01
    obj = PS(ref.path_embedding.Identity(l + r), ref.path_distance.RelativeMSE(), ds,
             ref.path_embedding.ImputationContext((l, c, r)))
    d, paths, idx = obj.shadow(x, k=40, n_splits=2, cuda=False)
    emb = ref.path_embedding.Foveal(1.15, 0.9, l + r)
    objf = PS(emb, ref.path_distance.RelativeMSE(), ds, ref.path_embedding.ImputationContext((l, c, r)))
    df, pathsf, idxf = objf.shadow(x, k=40, n_splits=1, cuda=False)
    np.savez_compressed(OUT / "imputation_R24_T700.npz", dataset=ds.numpy(), x_context=x.numpy(),
                        portion=np.array([l, c, r], np.int64), distances=d, paths=paths, indices=idx,
                        foveal=np.array([1.15, 0.9, l + r]), foveal_distances=df, foveal_indices=idxf)
    print("imputation", d.shape, paths.shape, idx.shape, df.shape)

    g = torch.Generator().manual_seed(21)
    ds = torch.randn(16, 3, 500, generator=g) * 0.01
    x = torch.randn(2, 1, 25, generator=g) * 0.01
    obj = PS(ref.path_embedding.Identity(25), ref.path_distance.RelativeMSE(), ds,
             ref.path_embedding.CrossChannelContext(2))
    d, paths, idx = obj.shadow(x, k=30, n_splits=1, cuda=False)
    np.savez_compressed(OUT / "crosschannel_R16_C3_T500.npz", dataset=ds.numpy(), x_context=x.numpy(),
                        out_channels=np.array([2], np.int64), distances=d, paths=paths, indices=idx)
    print("cross-channel", d.shape, paths.shape, idx.shape)


if __name__ == "__main__":
    main()
