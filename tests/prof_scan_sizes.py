"""Builder tool: scan-kernel time against the number of rows (slope = steady-state rate, intercept = the fixed
cost of a launch: tables, first copies, seeding rendezvous, tail).  usage: python tests/prof_scan_sizes.py"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from shadowing_b200 import _lib  # noqa: E402

T, W, H, K = 4096, 252, 20, 1024
L = _lib.lib()
g = torch.Generator().manual_seed(0)
q = (torch.randn(64, W, generator=g) * 0.01).cuda()
for R in [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384, 32768, 65536]:
    rows = (torch.randn(R, T, generator=g) * 0.01).cuda()
    aux = _lib.fft_prepare(rows, T, W, H)
    ws = None
    for i in range(3):
        d, idx, ws = _lib.scan_topk(rows, T, q[i:i + 1], H, K, 0, _lib.PSH_MODE_FFT, ws, aux)
    torch.cuda.synchronize()
    n = 20
    L.psh_profile_begin()
    for i in range(n):
        d, idx, ws = _lib.scan_topk(rows, T, q[3 + i:4 + i], H, K, 0, _lib.PSH_MODE_FFT, ws, aux)
    ms = (ctypes.c_double * 3)()
    cnt = (ctypes.c_uint64 * 3)()
    L.psh_profile_end(ms, cnt, 3)
    torch.cuda.synchronize()
    print(f"R={R:6d}  scan {ms[0] / n:.4f} ms ({cnt[0] // n} launches)  select/rerank {ms[1] / n:.4f} ms ({cnt[1] // n})", flush=True)
    del rows, aux, ws
