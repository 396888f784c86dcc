import sys, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
import bench, shadowing_b200 as sb
from shadowing_b200 import _lib
g = torch.Generator().manual_seed(1234); ds = torch.randn(bench.R_FULL, 1, bench.T, generator=g) * 0.01
qs = bench.make_queries(60); qp = qs.clone().pin_memory()
obj = sb.PathShadowing(sb.Identity(bench.W), sb.RelativeMSE(), ds, sb.PredictionContext(bench.H), device="cuda:0")
for i in range(5): obj.shadow(qp[i:i+1], k=1024)
torch.cuda.synchronize()
def T(f, n=30):
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(n): f(i)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
rows, Tt = obj._resident_rows()
qd = qs.cuda()
print("shadow e2e      ", T(lambda i: obj.shadow(qp[i:i+1], k=1024)))
print("shadow_device   ", T(lambda i: obj.shadow_device(qp[i:i+1], k=1024)))
print("_scan_device(dev q)", T(lambda i: obj._scan_device(qd[i:i+1], rows, Tt, 1024)))
mode, aux = obj._mode_and_aux(rows, Tt, bench.W, bench.H)
q2 = qd[:, 0, :].contiguous()
print("_lib.scan_topk  ", T(lambda i: _lib.scan_topk(rows, Tt, q2[i:i+1], bench.H, 1024, 0, mode, obj._workspace, aux)))
d, p, ix = obj.shadow_device(qp[0:1], k=1024)
print("gather          ", T(lambda i: _lib.gather_paths(rows, Tt, ix, 272, 0)))
print("h2d q           ", T(lambda i: qp[i:i+1][:, 0, :].to("cuda:0", non_blocking=True).contiguous()))
import cProfile, pstats, io
pr = cProfile.Profile()
pr.enable()
for i in range(200): obj.shadow(qp[i % 50:i % 50 + 1], k=1024)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:6000])
