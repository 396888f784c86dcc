#!/usr/bin/env python
"""bench.py -- the path-shadowing scan on B200: windows/sec on BASELINE.json configs[1]
(R=32768 x T=4096 synthetic Gaussian log-returns, W=252, H=20, k=1024, one query date per step).

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)
  python bench.py --scaling strong ...                     (the FIXED 32768-row ensemble split over the N ranks)
  python bench.py --config cfg3 | cfg4 ...                 (BASELINE configs[2] / configs[3]; extra lines for profiles/)

A "step" is one shadow scan of one query over the resident ensemble.  Prints ONE JSON line.
  value   : windows/s with the queries already in HBM: the K scans of the timed region are enqueued
            back to back through the C ABI (psh_scan_topk_f32 | PSH_FLAG_NOSYNC), alternating between
            three streams with their own workspaces (query i+1's preparation and scan overlap query i's
            re-rank, select and exchange; the scans leave a few SMs to those small kernels) at EVERY N, and verified by ONE overflow check per workspace at
            the end -- no host round trip between queries (CUDA events on the caller's stream, which joins
            both streams before the closing event)
  e2e     : windows/s through PathShadowing.shadow() with HOST numpy in/out, one call per step (pinned
            H2D of the query, scan, gather, D2H of distances+paths+indices, one synchronisation)
  roofline: the scan kernel (ONE launch per query: it seeds its own threshold) against the measured HBM
            peak (MEASURED_PEAKS.json); algorithmic bytes = the shard streamed once
  cpu_baseline: the reference's own `cuda=False` path (unmodified package from baseline/_ref, torch on all
            host threads) on a stated subsample -- kind "reference" -- with the C oracle port beside it
  parity_checked: step 0's (distances, indices) of the timed pipeline compared bit for bit with the CPU
            oracle: at N = 1 against a full scan of the shard; at N > 1 every rank checks the merged
            result restricted to ITS rows against a full scan of its shard (together: the whole result)
  host_enqueue_ms_per_step: host time to enqueue one step (the device loop is GPU-bound while this
            stays below ms_per_step)
N > 1 (default, weak scaling): each rank holds its own 32768-row shard (ensemble = N*32768 rows); the
per-rank top-k are exchanged over NVLink peer memory and merged by one kernel per step (NCCL all-gather
+ merge kernel if the peer mapping is unavailable).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

R_FULL, T, W, H, K_NEIGH = 32768, 4096, 252, 20, 1024
ALG_FLOP_PER_WINDOW = 3 * W           # exact mode: sub, mul, add per element
METRIC = "shadowing windows/sec (R=32768xT=4096, W=252, k=1024)"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md's clocks
    line), through NVML in-process (10 ms period); falls back to polling nvidia-smi."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._halt = threading.Event()
        self.period = float(os.environ.get("BENCH_CLOCK_PERIOD_S", "0.01"))
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = int(get(self.h))
        for name, const in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                            ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                            ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                            ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
            if bits & int(getattr(n, const, 0)):
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 8:
            self.sm.append(int(float(f[1])))
            self.mx.append(int(float(f[2])))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(self.period if self.nvml is not None else max(self.period, 0.1))

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(sm), "power_w_max": max(self.power) if self.power else None,
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# workloads (synthetic, SURVEY.md section 8d: CPU-generated so that oracle and GPU see identical bits)
# ---------------------------------------------------------------------------------------------
def workload(args, world: int):
    """(rows per rank, T, description).  cfg2 weak: 32768 rows per rank; cfg2 strong: 32768 rows in total;
    cfg4: 262144 rows x 8192 in total (BASELINE configs[3])."""
    if args.config == "cfg4":
        if 262144 % world:
            raise SystemExit("cfg4 needs a rank count that divides 262144")
        return 262144 // world, 8192, "BASELINE configs[3]: R=262144xT=8192 sharded over the ranks"
    if args.scaling == "strong":
        if R_FULL % world:
            raise SystemExit("strong scaling needs a rank count that divides 32768")
        return R_FULL // world, T, "BASELINE configs[1], the FIXED ensemble R=32768xT=4096 split over the ranks"
    return R_FULL, T, "BASELINE configs[1]: R=32768xT=4096 per GPU"


def make_rows(args, rank: int, world: int, rows: int, t_len: int):
    import torch
    if args.config != "cfg4" and args.scaling == "strong":
        g = torch.Generator().manual_seed(0)       # the N = 1 ensemble; this rank keeps its block of rows
        full = torch.randn(R_FULL, t_len, generator=g, dtype=torch.float32) * 0.01
        return full[rank * rows:(rank + 1) * rows].clone()
    g = torch.Generator().manual_seed(0 + rank)    # per-rank generation with seed 0 + rank
    return torch.randn(rows, t_len, generator=g, dtype=torch.float32) * 0.01


def make_queries(n: int):
    import torch
    g = torch.Generator().manual_seed(1)
    return torch.randn(n, 1, W, generator=g, dtype=torch.float32) * 0.01


def host_cores() -> int:
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arms are told the
    thread count explicitly, so they always run on all the cores they can use)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ---------------------------------------------------------------------------------------------
# CPU arms: the reference's own cuda=False path (when its package is present) and the C oracle port
# ---------------------------------------------------------------------------------------------
def reference_available() -> bool:
    from oracle import ref_loader
    return ref_loader.available()


def time_real_reference(ds_np: np.ndarray, q_np: np.ndarray, rows: int, n_splits: int, repeats: int = 1):
    """The UNMODIFIED reference (RudyMorel/shadowing, PathShadowing.shadow(cuda=False),
    path_shadowing.py:181) on the first `rows` trajectories; returns (seconds per call, windows, cores)."""
    import torch
    from oracle import ref_loader
    ref = ref_loader.load()
    cores = host_cores()
    torch.set_num_threads(cores)
    t_len = ds_np.shape[-1]
    sub = np.ascontiguousarray(ds_np[:rows]).reshape(rows, 1, t_len)
    obj = ref.path_shadowing.PathShadowing(ref.path_embedding.Identity(W), ref.path_distance.RelativeMSE(), sub,
                                           ref.path_embedding.PredictionContext(H))
    k = min(K_NEIGH, (rows // n_splits) * (t_len - W - H + 1))
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        obj.shadow(q_np.reshape(1, 1, W), k=k, n_splits=n_splits, cuda=False)
        best = min(best, time.perf_counter() - t0)
    return best, rows * (t_len - W - H + 1), cores, k


def time_port(ds_np: np.ndarray, q_np: np.ndarray, rows: int):
    from oracle import oracle
    cores = host_cores()
    t_len = ds_np.shape[-1]
    t0 = time.perf_counter()
    oracle.shadow_topk(ds_np[:rows], q_np, K_NEIGH, H, nthreads=cores)
    return time.perf_counter() - t0, rows * (t_len - W - H + 1), cores


def cpu_baseline_run(ds_np: np.ndarray, q_np: np.ndarray):
    """Bounded CPU sample of the same workload (10-30 s): the real reference on a row subsample when its
    package is present (kind "reference"), the C oracle port on all rows next to it."""
    dt_p, win_p, cores = time_port(ds_np, q_np, ds_np.shape[0])
    port = {"value": win_p / dt_p, "unit": "windows/s", "cores": cores,
            "sample": f"all {ds_np.shape[0]} rows x T={ds_np.shape[-1]}, one query: {win_p} windows in {dt_p:.2f} s "
                      f"(oracle/shadow_oracle.c, OpenMP, {cores} threads)"}
    if reference_available():
        try:
            time_real_reference(ds_np, q_np, 64, 1)                        # warm-up: imports, thread pool
            dt, win, cores, k = time_real_reference(ds_np, q_np, 2048, 16)
            return {"value": win / dt, "unit": "windows/s", "cores": cores, "kind": "reference",
                    "sample": f"UNMODIFIED reference PathShadowing.shadow(cuda=False, n_splits=16), torch {cores} threads, "
                              f"first 2048 of {ds_np.shape[0]} rows x T={ds_np.shape[-1]}, one query, W={W}, k={k}: "
                              f"{win} windows in {dt:.2f} s", "port": port}
        except Exception as e:   # the reference package is there but cannot run (missing dependency ...)
            port["reference_error"] = repr(e)[:200]
    port["kind"] = "port"
    return port


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = host_cores()
    g = torch.Generator().manual_seed(0)
    ds = (torch.randn(R_FULL, T, generator=g, dtype=torch.float32) * 0.01).numpy()
    qs = make_queries(args.steps + args.warmup).numpy()
    tp = T - W - H + 1
    if reference_available():
        # the reference's own code, bounded sample per step (~2-4 s on 16 threads)
        rows, n_splits = 1024, 8
        time_real_reference(ds, qs[0], 64, 1)
        for i in range(args.warmup):
            time_real_reference(ds, qs[i], rows, n_splits)
        t0 = time.perf_counter()
        for i in range(args.steps):
            _, win, cores, k = time_real_reference(ds, qs[args.warmup + i], rows, n_splits)
        dt = time.perf_counter() - t0
        kind = "reference"
        sample = (f"each step = UNMODIFIED reference PathShadowing.shadow(cuda=False, n_splits={n_splits}) on the first "
                  f"{rows} of {R_FULL} rows x T={T}, one query, k={k}, torch {cores} threads (path_shadowing.py:181)")
    else:
        from oracle import oracle
        rows = min(256 * max(cores // 8, 1), R_FULL)
        t0 = time.perf_counter()
        oracle.shadow_topk(ds[:rows], qs[0], K_NEIGH, H, nthreads=cores)
        rate = rows * tp / (time.perf_counter() - t0)
        rows = int(min(R_FULL, max(rows, rate * 2.0 / tp)))
        for i in range(args.warmup):
            oracle.shadow_topk(ds[:rows], qs[i], K_NEIGH, H, nthreads=cores)
        t0 = time.perf_counter()
        for i in range(args.steps):
            oracle.shadow_topk(ds[:rows], qs[args.warmup + i], K_NEIGH, H, nthreads=cores)
        dt = time.perf_counter() - t0
        kind = "port"
        sample = (f"each step = first {rows} of {R_FULL} rows x T={T}, one query (oracle/shadow_oracle.c port of "
                  f"path_shadowing.py:97-179, OpenMP {cores} threads; the reference package is not installed)")
    val = rows * tp * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: R=32768xT=4096 Gaussian dlnx, W=252, H=20, k=1024, "
                               "Identity+RelativeMSE, one query per step", "sample_rows": rows,
                   "arm": "the reference's own cuda=False path" if kind == "reference" else "C/OpenMP port of the reference"},
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def check_parity(dist_dev, idx_dev, ds_np, q_np, rank: int, world: int, rows: int, tp: int):
    """Bit-exact comparison of a merged (dist, idx) with the oracle, restricted to this rank's rows:
    every returned record whose trajectory lives here must be in the oracle's scan of the local shard with
    the same distance bits, and every local record ordered before the k-th merged record must have been
    returned.  Over all ranks this covers the whole result."""
    from oracle import oracle
    d = dist_dev.detach().cpu().numpy()[0]
    idx = idx_dev.detach().cpu().numpy()[0]
    lo = rank * rows
    k_loc = min(K_NEIGH, rows * tp)
    do, io = oracle.shadow_topk(ds_np, q_np, k_loc, H, row_offset=lo, nthreads=host_cores())
    do, io = do[0], io[0]
    if world == 1:
        return bool(np.array_equal(d.view(np.uint32), do.view(np.uint32)) and np.array_equal(idx, io))
    flat = lambda ii: ii[:, 0].astype(np.int64) * tp + ii[:, 1].astype(np.int64)
    dk, fk = int(d[-1:].view(np.uint32)[0]), int(flat(idx[-1:])[0])          # the k-th merged record
    mine = (idx[:, 0] >= lo) & (idx[:, 0] < lo + rows)
    got = set(zip(d[mine].view(np.uint32).tolist(), map(tuple, idx[mine].tolist())))
    dbo, fo = do.view(np.uint32).astype(np.int64), flat(io)
    before = (dbo < dk) | ((dbo == dk) & (fo <= fk))                          # local records ordered up to it
    want = set(zip(do[before].view(np.uint32).tolist(), map(tuple, io[before].tolist())))
    return bool(got == want and (np.diff(d.view(np.uint32).astype(np.int64)) >= 0).all())


def run_ours(args):
    import torch
    import torch.distributed as dist
    import shadowing_b200 as sb
    from shadowing_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    rows_n, t_len, wl_name = workload(args, world)
    tp = t_len - W - H + 1
    ds_host = make_rows(args, rank, world, rows_n, t_len)
    nq = args.steps + args.warmup
    qs_host = make_queries(max(nq, 256))
    qs_pinned = qs_host.clone().pin_memory()
    obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds_host, sb.PredictionContext(H), device=dev,
                           row_offset=rank * rows_n, process_group=pg, scan_mode=args.mode)
    rows, _ = obj._resident_rows()
    obj._pipe_streams = max(1, args.streams)       # the same pipeline at every N
    qs_dev = qs_host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "cfg3":
        return run_cfg3(args, obj, qs_host, ds_host, rank, world, rows_n, tp, barrier, dev)

    def step_device(i):
        # enqueue only: the K scans of the timed region form one pipeline on the stream(s) (no host
        # round trip between queries); `_check_pipeline` synchronises once and verifies that no
        # scan overflowed its candidate buffers (it would have to be repeated)
        return obj._scan_device(qs_dev[i:i + 1], rows, t_len, K_NEIGH, nosync=True)

    def step_e2e(i):
        return obj.shadow(qs_pinned[i:i + 1], k=K_NEIGH)

    # ---------------- device-resident timing (value) ----------------
    for i in range(args.warmup):
        step_device(i)
    obj._check_pipeline()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    th0 = time.perf_counter()
    first = None
    for i in range(args.steps):
        out = step_device(args.warmup + i)
        if i == 0:
            first = out
    host_enqueue_ms = (time.perf_counter() - th0) * 1e3 / args.steps   # host time to enqueue one step
    obj._check_pipeline()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0

    # ---------------- parity of the timed pipeline's first step (every rank, its own rows) ----------------
    ok = check_parity(first[0], first[1], ds_host.numpy(), qs_host[args.warmup].numpy(), rank, world, rows_n, tp)

    # ---------------- end-to-end timing (host buffers in and out) ----------------
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        d_np, p_np, i_np = step_e2e(args.warmup + i)
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 if world == 1 else 0.0)
    clocks = sampler.stop()

    # ---------------- per-kernel timing (roofline) ----------------
    # one stream: the roofline wants each kernel's own duration, not its duration next to the
    # neighbouring query's kernels
    pipe_streams = obj._pipe_streams
    obj._pipe_streams = 1
    L = _lib.lib()
    for i in range(2):
        step_device(i)
    obj._check_pipeline()
    L.psh_profile_begin()
    for i in range(args.steps):
        step_device(args.warmup + i)
    ms_kind = (ctypes.c_double * 3)()
    n_kind = (ctypes.c_uint64 * 3)()
    L.psh_profile_end(ms_kind, n_kind, 3)
    obj._check_pipeline()
    torch.cuda.synchronize()

    t = torch.tensor([ms_dev, ms_e2e, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, bad = float(t[0]), float(t[1]), float(t[2])

    eff_mode = "fft" if args.mode == "auto" else args.mode
    if rank == 0:
        windows_per_step = world * rows_n * tp
        value = windows_per_step * args.steps / (ms_dev * 1e-3)
        e2e = windows_per_step * args.steps / (ms_e2e * 1e-3)
        hbm_peak, peak_kind, sm_max = peaks()
        scan_ms_per_step = ms_kind[0] / args.steps
        alg_bytes = rows_n * t_len * 4  # the shard streamed once per query pass (4.283 B/window at cfg2)
        ach = alg_bytes / (scan_ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists() and args.config == "cfg2" and rows_n == R_FULL:
            try:
                traffic = json.loads(tpath.read_text()).get("scan_dram_bytes_per_step")
            except Exception:
                traffic = None
        sm_clk = (clocks.get("sm_mhz") or sm_max) * 1e6
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        metric = METRIC if args.config == "cfg2" else f"shadowing windows/sec ({wl_name})"
        out = {
            "metric": metric, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.scaling == "strong" or args.config == "cfg4") else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "parity_checked": bad == 0.0,
            "config": {"workload": f"{wl_name}, Gaussian dlnx, W=252, H=20, k=1024, Identity+RelativeMSE, one query date per step",
                       "rows_per_gpu": rows_n, "T": t_len, "scan_mode": eff_mode,
                       "l2": f"{rows_n * t_len * 4 >> 20} MiB shard per GPU vs 126 MB L2 (every query streams the whole shard)",
                       "dataset": "resident in HBM (uploaded once at construction)",
                       "timing": f"K enqueue-only scans pipelined on {pipe_streams} stream(s) + one overflow check (value); "
                                 "one synchronous shadow() per step (e2e)",
                       "parity": "step 0 of the timed pipeline vs the C oracle, bit-exact, every rank on its own rows",
                       "parallelism": f"rows sharded x{world}, peer-memory all-gather fused with the merge of per-GPU top-k"},
            "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": W * 4,
                    "d2h_bytes_per_step": int(d_np.nbytes + p_np.nbytes + i_np.nbytes),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": "fft_scan_warp_kernel (one warp per 1024-point transform; one launch per query, seeds its own threshold)",
                         "kernel_ms_per_step": scan_ms_per_step, "kernel_launches_per_step": n_kind[0] / args.steps,
                         "select_ms_per_step": ms_kind[1] / args.steps,
                         "merge_ms_per_step": ms_kind[2] / args.steps,
                         "alg_bytes_per_step": alg_bytes,
                         "note": "traffic: dram__bytes_read+write of the same kernel from the round's ncu --set full "
                                 "capture (profiles/traffic.json); achieved: algorithmic bytes / CUDA-event duration"},
        }
        if world == 1 and not args.no_cpu and args.config == "cfg2":
            out["cpu_baseline"] = cpu_baseline_run(ds_host.numpy(), qs_host[0].numpy())
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_cfg3(args, obj, qs_host, ds_host, rank, world, rows_n, tp, barrier, dev):
    """BASELINE configs[2]: 256 query dates per step, predict_from_paths realised variance Ts=[5,10,20],
    softmax eta=0.1, everything on the device (only the (256, 3) predictions leave it)."""
    import torch
    import shadowing_b200 as sb
    from oracle import oracle
    B = 256
    q = qs_host[:B]
    rv = sb.RealizedVariance([5, 10, 20], vol=False)
    for _ in range(max(1, min(args.warmup, 3))):
        pred, pstd = obj.predict(q, k=K_NEIGH, to_predict=rv, eta=0.1, proba_name="softmax")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = max(1, min(args.steps, 5))
    for _ in range(steps):
        pred, pstd = obj.predict(q, k=K_NEIGH, to_predict=rv, eta=0.1, proba_name="softmax")
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    # sampled parity: 4 queries against full oracle scans, their predictions against the numpy restatement
    ok = True
    if world == 1:
        dsn = ds_host.numpy()
        for b in (0, 31, 32, 255):
            d, paths, idx = obj.shadow(q[b:b + 1], k=K_NEIGH)
            do, po, io = oracle.shadow(dsn.reshape(rows_n, 1, -1), q[b:b + 1].numpy(), K_NEIGH, H)
            mo, so = oracle.predict_from_paths(do, po, H, [5, 10, 20], False, "softmax", 0.1)
            ok = ok and np.array_equal(d.view(np.uint32), do.view(np.uint32)) and np.array_equal(idx, io)
            ok = ok and np.allclose(pred[b], mo[0], rtol=1e-6, atol=0) and np.allclose(pstd[b], so[0], rtol=1e-5, atol=0)
    if rank == 0:
        print(json.dumps({
            "metric": "shadowing query-windows/sec (BASELINE configs[2]: 256 query dates, R=32768xT=4096, W=252, k=1024, "
                      "predict_from_paths RV Ts=[5,10,20], softmax eta=0.1)",
            "value": B * world * rows_n * tp / (ms * 1e-3), "unit": "windows/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms, "ms_per_query": ms / B, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "parity_checked": bool(ok),
            "config": {"workload": "one PathShadowing.predict() call with 256 contexts per step; sampled queries vs the C oracle "
                                   "(indices bit-exact, predictions 1e-6)", "rows_per_gpu": rows_n}}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fft", "filter", "exact"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, the driver's metric): 32768 rows per GPU; strong: 32768 rows in total")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("PSH_STREAMS", "3")),
                    help="streams the pipelined device loop alternates between (1: the caller's stream)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
