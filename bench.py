#!/usr/bin/env python
"""bench.py -- the path-shadowing scan on B200: windows/sec on BASELINE.json configs[1]
(R=32768 x T=4096 synthetic Gaussian log-returns, W=252, H=20, k=1024, one query date per step).

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm: oracle port)

A "step" is one shadow scan of one query over the resident ensemble.  Prints ONE JSON line.
  value   : windows/s with the queries already in HBM: the K scans of the timed region are enqueued
            back to back through the C ABI (psh_scan_topk_f32 | PSH_FLAG_NOSYNC), alternating between
            two streams with their own workspaces (query i+1's prologue and main launch overlap query
            i's re-rank and select), and verified by ONE psh_scan_overflowed per workspace at the end
            -- no host round trip between queries (CUDA events on the caller's stream, which joins
            both streams before the closing event)
  e2e     : windows/s through PathShadowing.shadow() with HOST numpy in/out, one call per step (pinned
            H2D of the query, scan, gather, D2H of distances+paths+indices, one synchronisation)
  roofline: the scan kernels (seed + main launch) against the measured HBM peak
            (MEASURED_PEAKS.json) -- plus the FP32-issue roof
  cpu_baseline: the C oracle (port of the reference algorithm) on this box's host cores,
            bounded row sample.
  host_enqueue_ms_per_step: host time to enqueue one step (the device loop is GPU-bound while this
            stays below ms_per_step)
N > 1: each rank holds its own 32768-row shard (weak scaling, ensemble = N*32768 rows); the per-rank
top-k are exchanged over NVLink peer memory and merged by one kernel per step (NCCL all-gather +
merge kernel if the peer mapping is unavailable).
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

R_PER_GPU, T, W, H, K_NEIGH = 32768, 4096, 252, 20, 1024
TP = T - W - H + 1
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


def make_shard(rank: int):
    import torch
    g = torch.Generator().manual_seed(0 + rank)
    ds = torch.randn(R_PER_GPU, T, generator=g, dtype=torch.float32) * 0.01
    return ds


def make_queries(n: int):
    import torch
    g = torch.Generator().manual_seed(1)
    return torch.randn(n, 1, W, generator=g, dtype=torch.float32) * 0.01


def host_cores() -> int:
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the oracle is told the
    thread count explicitly, so the CPU arm always runs on all the cores it can use)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_run(ds_np: np.ndarray, q_np: np.ndarray, target_s: float = 12.0):
    """C oracle (all host threads) on a bounded sample of rows of the same workload."""
    from oracle import oracle
    cores = host_cores()
    rows = min(256 * max(cores // 8, 1), ds_np.shape[0])
    t0 = time.perf_counter()
    oracle.shadow_topk(ds_np[:rows], q_np, K_NEIGH, H, nthreads=cores)
    dt = time.perf_counter() - t0
    rate = rows * TP / dt
    rows2 = int(min(ds_np.shape[0], max(rows, rate * target_s / TP)))
    t0 = time.perf_counter()
    oracle.shadow_topk(ds_np[:rows2], q_np, K_NEIGH, H, nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": rows2 * TP / dt, "unit": "windows/s", "cores": cores, "kind": "port",
            "sample": f"first {rows2} of {ds_np.shape[0]} rows x T={T}, one query, W={W}, k={K_NEIGH}: "
                      f"{rows2 * TP} windows in {dt:.2f} s (oracle/shadow_oracle.c, OpenMP, {cores} threads)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cores = host_cores()
    ds = make_shard(0).numpy()
    qs = make_queries(args.steps + args.warmup).numpy()
    # bounded sample per step: ~2 s of CPU work, fixed across steps
    rows = min(256 * max(cores // 8, 1), R_PER_GPU)
    t0 = time.perf_counter()
    oracle.shadow_topk(ds[:rows], qs[0], K_NEIGH, H, nthreads=cores)
    rate = rows * TP / (time.perf_counter() - t0)
    rows = int(min(R_PER_GPU, max(rows, rate * 2.0 / TP)))
    for i in range(args.warmup):
        oracle.shadow_topk(ds[:rows], qs[i], K_NEIGH, H, nthreads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        oracle.shadow_topk(ds[:rows], qs[args.warmup + i], K_NEIGH, H, nthreads=cores)
    dt = time.perf_counter() - t0
    val = rows * TP * args.steps / dt
    sample = (f"each step = first {rows} of {R_PER_GPU} rows x T={T}, one query (oracle/shadow_oracle.c port of "
              f"path_shadowing.py:97-179, OpenMP {cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: R=32768xT=4096 Gaussian dlnx, W=252, H=20, k=1024, "
                               "Identity+RelativeMSE, one query per step", "sample_rows": rows},
        "cpu_baseline": {"value": val, "unit": "windows/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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

    ds_host = make_shard(rank)
    nq = args.steps + args.warmup
    qs_host = make_queries(nq)
    qs_pinned = qs_host.clone().pin_memory()
    obj = sb.PathShadowing(sb.Identity(W), sb.RelativeMSE(), ds_host, sb.PredictionContext(H), device=dev,
                           row_offset=rank * R_PER_GPU, process_group=pg, scan_mode=args.mode)
    rows, _ = obj._resident_rows()
    obj._pipe_streams = max(1, args.streams) if world == 1 else 1
    qs_dev = qs_host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        # enqueue only: the K scans of the timed region form one pipeline on the stream (no host
        # round trip between queries); `_check_pipeline` synchronises once and verifies that no
        # scan overflowed its candidate buffers (it would have to be repeated)
        return obj._scan_device(qs_dev[i:i + 1], rows, T, K_NEIGH, nosync=True)

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
    for i in range(args.steps):
        step_device(args.warmup + i)
    host_enqueue_ms = (time.perf_counter() - th0) * 1e3 / args.steps   # host time to enqueue one step
    obj._check_pipeline()
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0

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
    L.psh_profile_begin()
    for i in range(args.steps):
        step_device(args.warmup + i)
    ms_kind = (ctypes.c_double * 3)()
    n_kind = (ctypes.c_uint64 * 3)()
    L.psh_profile_end(ms_kind, n_kind, 3)
    obj._check_pipeline()
    torch.cuda.synchronize()

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0]), float(t[1])

    eff_mode = "fft" if args.mode == "auto" else args.mode
    if rank == 0:
        windows_per_step = world * R_PER_GPU * TP
        value = windows_per_step * args.steps / (ms_dev * 1e-3)
        e2e = windows_per_step * args.steps / (ms_e2e * 1e-3)
        hbm_peak, peak_kind, sm_max = peaks()
        scan_ms_per_step = ms_kind[0] / args.steps
        alg_bytes = R_PER_GPU * T * 4  # the shard streamed once per query pass (4.283 B/window)
        ach = alg_bytes / (scan_ms_per_step * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get("scan_dram_bytes_per_step")
            except Exception:
                traffic = None
        eff_mode = "fft" if args.mode == "auto" else args.mode
        # fp32 lane-ops per window actually issued by the scan flavour (exact: sub+mul+add per
        # element; filter: one FMA per element; fft: ~1.6 k FP instructions per thread per pair of
        # trajectories / 7650 windows, counted from SASS)
        flop_per_win = {"exact": ALG_FLOP_PER_WINDOW, "filter": W, "fft": 27}[eff_mode]
        sm_clk = (clocks.get("sm_mhz") or sm_max) * 1e6
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_rate = R_PER_GPU * TP * flop_per_win / (scan_ms_per_step * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: R=32768xT=4096 Gaussian dlnx per GPU, W=252, H=20, "
                                   "k=1024, Identity+RelativeMSE, one query date per step",
                       "rows_per_gpu": R_PER_GPU, "scan_mode": eff_mode,
                       "l2": "512 MiB shard per GPU > 126 MB L2 (inputs larger than L2)",
                       "dataset": "resident in HBM (uploaded once at construction)",
                       "timing": f"K enqueue-only scans pipelined on {pipe_streams} stream(s) + one overflow check (value); "
                                 "one synchronous shadow() per step (e2e)",
                       "parallelism": f"rows sharded x{world}, peer-memory all-gather fused with the merge of per-GPU top-k"},
            "e2e": {"value": e2e, "unit": "windows/s", "h2d_bytes_per_step": W * 4,
                    "d2h_bytes_per_step": int(d_np.nbytes + p_np.nbytes + i_np.nbytes),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": "scan kernels of one step (all chunk launches)",
                         "kernel_ms_per_step": scan_ms_per_step, "kernel_launches_per_step": n_kind[0] / args.steps,
                         "select_ms_per_step": ms_kind[1] / args.steps,
                         "merge_ms_per_step": ms_kind[2] / args.steps,
                         "alg_bytes_per_step": alg_bytes,
                         "fp32": {"note": "binding roof (SURVEY 8d): lane-ops/s vs SMs*128*clock",
                                  "flop_per_window": flop_per_win, "achieved_tlaneops": fp32_rate / 1e12,
                                  "peak_tlaneops_at_sampled_clock": n_sm * 128 * sm_clk / 1e12,
                                  "frac": fp32_rate / (n_sm * 128 * sm_clk)}},
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_run(ds_host.numpy(), qs_host[0].numpy())
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fft", "filter", "exact"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("PSH_STREAMS", "2")),
                    help="streams the pipelined device loop alternates between (N = 1 only; 1: the caller's stream)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
