#!/usr/bin/env python
"""bench.py - BASELINE.json metric: MPC QP solves/sec (Go2, horizon 10, batch N) on 1/2/4/8 B200.

Workload (config.workload): BASELINE.json configs[1] = batched condensed Go2 trot MPC, horizon 10, 4096 randomised
states per GPU, ADMM kernel.  One "step" = one pass of the hot path (schedule->QP assembly->solve->GRFs) over one
batch.  `value` counts only CONVERGED problems (status 0 and all four KKT inf-norms <= 1e-6, FP64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--horizon H]
For N > 1 launch with torchrun (one rank per GPU); problems are sharded as independent slices, no collective on the
solve path (weak scaling: every rank gets its own batch of B problems).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KKT_TOL = 1e-6
# BASELINE.json configs: [1] is the configuration the metric is quoted on (default); [2] is the Riccati config.
WORKLOADS = {
    "config2": dict(name="go2_trot_condensed_mpc_h10_admm", horizon=10, batch=4096, gaits=("TROT",), solver="CONDENSED_ADMM",
                    seed=20261019 + 2, kernel="mpc_admm_kernel<64,60>"),
    "config3": dict(name="go2_mixed_sparse_mpc_h16_riccati", horizon=16, batch=16384, gaits=("TROT", "PACE", "BOUND"),
                    solver="SPARSE_RICCATI", seed=20261019 + 3, kernel="mpc_riccati_kernel<64>"),
    "config5": dict(name="go2_1M_mixed_sweep_sharded", horizon=0, batch=1048576, gaits=(), solver="ADMM(N=10 trot) + AUTO(N=16 mixed: condensed kernel where n <= 156, Riccati otherwise)",
                    seed=20261019 + 5, kernel="mpc_admm_kernel<64,60> + mpc_riccati_kernel<64>"),
    "config4": dict(name="go2_wbc_reduced_tsid_qp", horizon=10, batch=16384, gaits=("TROT",), solver="WBC_DENSE_IPM",
                    seed=20261019 + 4, kernel="wbc_qp_kernel"),
}


def algorithmic_bytes(N):
    """minimal MPC-level I/O per solve, FP64: inputs + contact bytes + GRFs + predicted states + info."""
    return 8 * (26 + 13 * (N + 1) + 12 * N) + 4 * N + 8 * (12 * N + 13 * (N + 1)) + 48


def algorithmic_flops(N, n_free, iters, rounds):
    """FP64 FLOPs of the algorithm as implemented: closed-form assembly + Gauss-Jordan K^-1 (2 n^3) +
    ADMM iterations (2 n^2 + ~60 n each) + per polish round (reduced assembly 24 n^2, GJ 2 nr^3 ~ 2 (0.7 n)^3,
    three H mat-vecs 6 n^2)."""
    n = float(n_free)
    return 20 * n * n + 2 * n ** 3 + iters * (2 * n * n + 60 * n) + rounds * (30 * n * n + 2 * (0.7 * n) ** 3)


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML in-process, one
    sample every 2 ms (the timed region of the headline workload is tens of milliseconds - too short for nvidia-smi's
    100 ms loop, which stays as the fallback when the NVML binding is unavailable)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.stop, self.nvml = [], None, index, False, None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        try:
            mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self.stop:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append([str(sm), str(mx if mx is not None else sm), "0"] +
                                 ["Active" if rs & bit else "Not Active" for _, bit in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def sample_now(self):
        """One synchronous sample (called while the timed launches are queued and the GPU is busy: the polling thread can be
        starved by the launching thread in a region that only lasts tens of milliseconds)."""
        n = self.nvml
        if n is None:
            return
        try:
            sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
            rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append([str(sm), str(mx), "0"] + ["Active" if rs & b else "Not Active" for b in (0x8, 0x40, 0x20, 0x4)])
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.nvml is not None:
            self.stop = True
            self.t.join(timeout=2)
            return
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by scripts/summarise_profile.py); None if that kernel was not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------------------------------- reference arm
def cpu_reference(N, rec, contact, budget_s, threads=None):
    """Times the restated reference CPU path (oracle/: C port of the OSQP-style condensed solve with OpenMP when
    built, else the numpy port) on a bounded sample of the same workload.  Returns dict for `cpu_baseline`."""
    from oracle import cpu_baseline as cb

    return cb.time_sample(N, rec, contact, budget_s, threads)


def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return
    from dfki_quad_b200 import sequencer  # host-side input generator only
    from oracle import mpc_oracle as mo
    from oracle import cpu_baseline as cb

    wl = WORKLOADS[args.workload]
    N, B = args.horizon or wl["horizon"], args.batch or wl["batch"]
    batch = cb.make_batch_cpu(B, N, seed=wl["seed"], gaits=wl["gaits"])
    per_step_budget = 8.0
    total_solved, total_t = 0, 0.0
    res = None
    for s in range(args.warmup + args.steps):
        res = cb.time_sample(N, batch["rec"], batch["contact"], per_step_budget, None)
        if s >= args.warmup:
            total_solved += res["solved"]
            total_t += res["seconds"]
    value = total_solved / total_t
    line = {
        "impl": "reference", "metric": "MPC QP solves/sec (Go2, horizon 10, batch N)", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "horizon": N, "batch_per_gpu": B, "gait": "/".join(wl["gaits"]), "solver": res["solver"]},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------- B200 arm
def closed_loop_extra(N, n, dev, steps=30):
    """4096 trotting single-rigid-body robots with the sequencer kernel + QP solve in the loop, warm_start = 2 (the reference runs
    with mpc_warm_start = 1): device time of the solve kernels per control period.  Consecutive problems share their active
    set, so the previous set goes straight to the polish and ADMM is skipped."""
    from dfki_quad_b200 import BatchedMPC, rollout, sequencer

    cfg = sequencer.GO2_MPC
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", warm_start=2,
                     horizon=N, dt=cfg["dt"], max_batch=n, device=dev)
    mpc.set_model(**sequencer.go2_seq_model())
    tgt = dict(hybrid_x_dot=0.4, hybrid_y_dot=0.1, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    h = rollout.closed_loop(mpc, n, steps, tgt, gait="TROT", seed=3, yaw0=0.0)
    ms = float(np.mean(h["kernel_ms"][5:]))
    return {"value": n * float(h["ok"][5:].mean()) / (ms * 1e-3), "unit": "solves/s", "kernel_ms_per_period": ms, "robots": n,
            "control_periods": steps, "converged_fraction": float(h["ok"].mean()), "mean_admm_iterations": float(h["iters"][5:].mean()),
            "mean_polish_rounds": float(h["rounds"][5:].mean()),
            "note": "device time of the solve kernels only; plant integration is host-side numpy and not part of any metric"}


def run_b200(args):
    import ctypes

    import torch

    from dfki_quad_b200 import BatchedMPC, _lib, sequencer

    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    torch.cuda.set_device(dev)
    wl = WORKLOADS[args.workload]
    N, B = args.horizon or wl["horizon"], args.batch or wl["batch"]
    cfg = sequencer.GO2_MPC
    # every rank gets its own batch (weak scaling): seed offset by rank
    batch = sequencer.random_batch(B, N, seed=wl["seed"] + 1000 * rank, gaits=wl["gaits"], device=dev)
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], wl["solver"],
                     horizon=N, dt=cfg["dt"], max_batch=B, device=dev)
    ext = torch.cuda.ExternalStream(mpc.stream(), device=dev)
    rec_h = torch.from_numpy(batch["rec"]).pin_memory()
    ct_h = torch.from_numpy(batch["contact"]).pin_memory()
    with torch.cuda.stream(ext):
        d_in = rec_h.to(dev, non_blocking=True)
        d_ct = ct_h.to(dev, non_blocking=True)
        d_f = torch.zeros((B, N, 12), dtype=torch.float64, device=dev)
        d_x = torch.zeros((B, N + 1, 13), dtype=torch.float64, device=dev)
        d_info = torch.zeros((B, 48), dtype=torch.uint8, device=dev)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    ext.synchronize()

    def step_device():
        mpc.solve_device(B, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    evs = []
    kernel_ms = []
    with ClockSampler(dev) as clocks:
        barrier()
        for _ in range(args.steps):
            with torch.cuda.stream(ext):
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                step_device()
                e1.record(ext)
            evs.append((e0, e1))
            clocks.sample_now()
        barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = float(sum(step_ms))
    info = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=np.dtype(
        [("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))]))
    conv = (info["status"] == 0) & (info["kkt"].max(axis=1) <= KKT_TOL)
    n_conv = int(conv.sum())
    launches_per_step = mpc.last_timing()[1]

    # ---- end-to-end through the host-buffer C ABI call (`e2e`) ----------------------------------------------------
    # host buffers are page-locked (the contract's "pinned host memory"): the library then DMAs straight from / into them
    f_h = torch.zeros((B, N, 12), dtype=torch.float64).pin_memory().numpy()
    x_h = torch.zeros((B, N + 1, 13), dtype=torch.float64).pin_memory().numpy()
    rec_np, ct_np = rec_h.numpy(), ct_h.numpy()
    for _ in range(max(1, args.warmup // 2)):
        mpc.solve_records(rec_np, ct_np, f_h, x_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, inf = mpc.solve_records(rec_np, ct_np, f_h, x_h)  # returns after the results are in f_h / x_h / inf
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # same inputs and a deterministic kernel every step: the convergence count of the last step holds for all of them
    conv_e2e = args.steps * int((inf.success & (inf.kkt.max(axis=1) <= KKT_TOL)).sum())

    # ---- same, entering one level earlier: (state, Target, gait, phase) in, sequencer kernel on the device --------------
    mpc.set_model(**sequencer.go2_seq_model())
    seq_np = torch.from_numpy(batch["seq_in"]).pin_memory().numpy()
    for _ in range(max(1, args.warmup // 2)):
        mpc.sequence_and_solve(seq_np, None, f_h, x_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, inf = mpc.sequence_and_solve(seq_np, None, f_h, x_h)
    torch.cuda.synchronize(dev)
    seq_s = time.perf_counter() - t0
    conv_seq = args.steps * int((inf.success & (inf.kkt.max(axis=1) <= KKT_TOL)).sum())

    # ---- max over ranks, whole-job aggregate -----------------------------------------------------------------------
    stats = torch.tensor([total_ms, e2e_s, seq_s], dtype=torch.float64, device=dev)
    counts = torch.tensor([n_conv * args.steps, conv_e2e, B * args.steps, conv_seq], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_s_max, seq_s_max = stats.tolist()
    conv_all, conv_e2e_all, issued_all, conv_seq_all = counts.tolist()
    value = conv_all / (total_ms_max * 1e-3)
    e2e_value = conv_e2e_all / e2e_s_max

    if rank == 0:
        peak, peak_src = measured_peaks()
        try:
            fp64_peak = _lib.fp64_peak_tflops(dev)
        except Exception:
            fp64_peak = None
        traffic = profiled_traffic(wl["kernel"])
        ab = algorithmic_bytes(N)
        # dominant kernel = mpc_admm_kernel<64,60> (all trot problems land in the n<=60 bin); the step's event time
        # also contains the bin kernel and two empty-bin launches (a few microseconds)
        kern_s = (total_ms / args.steps) * 1e-3
        achieved = ab * B / kern_s / 1e9
        if wl["solver"] == "CONDENSED_ADMM":
            fl = float(sum(algorithmic_flops(N, r["n_free"], r["iterations"], r["polish_rounds"]) for r in info))
        else:  # Riccati IPM: per iteration N stages x ~9.6 kFLOP structured matrix sweep + 2 vector sweeps (~1.2 kFLOP each)
            fl = float(info["iterations"].sum()) * N * (9600.0 + 2 * 1200.0)
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cpu = cpu_reference(N, batch["rec"], batch["contact"], args.cpu_budget)
            except Exception as e:  # the CPU baseline is a reported extra, never a reason to lose the GPU line
                cpu = {"error": repr(e)}
        line = {
            "metric": "MPC QP solves/sec (Go2, horizon 10, batch N)", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "horizon": N, "batch_per_gpu": B, "gait": "/".join(wl["gaits"]), "solver": wl["solver"],
                       "l2": "256 MiB flush between timed iterations", "converged_fraction": conv_all / issued_all,
                       "mean_iterations": float(info["iterations"].mean()), "mean_polish_rounds": float(info["polish_rounds"].mean()),
                       "kkt_max": float(info["kkt"][conv].max()) if n_conv else None, "parallelism": f"shard{world}"},
            "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(B * (8 * (26 + 13 * (N + 1) + 12 * N) + 4 * N)),
                    "d2h_bytes_per_step": int(B * (8 * (12 * N + 13 * (N + 1)) + 48))},
            "e2e_from_sequencer_inputs": {
                "value": conv_seq_all / seq_s_max, "unit": "solves/s", "h2d_bytes_per_step": int(B * 8 * 50),
                "d2h_bytes_per_step": int(B * (8 * (12 * N + 13 * (N + 1)) + 48)),
                "note": "b200qp_mpc_sequence_solve_batch: gait sequencer (contact schedule, trajectory planner, Raibert "
                        "foot steps) runs as a kernel, host sends 400 B/robot"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": wl["kernel"],
                         "algorithmic_bytes_per_solve": ab,
                         "note": "latency/FP64-bound on-chip solve: HBM fraction is tiny by design (SURVEY.md 8d)"},
            "fp64": {"flops_per_step": fl, "achieved_tflops": fl / kern_s / 1e12, "measured_peak_tflops": fp64_peak,
                     "frac": fl / kern_s / 1e12 / fp64_peak if fp64_peak else None, "nominal_peak_tflops": 37.0,
                     "note": "algorithmic FP64 FLOPs of the method as implemented / step time; peak = DFMA microbenchmark on this GPU"},
        }
        if wl["solver"] == "CONDENSED_ADMM" and not args.no_closed_loop:
            try:  # reported extra, outside every timed region above: the same kernels in closed loop with hot start
                line["closed_loop_hot_start"] = closed_loop_extra(N, min(B, 4096), dev)
            except Exception as e:
                line["closed_loop_hot_start"] = {"error": repr(e)}
        if cpu is not None and "error" not in cpu:
            line["cpu_baseline"] = {"value": cpu["solved"] / cpu["seconds"], "unit": "solves/s", "cores": cpu["cores"],
                                    "kind": cpu["kind"], "sample": cpu["sample"]}
        elif cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_b200_wbc(args):
    """BASELINE.json configs[3]: batched WBC task-space QP (reported beside the headline metric, not instead of it)."""
    import torch

    from dfki_quad_b200 import wbc

    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    torch.cuda.set_device(dev)
    wl = WORKLOADS["config4"]
    B = args.batch or wl["batch"]
    b = wbc.random_wbc_batch(B, seed=wl["seed"] + 1000 * rank, device=dev)
    solver = wbc.BatchedWBCSolver(max_batch=B, device=dev)
    H, g, A, lo, hi = solver.abi_arrays(b["qp"])
    ext = torch.cuda.ExternalStream(solver.stream(), device=dev)
    with torch.cuda.stream(ext):
        dH, dg, dA, dlo, dhi = (torch.from_numpy(a).to(dev) for a in (H, g, A, lo, hi))
        dx = torch.zeros((B, 30), dtype=torch.float64, device=dev)
        dinfo = torch.zeros((B, 48), dtype=torch.uint8, device=dev)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ext.synchronize()

    def step():
        solver.solve_device(B, dH.data_ptr(), dg.data_ptr(), dA.data_ptr(), dlo.data_ptr(), dhi.data_ptr(), dx.data_ptr(),
                            dinfo.data_ptr(), sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    evs = []
    with ClockSampler(dev) as clocks:
        barrier()
        for _ in range(args.steps):
            with torch.cuda.stream(ext):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                step()
                e1.record(ext)
            evs.append((e0, e1))
            clocks.sample_now()
        barrier()
    total_ms = float(sum(a.elapsed_time(c) for a, c in evs))
    info = np.frombuffer(dinfo.cpu().numpy().tobytes(), dtype=np.dtype(
        [("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))]))
    conv = (info["status"] == 0) & (info["kkt"].max(axis=1) <= KKT_TOL)
    # e2e: page-locked host buffers (the contract's "pinned host memory"), so the library DMAs straight from / into them
    H, g, A, lo, hi = (torch.from_numpy(a).pin_memory().numpy() for a in (H, g, A, lo, hi))
    xh = torch.zeros((B, 30), dtype=torch.float64).pin_memory().numpy()
    solver.solve_abi(H, g, A, lo, hi, xh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, inf = solver.solve_abi(H, g, A, lo, hi, xh)  # returns after x and the info records are in host memory
    e2e_s = time.perf_counter() - t0
    # same inputs and a deterministic kernel every step: the convergence count of the last step holds for all of them
    conv_e2e = args.steps * int(((inf["status"] == 0) & (inf["kkt"].max(axis=1) <= KKT_TOL)).sum())
    stats = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    counts = torch.tensor([float(conv.sum()) * args.steps, conv_e2e, B * args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, peak_src = measured_peaks()
        ab = 8 * (900 + 30 + 46 * 30 + 2 * 46 + 30) + 48
        kern_s = total_ms / args.steps * 1e-3
        line = {
            "metric": "WBC task-space QP solves/sec (Go2 reduced TSID, 30 vars, 18 eq, 28 ineq)", "value": counts[0].item() / (stats[0].item() * 1e-3),
            "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": stats[0].item() / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "batch_per_gpu": B, "solver": wl["solver"], "l2": "256 MiB flush between timed iterations",
                       "converged_fraction": counts[0].item() / counts[2].item(), "mean_iterations": float(info["iterations"].mean()),
                       "kkt_max": float(info["kkt"][conv].max()), "parallelism": f"shard{world}"},
            "e2e": {"value": counts[1].item() / stats[1].item(), "unit": "solves/s", "h2d_bytes_per_step": int(B * (ab - 48 - 240) + B * 240),
                    "d2h_bytes_per_step": int(B * (240 + 48))},
            "gpu_launches": args.steps, "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": ab * B / kern_s / 1e9, "peak": peak, "unit": "GB/s", "frac": ab * B / kern_s / 1e9 / peak,
                         "traffic": None, "peak_source": peak_src, "kernel": wl["kernel"], "algorithmic_bytes_per_solve": ab},
        }
        if not args.no_cpu_baseline:
            from oracle import qp_oracle as qo

            t0 = time.perf_counter()
            k = 0
            qp = b["qp"]
            while time.perf_counter() - t0 < args.cpu_budget and k < B:
                qo.solve(qp.H[k], qp.g[k], qp.A[k, :18], qp.lower_y[k, :18], qp.A[k, 18:], qp.lower_y[k, 18:], qp.upper_y[k, 18:])
                k += 1
            line["cpu_baseline"] = {"value": k / (time.perf_counter() - t0), "unit": "solves/s", "cores": 1, "kind": "port",
                                    "sample": f"first {k} problems, numpy IPM + active-set oracle, 1 thread"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_b200_sweep(args):
    """BASELINE.json configs[4]: 1 048 576 problems = config-2 union config-3 distributions, cut into contiguous slices over
    the GPUs (STRONG scaling: the total is fixed), no collective on the solve path.  Each rank generates one 32 768-problem
    chunk per distribution and re-solves it for every chunk of its slice (input generation is host-side numpy and is not
    part of the metric); timing = CUDA events around every chunk solve, summed, max over ranks."""
    import torch

    from dfki_quad_b200 import BatchedMPC, sequencer
    from dfki_quad_b200.shard import shard_range

    rank, local, world = dist_env()
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local
    torch.cuda.set_device(dev)
    wl = WORKLOADS["config5"]
    total = args.batch or wl["batch"]
    lo, hi = shard_range(total, rank, world)
    mine = hi - lo
    chunk = 32768
    cfg = sequencer.GO2_MPC
    parts = []
    for tag, N, gaits, solver, seed in (("admm", 10, ("TROT",), "CONDENSED_ADMM", wl["seed"]), ("auto16", 16, ("TROT", "PACE", "BOUND"), "AUTO", wl["seed"] + 1)):
        b = sequencer.random_batch(chunk, N, seed=seed + 1000 * rank, gaits=gaits, device=dev)
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"],
                         max_batch=chunk, device=dev)
        ext = torch.cuda.ExternalStream(mpc.stream(), device=dev)
        with torch.cuda.stream(ext):
            d = dict(inp=torch.from_numpy(b["rec"]).to(dev), ct=torch.from_numpy(b["contact"]).to(dev),
                     f=torch.zeros((chunk, N, 12), dtype=torch.float64, device=dev), x=torch.zeros((chunk, N + 1, 13), dtype=torch.float64, device=dev),
                     info=torch.zeros((chunk, 48), dtype=torch.uint8, device=dev))
        ext.synchronize()
        parts.append((tag, N, mpc, ext, d))
    half = mine // 2
    plan = [(parts[0], half), (parts[1], mine - half)]  # first half of the slice: config-2 problems, second half: config-3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_pass(timed):
        evs, conv = [], 0
        for (tag, N, mpc, ext, d), count in plan:
            done = 0
            while done < count:
                nb = min(chunk, count - done)
                with torch.cuda.stream(ext):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(ext)
                    mpc.solve_device(nb, d["inp"].data_ptr(), d["ct"].data_ptr(), d["f"].data_ptr(), d["x"].data_ptr(), d["info"].data_ptr(), sync=False)
                    e1.record(ext)
                evs.append((e0, e1))
                if timed:
                    ext.synchronize()
                    info = np.frombuffer(d["info"][:nb].cpu().numpy().tobytes(), dtype=np.dtype(
                        [("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))]))
                    conv += int(((info["status"] == 0) & (info["kkt"].max(axis=1) <= KKT_TOL)).sum())
                done += nb
        torch.cuda.synchronize(dev)
        return float(sum(a.elapsed_time(b) for a, b in evs)), conv

    for _ in range(max(1, args.warmup - 2)):
        one_pass(False)
    barrier()
    tot_ms, conv = 0.0, 0
    with ClockSampler(dev) as clocks:
        barrier()
        for _ in range(args.steps):
            ms, c = one_pass(True)
            tot_ms += ms
            conv += c
        barrier()
    stats = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
    counts = torch.tensor([conv, mine * args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {"metric": "MPC QP solves/sec (Go2, 1M-problem mixed sweep: horizon 10 trot + horizon 16 mixed gaits, solver AUTO)",
                "value": counts[0].item() / (stats[0].item() * 1e-3), "unit": "solves/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": stats[0].item() / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl["name"], "total_problems": total, "chunk": chunk, "chunk_reuse": True, "solver": wl["solver"],
                           "l2": "inputs of one chunk (75-115 MB) plus outputs exceed nothing; no flush (chunks stream back to back)",
                           "converged_fraction": counts[0].item() / counts[1].item(), "parallelism": f"shard{world}"},
                "gpu_launches": None, "clocks": clocks.summary()}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (0 = the workload's own)")
    ap.add_argument("--horizon", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the closed-loop hot-start extra")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config4":
        run_b200_wbc(args)
    elif args.workload == "config5":
        run_b200_sweep(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
