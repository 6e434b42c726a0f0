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
                    solver="SPARSE_RICCATI", seed=20261019 + 3, kernel="mpc_pcw_kernel<16,1>"),
    "config5": dict(name="go2_1M_mixed_sweep_sharded", horizon=0, batch=1048576, gaits=(), solver="ADMM(N=10 trot) + AUTO(N=16 mixed: condensed kernel where n <= 156, Riccati otherwise)",
                    seed=20261019 + 5, kernel="mpc_admm_kernel<64,60> + mpc_pcw_kernel<16,1>"),
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
        "impl": "reference", "metric": f"MPC QP solves/sec (Go2, horizon {N}, batch N)", "value": value, "unit": "solves/s",
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
        fp64_frac = fl / kern_s / 1e12 / fp64_peak if fp64_peak else None
        line = {
            "metric": f"MPC QP solves/sec (Go2, horizon {N}, batch N)", "value": value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps,
            "step_ms": {"min": float(min(step_ms)), "median": float(np.median(step_ms)), "max": float(max(step_ms)),
                        "note": "per-step device time on rank 0 (CUDA events on the handle's stream)"},
            "higher_is_better": True,
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
                         "algorithmic_bytes_per_solve": ab, "binding_ceiling": "fp64 FMA pipe / shared-memory latency (on-chip working set)",
                         "fp64_frac": fp64_frac,
                         "note": "HBM is NOT the binding ceiling of this kernel (SURVEY.md 8d): the on-chip solve is bound by the FP64 "
                                 "pipe and shared-memory latency, so the contract's HBM fraction is tiny by design and fp64_frac "
                                 "(algorithmic FLOPs / step time / measured DFMA peak, details under `fp64`) is the number to read. "
                                 "algorithmic_bytes uses the library's 48-byte info record (status, iterations, polish rounds, n_free, "
                                 "KKT 4-tuple); SURVEY 8d's formula with a 16-byte record gives 32 bytes less per solve (0.7 %)."},
            "fp64": {"flops_per_step": fl, "achieved_tflops": fl / kern_s / 1e12, "measured_peak_tflops": fp64_peak,
                     "frac": fp64_frac, "nominal_peak_tflops": 37.0,
                     "note": "algorithmic FP64 FLOPs of the method as implemented / step time; peak = DFMA microbenchmark on this GPU"},
        }
        if wl["solver"] == "CONDENSED_ADMM" and not args.no_closed_loop:
            try:  # reported extra, outside every timed region above: the same kernels in closed loop with hot start
                line["closed_loop_hot_start"] = closed_loop_extra(N, min(B, 4096), dev)
            except Exception as e:
                line["closed_loop_hot_start"] = {"error": repr(e)}
        if cpu is not None and "error" not in cpu:
            line["cpu_baseline"] = {"value": cpu["solved"] / cpu["seconds"], "unit": "solves/s", "cores": cpu["cores"],
                                    "kind": cpu["kind"], "sample": cpu["sample"], "solver": cpu.get("solver"),
                                    "note": "any GPU/CPU ratio depends on this box's host core count - quote it with `cores`"}
            if not args.no_extras:
                try:
                    line["cpu_baseline"]["more"] = cpu_extras(N, batch["rec"], batch["contact"], wl)
                except Exception as e:
                    line["cpu_baseline"]["more"] = {"error": repr(e)}
        elif cpu is not None:
            line["cpu_baseline"] = cpu
    # ---- the other BASELINE.json configurations, budgeted (a few seconds each), every rank takes part ---------------------
    extras = None
    if args.workload == "config2" and not args.no_extras:
        extras = {}
        for name, fn in (("config3", extra_config3), ("config4", extra_config4), ("config5_strong", extra_config5), ("fp32", extra_fp32),
                         ("closed_loop_on_device", extra_closed_loop_device)):
            try:
                extras[name] = fn(args, dev, rank, world)
            except Exception as e:  # an extra never costs the headline line
                extras[name] = {"error": repr(e)}
    if rank == 0:
        if extras is not None:
            line["other_configs"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------- extras
INFO_DT = np.dtype([("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))])


def _reduce(world, dev, maxes, sums):
    import torch

    if world == 1:
        return list(maxes), list(sums)
    import torch.distributed as dist

    a = torch.tensor(maxes, dtype=torch.float64, device=dev)
    b = torch.tensor(sums, dtype=torch.float64, device=dev)
    dist.all_reduce(a, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return a.tolist(), b.tolist()


def cpu_extras(N, rec, contact, wl):
    """More of the restated reference CPU path on the headline batch (rank 0, bounded): OSQP's library-default penalty, the
    Riccati IPM port in its three modes, and config 1 = single-problem single-thread latency."""
    from oracle import c_oracle, mpc_oracle as mo

    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    cores = os.cpu_count() or 1
    m = min(rec.shape[0], 64 * cores)
    out = {}
    t0 = time.perf_counter()
    _, _, inf, solved = c_oracle.solve_batch(N, rec[:m], contact[:m], params, threads=cores, rho=0.1)
    out["admm_osqp_default_rho_0.1"] = {"value": solved / (time.perf_counter() - t0), "unit": "solves/s", "cores": cores,
                                        "mean_iterations": float(inf["iters"].mean()), "sample": f"first {m} problems"}
    for mode in ("SPEED", "BALANCE", "parity"):
        t0 = time.perf_counter()
        _, _, inf, solved = c_oracle.riccati_batch(N, rec[:m], contact[:m], params, mode=mode, threads=cores)
        out[f"riccati_ipm_{mode}"] = {"value": m / (time.perf_counter() - t0), "unit": "solves/s", "cores": cores,
                                      "reached_tolerance": solved / m, "mean_iterations": float(inf["iters"].mean()),
                                      "sample": f"first {m} problems",
                                      "tolerance_cap": dict(zip(("tol", "max_iter"), c_oracle.RICCATI_MODES[mode]))}
    # BASELINE.json configs[0]: one trot problem at a time on one thread
    lat = {}
    k = min(200, rec.shape[0])
    for name, fn in (("admm_tuned", lambda i: c_oracle.solve_batch(N, rec[i:i + 1], contact[i:i + 1], params, threads=1)),
                     ("admm_osqp_default", lambda i: c_oracle.solve_batch(N, rec[i:i + 1], contact[i:i + 1], params, threads=1, rho=0.1)),
                     ("riccati_ipm_SPEED", lambda i: c_oracle.riccati_batch(N, rec[i:i + 1], contact[i:i + 1], params, mode="SPEED", threads=1)),
                     ("riccati_ipm_BALANCE", lambda i: c_oracle.riccati_batch(N, rec[i:i + 1], contact[i:i + 1], params, mode="BALANCE", threads=1))):
        ts = []
        for i in range(k):
            t0 = time.perf_counter()
            fn(i)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts[5:]) * 1e3
        lat[name] = {"median_ms": float(np.median(ts)), "p99_ms": float(np.percentile(ts, 99)), "problems": int(len(ts))}
    out["config1_single_problem_latency_1_thread"] = lat
    return out


def _mpc_pass(mpc, ext, dev, B, N, rec, contact, steps, warmup, flush):
    """device-resident + host-buffer timing of one handle on one batch -> dict of raw numbers (per rank)."""
    import torch

    rec_h = torch.from_numpy(rec).pin_memory()
    ct_h = torch.from_numpy(contact).pin_memory()
    with torch.cuda.stream(ext):
        d_in, d_ct = rec_h.to(dev, non_blocking=True), ct_h.to(dev, non_blocking=True)
        d_f = torch.zeros((B, N, 12), dtype=torch.float64, device=dev)
        d_x = torch.zeros((B, N + 1, 13), dtype=torch.float64, device=dev)
        d_info = torch.zeros((B, 48), dtype=torch.uint8, device=dev)
    ext.synchronize()
    for _ in range(warmup):
        mpc.solve_device(B, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=False)
    ext.synchronize()
    evs = []
    for _ in range(steps):
        with torch.cuda.stream(ext):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            mpc.solve_device(B, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=False)
            e1.record(ext)
        evs.append((e0, e1))
    ext.synchronize()
    ms = float(sum(a.elapsed_time(b) for a, b in evs))
    info = np.frombuffer(d_info.cpu().numpy().tobytes(), dtype=INFO_DT)
    conv = (info["status"] == 0) & (info["kkt"].max(axis=1) <= KKT_TOL)
    f_h = torch.zeros((B, N, 12), dtype=torch.float64).pin_memory().numpy()
    x_h = torch.zeros((B, N + 1, 13), dtype=torch.float64).pin_memory().numpy()
    mpc.solve_records(rec_h.numpy(), ct_h.numpy(), f_h, x_h)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, inf = mpc.solve_records(rec_h.numpy(), ct_h.numpy(), f_h, x_h)
    e2e_s = time.perf_counter() - t0
    conv_e2e = int((inf.success & (inf.kkt.max(axis=1) <= KKT_TOL)).sum())
    return dict(ms=ms, conv=int(conv.sum()), e2e_s=e2e_s, conv_e2e=conv_e2e, info=info)


def extra_config3(args, dev, rank, world, steps=3):
    """BASELINE.json configs[2]: sparse Go2 MPC, horizon 16, 16384 mixed trot / pace / bound problems per GPU, Riccati IPM kernel."""
    import torch

    from dfki_quad_b200 import BatchedMPC, _lib, sequencer

    wl = WORKLOADS["config3"]
    N, B = wl["horizon"], wl["batch"]
    cfg = sequencer.GO2_MPC
    batch = sequencer.random_batch(B, N, seed=wl["seed"] + 1000 * rank, gaits=wl["gaits"], device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {}
    raw = {}
    # riccati_kernel: the warp-per-problem Riccati IPM (mpc_pcw.cu), its default pairs stages (cond_N = N / 2); the cond_N = N run is the
    # same kernel with single-stage blocks (the fully sparse end of the reference's cond_N axis, mpc.cpp:219-225)
    for key, solver, kw in (("riccati_kernel", "SPARSE_RICCATI", {}), ("riccati_kernel_without_second_opinion", "SPARSE_RICCATI", {}),
                            ("riccati_kernel_cond_N_16", "SPARSE_RICCATI", dict(condensing_N=16)),
                            ("partial_condensing_cond_N_4", "PARTIAL_CONDENSING", dict(condensing_N=4)), ("auto", "AUTO", {})):
        if key == "riccati_kernel_without_second_opinion":
            os.environ["B200QP_IPM_REFINE"] = "0"  # read at handle creation: the interior-point kernels alone (error tail of DESIGN 3.3b)
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"],
                         max_batch=B, device=dev, **kw)
        os.environ.pop("B200QP_IPM_REFINE", None)
        ext = torch.cuda.ExternalStream(mpc.stream(), device=dev)
        r = _mpc_pass(mpc, ext, dev, B, N, batch["rec"], batch["contact"], steps, 3, flush)
        mpc.close()
        (ms, e2e_s), (conv, conv_e2e) = _reduce(world, dev, [r["ms"], r["e2e_s"]], [r["conv"] * steps, r["conv_e2e"] * steps])
        raw[key] = r
        out[key] = {"value": conv / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms / steps,
                    "e2e": {"value": conv_e2e / e2e_s, "unit": "solves/s"}, "converged_fraction": conv / (B * steps * world),
                    "mean_iterations": float(r["info"]["iterations"].mean())}
    if rank == 0:
        info = raw["riccati_kernel"]["info"]
        fl = float(info["iterations"].sum()) * N * (9600.0 + 2 * 1200.0)
        kern_s = raw["riccati_kernel"]["ms"] / steps * 1e-3
        try:
            pk = _lib.fp64_peak_tflops(dev)
            out["riccati_kernel"]["fp64"] = {"flops_per_step": fl, "achieved_tflops": fl / kern_s / 1e12, "frac": fl / kern_s / 1e12 / pk}
        except Exception:
            pass
        peak, _ = measured_peaks()
        out["riccati_kernel"]["roofline_hbm_frac"] = algorithmic_bytes(N) * B / kern_s / 1e9 / peak
        if not args.no_cpu_baseline:
            from oracle import c_oracle, mpc_oracle as mo

            params = dict(mo.GO2_PARAMS)
            params["weights"] = params["w_move"]
            cores = os.cpu_count() or 1
            m = min(B, 96 * cores)
            cb = {}
            for mode in ("SPEED", "BALANCE", "parity"):
                t0 = time.perf_counter()
                _, _, inf, solved = c_oracle.riccati_batch(N, batch["rec"][:m], batch["contact"][:m], params, mode=mode, threads=cores)
                cb[f"riccati_ipm_{mode}"] = {"value": m / (time.perf_counter() - t0), "reached_tolerance": solved / m,
                                             "mean_iterations": float(inf["iters"].mean())}
            t0 = time.perf_counter()
            _, _, inf, solved = c_oracle.solve_batch(N, batch["rec"][:m // 4], batch["contact"][:m // 4], params, threads=cores)
            cb["dense_admm_port"] = {"value": solved / (time.perf_counter() - t0)}
            out["cpu_baseline"] = {"unit": "solves/s", "cores": cores, "kind": "port", "sample": f"first {m} problems of rank 0's batch",
                                   "value": cb["riccati_ipm_parity"]["value"], "modes": cb,
                                   "note": "value = C Riccati IPM port (restatement of the reference's default PARTIAL_CONDENSING_HPIPM plan) "
                                           "at the GPU kernels' tolerance; SPEED / BALANCE-like caps beside it"}
    out["config"] = {"workload": wl["name"], "horizon": N, "batch_per_gpu": B, "gaits": "/".join(wl["gaits"]), "steps": steps,
                     "l2": "256 MiB flush between timed iterations"}
    return out


def extra_closed_loop_device(args, dev, rank, world, n=4096, periods=20):
    """n robots x `periods` control periods with sequencer + MPC (+ 5 WBC solves per period) in the loop, everything
    device-resident (b200qp_mpc_rollout); wall clock of the call, i.e. including the one H2D of the initial and D2H of the final states."""
    from dfki_quad_b200 import BatchedMPC, gait as gaitmod, pack_seq2, sequencer as sq, wbc

    rng = np.random.Generator(np.random.PCG64(20261019 + 40 + rank))
    c = sq.GO2_MPC
    m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "CONDENSED_ADMM", horizon=10, dt=c["dt"], max_batch=n, device=dev,
                   warm_start=1)
    m.set_model(sq.GO2_MASS, sq.GO2_INERTIA, sq.GO2_COM, sq.GO2_SHOULDERS, sq.RAIBERT_K, sq.RAIBERT_G, sq.MAX_LEG_LENGTH)
    w = wbc.BatchedWBCSolver(max_batch=n, device=dev)
    w.scene_configure()
    yaw = rng.uniform(-np.pi, np.pi, n)
    quat = sq._yaw_quat(yaw)
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.full(n, 0.30)], axis=-1)
    feet = pos[:, None, :] + np.einsum("nij,lj->nli", sq.quat_to_rot(quat), sq.GO2_SHOULDERS)
    feet[:, :, 2] = 0.0
    period, duty, offset = gaitmod.gait_arrays(np.full(n, gaitmod.GAIT_NAMES.index("TROT")))
    target = dict(hybrid_x_dot=0.4, hybrid_y_dot=0.1, wz=0.3, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    seq0 = pack_seq2(np.concatenate([quat, pos, np.zeros((n, 6))], axis=1), feet, rng.integers(0, 500_000_000, n).astype(np.float64), target,
                     (period, duty, offset))
    out = {}
    for name, wb, cdt, sub in (("mpc_only_period_50ms", None, 0.0, 10), ("mpc_100hz_wbc_500hz", w, 0.01, 5)):
        m.sequencer_configure("simple", early_contact_detection=0)
        m.SetWarmStart(1)
        m.rollout(seq0.copy(), 2, wbc=wb, plant_inertia="reference_model", history=False, control_dt=cdt, substeps=sub)  # warm-up
        m.sequencer_configure("simple", early_contact_detection=0)
        m.SetWarmStart(1)
        dt_s = 1e30
        for _ in range(3):  # best of three calls: one call is 40 - 450 ms of wall clock, a host hiccup would dominate it
            m.sequencer_configure("simple", early_contact_detection=0)
            m.SetWarmStart(1)
            seq = seq0.copy()
            t0 = time.perf_counter()
            hist = m.rollout(seq, periods, wbc=wb, plant_inertia="reference_model", control_dt=cdt, substeps=sub)
            dt_s = min(dt_s, time.perf_counter() - t0)
        (t_m,), _ = _reduce(world, dev, [dt_s], [0])
        ok = float((hist[:, :, 11] == 0).mean())
        out[name] = {"robot_periods_per_s": world * n * periods / t_m, "mpc_solves_per_s": world * n * periods / t_m,
                     "wbc_solves_per_s": (world * n * periods * 5 / t_m) if wb is not None else 0.0, "seconds": t_m, "mpc_converged_fraction": ok,
                     "wbc_converged_fraction": float((hist[:, :, 14] == 0).mean()) if wb is not None else None,
                     "height_error_max_m": float(np.abs(hist[-1, :, 2] - 0.30).max()), "control_period_s": cdt if cdt > 0 else 0.05,
                     "real_time_factor": (n * periods * (cdt if cdt > 0 else 0.05)) / t_m / n}
    out["config"] = {"robots_per_gpu": n, "control_periods": periods, "gait": "TROT", "horizon": 10, "warm_start": 1,
                     "note": "real_time_factor = simulated seconds per wall second for the whole batch; host work inside the loop: none (kernel launches only)"}
    m.close()
    w.close()
    return out


def extra_config4(args, dev, rank, world, steps=3):
    """BASELINE.json configs[3]: WBC task-space QP, 16384 per GPU."""
    import torch

    from dfki_quad_b200 import wbc

    wl = WORKLOADS["config4"]
    B = wl["batch"]
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
        solver.solve_device(B, dH.data_ptr(), dg.data_ptr(), dA.data_ptr(), dlo.data_ptr(), dhi.data_ptr(), dx.data_ptr(), dinfo.data_ptr(), sync=False)

    for _ in range(3):
        step()
    ext.synchronize()
    evs = []
    for _ in range(steps):
        with torch.cuda.stream(ext):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            step()
            e1.record(ext)
        evs.append((e0, e1))
    ext.synchronize()
    ms = float(sum(a.elapsed_time(c) for a, c in evs))
    info = np.frombuffer(dinfo.cpu().numpy().tobytes(), dtype=INFO_DT)
    conv = int(((info["status"] == 0) & (info["kkt"].max(axis=1) <= KKT_TOL)).sum())
    Hp, gp, Ap, lop, hip = (torch.from_numpy(a).pin_memory().numpy() for a in (H, g, A, lo, hi))
    xh = torch.zeros((B, 30), dtype=torch.float64).pin_memory().numpy()
    solver.solve_abi(Hp, gp, Ap, lop, hip, xh)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, inf = solver.solve_abi(Hp, gp, Ap, lop, hip, xh)
    e2e_s = time.perf_counter() - t0
    conv_e2e = int(((inf["status"] == 0) & (inf["kkt"].max(axis=1) <= KKT_TOL)).sum())
    # the same QPs entered as robot states (576 B each): scene assembled on the device (kernel K6), then the same QP kernel
    mb = b["mpc_batch"]
    scene = torch.from_numpy(wbc.pack_scene(mb["pos"], mb["quat"], mb["vel"], mb["omega"], b["q"], b["qd"], b["contacts"], b["a_body"], b["a_foot"],
                                            b["f_ref"])).pin_memory().numpy()
    solver.scene_configure()
    tau_h = np.zeros((B, 12))
    solver.scene_solve(scene, xh, tau_h)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, inf_s = solver.scene_solve(scene, xh, tau_h)
    scene_s = time.perf_counter() - t0
    conv_scene = int(((inf_s["status"] == 0) & (inf_s["kkt"].max(axis=1) <= KKT_TOL)).sum())
    (ms_m, e2e_m, scene_m), (c1, c2, c3) = _reduce(world, dev, [ms, e2e_s, scene_s], [conv * steps, conv_e2e * steps, conv_scene * steps])
    out = {"metric": "WBC task-space QP solves/sec (Go2 reduced TSID, 30 vars, 18 eq, 28 ineq)", "value": c1 / (ms_m * 1e-3), "unit": "solves/s",
           "ms_per_step": ms_m / steps, "e2e": {"value": c2 / e2e_m, "unit": "solves/s", "h2d_bytes_per_step": int(B * 19512), "d2h_bytes_per_step": int(B * 288)},
           "converged_fraction": c1 / (B * steps * world), "mean_iterations": float(info["iterations"].mean()),
           "kkt_max_converged": float(info["kkt"][(info["status"] == 0)].max()),
           "e2e_scene_on_device": {"value": c3 / scene_m, "unit": "solves/s", "h2d_bytes_per_step": int(B * 576), "d2h_bytes_per_step": int(B * (240 + 96 + 48)),
                                   "note": "b200qp_wbc_scene_solve_batch: robot state in (576 B), rigid-body terms + TSID scene assembled by kernel K6, x + joint torques out"},
           "config": {"workload": wl["name"], "batch_per_gpu": B, "steps": steps, "l2": "256 MiB flush between timed iterations"}}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import c_oracle

        cores = os.cpu_count() or 1
        m = min(B, 512 * cores)
        t0 = time.perf_counter()
        _, inf2, solved = c_oracle.wbc_batch(H[:m], g[:m], A[:m], lo[:m], hi[:m], 18, threads=cores)
        out["cpu_baseline"] = {"value": solved / (time.perf_counter() - t0), "unit": "solves/s", "cores": cores, "kind": "port",
                               "solved_fraction": solved / m, "sample": f"first {m} QPs of rank 0's batch",
                               "solver": "C dense primal-dual IPM (oracle/wbc_ref.c), OpenMP over the batch"}
    return out


def extra_config5(args, dev, rank, world):
    """BASELINE.json configs[4]: 1 048 576 DISTINCT problems (half config-2, half config-3 distribution), contiguous slices over
    the GPUs (STRONG scaling: the total is fixed), host buffers end to end.  Every robot enters as a sequencer input
    (state, feet, Target, gait, phase: 400 B) - trajectory planner, contact schedule and Raibert footholds run as a kernel -
    and its GRFs + predicted states come back to pinned host memory.  `value` = problems / sum of kernel time (max over ranks),
    `e2e` = problems / wall time of the host-buffer calls (max over ranks)."""
    import torch

    from dfki_quad_b200 import BatchedMPC, sequencer
    from dfki_quad_b200.shard import shard_range

    total = 1 << 20
    lo_i, hi_i = shard_range(total, rank, world)
    mine = hi_i - lo_i
    chunk = 65536
    cfg = sequencer.GO2_MPC
    res = []
    kern_ms = wall = 0.0
    conv = 0
    sample = None
    for tag, N, gaits, solver, lo_p, hi_p in (("h10_trot", 10, ("TROT",), "CONDENSED_ADMM", 0, total // 2),
                                               ("h16_mixed", 16, ("TROT", "PACE", "BOUND"), "AUTO", total // 2, total)):
        a, b = max(lo_i, lo_p), min(hi_i, hi_p)
        if b <= a:
            continue
        # distinct problems: the global problem index seeds nothing twice (one generator stream per 65536-problem chunk)
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"],
                         max_batch=chunk, device=dev)
        mpc.set_model(**sequencer.go2_seq_model())
        f_h = torch.zeros((chunk, N, 12), dtype=torch.float64).pin_memory().numpy()
        x_h = torch.zeros((chunk, N + 1, 13), dtype=torch.float64).pin_memory().numpy()
        seq_pin = torch.zeros((chunk, 50), dtype=torch.float64).pin_memory().numpy()
        first = True
        for c0 in range(a, b, chunk):
            n = min(chunk, b - c0)
            seq_pin[:n] = sequencer.random_seq_inputs(n, seed=WORKLOADS["config5"]["seed"] + c0 // 4096 + 7, gaits=gaits)
            if first:  # warm-up on the first chunk (untimed), and a CPU sample of the same problems
                mpc.sequence_and_solve(seq_pin[:n], None, f_h[:n], x_h[:n])
                if rank == 0 and sample is None and not args.no_cpu_baseline:
                    sample = (N,) + mpc.sequence_records(seq_pin[: min(n, 4096)])
                first = False
            t0 = time.perf_counter()
            _, _, inf = mpc.sequence_and_solve(seq_pin[:n], None, f_h[:n], x_h[:n])
            wall += time.perf_counter() - t0
            kern_ms += inf.total_solver_time * 1e3
            conv += int((inf.success & (inf.kkt.max(axis=1) <= KKT_TOL)).sum())
        mpc.close()
    (kern_m, wall_m), (conv_all,) = _reduce(world, dev, [kern_ms, wall], [conv])
    out = {"metric": "MPC QP solves/sec (Go2, 1M-problem mixed sweep: horizon 10 trot + horizon 16 trot/pace/bound)", "scaling": "strong",
           "value": conv_all / (kern_m * 1e-3), "unit": "solves/s", "e2e": {"value": conv_all / wall_m, "unit": "solves/s",
                                                                           "h2d_bytes_per_step": int(total * 400),
                                                                           "d2h_bytes_per_step": int(total // 2 * (2152 + 3368))},
           "converged_fraction": conv_all / total, "seconds_for_the_sweep_e2e": wall_m,
           "config": {"workload": WORKLOADS["config5"]["name"], "total_problems": total, "distinct_problems": True, "chunk": chunk,
                      "solver": "CONDENSED_ADMM (horizon 10) + AUTO (horizon 16)", "entry": "b200qp_mpc_sequence_solve_batch, pinned host buffers",
                      "l2": "every chunk is new data: 26 MB in, 138 - 221 MB out per chunk (larger than L2)", "parallelism": f"shard{world}"}}
    if rank == 0 and sample is not None:
        from oracle import c_oracle, mpc_oracle as mo

        params = dict(mo.GO2_PARAMS)
        params["weights"] = params["w_move"]
        N, rec, ct = sample
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        _, _, _, solved = c_oracle.solve_batch(N, rec, ct, params, threads=cores)
        out["cpu_baseline"] = {"value": solved / (time.perf_counter() - t0), "unit": "solves/s", "cores": cores, "kind": "port",
                               "sample": f"{rec.shape[0]} problems of the horizon-{N} half, C ADMM port"}
    return out


def extra_fp32(args, dev, rank, world):
    """North star: "with the FP32 error stated separately".  The condensed kernel is run on rank 0's config-2 batch with every
    stored intermediate rounded to binary32 (b200qp_mpc_config.precision): (1) FP32 K^-1 + ADMM iterations on FP64 problem data
    with the polish in FP64, (2) FP32 everywhere (assembly, factor, iterations, polish).  Errors are against the FP64 path on the same problems; KKT by the C checker in FP64."""
    if rank != 0:
        return None
    from dfki_quad_b200 import BatchedMPC, sequencer

    wl = WORKLOADS["config2"]
    N, B = wl["horizon"], wl["batch"]
    cfg = sequencer.GO2_MPC
    batch = sequencer.random_batch(B, N, seed=wl["seed"], gaits=wl["gaits"], device=dev)
    out = {}
    ref = None
    for key, prec in (("fp64", 0), ("fp32_factor_and_admm_fp64_assembly_and_polish", 1), ("fp32_everywhere", 2)):
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N,
                         dt=cfg["dt"], max_batch=B, device=dev, precision=prec)
        f, x, info = mpc.solve_records(batch["rec"], batch["contact"])
        mpc.close()
        if prec == 0:
            ref = f
            continue
        ok = info.success
        sc = np.maximum(1.0, np.abs(ref).reshape(B, -1).max(axis=1))
        err = np.abs(f - ref).reshape(B, -1).max(axis=1) / sc
        d = {"solved_fraction": float(ok.mean()), "mean_admm_iterations": float(info.acados_num_iter.mean()),
             "mean_polish_rounds": float(info.polish_rounds.mean()),
             "grf_rel_error_vs_fp64": {"median": float(np.median(err[ok])), "p99": float(np.percentile(err[ok], 99)), "max": float(err[ok].max())}}
        if not args.no_cpu_baseline:
            try:
                from oracle import c_oracle, mpc_oracle as mo

                params = dict(mo.GO2_PARAMS)
                params["weights"] = params["w_move"]
                kkt = c_oracle.kkt_batch(N, batch["rec"][ok], batch["contact"][ok], params, f[ok], x[ok])
                d["kkt_fp64_checker"] = {"median": [float(v) for v in np.median(kkt, axis=0)], "max": [float(v) for v in kkt.max(axis=0)],
                                         "order": "stationarity, dynamics, inequality, complementarity"}
            except Exception as e:
                d["kkt_fp64_checker"] = {"error": repr(e)}
        out[key] = d
    out["note"] = ("rounding study: the FP64 kernel with every stored intermediate rounded to binary32; no FP32 build is shipped - the "
                   "reference computes in double and the 1e-5 GRF / 1e-6 KKT bar is an FP64 bar")
    return out


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
    ap.add_argument("--no-extras", action="store_true", help="skip the budgeted runs of the other BASELINE configs (config 3, 4, 5)")
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
