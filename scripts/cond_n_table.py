"""The reference's cond_N axis (MPC::MPC condensing_N, mpc.cpp:219-225; experiment axis solver_experiments.py:619-640) on the GPU:
converged solves/s for a trot batch at N = 10 / 16 / 20 from the dense end (full condensing, ADMM kernel) over the block-condensed
Riccati IPM (cond_N blocks) to the sparse end (cond_N = N), block kernel and warp kernel.  Usage: python scripts/cond_n_table.py [batch]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from dfki_quad_b200 import BatchedMPC, sequencer  # noqa: E402


def mk(N, n, solver, **kw):
    c = sequencer.GO2_MPC
    return BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], solver, horizon=N, dt=c["dt"], max_batch=n, **kw)


def timed(m, rec, ct, reps=3):
    m.solve_records(rec, ct)
    best = 1e9
    for _ in range(reps):
        f, x, info = m.solve_records(rec, ct)
        best = min(best, info.total_solver_time)
    m.close()
    return f, info, best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    print("| N | gait | dense: full condensing + ADMM | " + " | ".join(f"block kernel cond_N={c}" for c in ("1", "2", "4", "5", "N/2", "N")) +
          " | warp kernel cond_N=N/2 | warp kernel cond_N=N |")
    print("|---|---|---|" + "---|" * 8)
    for N, gaits in ((10, ("TROT",)), (16, ("TROT",)), (20, ("TROT",)), (16, ("TROT", "PACE", "BOUND"))):
        b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=gaits)
        row = []
        try:
            f0, i0, t = timed(mk(N, n, "CONDENSED_ADMM"), b["rec"], b["contact"])
            row.append(f"{i0.success.sum() / t / 1e3:.0f}k ({i0.success.mean():.3f})")
        except Exception as e:  # noqa: BLE001
            f0 = None
            row.append("n > 156: refused")
        ref = None
        for cN in (1, 2, 4, 5, N // 2, N):
            f, info, t = timed(mk(N, n, "PARTIAL_CONDENSING", condensing_N=cN), b["rec"], b["contact"])
            ref = f if ref is None else ref
            row.append(f"{info.success.sum() / t / 1e3:.0f}k")
        for cN in (N // 2, N):
            f, info, t = timed(mk(N, n, "SPARSE_RICCATI", condensing_N=cN), b["rec"], b["contact"])
            ok = info.success
            sc = np.maximum(1.0, np.abs(ref).reshape(n, -1).max(axis=1))
            err = (np.abs(f - ref).reshape(n, -1).max(axis=1) / sc)[ok].max()
            row.append(f"{ok.sum() / t / 1e3:.0f}k (it {info.acados_num_iter.mean():.1f}, diff {err:.0e})")
        print(f"| {N} | {'/'.join(gaits)} | " + " | ".join(row) + " |", flush=True)


if __name__ == "__main__":
    main()
