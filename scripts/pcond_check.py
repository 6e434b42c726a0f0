"""Partial-condensing kernel: same optimum and iterate count for every cond_N, timing over the cond_N axis (GPU)."""
import sys
import time

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
    return f, x, info, best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for N, gaits in ((10, ("TROT",)), (16, ("TROT", "PACE", "BOUND")), (10, ("STAND",)), (20, ("TROT", "STAND"))):
        b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=gaits)
        fr, xr, ir, tr = timed(mk(N, n, "SPARSE_RICCATI"), b["rec"], b["contact"])
        print(f"N={N} {gaits}: sparse Riccati kernel {n / tr / 1e3:8.1f} k/s  ok {ir.success.mean():.4f} it {ir.acados_num_iter.mean():.2f}")
        for cN in sorted({N, 5, 4, 2, 1} | ({8} if N == 16 else set()), reverse=True):
            try:
                f, x, info, t = timed(mk(N, n, "PARTIAL_CONDENSING", condensing_N=cN), b["rec"], b["contact"])
            except Exception as e:  # noqa: BLE001
                print(f"   cond_N={cN}: {e}")
                continue
            ok = info.success & ir.success
            sc = np.maximum(1.0, np.abs(fr).reshape(n, -1).max(axis=1))
            err = (np.abs(f - fr).reshape(n, -1).max(axis=1) / sc)[ok].max() if ok.any() else np.nan
            print(f"   cond_N={cN:2d}: {n / t / 1e3:8.1f} k/s  ok {info.success.mean():.4f} it {info.acados_num_iter.mean():.2f} "
                  f"max rel GRF diff vs sparse {err:.2e}  kkt {info.kkt[info.success].max() if info.success.any() else np.nan:.1e}")


if __name__ == "__main__":
    main()
