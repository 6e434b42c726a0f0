"""Warp-per-problem Riccati kernel (solver SPARSE_RICCATI) against the block kernel (B200QP_RICCATI_LEGACY=1) and the C port."""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")


def run(tag):
    from dfki_quad_b200 import BatchedMPC, sequencer
    from oracle import c_oracle, mpc_oracle as mo

    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    c = sequencer.GO2_MPC
    out = []
    for N, gaits, n in ((16, ("TROT", "PACE", "BOUND"), 16384), (10, ("TROT",), 4096), (10, ("STAND", "STATIC_WALK"), 4096), (20, ("TROT", "STAND"), 4096)):
        b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=gaits)
        Fc, Xc, ci, solved = c_oracle.solve_batch(N, b["rec"], b["contact"], params, kkt_tol=1e-10)
        sc = np.maximum(1.0, np.abs(Fc).reshape(n, -1).max(axis=1))
        for cN in ((N,) if tag == "legacy" else (N, N // 2)):
            m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI", horizon=N, dt=c["dt"], max_batch=n,
                           cond_N=cN)
            m.solve_records(b["rec"], b["contact"])
            best = 1e9
            for _ in range(3):
                f, x, info = m.solve_records(b["rec"], b["contact"])
                best = min(best, info.total_solver_time)
            err = np.abs(f - Fc).reshape(n, -1).max(axis=1) / sc
            ok = info.success
            print(f"{tag:7s} N={N} {'/'.join(gaits):20s} cond_N={cN:2d}: {n / best / 1e3:8.1f} k/s ok {ok.mean():.4f} it {info.acados_num_iter.mean():.2f} "
                  f"GRF err vs C port: max {err[ok].max():.2e} n>1e-5 {(err[ok] > 1e-5).sum()}  kkt max {info.kkt[ok].max():.1e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        run("warp")
        env = dict(os.environ, B200QP_RICCATI_LEGACY="1")
        subprocess.run([sys.executable, __file__, "legacy"], env=env, check=False)
