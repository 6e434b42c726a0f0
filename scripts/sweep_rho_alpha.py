"""Does the best initial ADMM penalty scale with the input weight alpha?  (trot N = 10, 4096 problems)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 10, 4096
b = sequencer.random_batch(n, N, seed=20261021)
for alpha in (1e-6, 1e-5, 1e-4, 1e-3):
    for rho in (5e-5, 0.5 * alpha, 5 * alpha, 50 * alpha):
        mpc = BatchedMPC(alpha, cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N, dt=cfg["dt"],
                         max_batch=n, admm_rho=rho)
        ts = []
        for _ in range(4):
            f, x, info = mpc.solve_records(b["rec"], b["contact"])
            ts.append(mpc.last_timing()[0])
        ms = float(np.median(ts[1:]))
        print(f"alpha {alpha:7.0e} rho0 {rho:8.1e}: {ms:7.3f} ms {n / ms / 1e3:6.3f} M/s iters {info.acados_num_iter.mean():6.1f} max {info.acados_num_iter.max():5d} "
              f"rounds {info.polish_rounds.mean():4.2f} ok {info.success.mean():.4f}", flush=True)
        mpc.close()
