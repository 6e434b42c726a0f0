"""Warm start of the Riccati interior-point kernel (solver SPARSE_RICCATI, warm_start = 1): iterations and solutions on a batch whose
states moved a little since the previous call, against the cold solve of the same problems.  Usage: python scripts/ipm_warm_check.py"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from dfki_quad_b200 import BatchedMPC, sequencer  # noqa: E402

N, n = 16, 8192
c = sequencer.GO2_MPC
b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=("TROT", "PACE", "BOUND"))
rng = np.random.default_rng(1)
rec2 = b["rec"].copy()
rec2[:, 4:7] += rng.normal(0, 3e-3, (n, 3))     # position
rec2[:, 7:13] += rng.normal(0, 2e-2, (n, 6))    # twist
cold = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI", horizon=N, dt=c["dt"], max_batch=n)
cold.solve_records(rec2, b["contact"])
fc, xc, ic = cold.solve_records(rec2, b["contact"])
print(f"cold: ok {ic.success.mean():.4f} iterations {ic.acados_num_iter.mean():.2f} time {ic.total_solver_time * 1e3:.2f} ms")
for mu in (os.environ.get("MUS", "1 0.1 0.01 0.001")).split():
    os.environ["B200QP_IPM_WARM_MU"] = mu
    w = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI", horizon=N, dt=c["dt"], max_batch=n,
                   warm_start=1)
    w.solve_records(b["rec"], b["contact"])
    f, x, i = w.solve_records(rec2, b["contact"])
    sc = np.maximum(1.0, np.abs(fc).reshape(n, -1).max(axis=1))
    ok = i.success & ic.success
    err = np.abs(f - fc).reshape(n, -1).max(axis=1) / sc
    print(f"warm mu0={mu}: ok {i.success.mean():.4f} iterations {i.acados_num_iter.mean():.2f} time {i.total_solver_time * 1e3:.2f} ms  GRF diff vs cold max {err[ok].max():.2e}")
    w.close()
