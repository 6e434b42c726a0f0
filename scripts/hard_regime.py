import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 20, 1024
b = sequencer.random_batch(n, N, seed=1080, gaits=("FLYING_TROT", "PRONK"))
for solver in ("CONDENSED_ADMM", "AUTO", "SPARSE_RICCATI"):
    mpc = BatchedMPC(3.1e-6, cfg["state_weights_move"], 0.9, 0.0, 120.0, solver, horizon=N, dt=cfg["dt"], max_batch=n)
    f, x, info = mpc.solve_records(b["rec"], b["contact"])
    print(solver, "ok", info.success.mean(), "codes", np.unique(info.return_code, return_counts=True), "iters max", info.acados_num_iter.max(), "ms", mpc.last_timing()[0])
