import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 16, 16
batch = sequencer.random_batch(n, N, seed=316, gaits=("TROT", "PACE", "BOUND"))
for mi in (12, 14, 16, 18, 20, 22, 25, 30, 40):
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=n, ipm_max_iter=mi)
    f, x, info = mpc.solve_records(batch["rec"], batch["contact"])
    print(mi, "status", info.return_code[[1, 9]], "iters", info.acados_num_iter[[1, 9]], "kkt", info.kkt[[1, 9]][:, [0, 3]].ravel(), "all iters", info.acados_num_iter.tolist())
