import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 10, 4096
b = sequencer.random_batch(n, N, seed=20261021, gaits=tuple(os.environ.get("GAITS", "TROT").split(",")))
for alpha, sigma in [(1.6, 1e-6), (1.4, 1e-6), (1.7, 1e-6), (1.8, 1e-6), (1.9, 1e-6), (1.95, 1e-6), (1.8, 1e-5), (1.8, 1e-7), (1.6, 1e-5)]:
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N,
                     dt=cfg["dt"], max_batch=n, admm_alpha=alpha, admm_sigma=sigma)
    ts = []
    for _ in range(6):
        f, x, info = mpc.solve_records(b["rec"], b["contact"])
        ts.append(mpc.last_timing()[0])
    ms = float(np.median(ts[2:]))
    print(f"relax {alpha:4.2f} sigma {sigma:6.0e}: {ms:6.3f} ms {n / ms / 1e3:6.3f} M/s iters {info.acados_num_iter.mean():5.1f} rounds {info.polish_rounds.mean():4.2f} ok {info.success.mean():.4f}", flush=True)
    mpc.close()
