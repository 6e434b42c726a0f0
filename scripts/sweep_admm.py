"""Development: throughput of the condensed kernel over the ADMM stopping tolerance / check interval / rho0 (config 2)."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = int(os.environ.get("HORIZON", "10")), int(os.environ.get("BATCH", "4096"))
GAITS = tuple(os.environ.get("GAITS", "TROT").split(","))
b = sequencer.random_batch(n, N, seed=20261021, gaits=GAITS)
grid = list(itertools.product([1e-3, 3e-4, 1e-4, 3e-5], [5, 10], [1e-3, 3e-3]))
if len(sys.argv) > 1:
    grid = [tuple(float(v) for v in a.split(",")) for a in sys.argv[1:]]
for eps, ce, rho in grid:
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N,
                     dt=cfg["dt"], max_batch=n, admm_eps=eps, admm_check_every=int(ce), admm_rho=rho)
    ts = []
    for _ in range(6):
        f, x, info = mpc.solve_records(b["rec"], b["contact"])
        ts.append(mpc.last_timing()[0])
    ms = float(np.median(ts[2:]))
    print(f"eps {eps:7.0e} check {int(ce):3d} rho0 {rho:7.0e}: {ms:6.3f} ms  {n / ms / 1e3:6.3f} M/s  iters {info.acados_num_iter.mean():6.1f} "
          f"rounds {info.polish_rounds.mean():5.2f} ok {info.success.mean():.4f}", flush=True)
    mpc.close()
