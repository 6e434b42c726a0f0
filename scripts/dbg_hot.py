import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 10, 1024
mk = lambda **o: BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N, dt=cfg["dt"], max_batch=n, **o)
batch = sequencer.random_batch(n, N, seed=91, gaits=("TROT", "WALKING_TROT", "STAND"))
other = sequencer.random_batch(n, N, seed=92, gaits=("TROT", "PACE"))
cold, hot = mk(), mk(warm_start=2)
hot.solve_records(batch["rec"], batch["contact"])
f3, x3, i3 = hot.solve_records(other["rec"], other["contact"])
f3c, _, i3c = cold.solve_records(other["rec"], other["contact"])
scale = np.maximum(1.0, np.abs(f3c).reshape(n, -1).max(axis=1))
d = np.abs(f3 - f3c).reshape(n, -1).max(axis=1) / scale
print("ok", i3.success.mean(), "max rel diff", d.max(), "count>1e-5", (d > 1e-5).sum(), "iters hot", i3.acados_num_iter.mean(), "cold", i3c.acados_num_iter.mean())
print("hot rounds mean", i3.polish_rounds.mean(), "rounds when no admm", i3.polish_rounds[i3.acados_num_iter == 0].mean(), "frac no admm", (i3.acados_num_iter == 0).mean(), "hot ms", hot.last_timing()[0], "cold ms", cold.last_timing()[0])
bad = np.nonzero(d > 1e-5)[0][:6]
for p in bad:
    print(p, d[p], "hot kkt", i3.kkt[p], "iters", i3.acados_num_iter[p], "rounds", i3.polish_rounds[p], "| cold kkt", i3c.kkt[p], i3c.acados_num_iter[p], i3c.polish_rounds[p])
