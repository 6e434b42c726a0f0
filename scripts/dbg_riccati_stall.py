import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_riccati_stall.npz"))
N = int(d["N"]); cfg = sequencer.GO2_MPC
rec, ct = d["rec"][None], d["contact"][None]
for kw in (dict(), dict(ipm_max_iter=200), dict(kkt_tol=1e-5)):
    for alpha in (float(d["alpha"]), 2 * float(d["alpha"]), 1e-5):
        mpc = BatchedMPC(alpha, d["w"], float(d["mu"]), float(d["fmin"]), float(d["fmax"]), "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=1, **kw)
        f, x, info = mpc.solve_records(rec, ct)
        print(kw, f"alpha {alpha:.2e}", "status", int(info.return_code[0]), "iters", int(info.acados_num_iter[0]), "kkt", info.kkt[0].tolist())
# perturb the state slightly
rng = np.random.default_rng(0)
mpc = BatchedMPC(float(d["alpha"]), d["w"], float(d["mu"]), float(d["fmin"]), float(d["fmax"]), "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=64)
recs = np.repeat(rec, 64, axis=0); recs[1:, 4:13] += rng.normal(0, 1e-3, (63, 9))
f, x, info = mpc.solve_records(recs, np.repeat(ct, 64, axis=0))
print("perturbed: ok", info.success.mean(), "statuses", np.unique(info.return_code, return_counts=True))
