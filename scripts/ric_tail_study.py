"""Which config-3 problems carry the IPM path's error tail: final KKT residual of the warp Riccati kernel against the relative GRF error
(reference: the C port tightened to KKT 1e-10)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from dfki_quad_b200 import BatchedMPC, sequencer
from oracle import c_oracle, mpc_oracle as mo
N, n = 16, 16384
c = sequencer.GO2_MPC
b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=("TROT", "PACE", "BOUND"))
params = dict(mo.GO2_PARAMS); params["weights"] = params["w_move"]
Fc, Xc, ci, solved = c_oracle.solve_batch(N, b["rec"], b["contact"], params, kkt_tol=1e-10)
m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI", horizon=N, dt=c["dt"], max_batch=n)
f, x, info = m.solve_records(b["rec"], b["contact"])
sc = np.maximum(1.0, np.abs(Fc).reshape(n, -1).max(axis=1))
err = np.abs(f - Fc).reshape(n, -1).max(axis=1) / sc
k = info.kkt.max(axis=1)
for thr in (1e-10, 3e-10, 1e-9, 3e-9, 1e-8):
    sel = k > thr
    print(f"kkt > {thr:.0e}: {sel.sum():5d} problems; of the {int((err > 1e-5).sum())} with err > 1e-5 it holds {(sel & (err > 1e-5)).sum()}; max err outside {err[~sel].max():.2e}")
print("iterations of err>1e-5:", info.acados_num_iter[err > 1e-5], "kkt", k[err > 1e-5])
