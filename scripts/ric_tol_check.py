import sys, numpy as np
sys.path.insert(0, "/root/repo")
from dfki_quad_b200 import BatchedMPC, sequencer
from oracle import c_oracle, mpc_oracle as mo
params = dict(mo.GO2_PARAMS); params["weights"] = params["w_move"]
N, n = 16, 16384
b = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=("TROT", "PACE", "BOUND"))
Fc, Xc, ci, solved = c_oracle.solve_batch(N, b["rec"], b["contact"], params, kkt_tol=1e-10)
F2, X2, i2, s2 = c_oracle.riccati_batch(N, b["rec"], b["contact"], params, mode="parity")
sc = np.maximum(1.0, np.abs(Fc).reshape(n, -1).max(axis=1))
e2 = np.abs(F2 - Fc).reshape(n, -1).max(axis=1) / sc
print("C riccati parity vs C ADMM port: solved", s2, "max err", e2[i2["status"] == 0].max(), "n>1e-5", int((e2[i2["status"] == 0] > 1e-5).sum()))
c = sequencer.GO2_MPC
for solver, kt in (("SPARSE_RICCATI", 1e-7), ("SPARSE_RICCATI", 1e-8), ("SPARSE_RICCATI", 1e-9), ("AUTO", 1e-7)):
    m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], solver, horizon=N, dt=c["dt"], max_batch=n, kkt_tol=kt)
    m.solve_records(b["rec"], b["contact"])
    f, x, info = m.solve_records(b["rec"], b["contact"])
    err = np.abs(f - Fc).reshape(n, -1).max(axis=1) / sc
    bad = np.nonzero(err > 1e-5)[0]
    print(solver, kt, "ok", info.success.mean(), "it", info.acados_num_iter.mean(), "k/s", n / info.total_solver_time / 1e3, "max err", err.max(), "n>1e-5", len(bad), bad[:6], "gaits", b["gait_ids"][bad[:6]])
m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI", horizon=N, dt=c["dt"], max_batch=n)
f, x, info = m.solve_records(b["rec"], b["contact"])
err = np.abs(f - Fc).reshape(n, -1).max(axis=1) / sc
order = np.argsort(-err)[:12]
for p in order:
    print(p, "err %.2e" % err[p], "kkt", info.kkt[p], "it", info.acados_num_iter[p], "C-ipm status", i2["status"][p], "its", i2["iters"][p], "kkt", i2["kkt"][p][[0, 3]])
print("fraction with kkt > 1e-9:", (info.kkt.max(axis=1) > 1e-9).mean(), " > 1e-8:", (info.kkt.max(axis=1) > 1e-8).mean())
