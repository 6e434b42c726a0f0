"""Development: per-phase cycle accounting of mpc_riccati_kernel (B200QP_PHASE_CYCLES=1)."""
import ctypes, os, sys
os.environ["B200QP_PHASE_CYCLES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, _lib
cfg = sequencer.GO2_MPC
N, n = 16, 4096
b = sequencer.random_batch(n, N, seed=20261022, gaits=("TROT", "PACE", "BOUND"))
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=n)
for _ in range(2):
    f, x, info = mpc.solve_records(b["rec"], b["contact"])
out = np.zeros((n, 10), dtype=np.int64)
lib = _lib.load()
lib.b200qp_dev_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
assert lib.b200qp_dev_phase_cycles(mpc._h, n, out.ctypes.data) == 0
names = ["setup", "rollout+adjoint", "residuals+reduce", "R~ + PB,T + Huu,Hux", "chol+Linv | Pn", "Y, K, P update", "rt + (to sweeps)", "vector sweeps", "chol loop (warp0)", "Linv (warp0)"]
tot = out.sum(axis=1).mean()
print("mean cycles per problem (thread 0):", tot, " kernel ms", mpc.last_timing(), "iters", info.acados_num_iter.mean())
for i, nm in enumerate(names):
    print(f"  {nm:26s} {out[:, i].mean():10.0f}  {100 * out[:, i].mean() / tot:5.1f}%")
