"""Development: per-phase cycle accounting of mpc_pcw_kernel (B200QP_PHASE_CYCLES=1).  usage: [cond_N] [N] [gaits,..]"""
import ctypes, os, sys
os.environ["B200QP_PHASE_CYCLES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, _lib
cfg = sequencer.GO2_MPC
cN = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gaits = tuple(sys.argv[3].split(",")) if len(sys.argv) > 3 else ("TROT", "PACE", "BOUND")
n = 4096
b = sequencer.random_batch(n, N, seed=20261022, gaits=gaits)
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "SPARSE_RICCATI", cond_N=cN, horizon=N, dt=cfg["dt"], max_batch=n)
for _ in range(2):
    f, x, info = mpc.solve_records(b["rec"], b["contact"])
out = np.zeros((n, 10), dtype=np.int64)
lib = _lib.load()
lib.b200qp_dev_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p]
assert lib.b200qp_dev_phase_cycles(mpc._h, n, out.ctypes.data) == 0
names = ["setup + block constants", "rollout + adjoint", "residuals + reduce", "barrier weights, T1, PA, PB, Huu, Hux, Hxx", "Cholesky", "L^-1", "PA + PB (first half of assembly)", "vector sweeps (2 passes)", "steps in (s, z)", "(C) weights + rhs"]
tot = out.sum(axis=1).mean()
print(f"cond_N {cN} N {N} {gaits}: mean cycles per problem (thread 0): {tot:.0f}  kernel ms {info.total_solver_time * 1e3:.3f}  iters {info.acados_num_iter.mean():.2f}  {n / info.total_solver_time / 1e3:.1f} k/s")
for i, nm in enumerate(names):
    print(f"  {nm:44s} {out[:, i].mean():10.0f}  {100 * out[:, i].mean() / tot:5.1f}%")
