"""Bit-exact repeatability of every kernel path: each configuration is solved 20 times and compared with the first solve."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, wbc
cfg = sequencer.GO2_MPC
REPS = int(os.environ.get("REPS", "20"))
def mk(N, n, solver, **o):
    return BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"], max_batch=n, **o)
cases = [("ADMM trot N=10", 10, 4096, "CONDENSED_ADMM", ("TROT",)),
         ("ADMM 8 gaits N=10", 10, 4096, "CONDENSED_ADMM", ("STAND", "STATIC_WALK", "WALKING_TROT", "TROT", "FLYING_TROT", "PACE", "BOUND", "PRONK")),
         ("ADMM N=16 mixed", 16, 4096, "CONDENSED_ADMM", ("TROT", "PACE", "BOUND")),
         ("ADMM N=13 stand (156 bin)", 13, 1024, "CONDENSED_ADMM", ("STAND",)),
         ("Riccati N=16 mixed", 16, 4096, "SPARSE_RICCATI", ("TROT", "PACE", "BOUND")),
         ("AUTO N=16 incl. stand", 16, 2048, "AUTO", ("TROT", "STAND", "STATIC_WALK"))]
for name, N, n, solver, gaits in cases:
    b = sequencer.random_batch(n, N, seed=99, gaits=gaits)
    mpc = mk(N, n, solver)
    f0, x0, i0 = mpc.solve_records(b["rec"], b["contact"])
    bad = 0
    for rep in range(REPS):
        f, x, i = mpc.solve_records(b["rec"], b["contact"])
        d = int(((f != f0).reshape(n, -1).any(axis=1) | (x != x0).reshape(n, -1).any(axis=1) | (i.return_code != i0.return_code)).sum())
        bad += d
    print(f"{name:28s} ok {i0.success.mean():.4f}  differing problem-solves in {REPS} repetitions: {bad}", flush=True)
    mpc.close()
# sequencer + solve, hot start (second and later solves all start from the same stored solution of the previous call)
b = sequencer.random_batch(2048, 10, seed=5, gaits=("TROT", "WALKING_TROT"))
mpc = mk(10, 2048, "CONDENSED_ADMM")
mpc.set_model(**sequencer.go2_seq_model())
f0, x0, i0 = mpc.sequence_and_solve(b["seq_in"])
bad = sum(int((mpc.sequence_and_solve(b["seq_in"])[0] != f0).reshape(2048, -1).any(axis=1).sum()) for _ in range(REPS))
print(f"{'sequencer + ADMM':28s} differing: {bad}", flush=True)
# WBC
bw = wbc.random_wbc_batch(4096, seed=3)
solver = wbc.BatchedWBCSolver(max_batch=4096)
H, g, A, lo, hi = solver.abi_arrays(bw["qp"])
x0, inf0 = solver.solve_abi(H, g, A, lo, hi)
x0 = x0.copy()
bad = 0
for _ in range(REPS):
    x, inf = solver.solve_abi(H, g, A, lo, hi)
    bad += int(((x != x0).any(axis=1) | (inf["status"] != inf0["status"])).sum())
print(f"{'WBC':28s} differing: {bad}", flush=True)
