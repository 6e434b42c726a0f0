import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N, n = 10, 4096
batch = sequencer.random_batch(n, N, seed=314, gaits=("TROT", "PACE", "WALKING_TROT"))
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N, dt=cfg["dt"], max_batch=n)
f_ref, x_ref, i_ref = mpc.solve_records(batch["rec"], batch["contact"])
f_ref2, _, _ = mpc.solve_records(batch["rec"], batch["contact"])
print("pageable twice equal:", np.array_equal(f_ref, f_ref2))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
rec_p, ct_p = pin(batch["rec"]), pin(batch["contact"])
f_p, x_p = pin(np.zeros((n, N, 12))), pin(np.zeros((n, N + 1, 13)))
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
for rep in range(30):
    f_p[:] = 0; x_p[:] = 0
    if mode == "pin_in":   # pinned inputs, pageable outputs
        f, x, info = mpc.solve_records(rec_p, ct_p)
    elif mode == "pin_out":  # pageable inputs, pinned outputs
        f, x, info = mpc.solve_records(batch["rec"], batch["contact"], f_p, x_p)
    else:
        f, x, info = mpc.solve_records(rec_p, ct_p, f_p, x_p)
    bad = np.nonzero((f != f_ref).reshape(n, -1).any(axis=1) | (x != x_ref).reshape(n, -1).any(axis=1))[0]
    if len(bad):
        p = bad[0]
        print(f"rep {rep} mode {mode}: {len(bad)} problems differ, first {bad[:8]}, max |df| {np.abs(f - f_ref).max():.3e} max |dx| {np.abs(x - x_ref).max():.3e} "
              f"iters {info.acados_num_iter[p]} vs {i_ref.acados_num_iter[p]} rounds {info.polish_rounds[p]} vs {i_ref.polish_rounds[p]} f zero rows {(f[p] == 0).all()}")
print("done", mode)
