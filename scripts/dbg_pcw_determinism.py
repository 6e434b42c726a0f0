"""Does the warp Riccati kernel give the same answer for the same batch regardless of what ran before it (stale scratch)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, gait as gaitmod
cfg = sequencer.GO2_MPC
rng = np.random.default_rng(7)
n = 256
cases = [(10, ("TRAVERSE_GALLOP", "BOUND"), 1.2, 5.0, 250.0, 5.2e-4, 1000 + 24),
         (16, ("TROT", "PACE", "BOUND"), 0.6, 0.0, 400.0, 1e-5, 3)]
ref = {}
for rep in range(int(os.environ.get("REPS", "30"))):
    # something else first: changes what the freshly allocated scratch contains
    N0 = int(rng.choice([5, 8, 13, 16, 20]))
    g0 = tuple(rng.choice(gaitmod.GAIT_NAMES, size=2, replace=False))
    b0 = sequencer.random_batch(n, N0, seed=int(rng.integers(0, 1000)), gaits=g0)
    m0 = BatchedMPC(10 ** rng.uniform(-6, -3), np.array(cfg["state_weights_move"]) * np.exp(rng.normal(0, 1, 12)), 0.6, 0.0, 400.0,
                    str(rng.choice(["CONDENSED_ADMM", "SPARSE_RICCATI", "AUTO"])), horizon=N0, dt=cfg["dt"], max_batch=n)
    m0.solve_records(b0["rec"], b0["contact"])
    m0.close()
    for ci, (N, gaits, mu, fmin, fmax, alpha, seed) in enumerate(cases):
        b = sequencer.random_batch(n, N, seed=seed, gaits=gaits)
        w = np.array(cfg["state_weights_move"], float)
        m = BatchedMPC(alpha, w, mu, fmin, fmax, "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=n)
        f, x, info = m.solve_records(b["rec"], b["contact"])
        m.close()
        key = (f.tobytes(), info.return_code.tobytes(), info.acados_num_iter.tobytes())
        if ci not in ref:
            ref[ci] = (key, f.copy(), info.return_code.copy(), info.acados_num_iter.copy())
            print(f"case {ci}: ok {info.success.mean():.4f} iterations {info.acados_num_iter.mean():.2f}", flush=True)
        elif key != ref[ci][0]:
            d = np.nonzero((f != ref[ci][1]).reshape(n, -1).any(axis=1) | (info.return_code != ref[ci][2]))[0]
            print(f"rep {rep} case {ci}: DIFFERS in problems {d[:8]} status {info.return_code[d[:8]]} (first run {ref[ci][2][d[:8]]}) "
                  f"iterations {info.acados_num_iter[d[:8]]} vs {ref[ci][3][d[:8]]} max |df| {np.abs(f - ref[ci][1]).max():.3e}", flush=True)
print("done")
