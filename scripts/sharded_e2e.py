"""One host array, all GPUs of the box, ONE library call (b200qp_mpc_solve_batch_sharded): end-to-end solves/s with pinned host
buffers, against G x the single-GPU end-to-end rate of the same call.  usage: sharded_e2e.py [problems_per_gpu]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from dfki_quad_b200 import ShardedMPC, _lib, sequencer  # noqa: E402

per = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
G = _lib.load().b200qp_device_count()
N = 10
cfg = sequencer.GO2_MPC
base = sequencer.random_batch(8192, N, seed=5, gaits=("TROT",))
res = {}
for g in sorted({1, G}):
    n = per * g
    reps = -(-n // 8192)
    rec = torch.from_numpy(np.tile(base["rec"], (reps, 1))[:n].copy()).pin_memory().numpy()
    rec[:, 4:6] += np.random.default_rng(1).uniform(-1e-3, 1e-3, (n, 2))  # distinct problems
    ct = torch.from_numpy(np.tile(base["contact"], (reps, 1, 1))[:n].copy()).pin_memory().numpy()
    f = torch.zeros((n, N, 12), dtype=torch.float64).pin_memory().numpy()
    x = torch.zeros((n, N + 1, 13), dtype=torch.float64).pin_memory().numpy()
    sh = ShardedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N, dt=cfg["dt"],
                    max_batch=n, devices=tuple(range(g)))
    sh.solve_records(rec, ct, f, x)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        _, _, info = sh.solve_records(rec, ct, f, x)
        best = min(best, time.perf_counter() - t0)
    ok = (info["status"] == 0) & (info["kkt"].max(axis=1) <= 1e-6)
    res[g] = ok.sum() / best
    print(f"{g} GPU(s): {n} problems in one call, {best * 1e3:.2f} ms, {res[g] / 1e6:.3f} M converged solves/s end to end")
    sh.close()
if G > 1:
    print(f"efficiency vs {G} x the 1-GPU e2e rate: {res[G] / (G * res[1]):.3f}")
