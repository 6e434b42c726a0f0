"""Wall-clock latency of one host-buffer call (pinned buffers) vs batch size, cold and hot start (DESIGN numbers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
N = 10
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
for n in (1, 16, 256, 1024, 4096):
    b = sequencer.random_batch(n, N, seed=5)
    rec, ct = pin(b["rec"]), pin(b["contact"])
    rec2 = pin(b["rec"] + 0.0)
    rec2[:, 10:13] += 0.01
    f, x = pin(np.zeros((n, N, 12))), pin(np.zeros((n, N + 1, 13)))
    row = [f"n={n:5d}"]
    for mode in (0, 2):
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", warm_start=mode,
                         horizon=N, dt=cfg["dt"], max_batch=n)
        for _ in range(5):
            mpc.solve_records(rec, ct, f, x)
        ts, ks = [], []
        for i in range(40):
            r = rec2 if i % 2 else rec  # alternate two nearby problems so that hot start has something to reuse
            t0 = time.perf_counter()
            _, _, info = mpc.solve_records(r, ct, f, x)
            ts.append(time.perf_counter() - t0)
            ks.append(info.total_solver_time)
        assert info.success.all()
        row.append(f"{'cold' if mode == 0 else 'hot '}: call {1e3 * np.median(ts):6.3f} ms (kernels {1e3 * np.median(ks):6.3f} ms)")
        mpc.close()
    print("  ".join(row), flush=True)
