"""Dense <-> sparse trade-off on the GPU (the two ends of the reference's cond_N axis): throughput of the condensed
ADMM kernel (cond_N = 1 block, everything condensed) and of the Riccati IPM kernel (cond_N = N, fully sparse) over the horizon."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer
cfg = sequencer.GO2_MPC
n = 8192
print("| gait | N | free variables | condensed ADMM | sparse Riccati | ratio |\n|---|---|---|---|---|---|")
for gait in ("TROT", "STAND"):
    for N in (4, 6, 8, 10, 12, 16, 20):
        b = sequencer.random_batch(n, N, seed=100 + N, gaits=(gait,))
        nfree = int(3 * b["contact"].reshape(n, -1).sum(axis=1).max())
        res = {}
        for solver in ("CONDENSED_ADMM", "SPARSE_RICCATI"):
            mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"], max_batch=n)
            ts = []
            for _ in range(4):
                f, x, info = mpc.solve_records(b["rec"], b["contact"])
                ts.append(mpc.last_timing()[0])
            ok = float(info.success.mean())
            res[solver] = (n * ok / (np.median(ts[1:]) * 1e-3), ok)
            mpc.close()
        a, r = res["CONDENSED_ADMM"], res["SPARSE_RICCATI"]
        fa = f"{a[0] / 1e3:7.0f} k/s" if a[1] > 0.99 else f"does not fit (n > 156)"
        print(f"| {gait} | {N} | {nfree} | {fa} | {r[0] / 1e3:7.0f} k/s | {a[0] / r[0]:.1f}× |" if a[1] > 0.99 else f"| {gait} | {N} | {nfree} | {fa} | {r[0] / 1e3:7.0f} k/s | — |", flush=True)
