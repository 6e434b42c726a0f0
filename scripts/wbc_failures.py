"""Dumps the WBC QPs (BASELINE config 4 distribution) that the device IPM does not bring to its tolerance, for offline study.
Usage (GPU box): python scripts/wbc_failures.py [batches]  -> gpurun_out/wbc_fail.npz"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfki_quad_b200 import wbc  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 16384
keep = {k: [] for k in ("H", "g", "A", "lo", "hi", "status", "iters", "kkt", "seed", "index")}
solver = wbc.BatchedWBCSolver(max_batch=B)
tot = 0
for s in range(nb):
    seed = 20261019 + 4 + 1000 * s
    b = wbc.random_wbc_batch(B, seed=seed)
    H, g, A, lo, hi = solver.abi_arrays(b["qp"])
    x, info = solver.solve_abi(H, g, A, lo, hi)
    st = np.asarray(info["status"])
    bad = np.nonzero(st != 0)[0]
    tot += bad.size
    print(f"seed {seed}: {bad.size} of {B} not converged; status {np.bincount(st)}; iterations mean {np.mean(info["iterations"]):.2f} max {np.max(info["iterations"])}, {int((np.asarray(info["iterations"]) > 20).sum())} second attempts; kkt max {np.max(np.asarray(info["kkt"])[st == 0]):.2e}",
          flush=True)
    good = np.nonzero(st == 0)[0][:8]
    for i in list(bad) + list(good):
        for k, v in (("H", H[i]), ("g", g[i]), ("A", A[i]), ("lo", lo[i]), ("hi", hi[i]), ("status", st[i]),
                     ("iters", info["iterations"][i]), ("kkt", info["kkt"][i]), ("seed", seed), ("index", i)):
            keep[k].append(v)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/wbc_fail.npz", **{k: np.array(v) for k, v in keep.items()})
print("not converged in total:", tot, "of", nb * B)
