import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import wbc
n = 16384
b = wbc.random_wbc_batch(n, seed=20261019 + 4)
x, info = wbc.BatchedWBCSolver(max_batch=n).solve(b["qp"])
bad = np.nonzero(info["status"] != 0)[0]
print("fail", len(bad), np.unique(info["status"], return_counts=True), "iters mean", info["iterations"].mean(), "max", info["iterations"].max())
for p in bad[:12]:
    print(p, info["status"][p], info["iterations"][p], info["kkt"][p], "|g| %.0f" % np.abs(b["qp"].g[p]).max(), "contacts", b["contacts"][p])
np.savez("gpurun_out/wbc_fail.npz", idx=bad, H=b["qp"].H[bad[:8]], g=b["qp"].g[bad[:8]], A=b["qp"].A[bad[:8]], lo=b["qp"].lower_y[bad[:8]], hi=b["qp"].upper_y[bad[:8]])
