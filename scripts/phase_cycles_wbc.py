"""Per-phase clock64 accounting of the WBC kernel (thread 0 of every block).  Needs a library built with -DB200QP_WBC_TIMING (it writes the
counters into x instead of the solution): make -C dfki_quad_b200/csrc EXTRA=-DB200QP_WBC_TIMING, B200QP_LIB=<that build>."""
import sys
import numpy as np
sys.path.insert(0, ".")
from dfki_quad_b200 import wbc
B = 16384
b = wbc.random_wbc_batch(B, seed=20261023)
s = wbc.BatchedWBCSolver(max_batch=B)
x, info = s.solve(b["qp"])
x, info = s.solve(b["qp"])
c = x[:, :10].mean(axis=0)
names = ["load + AL form", "residuals + bookkeeping", "Phi build", "Cholesky + inverse (Phi)", "Y, S", "Cholesky + inverse (S)", "start point (iteration 0)",
         "predictor / corrector (solves, refinement, steps)", "-", "-"]
tot = c.sum()
print(f"mean cycles per QP (thread 0): {tot:.0f}; iterations {info['iterations'].mean():.2f}")
for n, v in zip(names, c):
    if v > 0:
        print(f"  {n:52s} {v:10.0f}  {100 * v / tot:5.1f}%")
