import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_rollout import _setup
n, periods = 96, 16
m, seq = _setup(n, seed=1)
s = seq.copy()
h = m.rollout(s, periods)
bad = np.nonzero(np.abs(h[-1, :, 2] - 0.30) > 0.03)[0]
print("bad robots", bad)
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for r in list(bad[:2]) + [0]:
    print("robot", r, "yaw0", 2 * np.arctan2(seq[r, 2], seq[r, 3]), "t0", seq[r, 25])
    for k in range(periods):
        print(k, "pos", h[k, r, 0:3], "vel", h[k, r, 3:6], "fz", h[k, r, 10], "st", h[k, r, 11:14])
