import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, rollout
cfg = sequencer.GO2_MPC
n = 4
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=10, dt=cfg["dt"], max_batch=n)
mpc.set_model(**sequencer.go2_seq_model())
gait = sys.argv[1] if len(sys.argv) > 1 else "STAND"
vx = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
tgt = dict(hybrid_x_dot=vx, hybrid_y_dot=0.0, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
h = rollout.closed_loop(mpc, n, 60, tgt, gait=gait, seed=3, yaw0=(0.0 if len(sys.argv) < 4 else None), plant_inertia=("physical" if len(sys.argv) < 4 else sys.argv[3]))
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for t in range(0, 61, 4):
    print(t, "pos", h["pos"][t, 0], "vel", h["vel"][t, 0], "q", h["quat"][t, 0], "fz", h["fz"][t - 1, 0] if t else None, "ok", h["ok"][t - 1, 0] if t else None)
