"""Closed-loop cost of the three warm-start modes: 4096 trotting robots, 60 control periods (development / DESIGN numbers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, rollout
cfg = sequencer.GO2_MPC
n, steps = 4096, 60
tgt = dict(hybrid_x_dot=0.4, hybrid_y_dot=0.1, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
for mode in (0, 1, 2):
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", warm_start=mode,
                     horizon=10, dt=cfg["dt"], max_batch=n)
    mpc.set_model(**sequencer.go2_seq_model())
    h = rollout.closed_loop(mpc, n, steps, tgt, gait="TROT", seed=3, yaw0=0.0)
    print(f"warm_start {mode}: ok {h['ok'].mean():.4f}  ADMM iters {h['iters'][5:].mean():5.1f}  polish rounds {h['rounds'][5:].mean():4.2f}  "
          f"no-ADMM fraction {(h['iters'][5:] == 0).mean():.3f}  kernel {np.mean(h['kernel_ms'][5:]):.3f} ms/step "
          f"= {n / np.mean(h['kernel_ms'][5:]) / 1e3:.2f} M solves/s  final |v_x - 0.4| {np.abs(h['vel'][-10:, :, 0].mean(axis=0) - 0.4).max():.3f}")
