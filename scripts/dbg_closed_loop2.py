import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer as sq, rollout, gait as gaitmod
from dfki_quad_b200.mpc import euler_to_quaternion
cfg = sq.GO2_MPC
n = 2
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=10, dt=cfg["dt"], max_batch=n)
mpc.set_model(**sq.go2_seq_model())
np.set_printoptions(precision=5, suppress=True, linewidth=220)
quat = euler_to_quaternion(np.array([[0.05, -0.03, 0.4], [0.0, 0.0, 0.0]]))
pos = np.array([[0.1, 0.2, 0.30], [0, 0, 0.3]]); omega = np.array([[0.1, 0.2, -0.1], [0, 0, 0.0]]); vel = np.array([[0.1, -0.1, 0.05], [0, 0, 0.0]])
yaw = sq.yaw_from_quaternion(quat)
feet = pos[:, None, :] + np.einsum("nij,lj->nli", sq.quat_to_rot(sq._yaw_quat(yaw)), sq.GO2_SHOULDERS); feet[:, :, 2] = 0
gid = np.full(n, gaitmod.GAIT_NAMES.index("STAND")); period, duty, offset = gaitmod.gait_arrays(gid)
tgt = {k: np.zeros(n) for k in sq.TARGET_FIELDS}; tgt["z"][:] = 0.3
seq_in = sq.pack_sequencer_inputs(quat, pos, omega, vel, feet, vel, np.zeros((n, 4)), tgt, np.zeros(n), period, duty, offset)
forces, xpred, info = mpc.sequence_and_solve(seq_in)
print("ok", info.success, "f0", forces[0, 0].reshape(4, 3))
q2, p2, w2, v2 = rollout.srbd_step(quat, pos, omega, vel, feet, forces[:, 0].reshape(n, 4, 3), 0.05, substeps=50)
from oracle import mpc_oracle as mo
print("pred x1 (rpy,p,w,v):", xpred[0, 1])
print("plant: rpy", mo.quaternion_to_euler(q2[0]), "p", p2[0], "w", w2[0], "v", v2[0])
print("x0:", xpred[0, 0])
