"""Scalar (one robot at a time) restatement of the reference's gait-sequencer input side.  TEST INFRASTRUCTURE ONLY
(see oracle/mpc_oracle.py header).  Pinned on the compiled reference sequencer (oracle/_ref, tests/test_ref.py: records 2e-16 on the
first call); the multi-call behaviour is checked directly against live reference objects (tests/seq_cases.py).

Follows  SimpleGaitSequencer::GetGaitSequence   simple_gait_sequencer.cpp:43-60
         MPCTrajectoryPlanner::plan_trajectory  mpc_trajectory_planner.cpp:28-140 (+ plan_desired_trajectory :142-264,
                                                first call: sequence_initialized_ == false so no latching)
         RaibertFootStepPlanner                 raibert_foot_step_planner.cpp:43-169
for the Target activation the node uses (mit_controller_node.cpp:145-165): hybrid_x_dot, hybrid_y_dot, wz, z,
roll, pitch, wx, wy active; yaw, x, y, z_dot inactive.
"""
from __future__ import annotations

import math
import numpy as np

from . import mpc_oracle as mo

RAIBERT_K = 0.03  # mit_controller_sim_go2.yaml: raibert.k
RAIBERT_G = 9.81  # raibert_foot_step_planner.hpp default g
# FK(q=0).z = -(l1 + l2) = -0.426 -> max_leg_length = 0.426 - 0.02 (raibert_foot_step_planner.cpp:36-40)
MAX_LEG_LENGTH = 0.213 + 0.213 - 0.02


def yaw_quat(yaw):
    return np.array([0.0, 0.0, math.sin(0.5 * yaw), math.cos(0.5 * yaw)])


def slerp_half(a, b):
    """Eigen::Quaternion::slerp(0.5, other)."""
    d = float(a @ b)
    absd = abs(d)
    one = 1.0 - np.finfo(float).eps
    if absd >= one:
        s0, s1 = 0.5, 0.5
    else:
        th = math.acos(absd)
        st = math.sin(th)
        s0 = math.sin(0.5 * th) / st
        s1 = math.sin(0.5 * th) / st
    if d < 0:
        s1 = -s1
    return s0 * a + s1 * b


def rotate(q, v):
    return mo.quat_to_rot(q) @ v


def plan_trajectory(steps, dt, quat, pos, vel, twist, tgt, min_foot_height):
    """Returns dict of reference_* and desired_* arrays of length `steps` (rows 0..steps-1)."""
    position = np.array(pos, dtype=float)
    orientation = np.array(quat, dtype=float)
    velocity = np.array(vel, dtype=float)
    tw = np.array(twist, dtype=float)
    eul = mo.quaternion_to_euler(orientation)
    ref_pos, ref_quat, ref_vel, ref_tw = [position.copy()], [orientation.copy()], [velocity.copy()], [tw.copy()]
    for i in range(1, steps):
        tw = np.array([tgt["wx"], tgt["wy"], tgt["wz"]])
        eul = eul + dt * tw
        eul[1] = tgt["pitch"]
        eul[0] = tgt["roll"]
        orientation = mo.euler_to_quaternion(eul)
        hyb = np.array([tgt["hybrid_x_dot"], tgt["hybrid_y_dot"], 0.0])
        velocity = rotate(orientation, hyb)
        mid = slerp_half(yaw_quat(mo.yaw_from_quaternion(ref_quat[i - 1])), yaw_quat(mo.yaw_from_quaternion(orientation)))
        position = position + rotate(mid, dt * hyb)
        position[2] = tgt["z"] + min_foot_height
        ref_pos.append(position.copy())
        ref_quat.append(orientation.copy())
        ref_vel.append(velocity.copy())
        ref_tw.append(tw.copy())
    # plan_desired_trajectory, first call (no latching): same recursion driven by the stored twists/velocities
    position = np.array(pos, dtype=float)
    eul = mo.quaternion_to_euler(np.array(quat, dtype=float))
    des_pos, des_quat = [position.copy()], [np.array(quat, dtype=float)]
    for i in range(1, steps):
        eul = eul + dt * ref_tw[i]
        eul[1] = tgt["pitch"]
        eul[0] = tgt["roll"]
        orientation = mo.euler_to_quaternion(eul)
        hyb = np.array([tgt["hybrid_x_dot"], tgt["hybrid_y_dot"], 0.0])
        mid = slerp_half(yaw_quat(mo.yaw_from_quaternion(des_quat[i - 1])), yaw_quat(mo.yaw_from_quaternion(orientation)))
        position = position + rotate(mid, dt * hyb)
        position[2] = tgt["z"] + min_foot_height
        des_pos.append(position.copy())
        des_quat.append(orientation.copy())
    return dict(ref_pos=np.array(ref_pos), ref_quat=np.array(ref_quat), ref_vel=np.array(ref_vel), ref_twist=np.array(ref_tw),
                des_pos=np.array(des_pos), des_quat=np.array(des_quat))


def raibert_foot_positions(steps, contact, traj, cur_foot_pos, v_filtered, t_stance, last_foot_heights, shoulders=mo.GO2_SHOULDERS):
    """get_foot_position_sequence with z_on_plane = false.  contact: [steps,4]; cur_foot_pos: [4,3] world."""
    fp = np.zeros((steps, 4, 3))
    resume = [True] * 4
    heights = list(last_foot_heights)
    for i in range(steps):
        for leg in range(4):
            if contact[i][leg]:
                if i == 0:
                    fp[i, leg] = cur_foot_pos[leg]
                    heights[leg] = cur_foot_pos[leg][2]
                elif resume[leg]:
                    fp[i, leg] = fp[i - 1, leg]
                else:
                    position = traj["ref_pos"][i]
                    orient = yaw_quat(mo.yaw_from_quaternion(traj["ref_quat"][i]))
                    v_cmd = traj["ref_vel"][i]
                    w_cmd = traj["ref_twist"][i]
                    p_sh = position + rotate(orient, shoulders[leg])
                    h = position[2] - min(heights)
                    h = min(max(h, 0.0), MAX_LEG_LENGTH)
                    p = p_sh + 0.5 * math.sqrt(h / RAIBERT_G) * np.cross(v_filtered, w_cmd) \
                        + ((t_stance[leg] / 2.0) * v_filtered + RAIBERT_K * (v_filtered - v_cmd))
                    p[2] = heights[leg]
                    dif = p - p_sh
                    if np.linalg.norm(dif) > MAX_LEG_LENGTH and abs(dif[2]) <= MAX_LEG_LENGTH:
                        nd = np.zeros(3)
                        if abs(dif[0]) > abs(dif[1]):
                            nd[0] = -math.sqrt(-(dif[2] - MAX_LEG_LENGTH) * (dif[2] + MAX_LEG_LENGTH) / ((dif[1] / dif[0]) ** 2 + 1))
                            nd[1] = (dif[1] / dif[0]) * nd[0]
                        else:
                            nd[1] = -math.sqrt(-(dif[2] - MAX_LEG_LENGTH) * (dif[2] + MAX_LEG_LENGTH) / ((dif[0] / dif[1]) ** 2 + 1))
                            nd[0] = (dif[0] / dif[1]) * nd[1]
                        nd[2] = dif[2]
                        p = p_sh + nd
                    fp[i, leg] = p
                resume[leg] = True
            else:
                resume[leg] = False
    return fp


def make_problem(N, gait, phase, state, tgt, cur_foot_pos, params, measured_contacts=None):
    """One full SimpleGaitSequencer pass -> (record, contacts[N,4]).  state: dict quat,pos,omega,vel."""
    T, duty, off = mo.GAITS[gait]
    dt = params["dt"]
    steps = N + 1
    contact, _ = mo.gait_update_sequence(T, duty, off, dt, phase, steps, measured_contacts)
    min_h = float(min(cur_foot_pos[l][2] for l in range(4)))
    traj = plan_trajectory(steps, dt, state["quat"], state["pos"], state["vel"], state["omega"], tgt, min_h)
    t_stance = [duty * T] * 4
    fp = raibert_foot_positions(steps, contact, traj, cur_foot_pos, np.asarray(state["vel"], float), t_stance, [min_h] * 4)
    rec = mo.pack_problem(N, state["quat"], state["pos"], state["omega"], state["vel"], mo.GO2_MASS, mo.GO2_INERTIA,
                          mo.GO2_COM, traj["des_quat"], traj["des_pos"], traj["ref_twist"], traj["ref_vel"], fp)
    return rec, contact[:N]


def random_problem(rng, N, gait="TROT", params=None):
    """Synthetic state/command distribution of SURVEY.md section 8d config 2."""
    roll, pitch = rng.uniform(-0.2, 0.2, 2)
    yaw = rng.uniform(-math.pi, math.pi)
    quat = mo.euler_to_quaternion([roll, pitch, yaw])
    pos = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), 0.30 + rng.uniform(-0.03, 0.03)])
    omega = rng.uniform(-0.5, 0.5, 3)
    vel = np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.1, 0.1)])
    tgt = dict(hybrid_x_dot=rng.uniform(-0.5, 0.5), hybrid_y_dot=rng.uniform(-0.5, 0.5), wz=rng.uniform(-0.7, 0.7),
               wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    phase = rng.uniform(0.0, 1.0)
    Ryaw = mo.quat_to_rot(yaw_quat(yaw))
    feet = np.array([pos + Ryaw @ mo.GO2_SHOULDERS[l] for l in range(4)])
    feet[:, 0:2] += rng.uniform(-0.05, 0.05, (4, 2))
    feet[:, 2] = 0.0
    state = dict(quat=quat, pos=pos, omega=omega, vel=vel)
    return make_problem(N, gait, phase, state, tgt, feet, params or mo.GO2_PARAMS), dict(state=state, tgt=tgt, phase=phase, feet=feet)
