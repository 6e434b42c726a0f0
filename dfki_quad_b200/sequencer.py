"""Batched host-side gait sequencer: reference trajectory + Raibert footholds for n robots at once (numpy, vectorised
over the batch; the per-step recursion of the reference is kept).

Mirrors  SimpleGaitSequencer::GetGaitSequence           simple_gait_sequencer.cpp:43-60
         MPCTrajectoryPlanner::plan_trajectory           mpc_trajectory_planner.cpp:28-140
         MPCTrajectoryPlanner::plan_desired_trajectory   mpc_trajectory_planner.cpp:142-264 (first call, no latching)
         RaibertFootStepPlanner::get_foot_position_sequence / get_foothold   raibert_foot_step_planner.cpp:43-169
for the Target activation of the node (mit_controller_node.cpp:145-165: hybrid_x_dot, hybrid_y_dot, wz, z, roll,
pitch, wx, wy active).  The contact schedule comes from the GPU (gait.contact_schedule, kernel K4).
This is the input generator of the benchmark; the QP itself never runs here.
"""
from __future__ import annotations

import numpy as np

from . import gait as gaitmod
from .mpc import euler_to_quaternion

GO2_MASS = 15.019  # quad_model_symbolic.cpp:15
GO2_INERTIA = np.array([[0.202755, 0.00012166, -0.0143742], [0.00012166, 0.490926, -3.12e-05],
                        [-0.0143742, -3.12e-05, 0.52697]])  # quad_model_symbolic.cpp:19
GO2_COM = np.zeros(3)  # quad_model_symbolic.cpp:30
GO2_SHOULDERS = np.array([[0.167, 0.1738, 0.0], [0.167, -0.1738, 0.0], [-0.197, 0.1738, 0.0],
                          [-0.197, -0.1738, 0.0]])  # mit_controller_sim_go2.yaml:33-38
GO2_MPC = dict(alpha=1e-5, mu=0.6, fmin=0.0, fmax=400.0, dt=0.05,
               state_weights_move=[20.0, 200.0, 50.0, 0.0001, 0.0001, 100.0, 0.2, 0.2, 0.8, 20.0, 20.0, 0.3],
               state_weights_stand=[30.0] * 6 + [1.0] * 6)  # mit_controller_sim_go2.yaml:23-32
RAIBERT_K = 0.03
RAIBERT_G = 9.81
MAX_LEG_LENGTH = 0.213 + 0.213 - 0.02  # raibert_foot_step_planner.cpp:36-40


def quat_to_rot(q):
    """Eigen toRotationMatrix, vectorised; q [..., 4] (x,y,z,w) -> R [..., 3, 3]."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz)
    R[..., 0, 1] = txy - twz
    R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz
    R[..., 1, 1] = 1 - (txx + tzz)
    R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy
    R[..., 2, 1] = tyz + twx
    R[..., 2, 2] = 1 - (txx + tyy)
    return R


def yaw_from_quaternion(q):
    R = quat_to_rot(q)
    return np.arctan2(R[..., 1, 0], R[..., 0, 0])


def quaternion_to_euler(q):
    """matrix_to_euler (quaternion_operations.hpp:106-124), regular branch + the two gimbal branches."""
    R = quat_to_rot(q)
    eps = np.finfo(float).eps
    m20 = R[..., 2, 0]
    pitch = -np.arcsin(np.clip(m20, -1, 1))
    yaw = np.arctan2(R[..., 1, 0], R[..., 0, 0])
    roll = np.arctan2(R[..., 2, 1], R[..., 2, 2])
    g1 = (1 - eps <= m20) & (m20 <= 1 + eps)
    g2 = (~g1) & (m20 <= -1 + eps)
    yaw = np.where(g1 | g2, 0.0, yaw)
    roll = np.where(g1, np.arctan2(-R[..., 0, 1], -R[..., 0, 2]), roll)
    roll = np.where(g2, np.arctan2(R[..., 0, 1], R[..., 0, 2]), roll)
    return np.stack([roll, pitch, yaw], axis=-1)


def _yaw_quat(yaw):
    z = np.zeros_like(yaw)
    return np.stack([z, z, np.sin(0.5 * yaw), np.cos(0.5 * yaw)], axis=-1)


def _slerp_half(a, b):
    """Eigen::Quaternion::slerp(0.5, b), vectorised."""
    d = np.sum(a * b, axis=-1)
    absd = np.abs(d)
    one = 1.0 - np.finfo(float).eps
    th = np.arccos(np.minimum(absd, 1.0))
    st = np.sin(th)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(absd >= one, 0.5, np.sin(0.5 * th) / st)
    s1 = np.where(d < 0, -s, s)
    return s[..., None] * a + s1[..., None] * b


def _rotate(q, v):
    return np.einsum("...ij,...j->...i", quat_to_rot(q), v)


def plan_trajectory(steps, dt, quat, pos, vel, twist, target, min_foot_height):
    """Batched plan_trajectory + plan_desired_trajectory.  target: dict of [n] arrays hybrid_x_dot, hybrid_y_dot,
    wz, wx, wy, roll, pitch, z."""
    n = quat.shape[0]
    out = {k: np.zeros((n, steps, d)) for k, d in
           (("ref_pos", 3), ("ref_quat", 4), ("ref_vel", 3), ("ref_twist", 3), ("des_pos", 3), ("des_quat", 4))}
    hyb = np.stack([target["hybrid_x_dot"], target["hybrid_y_dot"], np.zeros(n)], axis=-1)
    tw_cmd = np.stack([target["wx"], target["wy"], target["wz"]], axis=-1)
    for pref in ("ref", "des"):
        position = pos.copy()
        eul = quaternion_to_euler(quat)
        orientation = quat.copy()
        out[pref + "_pos"][:, 0] = position
        out[pref + "_quat"][:, 0] = orientation
        if pref == "ref":
            out["ref_vel"][:, 0] = vel
            out["ref_twist"][:, 0] = twist
        for i in range(1, steps):
            eul = eul + dt * tw_cmd
            eul[:, 1] = target["pitch"]
            eul[:, 0] = target["roll"]
            orientation = euler_to_quaternion(eul)
            prev = out[pref + "_quat"][:, i - 1]
            mid = _slerp_half(_yaw_quat(yaw_from_quaternion(prev)), _yaw_quat(yaw_from_quaternion(orientation)))
            position = position + _rotate(mid, dt * hyb)
            position[:, 2] = target["z"] + min_foot_height
            out[pref + "_pos"][:, i] = position
            out[pref + "_quat"][:, i] = orientation
            if pref == "ref":
                out["ref_vel"][:, i] = _rotate(orientation, hyb)
                out["ref_twist"][:, i] = tw_cmd
    return out


def raibert_foot_positions(steps, contact, traj, cur_foot_pos, v_filtered, t_stance, last_foot_heights,
                           shoulders=GO2_SHOULDERS):
    """Batched get_foot_position_sequence (z_on_plane = false).  contact [n,steps,4] uint8, cur_foot_pos [n,4,3]."""
    n = contact.shape[0]
    fp = np.zeros((n, steps, 4, 3))
    resume = np.ones((n, 4), dtype=bool)
    heights = np.array(last_foot_heights, dtype=float).copy()  # [n,4]
    L = MAX_LEG_LENGTH
    for i in range(steps):
        for leg in range(4):
            c = contact[:, i, leg] != 0
            if i == 0:
                fp[c, 0, leg] = cur_foot_pos[c, leg]
                heights[c, leg] = cur_foot_pos[c, leg, 2]
            else:
                keep = c & resume[:, leg]
                fp[keep, i, leg] = fp[keep, i - 1, leg]
                new = c & ~resume[:, leg]
                if new.any():
                    position = traj["ref_pos"][new, i]
                    orient = _yaw_quat(yaw_from_quaternion(traj["ref_quat"][new, i]))
                    v_cmd = traj["ref_vel"][new, i]
                    w_cmd = traj["ref_twist"][new, i]
                    v = v_filtered[new]
                    p_sh = position + _rotate(orient, np.broadcast_to(shoulders[leg], position.shape))
                    h = np.clip(position[:, 2] - heights[new].min(axis=1), 0.0, L)
                    p = p_sh + 0.5 * np.sqrt(h / RAIBERT_G)[:, None] * np.cross(v, w_cmd) \
                        + ((t_stance[new, leg] / 2.0)[:, None] * v + RAIBERT_K * (v - v_cmd))
                    p[:, 2] = heights[new, leg]
                    dif = p - p_sh
                    clip = (np.linalg.norm(dif, axis=1) > L) & (np.abs(dif[:, 2]) <= L)
                    if clip.any():
                        dx, dy, dz = dif[clip, 0], dif[clip, 1], dif[clip, 2]
                        num = -(dz - L) * (dz + L)
                        xbig = np.abs(dx) > np.abs(dy)
                        with np.errstate(divide="ignore", invalid="ignore"):
                            ndx = np.where(xbig, -np.sqrt(num / ((dy / dx) ** 2 + 1)), 0.0)
                            ndy = np.where(xbig, (dy / dx) * ndx, -np.sqrt(num / ((dx / dy) ** 2 + 1)))
                            ndx = np.where(xbig, ndx, (dx / dy) * ndy)
                        p[clip] = p_sh[clip] + np.stack([ndx, ndy, dz], axis=-1)
                    fp[new, i, leg] = p
            resume[:, leg] = c
    return fp


def random_batch(n, horizon, seed, gaits=("TROT",), dt=0.05, device=0, schedule_fn=None):
    """Synthetic randomised Go2 states / commands / gaits of SURVEY.md section 8d (config 2/3 distributions) ->
    dict with the packed input records, contact masks and the raw pieces.  Contact masks come from kernel K4
    (`schedule_fn(period, duty, offset, phase, steps, dt) -> (contact, swing)` overrides it for CPU-only tests)."""
    from .mpc import pack_records

    rng = np.random.Generator(np.random.PCG64(seed))
    N = horizon
    steps = N + 1
    roll = rng.uniform(-0.2, 0.2, n)
    pitch = rng.uniform(-0.2, 0.2, n)
    yaw = rng.uniform(-np.pi, np.pi, n)
    quat = euler_to_quaternion(np.stack([roll, pitch, yaw], axis=-1))
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.30 + rng.uniform(-0.03, 0.03, n)], axis=-1)
    omega = rng.uniform(-0.5, 0.5, (n, 3))
    vel = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.6, 0.6, n), rng.uniform(-0.1, 0.1, n)], axis=-1)
    target = dict(hybrid_x_dot=rng.uniform(-0.5, 0.5, n), hybrid_y_dot=rng.uniform(-0.5, 0.5, n),
                  wz=rng.uniform(-0.7, 0.7, n), wx=np.zeros(n), wy=np.zeros(n), roll=np.zeros(n), pitch=np.zeros(n),
                  z=np.full(n, 0.30))
    phase = rng.uniform(0.0, 1.0, n)
    gait_ids = np.array([gaitmod.GAIT_NAMES.index(g) for g in gaits])[rng.integers(0, len(gaits), n)]
    period, duty, offset = gaitmod.gait_arrays(gait_ids)
    Ryaw = quat_to_rot(_yaw_quat(yaw))
    feet = pos[:, None, :] + np.einsum("nij,lj->nli", Ryaw, GO2_SHOULDERS)
    feet[:, :, 0:2] += rng.uniform(-0.05, 0.05, (n, 4, 2))
    feet[:, :, 2] = 0.0
    # measured contacts = planned contacts at i = 0 -> the early-contact rule never fires
    if schedule_fn is None:
        contact, swing = gaitmod.contact_schedule(period, duty, offset, phase, steps, dt, device=device)
    else:
        contact, swing = schedule_fn(period, duty, offset, phase, steps, dt)
    min_h = feet[:, :, 2].min(axis=1)
    traj = plan_trajectory(steps, dt, quat, pos, vel, omega, target, min_h)
    t_stance = duty * period[:, None]
    fp = raibert_foot_positions(steps, contact, traj, feet, vel, t_stance, np.repeat(min_h[:, None], 4, axis=1))
    rec = pack_records(N, quat, pos, omega, vel, np.full(n, GO2_MASS), GO2_INERTIA, np.broadcast_to(GO2_COM, (n, 3)),
                       traj["des_quat"], traj["des_pos"], traj["ref_twist"], traj["ref_vel"], fp)
    seq_in = pack_sequencer_inputs(quat, pos, omega, vel, feet, vel, np.repeat(min_h[:, None], 4, axis=1), target, phase,
                                   period, duty, offset)
    return dict(rec=rec, contact=np.ascontiguousarray(contact[:, :N]), quat=quat, pos=pos, omega=omega, vel=vel,
                target=target, phase=phase, gait_ids=gait_ids, period=period, duty=duty, offset=offset, feet=feet,
                traj=traj, foot_pos=fp, swing_time=swing, seq_in=seq_in)


def random_seq_inputs(n, seed, gaits=("TROT",)):
    """The same synthetic state / command / gait distribution as random_batch, as sequencer inputs only ([n, 50], the 400 B
    per robot that b200qp_mpc_sequence_solve_batch takes): no host-side planning, cheap enough for million-problem sweeps."""
    rng = np.random.Generator(np.random.PCG64(seed))
    roll = rng.uniform(-0.2, 0.2, n)
    pitch = rng.uniform(-0.2, 0.2, n)
    yaw = rng.uniform(-np.pi, np.pi, n)
    quat = euler_to_quaternion(np.stack([roll, pitch, yaw], axis=-1))
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), 0.30 + rng.uniform(-0.03, 0.03, n)], axis=-1)
    omega = rng.uniform(-0.5, 0.5, (n, 3))
    vel = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.6, 0.6, n), rng.uniform(-0.1, 0.1, n)], axis=-1)
    target = dict(hybrid_x_dot=rng.uniform(-0.5, 0.5, n), hybrid_y_dot=rng.uniform(-0.5, 0.5, n),
                  wz=rng.uniform(-0.7, 0.7, n), wx=np.zeros(n), wy=np.zeros(n), roll=np.zeros(n), pitch=np.zeros(n),
                  z=np.full(n, 0.30))
    phase = rng.uniform(0.0, 1.0, n)
    gait_ids = np.array([gaitmod.GAIT_NAMES.index(g) for g in gaits])[rng.integers(0, len(gaits), n)]
    period, duty, offset = gaitmod.gait_arrays(gait_ids)
    Ryaw = quat_to_rot(_yaw_quat(yaw))
    feet = pos[:, None, :] + np.einsum("nij,lj->nli", Ryaw, GO2_SHOULDERS)
    feet[:, :, 0:2] += rng.uniform(-0.05, 0.05, (n, 4, 2))
    feet[:, :, 2] = 0.0
    return pack_sequencer_inputs(quat, pos, omega, vel, feet, vel, np.zeros((n, 4)), target, phase, period, duty, offset)


SEQ_IN_DOUBLES = 50  # B200QP_SEQ_IN_DOUBLES
TARGET_FIELDS = ("hybrid_x_dot", "hybrid_y_dot", "wz", "wx", "wy", "roll", "pitch", "z")


def pack_sequencer_inputs(quat, pos, omega, vel, feet, v_filtered, last_foot_heights, target, phase, period, duty, offset):
    """Per-robot input of the on-device sequencer (include/b200qp.h, B200QP_SEQ_IN_DOUBLES layout) -> [n, 50]."""
    n = len(phase)
    s = np.zeros((n, SEQ_IN_DOUBLES))
    s[:, 0:4] = quat
    s[:, 4:7] = pos
    s[:, 7:10] = omega
    s[:, 10:13] = vel
    s[:, 13:25] = np.asarray(feet).reshape(n, 12)
    s[:, 25:28] = v_filtered
    s[:, 28:32] = last_foot_heights
    for j, f in enumerate(TARGET_FIELDS):
        s[:, 32 + j] = target[f]
    s[:, 40] = phase
    s[:, 41] = period
    s[:, 42:46] = duty
    s[:, 46:50] = offset
    return s


def go2_seq_model():
    """b200qp_seq_model constants of the Go2 configuration (mit_controller_sim_go2.yaml / quad_model_symbolic.cpp)."""
    return dict(mass=GO2_MASS, inertia=np.asarray(GO2_INERTIA, float).reshape(3, 3), com=np.asarray(GO2_COM, float),
                shoulders=np.asarray(GO2_SHOULDERS, float), raibert_k=RAIBERT_K, raibert_g=RAIBERT_G,
                max_leg_length=MAX_LEG_LENGTH)
