"""Closed-loop rollout harness: a batch of single-rigid-body robots driven by the MPC path in
the loop, the way MITController drives one robot - every control period the gait sequencer is re-planned from the measured
state, the QP is solved and the first-stage ground-reaction forces are applied (mit_controller_node.cpp MPC thread:
UpdateState -> GetGaitSequence -> UpdateGaitSequence -> GetWrenchSequence).  The plant here is the same single rigid body
the MPC predicts with, integrated with small explicit steps on the host (numpy); it stands in for the simulator and is not
part of the accelerated path.  Feet are massless points: a stance foot stays where it touched down, a swing foot lands on
the foothold the Raibert planner (on the device) chose for its touch-down stage."""
from __future__ import annotations

import numpy as np

from . import gait as gaitmod
from . import sequencer as sq

GRAVITY = np.array([0.0, 0.0, -9.8067])  # MPC::GRAVITY_CONSTANT, mpc.hpp:22


def _quat_mul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def srbd_step(quat, pos, omega, vel, feet, forces, dt, substeps=10, mass=sq.GO2_MASS, inertia=sq.GO2_INERTIA,
              plant_inertia="physical"):
    """Integrate m v' = sum f + m g, Iw w' = sum (r - p) x f - w x Iw w, q' = 0.5 [w,0] q over dt with constant forces.
    quat [n,4] xyzw, feet/forces [n,4,3] (world).  Semi-implicit Euler, `substeps` per call.
    plant_inertia: "physical" -> Iw = R Ib R^T;  "reference_model" -> Iw = Rz(psi)^T Ib Rz(psi), the world inertia the
    reference's MPC model uses (MPC::GetIWorld, mpc.cpp:60-64, builds it from GetR(psi) = Rz(psi)^T, i.e. rotated the other
    way; reproduced in the kernels for parity).  With Go2's Ixx = 0.20 vs Iyy = 0.49 the two differ by a 2 psi rotation of
    the inertia ellipsoid, so a physical plant is only consistent with the MPC's prediction for headings near zero."""
    h = dt / substeps
    Ib = np.asarray(inertia, float)
    for _ in range(substeps):
        R = sq.quat_to_rot(quat)
        if plant_inertia == "reference_model":
            Rz = sq.quat_to_rot(sq._yaw_quat(sq.yaw_from_quaternion(quat)))
            Iw = np.swapaxes(Rz, 1, 2) @ Ib @ Rz
        else:
            Iw = R @ Ib @ np.swapaxes(R, 1, 2)
        tau = np.cross(feet - pos[:, None, :], forces).sum(axis=1)
        Iw_w = np.einsum("nij,nj->ni", Iw, omega)
        wdot = np.linalg.solve(Iw, (tau - np.cross(omega, Iw_w))[..., None])[..., 0]
        vdot = forces.sum(axis=1) / mass + GRAVITY
        omega = omega + h * wdot
        vel = vel + h * vdot
        pos = pos + h * vel
        dq = _quat_mul(np.concatenate([omega, np.zeros((len(omega), 1))], axis=1), quat)
        quat = quat + 0.5 * h * dq
        quat = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    return quat, pos, omega, vel


def closed_loop(mpc, n, steps, target, gait="TROT", seed=0, dt=0.05, v_filter_len=20, yaw0=None, plant_inertia="physical"):
    """Run `steps` control periods of length dt (= the MPC stage length) for n robots from a standing start.
    target: dict of the 8 active QuadControlTarget fields (scalars or [n]).  Returns the state history and solver stats."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = mpc.N
    gid = np.full(n, gaitmod.GAIT_NAMES.index(gait))
    period, duty, offset = gaitmod.gait_arrays(gid)
    yaw = rng.uniform(-np.pi, np.pi, n) if yaw0 is None else np.broadcast_to(np.asarray(yaw0, float), (n,)).copy()
    quat = sq._yaw_quat(yaw)
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.full(n, 0.30)], axis=-1)
    omega, vel = np.zeros((n, 3)), np.zeros((n, 3))
    Ryaw = sq.quat_to_rot(quat)
    feet = pos[:, None, :] + np.einsum("nij,lj->nli", Ryaw, sq.GO2_SHOULDERS)
    feet[:, :, 2] = 0.0
    heights = np.zeros((n, 4))
    tgt = {k: np.broadcast_to(np.asarray(target[k], float), (n,)).copy() for k in sq.TARGET_FIELDS}
    # SimpleGaitSequencer::UpdateState (simple_gait_sequencer.cpp:62-80): phase = fmod((t - t0) / T, 1), 0 for a pure standing gait;
    # every robot gets its own t0 so that the batch covers all phases
    t_since_t0 = rng.uniform(0.0, 1.0, n) * period
    phase = gaitmod.sequencer_phase(t_since_t0, period, duty)
    vhist = [vel.copy()]
    prev_contact = None
    hist = dict(pos=[pos.copy()], vel=[vel.copy()], quat=[quat.copy()], ok=[], fz=[], iters=[], rounds=[], kernel_ms=[])
    for _ in range(steps):
        vf = np.mean(vhist[-v_filter_len:], axis=0)
        seq_in = sq.pack_sequencer_inputs(quat, pos, omega, vel, feet, vf, heights, tgt, phase, period, duty, offset)
        rec, contact = mpc.sequence_records(seq_in)  # plan (needed below for the footholds) ...
        forces, _, info = mpc.sequence_and_solve(seq_in)  # ... and the fused sequencer + QP call that produces the GRFs
        ok = info.success
        f0 = np.where(ok[:, None, None], forces[:, 0].reshape(n, 4, 3), 0.0)
        c0 = contact[:, 0].astype(bool)
        # touch-down: a foot that was swinging and is planned in stance now sits on the foothold planned for this stage
        fp0 = rec[:, 26 + 13 * (N + 1): 26 + 13 * (N + 1) + 12].reshape(n, 4, 3)
        feet = np.where(c0[:, :, None], fp0, feet)
        if prev_contact is not None:
            landed = c0 & ~prev_contact
            heights = np.where(landed, feet[:, :, 2], heights)
        # swing feet travel with the hips towards the next planned foothold (first stage at which the leg is in stance)
        fp_all = rec[:, 26 + 13 * (N + 1):].reshape(n, N, 4, 3)
        first = np.argmax(contact.astype(bool), axis=1)  # [n,4] first stance stage (0 if never)
        nxt = np.take_along_axis(fp_all, first[:, None, :, None], axis=1)[:, 0]
        has = contact.astype(bool).any(axis=1)
        feet = np.where((~c0 & has)[:, :, None], nxt, feet)
        quat, pos, omega, vel = srbd_step(quat, pos, omega, vel, feet, f0 * c0[:, :, None], dt, plant_inertia=plant_inertia)
        t_since_t0 = t_since_t0 + dt
        phase = gaitmod.sequencer_phase(t_since_t0, period, duty)
        prev_contact = c0
        vhist.append(vel.copy())
        hist["pos"].append(pos.copy()); hist["vel"].append(vel.copy()); hist["quat"].append(quat.copy())
        hist["ok"].append(ok.copy()); hist["fz"].append(f0[:, :, 2].sum(axis=1)); hist["iters"].append(info.acados_num_iter.copy())
        hist["rounds"].append(info.polish_rounds.copy()); hist["kernel_ms"].append(info.total_solver_time * 1e3)
    return {k: np.array(v) for k, v in hist.items()}


# ---- host-stepped mirror of the device loop (b200qp_mpc_rollout) --------------------------------------------------------------
ABAD = np.array([0.1934, 0.0465])
HIP_Y, L1, L2 = 0.0955, 0.213, 0.213


def leg_ik(quat, pos, feet):
    """Closed-form inverse kinematics of the Go2 legs: foot positions (world) -> joint angles [n,12] and leg Jacobians (world) [n,4,3,3]
    (what rollout_wbc_target_kernel computes; checked against go2_model.dynamics' forward kinematics in the tests)."""
    n = pos.shape[0]
    R = sq.quat_to_rot(quat)
    q = np.zeros((n, 12))
    J = np.zeros((n, 4, 3, 3))
    for leg in range(4):
        sx, sy = (1.0 if leg < 2 else -1.0), (1.0 if leg % 2 == 0 else -1.0)
        rw = feet[:, leg] - pos
        b = np.einsum("nji,nj->ni", R, rw) - np.array([sx * ABAD[0], sy * ABAD[1], 0.0])
        bx, by, bz = b[:, 0], b[:, 1], b[:, 2]
        l0 = sy * HIP_Y
        hh = np.sqrt(np.maximum(by * by + bz * bz - l0 * l0, 1e-12))
        q0 = np.arctan2(hh * by + l0 * bz, l0 * by - hh * bz)
        d2 = np.clip(bx * bx + hh * hh, 1e-6, (L1 + L2) ** 2 * (1.0 - 1e-9))
        ck = np.clip((d2 - L1 * L1 - L2 * L2) / (2.0 * L1 * L2), -1.0, 1.0)
        q2 = -np.arccos(ck)
        q1 = np.arctan2(-bx, hh) - np.arctan2(L2 * np.sin(q2), L1 + L2 * np.cos(q2))
        q[:, 3 * leg], q[:, 3 * leg + 1], q[:, 3 * leg + 2] = q0, q1, q2
        c0, s0, s1, c1, s12, c12 = np.cos(q0), np.sin(q0), np.sin(q1), np.cos(q1), np.sin(q1 + q2), np.cos(q1 + q2)
        tx, tz, kx, kz = -L1 * s1 - L2 * s12, -L1 * c1 - L2 * c12, -L2 * s12, -L2 * c12
        z = np.zeros(n)
        a_foot = np.stack([tx, l0 * c0 - tz * s0, l0 * s0 + tz * c0], axis=1)
        h_foot = np.stack([tx, -tz * s0, tz * c0], axis=1)
        k_foot = np.stack([kx, -kz * s0, kz * c0], axis=1)
        ax0 = np.stack([z + 1.0, z, z], axis=1)
        ax1 = np.stack([z, c0, s0], axis=1)
        jb = np.stack([np.cross(ax0, a_foot), np.cross(ax1, h_foot), np.cross(ax1, k_foot)], axis=2)
        J[:, leg] = R @ jb
    return q, J


def wbc_target(quat, pos, omega, vel, feet, mask, x1, f0, gains):
    """The scene input rows rollout_wbc_target_kernel builds: joints from the leg IK with the feet at rest, body PD on the stage-1
    prediction, wrench references f0, contacts `mask` [n,4] bool."""
    from . import wbc
    from .mpc import euler_to_quaternion

    n = pos.shape[0]
    q, J = leg_ik(quat, pos, feet)
    rhs = -(vel[:, None, :] + np.cross(omega[:, None, :], feet - pos[:, None, :]))
    qd = np.linalg.solve(J, rhs[..., None])[..., 0].reshape(n, 12)
    kp, kd = np.asarray(gains["com_pose_Kp"], float), np.asarray(gains["com_pose_Kd"], float)
    a_body = np.concatenate([kd[:3] * (x1[:, 9:12] - vel) + kp[:3] * (x1[:, 3:6] - pos),
                             kd[3:] * (x1[:, 6:9] - omega) + kp[3:] * wbc._quat_err_rotvec(euler_to_quaternion(x1[:, 0:3]), quat)], axis=1)
    return wbc.pack_scene(pos, quat, vel, omega, q, qd, mask, a_body, np.zeros((n, 4, 3)), f0.reshape(n, 4, 3))


def closed_loop_stateful(mpc, seq2, periods, wbc_solver=None, substeps=10, plant_inertia="physical", wbc_per_mpc=5, gains=None, control_dt=0.0):
    """What b200qp_mpc_rollout does, with the plant, the foot bookkeeping and the MPC -> WBC coupling stepped on the HOST in numpy and one
    library call per solve (mpc.sequencer_step / solve_records, wbc_solver.scene_solve) - the reference the device loop is tested
    against.  seq2 [n,64] is updated in place; returns the same history array."""
    n, N, dt = seq2.shape[0], mpc.N, (control_dt if control_dt > 0 else 0.05)
    g = gains or dict(com_pose_Kp=[300.0] * 3 + [200.0] * 3, com_pose_Kd=[100.0] * 6, feet_pose_Kp=[1200.0] * 3, feet_pose_Kd=[500.0, 500.0, 400.0])
    hist = np.zeros((periods, n, 16))
    meas = np.ones((n, 4), dtype=np.uint8)
    forces_prev = np.zeros((n, N, 12))
    xpred_prev = np.zeros((n, N + 1, 13))
    for k in range(periods):
        rec, contact = mpc.sequencer_step(seq2, meas)
        forces, xpred, info = mpc.solve_records(rec, contact, forces_prev, xpred_prev)
        ok = info.success
        c0 = contact[:, 0].astype(bool)
        fp = rec[:, 26 + 13 * (N + 1):].reshape(n, N, 4, 3)
        has = contact.astype(bool).any(axis=1)
        first = np.argmax(contact.astype(bool), axis=1)
        nxt = np.take_along_axis(fp, first[:, None, :, None], axis=1)[:, 0]
        feet = seq2[:, 13:25].reshape(n, 4, 3)
        feet = np.where(has[:, :, None], nxt, feet)
        f0 = np.where((ok[:, None] & c0)[:, :, None], forces[:, 0].reshape(n, 4, 3), 0.0)
        seq2[:, 13:25] = feet.reshape(n, 12)
        meas = c0.astype(np.uint8)
        seq2[:, 55] = 16 + (c0 * np.array([1, 2, 4, 8])).sum(axis=1)
        quat, pos, omega, vel = seq2[:, 0:4].copy(), seq2[:, 4:7].copy(), seq2[:, 7:10].copy(), seq2[:, 10:13].copy()
        wst = np.full(n, -1.0)
        if wbc_solver is None:
            quat, pos, omega, vel = srbd_step(quat, pos, omega, vel, feet, f0, dt, substeps=substeps, plant_inertia=plant_inertia)
        else:
            for _ in range(wbc_per_mpc):
                scene = wbc_target(quat, pos, omega, vel, feet, c0, xpred[:, 1], f0.reshape(n, 12), g)
                x, _, winfo = wbc_solver.scene_solve(scene)
                wok = winfo["status"] == 0
                fw = np.where(wok[:, None, None], x[:, 18:].reshape(n, 4, 3), f0)
                quat, pos, omega, vel = srbd_step(quat, pos, omega, vel, feet, fw, dt / wbc_per_mpc, substeps=substeps // wbc_per_mpc,
                                                  plant_inertia=plant_inertia)
                wst = winfo["status"].astype(float)
        seq2[:, 0:4], seq2[:, 4:7], seq2[:, 7:10], seq2[:, 10:13] = quat, pos, omega, vel
        seq2[:, 25] += round(dt * 1e9)
        hist[k, :, 0:3], hist[k, :, 3:6], hist[k, :, 6:10] = pos, vel, quat
        hist[k, :, 10] = f0[:, :, 2].sum(axis=1)
        hist[k, :, 11], hist[k, :, 12], hist[k, :, 13], hist[k, :, 14] = info.return_code, info.acados_num_iter, info.polish_rounds, wst
        forces_prev, xpred_prev = forces, xpred
    return hist
