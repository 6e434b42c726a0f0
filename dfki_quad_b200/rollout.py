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
