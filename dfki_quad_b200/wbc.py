"""Whole-body-control task-space QP for a batch of robots: host-side scene assembly + the GPU QP solver (kernel K3).

Reference: `WBCArcOPT` (ws/src/controllers/src/mit_controller/wbc_arc_opt.cpp) hands tasks to ARC-OPT's
`AccelerationSceneReducedTSID`; `scene->update()` assembles a `QuadraticProgram` (H, g, A, lower/upper) and
`scene->solve(qp)` passes it to a `wbc::QPSolver` plugin (wbc_arc_opt.cpp:51-74,292-297).  The plugin is the drop-in
point: `BatchedWBCSolver.solve` is that call for n QPs at once (C ABI `b200qp_wbc_solve_batch`).  The scene assembly
below restates the reduced-TSID formulation from the upstream ARC-OPT documentation (not verifiable here, SURVEY.md
Appendix B) with the task set and weights of wbc_arc_opt.cpp:78-123 / mit_controller_sim_go2.yaml:62-72; it runs on
the host like the reference's, its rigid-body terms come from go2_model.dynamics (Pinocchio in the reference).

Decision vector x = [nu_dot (18), f_fl, f_fr, f_bl, f_br (3 each)]  (30);  18 equalities;  28 two-sided rows.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

NQ, NEQ, NIN = 30, 18, 28
INF = 1e20

# mit_controller_sim_go2.yaml:62-72
GO2_WBC = dict(mu=0.45, com_pose_weight=np.full(6, 10.0), com_pose_Kp=np.array([300.0, 300.0, 300.0, 200.0, 200.0, 200.0]),
               com_pose_Kd=np.full(6, 100.0), foot_force_weight=np.full(3, 30.0), foot_pose_weight=np.full(3, 10.0),
               feet_pose_Kp=np.full(3, 1200.0), feet_pose_Kd=np.array([500.0, 500.0, 400.0]), hessian_regularizer=1e-8)


@dataclass
class QuadraticProgram:
    """Batched analogue of ARC-OPT's wbc::QuadraticProgram: the first `neq` rows of A are equalities."""

    H: np.ndarray  # [n, nq, nq]
    g: np.ndarray  # [n, nq]
    A: np.ndarray  # [n, nc, nq]
    lower_y: np.ndarray  # [n, nc]
    upper_y: np.ndarray  # [n, nc]
    neq: int


def reduced_tsid_qp(dyn, contacts, a_body_des, a_foot_des, f_ref, tau_max, cfg=GO2_WBC):
    """Assemble the reduced-TSID QP for a batch.  dyn = go2_model.dynamics(...); contacts [n,4] bool;
    a_body_des [n,6] (linear, angular, world); a_foot_des [n,4,3]; f_ref [n,4,3] (MPC GRFs, stage 0)."""
    M, h, J, Jdnu = dyn["M"], dyn["h"], dyn["foot_J"], dyn["foot_Jdnu"]
    n = M.shape[0]
    on = np.asarray(contacts).astype(bool)
    H = np.zeros((n, NQ, NQ))
    g = np.zeros((n, NQ))
    idx = np.arange(NQ)
    H[:, idx, idx] = cfg["hessian_regularizer"]
    wb = cfg["com_pose_weight"]
    # body_pose task (CartesianAccelerationTask on base_link in world, wbc_arc_opt.cpp:82-91): J = [I6 0], Jdot nu = 0.
    # weights order (x,y,z,roll,pitch,yaw) = (linear, angular)
    H[:, idx[:6], idx[:6]] += wb
    g[:, :6] -= wb * a_body_des
    wfp, wff = cfg["foot_pose_weight"], cfg["foot_force_weight"]
    for f in range(4):
        sw = ~on[:, f]
        Jf = J[:, f]  # [n,3,18]
        # foot_pose task, active for feet NOT in contact (wbc_arc_opt.cpp:108-117)
        JW = Jf * wfp[None, :, None]
        H[:, :18, :18] += np.einsum("nij,nik->njk", JW, Jf) * sw[:, None, None]
        g[:, :18] += np.einsum("nij,ni->nj", JW, Jdnu[:, f] - a_foot_des[:, f]) * sw[:, None]
        # foot_contact wrench-forward task, active for feet in contact (wbc_arc_opt.cpp:95-104)
        s3 = 18 + 3 * f
        H[:, idx[s3:s3 + 3], idx[s3:s3 + 3]] += wff * on[:, f][:, None]
        g[:, s3:s3 + 3] -= wff * f_ref[:, f] * on[:, f][:, None]
    A = np.zeros((n, NEQ + NIN, NQ))
    lo = np.zeros((n, NEQ + NIN))
    hi = np.zeros((n, NEQ + NIN))
    # floating-base rows of the equation of motion: M_b nud - sum_f J_f[:, :6]' f_f = -h_b
    A[:, 0:6, 0:18] = M[:, 0:6, :]
    for f in range(4):
        A[:, 0:6, 18 + 3 * f:21 + 3 * f] = -np.transpose(J[:, f, :, 0:6], (0, 2, 1))
    lo[:, 0:6] = hi[:, 0:6] = -h[:, 0:6]
    mu = cfg["mu"]
    pyr = np.array([[1, 0, -mu], [1, 0, mu], [0, 1, -mu], [0, 1, mu]], float)
    for f in range(4):
        r0 = 6 + 3 * f
        st = on[:, f]
        # stance: rigid contact J_f nud = -Jdot nu ; swing: the contact wrench is removed (f_f = 0)
        A[:, r0:r0 + 3, 0:18] = J[:, f] * st[:, None, None]
        A[:, idx[r0:r0 + 3], idx[18 + 3 * f:21 + 3 * f]] = (~st)[:, None] * 1.0
        lo[:, r0:r0 + 3] = hi[:, r0:r0 + 3] = -Jdnu[:, f] * st[:, None]
        # friction pyramid rows (inactive for swing feet)
        p0 = NEQ + 12 + 4 * f
        A[:, p0:p0 + 4, 18 + 3 * f:21 + 3 * f] = pyr
        lo[:, p0:p0 + 4] = np.where(st[:, None], np.array([-INF, 0.0, -INF, 0.0]), -INF)
        hi[:, p0:p0 + 4] = np.where(st[:, None], np.array([0.0, INF, 0.0, INF]), INF)
    # joint torque limits: tau = M_j nud + h_j - sum_f J_f[:, 6:]' f_f in [-tau_max, tau_max]
    A[:, NEQ:NEQ + 12, 0:18] = M[:, 6:, :]
    for f in range(4):
        A[:, NEQ:NEQ + 12, 18 + 3 * f:21 + 3 * f] = -np.transpose(J[:, f, :, 6:], (0, 2, 1))
    lo[:, NEQ:NEQ + 12] = -tau_max - h[:, 6:]
    hi[:, NEQ:NEQ + 12] = tau_max - h[:, 6:]
    return QuadraticProgram(H=H, g=g, A=A, lower_y=lo, upper_y=hi, neq=NEQ)


SCENE_IN_DOUBLES = 72  # B200QP_WBC_SCENE_IN_DOUBLES


def pack_scene(pos, quat, v_lin, omega, q, qd, contacts, a_body_des, a_foot_des, f_ref):
    """Input records of the on-device scene assembly (kernel K6, b200qp_wbc_scene_*): [n, 72] doubles = quat xyzw 0:4, pos 4:7,
    v 7:10, omega 10:13, q 13:25, qd 25:37, a_body_des 37:43, a_foot_des 43:55, f_ref 55:67, contact bit mask 67."""
    n = pos.shape[0]
    rec = np.zeros((n, SCENE_IN_DOUBLES))
    rec[:, 0:4], rec[:, 4:7], rec[:, 7:10], rec[:, 10:13] = quat, pos, v_lin, omega
    rec[:, 13:25], rec[:, 25:37], rec[:, 37:43] = q, qd, a_body_des
    rec[:, 43:55] = np.asarray(a_foot_des).reshape(n, 12)
    rec[:, 55:67] = np.asarray(f_ref).reshape(n, 12)
    rec[:, 67] = (np.asarray(contacts).astype(np.int64) * np.array([1, 2, 4, 8])).sum(axis=1)
    return rec


def joint_torques(dyn, x):
    """tau = M_j nud + h_j - sum_f J_f[:, 6:]' f_f  (the scene's solver_output -> effort mapping)."""
    nud, f = x[:, :18], x[:, 18:].reshape(-1, 4, 3)
    tau = np.einsum("nij,nj->ni", dyn["M"][:, 6:, :], nud) + dyn["h"][:, 6:]
    tau -= np.einsum("nfij,nfi->nj", dyn["foot_J"][:, :, :, 6:], f)
    return tau


def cartesian_pd(kp, kd, pos_err, vel_err, acc_ff=0.0):
    """wbc CartesianPosPDController: a = a_ff + Kd (v_ref - v) + Kp (x_ref - x)  (wbc_arc_opt.cpp:130-150,256-289)."""
    return acc_ff + kd * vel_err + kp * pos_err


class BatchedWBCSolver:
    """Mirror of a `wbc::QPSolver` plugin: solve(qp) -> solver_output for a batch (kernel K3, dense primal-dual IPM)."""

    def __init__(self, nq=NQ, neq=NEQ, nin=NIN, max_batch=16384, device=0, kkt_tol=1e-6, max_iter=40):
        self._lib = _lib.load()
        cfg = _lib.WbcConfig()
        cfg.nq, cfg.neq, cfg.nin, cfg.max_batch, cfg.device = nq, neq, nin, int(max_batch), int(device)
        cfg.kkt_tol, cfg.max_iter = float(kkt_tol), int(max_iter)
        self.cfg = cfg
        self._h = C.c_void_p()
        _lib.check(self._lib.b200qp_wbc_create(C.byref(cfg), C.byref(self._h)), "b200qp_wbc_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200qp_wbc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def abi_arrays(self, qp: QuadraticProgram):
        """Column-major (Eigen) layouts of the C ABI, as ARC-OPT's QuadraticProgram already holds them."""
        n = qp.H.shape[0]
        nq, nc = self.cfg.nq, self.cfg.neq + self.cfg.nin
        assert qp.neq == self.cfg.neq and qp.A.shape == (n, nc, nq)
        H = np.ascontiguousarray(np.transpose(qp.H, (0, 2, 1)))
        A = np.ascontiguousarray(np.transpose(qp.A, (0, 2, 1)))  # [n][nq][nc] == column-major nc x nq
        g = np.ascontiguousarray(qp.g)
        lo = np.ascontiguousarray(np.clip(qp.lower_y, -INF, INF))
        hi = np.ascontiguousarray(np.clip(qp.upper_y, -INF, INF))
        return H, g, A, lo, hi

    def solve(self, qp: QuadraticProgram):
        """-> (x [n,nq], info structured array).  Failed problems keep zeros."""
        return self.solve_abi(*self.abi_arrays(qp))

    def solve_abi(self, H, g, A, lo, hi, x=None):
        """b200qp_wbc_solve_batch on arrays that are already in the ABI layout (no host-side reshuffling)."""
        n, nq = g.shape
        if x is None:
            x = np.zeros((n, nq))
        info = (_lib.SolverInfo * max(n, 1))()
        rc = self._lib.b200qp_wbc_solve_batch(self._h, n, H.ctypes.data, g.ctypes.data, A.ctypes.data, lo.ctypes.data,
                                              hi.ctypes.data, x.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_wbc_solve_batch")
        dt = np.dtype([("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))])
        arr = np.frombuffer(info, dtype=np.uint8, count=48 * n).copy().view(dt)  # bytes copy + view: 10x faster than a structured copy
        return x, arr

    # -- scene assembled on the device (kernel K6) ----------------------------------------------------------------------
    def scene_configure(self, cfg=GO2_WBC, tau_max=None):
        """b200qp_wbc_scene_configure: weights / friction of the reduced-TSID scene (mit_controller_sim_go2.yaml:62-72)."""
        from . import go2_model as gm

        c = _lib.WbcSceneConfig()
        c.mu = float(cfg["mu"])
        c.com_pose_weight[:] = [float(v) for v in cfg["com_pose_weight"]]
        c.foot_force_weight[:] = [float(v) for v in cfg["foot_force_weight"]]
        c.foot_pose_weight[:] = [float(v) for v in cfg["foot_pose_weight"]]
        c.hessian_regularizer = float(cfg["hessian_regularizer"])
        c.tau_max[:] = [float(v) for v in (gm.TAU_MAX if tau_max is None else tau_max)]
        _lib.check(self._lib.b200qp_wbc_scene_configure(self._h, C.byref(c)), "b200qp_wbc_scene_configure")

    def scene_assemble(self, scene_in):
        """b200qp_wbc_scene_assemble (parity hook) -> H [n,900], g, A [n,1380], lower_y, upper_y (ABI layout), feet [n,24]."""
        scene_in = np.ascontiguousarray(scene_in, dtype=np.float64)
        n = scene_in.shape[0]
        out = [np.zeros((n, k)) for k in (900, 30, 1380, 46, 46, 24)]
        rc = self._lib.b200qp_wbc_scene_assemble(self._h, n, scene_in.ctypes.data, *[o.ctypes.data for o in out])
        _lib.check(rc, "b200qp_wbc_scene_assemble")
        return out

    def scene_solve(self, scene_in, x=None, tau=None):
        """b200qp_wbc_scene_solve_batch: robot states -> (x [n,30], tau [n,12], info)."""
        scene_in = np.ascontiguousarray(scene_in, dtype=np.float64)
        n = scene_in.shape[0]
        if scene_in.shape != (n, SCENE_IN_DOUBLES):
            raise ValueError("scene_in must be [n, 72]")
        x = np.zeros((n, 30)) if x is None else x
        tau = np.zeros((n, 12)) if tau is None else tau
        info = (_lib.SolverInfo * max(n, 1))()
        rc = self._lib.b200qp_wbc_scene_solve_batch(self._h, n, scene_in.ctypes.data, x.ctypes.data, tau.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_wbc_scene_solve_batch")
        dt = np.dtype([("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))])
        return x, tau, np.frombuffer(info, dtype=np.uint8, count=48 * n).copy().view(dt)

    def scene_solve_device(self, n, d_scene_in, d_x, d_tau, d_info, sync=True):
        rc = self._lib.b200qp_wbc_scene_solve_batch_device(self._h, n, d_scene_in, d_x, d_tau, d_info, 1 if sync else 0)
        _lib.check(rc, "b200qp_wbc_scene_solve_batch_device")

    def solve_device(self, n, d_H, d_g, d_A, d_lo, d_hi, d_x, d_info, sync=True):
        rc = self._lib.b200qp_wbc_solve_batch_device(self._h, n, d_H, d_g, d_A, d_lo, d_hi, d_x, d_info, 1 if sync else 0)
        _lib.check(rc, "b200qp_wbc_solve_batch_device")

    def stream(self) -> int:
        return int(self._lib.b200qp_wbc_stream(self._h) or 0)


def _quat_err_rotvec(q_ref, q):
    """Rotation vector of R_ref R^T (world frame) for (x,y,z,w) quaternions."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    qi = np.stack([-x, -y, -z, w], axis=1)
    ax, ay, az, aw = q_ref.T
    bx, by, bz, bw = qi.T
    d = np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                  aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], axis=1)
    d = d * np.where(d[:, 3:4] < 0, -1.0, 1.0)
    nv = np.linalg.norm(d[:, :3], axis=1, keepdims=True)
    ang = 2.0 * np.arctan2(nv, d[:, 3:4])
    return np.where(nv > 1e-12, d[:, :3] / np.maximum(nv, 1e-300) * ang, 2.0 * d[:, :3])


def random_wbc_batch(n, seed, device=0, mpc_horizon=10, cfg=GO2_WBC, schedule_fn=None, mpc_solution=None):
    """BASELINE.json configs[3]: n Go2 WBC QPs. States/commands as the MPC configs, joint angles =
    nominal stance + U(-0.15, 0.15), joint velocities U(-1, 1), stance set = gait mask at stage 0, wrench references =
    stage-0 GRFs of a config-2-style MPC solve on the GPU ("fed by the MPC GRFs", mit_controller_node.cpp:1000-1006,1173),
    body task reference = Cartesian PD on the MPC's stage-1 prediction, swing-foot references = Cartesian PD on random
    swing targets.  `mpc_solution=(forces, xpred)` overrides the GPU MPC solve (CPU-only tests)."""
    from . import go2_model as gm
    from . import sequencer
    from .mpc import BatchedMPC, euler_to_quaternion

    N = mpc_horizon
    b = sequencer.random_batch(n, N, seed, gaits=("TROT",), device=device, schedule_fn=schedule_fn)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    if mpc_solution is None:
        m = sequencer.GO2_MPC
        mpc = BatchedMPC(m["alpha"], m["state_weights_move"], m["mu"], m["fmin"], m["fmax"], "CONDENSED_ADMM", horizon=N,
                         dt=m["dt"], max_batch=n, device=device)
        forces, xpred, info = mpc.solve_records(b["rec"], b["contact"])
        mpc.close()
    else:
        forces, xpred = mpc_solution
    contacts = b["contact"][:, 0, :].astype(bool)
    f_ref = forces.reshape(n, N, 4, 3)[:, 0] * contacts[:, :, None]
    q = gm.nominal_joint_angles()[None, :] + rng.uniform(-0.15, 0.15, (n, 12))
    qd = rng.uniform(-1.0, 1.0, (n, 12))
    dyn = gm.dynamics(b["pos"], b["quat"], q, b["vel"], b["omega"], qd)
    # body task: PD on the stage-1 prediction (mit_controller_node.cpp:1000-1003; gains mit_controller_sim_go2.yaml:66-67)
    x1 = xpred[:, 1]
    q_ref = euler_to_quaternion(x1[:, 0:3])
    kp, kd = cfg["com_pose_Kp"], cfg["com_pose_Kd"]
    a_body = np.concatenate([cartesian_pd(kp[:3], kd[:3], x1[:, 3:6] - b["pos"], x1[:, 9:12] - b["vel"]),
                             cartesian_pd(kp[3:], kd[3:], _quat_err_rotvec(q_ref, b["quat"]), x1[:, 6:9] - b["omega"])], axis=1)
    # swing feet: PD on random swing targets around the current foot position
    tgt_p = dyn["foot_pos"] + rng.uniform(-0.01, 0.01, (n, 4, 3))
    tgt_v = dyn["foot_vel"] + rng.uniform(-0.05, 0.05, (n, 4, 3))
    a_foot = cartesian_pd(cfg["feet_pose_Kp"], cfg["feet_pose_Kd"], tgt_p - dyn["foot_pos"], tgt_v - dyn["foot_vel"])
    qp = reduced_tsid_qp(dyn, contacts, a_body, a_foot, f_ref, gm.TAU_MAX, cfg)
    return dict(qp=qp, dyn=dyn, contacts=contacts, f_ref=f_ref, a_body=a_body, a_foot=a_foot, q=q, qd=qd, mpc_batch=b)
