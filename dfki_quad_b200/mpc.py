"""Host-side mirror of the reference's MPC solver plugin for a *batch* of robots.

`BatchedMPC` keeps the method set of `class MPC : MPCInterface`
(ws/src/controllers/include/mit_controller/mpc.hpp:117-141, mpc_interface.hpp:22-33):
UpdateState / UpdateModel / UpdateGaitSequence / GetWrenchSequence + SetStateWeights / SetInputWeights / SetFmax /
SetMu, with every argument gaining a leading batch dimension.  All arithmetic runs in libb200qp.so (CUDA, sm_100a)
through the C ABI of include/b200qp.h; nothing here computes a QP on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

NX, NU, NLEGS = 13, 12, 4


def in_doubles(N: int) -> int:
    return 26 + 13 * (N + 1) + 12 * N


@dataclass
class SolverInformation:
    """Batched analogue of SolverInformation (mpc_interface.hpp:10-20)."""

    success: np.ndarray  # bool [n]
    return_code: np.ndarray  # int32 [n]
    acados_num_iter: np.ndarray  # int32 [n]
    polish_rounds: np.ndarray  # int32 [n]
    n_free: np.ndarray  # int32 [n]
    kkt: np.ndarray  # float64 [n, 4]  (stationarity, dynamics, inequality, complementarity)
    total_solver_time: float  # seconds, device time of the kernels of this call (all problems)
    kernel_launches: int


_INFO_DTYPE = np.dtype([("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"), ("kkt", "f8", (4,))])


@dataclass
class MPCPrediction:
    """Batched analogue of MPCPrediction (mpc_prediction.hpp:7-13)."""

    orientation: np.ndarray  # [n, N+1, 4] quaternion (x, y, z, w)
    position: np.ndarray  # [n, N+1, 3]
    angular_velocity: np.ndarray
    linear_velocity: np.ndarray
    raw_data: np.ndarray  # [n, N+1, 13]


def pack_records(N, quat, pos, omega, vel, mass, inertia, com, des_quat, des_pos, ref_twist, ref_vel, foot_pos):
    """Pack batched inputs into the flat record layout of include/b200qp.h.  inertia is [n,3,3] (row-major numpy)."""
    n = quat.shape[0]
    rec = np.empty((n, in_doubles(N)))
    rec[:, 0:4] = quat
    rec[:, 4:7] = pos
    rec[:, 7:10] = omega
    rec[:, 10:13] = vel
    rec[:, 13] = mass
    rec[:, 14:23] = np.transpose(np.broadcast_to(inertia, (n, 3, 3)), (0, 2, 1)).reshape(n, 9)
    rec[:, 23:26] = com
    st = rec[:, 26 : 26 + 13 * (N + 1)].reshape(n, N + 1, 13)
    st[:, :, 0:4] = des_quat[:, : N + 1]
    st[:, :, 4:7] = des_pos[:, : N + 1]
    st[:, :, 7:10] = ref_twist[:, : N + 1]
    st[:, :, 10:13] = ref_vel[:, : N + 1]
    rec[:, 26 + 13 * (N + 1) :] = foot_pos[:, :N].reshape(n, 12 * N)
    return rec


def euler_to_quaternion(e):
    """euler_to_quaternion (quaternion_operations.hpp:19-25), vectorised; e[..., (roll, pitch, yaw)] -> (x,y,z,w)."""
    r, p, y = e[..., 0] * 0.5, e[..., 1] * 0.5, e[..., 2] * 0.5
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    # qz * qy * qx
    w = cy * cp * cr + sy * sp * sr
    x = cy * cp * sr - sy * sp * cr
    yy = cy * sp * cr + sy * cp * sr
    z = sy * cp * cr - cy * sp * sr
    return np.stack([x, yy, z, w], axis=-1)


# b200qp_seq_options defaults: config/mit_controller_sim_go2.yaml:40-52, mit_controller_node.cpp:249-264 (AdaptiveGait arguments)
SEQ_OPTION_DEFAULTS = dict(
    raibert_filtersize=20, fix_standing_position=1, early_contact_detection=1, fix_position_distance_threshold=0.1,
    fix_position_angular_threshold=0.26, fix_position_velocity_threshold=0.1, swing_time=0.2, zero_velocity_threshold=0.03,
    standing_foot_position_threshold=0.08, min_v_cmd_factor=0.5, max_correction_cycles=2.0, correction_period=0.6,
    disturbance_correction=0.0, min_v=0.0, offset_delay=0.0, max_stride_length=0.0, phase_offset=(0.0, 0.5, 0.5, 0.0),
    control_dt=0.01, filter_size=10, switch_offsets=1, correct_all=0)
# Target fields in the order of the value block / activation bits of the stateful sequencer input (target.hpp:8-45)
TARGET_VALUES = ("x", "y", "z", "roll", "pitch", "yaw", "qx", "qy", "qz", "qw", "x_dot", "y_dot", "z_dot", "hybrid_x_dot",
                 "hybrid_y_dot", "hybrid_z_dot", "wx", "wy", "wz")
TARGET_FLAGS = ("x", "y", "z", "roll", "pitch", "yaw", "full_orientation", "x_dot", "y_dot", "z_dot", "hybrid_x_dot", "hybrid_y_dot",
                "hybrid_z_dot", "wx", "wy", "wz")


def pack_seq2(state13, feet, time_ns, target, gait=None):
    """Input records of the stateful sequencer (b200qp.h, B200QP_SEQ2_IN_DOUBLES).  state13 [n,13], feet [n,4,3], time_ns [n] (integer
    nanoseconds), target: dict field -> [n] array or scalar (`full_orientation` -> [n,4] xyzw); every given field is activated.
    gait = (period [n], duty [n,4], offset [n,4]) for the simple sequencer."""
    state13 = np.asarray(state13, dtype=np.float64)
    n = state13.shape[0]
    rec = np.zeros((n, 64))
    rec[:, 0:13] = state13
    rec[:, 13:25] = np.asarray(feet, dtype=np.float64).reshape(n, 12)
    rec[:, 25] = np.asarray(time_ns, dtype=np.float64)
    rec[:, 26 + 9] = 1.0
    act = 0
    for k, v in target.items():
        if k == "full_orientation":
            rec[:, 32:36] = v
        else:
            rec[:, 26 + TARGET_VALUES.index(k)] = v
        act |= 1 << TARGET_FLAGS.index(k)
    rec[:, 45] = float(act)
    if gait is not None:
        rec[:, 46] = gait[0]
        rec[:, 47:51] = gait[1]
        rec[:, 51:55] = gait[2]
    return rec


class BatchedMPC:
    """Drop-in for `MPC` over a batch.  Constructor arguments follow mpc.hpp:117-127; `solver` accepts the reference's
    acados plan names ("PARTIAL_CONDENSING_HPIPM", "PARTIAL_CONDENSING_OSQP", ...) or the kernel names
    "CONDENSED_ADMM" / "SPARSE_RICCATI" / "PARTIAL_CONDENSING" / "AUTO"; `condensing_N` is the reference's cond_N (used by the
    partial-condensing kernel: number of blocks, None = library default), `hpipm_mode` is accepted for signature compatibility
    (the GPU kernels always solve to `kkt_tol`), `warm_start` as in the reference (0 cold, 1 warm, 2 hot)."""

    def __init__(self, alpha, state_weights, mu, fmin, fmax, solver="CONDENSED_ADMM", condensing_N=None,
                 hpipm_mode="BALANCE", warm_start=0, *, horizon=10, dt=0.05, max_batch=4096, device=0, **opts):
        self._lib = _lib.load()
        if solver not in _lib.SOLVERS:
            raise ValueError(f"unknown solver {solver!r}")
        cfg = _lib.MpcConfig()
        cfg.horizon, cfg.solver, cfg.max_batch, cfg.device = int(horizon), _lib.SOLVERS[solver], int(max_batch), int(device)
        cfg.dt, cfg.alpha, cfg.mu, cfg.fmin, cfg.fmax = float(dt), float(alpha), float(mu), float(fmin), float(fmax)
        sw = np.asarray(state_weights, dtype=np.float64).reshape(12)
        for i in range(12):
            cfg.state_weights[i] = sw[i]
        cfg.cond_N = int(condensing_N) if condensing_N else 0  # MPC::MPC condensing_N (mpc.hpp:125)
        cfg.warm_start = int(warm_start)  # MPC::MPC warm_start (mpc.hpp:127): 0 cold, 1 warm, 2 hot
        for k, v in opts.items():
            if not hasattr(cfg, k):
                raise TypeError(f"unknown solver option {k!r}")
            setattr(cfg, k, v)
        self.cfg = cfg
        self.N = int(horizon)
        self.max_batch = int(max_batch)
        self._h = C.c_void_p()
        _lib.check(self._lib.b200qp_mpc_create(C.byref(cfg), C.byref(self._h)), "b200qp_mpc_create")
        self._state = None
        self._model = None
        self._gs = None
        self._last = None  # (forces, xpred) of the previous GetWrenchSequence: a failed problem keeps them (mpc.cpp:447-455)

    # -- lifetime ---------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200qp_mpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- MPCInterface -------------------------------------------------------------------------------------------------
    def UpdateState(self, orientation, position, angular_vel_world, linear_vel_world):
        """StateInterface getters used at mpc.cpp:332,392-395; arrays [n,4] (x,y,z,w), [n,3], [n,3], [n,3]."""
        self._state = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in
                            (orientation, position, angular_vel_world, linear_vel_world))

    def UpdateModel(self, mass, inertia, body_to_com):
        """ModelInterface getters used at mpc.cpp:335,357-358; scalars/arrays broadcast over the batch."""
        self._model = (mass, np.asarray(inertia, dtype=np.float64), np.asarray(body_to_com, dtype=np.float64))

    def UpdateGaitSequence(self, gait_sequence: dict):
        """GaitSequence fields read by the MPC (gait_sequence.hpp:21-29): contact_sequence [n,>=N,4],
        foot_position_sequence [n,>=N,4,3], desired_reference_trajectory_orientation [n,>=N+1,4],
        desired_reference_trajectory_position [n,>=N+1,3], reference_trajectory_twist, reference_trajectory_velocity."""
        self._gs = gait_sequence

    def GetWrenchSequence(self):
        """-> (forces [n,N,4,3], MPCPrediction, SolverInformation).  As in the reference, whose caller passes the same
        WrenchSequence / MPCPrediction objects every cycle, a problem whose solve fails keeps the wrench and prediction of
        the previous call (mpc.cpp:447-455) - zeros only before the first success.  The returned arrays are copies."""
        if self._state is None or self._model is None or self._gs is None:
            raise RuntimeError("UpdateState / UpdateModel / UpdateGaitSequence must be called first")
        q, p, w, v = self._state
        n = q.shape[0]
        gs = self._gs
        mass, inertia, com = self._model
        rec = pack_records(self.N, q, p, w, v, np.broadcast_to(mass, (n,)), inertia, np.broadcast_to(com, (n, 3)),
                           gs["desired_reference_trajectory_orientation"], gs["desired_reference_trajectory_position"],
                           gs["reference_trajectory_twist"], gs["reference_trajectory_velocity"],
                           gs["foot_position_sequence"])
        contact = np.ascontiguousarray(np.asarray(gs["contact_sequence"])[:, : self.N], dtype=np.uint8)
        if self._last is None or self._last[0].shape[0] != n:
            self._last = (np.zeros((n, self.N, 12)), np.zeros((n, self.N + 1, 13)))
        forces, xpred, info = self.solve_records(rec, contact, self._last[0], self._last[1])
        forces, xpred = forces.copy(), xpred.copy()
        com_b = np.broadcast_to(com, (n, 3))[:, None, :]
        pred = MPCPrediction(
            orientation=euler_to_quaternion(xpred[:, :, 0:3]),  # mpc.cpp:463
            position=xpred[:, :, 3:6] - com_b,  # mpc.cpp:464
            angular_velocity=xpred[:, :, 6:9].copy(), linear_velocity=xpred[:, :, 9:12].copy(), raw_data=xpred)
        return forces.reshape(n, self.N, 4, 3), pred, info

    # -- setters (mpc.hpp:138-141) ---------------------------------------------------------------------------------------
    def SetStateWeights(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.float64).reshape(12)
        _lib.check(self._lib.b200qp_mpc_set_state_weights(self._h, w.ctypes.data_as(C.POINTER(C.c_double))), "set_state_weights")

    def SetInputWeights(self, alpha):
        _lib.check(self._lib.b200qp_mpc_set_input_weights(self._h, float(alpha)), "set_input_weights")

    def SetFmax(self, fmax):
        _lib.check(self._lib.b200qp_mpc_set_fmax(self._h, float(fmax)), "set_fmax")

    def SetWarmStart(self, mode):
        """0 cold / 1 warm / 2 hot start (b200qp_mpc_set_warm_start); also forgets the stored solutions."""
        _lib.check(self._lib.b200qp_mpc_set_warm_start(self._h, int(mode)), "set_warm_start")

    def SetMu(self, mu):
        _lib.check(self._lib.b200qp_mpc_set_mu(self._h, float(mu)), "set_mu")

    # -- record-level entry points -------------------------------------------------------------------------------------
    def solve_records(self, rec, contact, forces=None, xpred=None):
        """b200qp_mpc_solve_batch on host arrays: rec [n, in_doubles(N)] float64, contact [n, N, 4] uint8."""
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        contact = np.ascontiguousarray(contact, dtype=np.uint8)
        n = rec.shape[0]
        if rec.shape != (n, in_doubles(self.N)) or contact.shape != (n, self.N, 4):
            raise ValueError("record / contact shape mismatch")
        if forces is None:
            forces = np.zeros((n, self.N, 12))
        if xpred is None:
            xpred = np.zeros((n, self.N + 1, 13))
        info = self._info_buf()
        rc = self._lib.b200qp_mpc_solve_batch(self._h, n, rec.ctypes.data, contact.ctypes.data, forces.ctypes.data,
                                              xpred.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_mpc_solve_batch")
        return forces, xpred, self._info(info, n)

    # -- on-device gait sequencer (b200qp_mpc_sequence_*) ---------------------------------------------------------------
    def set_model(self, mass, inertia, com, shoulders, raibert_k, raibert_g, max_leg_length):
        """MPC::UpdateModel + the planner constants for the sequencer entry points (b200qp_mpc_set_model)."""
        m = _lib.SeqModel()
        m.mass = float(mass)
        m.inertia[:] = [float(v) for v in np.asarray(inertia, float).reshape(9)]
        m.com[:] = [float(v) for v in np.asarray(com, float).reshape(3)]
        m.shoulders[:] = [float(v) for v in np.asarray(shoulders, float).reshape(12)]
        m.raibert_k, m.raibert_g, m.max_leg_length = float(raibert_k), float(raibert_g), float(max_leg_length)
        _lib.check(self._lib.b200qp_mpc_set_model(self._h, C.byref(m)), "b200qp_mpc_set_model")

    def _seq_args(self, seq_in, measured_contact):
        seq_in = np.ascontiguousarray(seq_in, dtype=np.float64)
        n = seq_in.shape[0]
        if seq_in.shape != (n, 50):
            raise ValueError("seq_in must be [n, 50]")
        mc = None
        if measured_contact is not None:
            mc = np.ascontiguousarray(measured_contact, dtype=np.uint8)
            if mc.shape != (n, 4):
                raise ValueError("measured_contact must be [n, 4]")
        return seq_in, mc, n

    def sequence_records(self, seq_in, measured_contact=None):
        """b200qp_mpc_sequence_batch: (state, feet, Target, gait, phase) -> (records, contact) as solve_records takes."""
        seq_in, mc, n = self._seq_args(seq_in, measured_contact)
        rec = np.zeros((n, in_doubles(self.N)))
        contact = np.zeros((n, self.N, 4), dtype=np.uint8)
        rc = self._lib.b200qp_mpc_sequence_batch(self._h, n, seq_in.ctypes.data, mc.ctypes.data if mc is not None else None,
                                                 rec.ctypes.data, contact.ctypes.data)
        _lib.check(rc, "b200qp_mpc_sequence_batch")
        return rec, contact

    def sequence_and_solve(self, seq_in, measured_contact=None, forces=None, xpred=None):
        """b200qp_mpc_sequence_solve_batch: sequencer + QP assembly + solve in one call."""
        seq_in, mc, n = self._seq_args(seq_in, measured_contact)
        if forces is None:
            forces = np.zeros((n, self.N, 12))
        if xpred is None:
            xpred = np.zeros((n, self.N + 1, 13))
        info = self._info_buf()
        rc = self._lib.b200qp_mpc_sequence_solve_batch(self._h, n, seq_in.ctypes.data,
                                                       mc.ctypes.data if mc is not None else None, forces.ctypes.data,
                                                       xpred.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_mpc_sequence_solve_batch")
        return forces, xpred, self._info(info, n)

    # -- the sequencers with their state across calls (b200qp_mpc_sequencer_*) ----------------------------------------------
    def sequencer_configure(self, mode="simple", **options):
        """b200qp_mpc_sequencer_configure: `mode` "simple" (SimpleGaitSequencer) or "adaptive" (AdaptiveGaitSequencer); options are
        the fields of b200qp_seq_options, defaults = config/mit_controller_sim_go2.yaml + mit_controller_node.cpp:250-264."""
        o = _lib.SeqOptions()
        vals = dict(SEQ_OPTION_DEFAULTS)
        unknown = set(options) - set(vals)
        if unknown:
            raise TypeError(f"unknown sequencer options {sorted(unknown)}")
        vals.update(options)
        o.mode = {"simple": 0, "adaptive": 1}[mode]
        for k, v in vals.items():
            if k == "phase_offset":
                o.phase_offset[:] = [float(x) for x in v]
            elif isinstance(getattr(o, k), int):
                setattr(o, k, int(v))
            else:
                setattr(o, k, float(v))
        _lib.check(self._lib.b200qp_mpc_sequencer_configure(self._h, C.byref(o)), "b200qp_mpc_sequencer_configure")

    def sequencer_reset(self):
        _lib.check(self._lib.b200qp_mpc_sequencer_reset(self._h), "b200qp_mpc_sequencer_reset")

    def _seq2_args(self, seq2_in, measured_contact):
        seq2_in = np.ascontiguousarray(seq2_in, dtype=np.float64)
        n = seq2_in.shape[0]
        if seq2_in.shape != (n, 64):
            raise ValueError("seq2_in must be [n, 64]")
        mc = None
        if measured_contact is not None:
            mc = np.ascontiguousarray(measured_contact, dtype=np.uint8)
            if mc.shape != (n, 4):
                raise ValueError("measured_contact must be [n, 4]")
        return seq2_in, mc, n

    def sequencer_step(self, seq2_in, measured_contact=None):
        """b200qp_mpc_sequencer_step: UpdateState + UpdateTarget + GetGaitSequence of every robot -> (records, contact)."""
        seq2_in, mc, n = self._seq2_args(seq2_in, measured_contact)
        rec = np.zeros((n, in_doubles(self.N)))
        contact = np.zeros((n, self.N, 4), dtype=np.uint8)
        rc = self._lib.b200qp_mpc_sequencer_step(self._h, n, seq2_in.ctypes.data, mc.ctypes.data if mc is not None else None,
                                                 rec.ctypes.data, contact.ctypes.data)
        _lib.check(rc, "b200qp_mpc_sequencer_step")
        return rec, contact

    def sequencer_step_solve(self, seq2_in, measured_contact=None, forces=None, xpred=None):
        """b200qp_mpc_sequencer_step_solve: the stateful sequencer + QP assembly + solve in one call."""
        seq2_in, mc, n = self._seq2_args(seq2_in, measured_contact)
        if forces is None:
            forces = np.zeros((n, self.N, 12))
        if xpred is None:
            xpred = np.zeros((n, self.N + 1, 13))
        info = self._info_buf()
        rc = self._lib.b200qp_mpc_sequencer_step_solve(self._h, n, seq2_in.ctypes.data, mc.ctypes.data if mc is not None else None,
                                                       forces.ctypes.data, xpred.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_mpc_sequencer_step_solve")
        return forces, xpred, self._info(info, n)

    # -- closed loop on the device (b200qp_mpc_rollout) -------------------------------------------------------------------
    def rollout(self, seq2, periods, wbc=None, substeps=10, plant_inertia="physical", wbc_per_mpc=5, gains=None, history=True, control_dt=0.0):
        """b200qp_mpc_rollout: `periods` control periods of n single-rigid-body robots with the sequencer + MPC (+ the WBC of the
        BatchedWBCSolver `wbc`, scene configured) in the loop, all on the device.  seq2 [n,64] is updated in place to the final
        state; returns the history [periods, n, 16] (pos, vel, quat, sum f_z, MPC status, iterations, polish rounds, WBC status)."""
        if not (isinstance(seq2, np.ndarray) and seq2.dtype == np.float64 and seq2.flags.c_contiguous and seq2.ndim == 2 and seq2.shape[1] == 64):
            raise ValueError("seq2 must be a C-contiguous float64 [n, 64] array (it is updated in place)")
        n = seq2.shape[0]
        o = _lib.RolloutOptions()
        o.substeps, o.wbc_per_mpc, o.control_dt = int(substeps), int(wbc_per_mpc), float(control_dt)
        o.plant_inertia = {"physical": 0, "reference_model": 1}[plant_inertia]
        g = gains or dict(com_pose_Kp=[300.0] * 3 + [200.0] * 3, com_pose_Kd=[100.0] * 6, feet_pose_Kp=[1200.0] * 3, feet_pose_Kd=[500.0, 500.0, 400.0])
        for k in ("com_pose_Kp", "com_pose_Kd", "feet_pose_Kp", "feet_pose_Kd"):
            getattr(o, k)[:] = [float(v) for v in g[k]]
        hist = np.zeros((periods, n, 16)) if history else None
        rc = self._lib.b200qp_mpc_rollout(self._h, wbc._h if wbc is not None else None, n, int(periods), seq2.ctypes.data, C.byref(o),
                                          hist.ctypes.data if hist is not None else None)
        _lib.check(rc, "b200qp_mpc_rollout")
        return hist

    def sequence_and_solve_device(self, n, d_seq_in, d_measured, d_forces, d_xpred, d_info, sync=True):
        rc = self._lib.b200qp_mpc_sequence_solve_batch_device(self._h, n, d_seq_in, d_measured, d_forces, d_xpred, d_info,
                                                              1 if sync else 0)
        _lib.check(rc, "b200qp_mpc_sequence_solve_batch_device")

    def solve_device(self, n, d_in, d_contact, d_forces, d_xpred, d_info, sync=True):
        """b200qp_mpc_solve_batch_device on raw device pointers (ints)."""
        rc = self._lib.b200qp_mpc_solve_batch_device(self._h, n, d_in, d_contact, d_forces, d_xpred, d_info, 1 if sync else 0)
        _lib.check(rc, "b200qp_mpc_solve_batch_device")

    def last_timing(self):
        ms, nl = C.c_float(), C.c_int32()
        _lib.check(self._lib.b200qp_mpc_last_timing(self._h, C.byref(ms), C.byref(nl)), "last_timing")
        return ms.value, nl.value

    def stream(self) -> int:
        return int(self._lib.b200qp_mpc_stream(self._h) or 0)

    def get_duals(self, n):
        lam_box = np.zeros((n, self.N, 12))
        lam_g = np.zeros((n, self.N, 16))
        _lib.check(self._lib.b200qp_mpc_get_duals(self._h, n, lam_box.ctypes.data, lam_g.ctypes.data), "get_duals")
        return lam_box, lam_g

    def get_condensed(self, index):
        nmax = 12 * self.N
        H = np.zeros((nmax, nmax), order="F")
        g = np.zeros(nmax)
        nf = self._lib.b200qp_mpc_get_condensed(self._h, int(index), H.ctypes.data, nmax, g.ctypes.data)
        if nf < 0:
            raise _lib.B200QPError(f"b200qp_mpc_get_condensed failed with code {nf}: {_lib.last_error()}")
        return np.array(H[:nf, :nf]), g[:nf].copy()

    def _info_buf(self):
        if getattr(self, "_ibuf", None) is None:
            self._ibuf = (_lib.SolverInfo * max(self.max_batch, 1))()  # reused: _info() copies out of it
        return self._ibuf

    def _info(self, info, n):
        # one bulk copy of the 48-byte records; the fields are views into that private copy
        arr = np.frombuffer(info, dtype=np.uint8, count=48 * n).copy().view(_INFO_DTYPE)  # (a structured .copy() is 10x slower)
        ms, nl = self.last_timing()
        return SolverInformation(success=arr["status"] == 0, return_code=arr["status"], acados_num_iter=arr["iterations"],
                                 polish_rounds=arr["polish_rounds"], n_free=arr["n_free"], kkt=arr["kkt"],
                                 total_solver_time=ms * 1e-3, kernel_launches=nl)


class ShardedMPC:
    """One batch over several GPUs of one box through ONE library call (b200qp_mpc_create_sharded /
    b200qp_mpc_solve_batch_sharded): contiguous slices, one host thread + stream + staging per GPU, no
    collective.  Constructor arguments as BatchedMPC plus `devices`; `max_batch` is the size of the whole batch."""

    def __init__(self, alpha, state_weights, mu, fmin, fmax, solver="CONDENSED_ADMM", condensing_N=None, warm_start=0, *,
                 devices=(0,), horizon=10, dt=0.05, max_batch=4096, **opts):
        self._lib = _lib.load()
        cfg = _lib.MpcConfig()
        cfg.horizon, cfg.solver, cfg.max_batch, cfg.device = int(horizon), _lib.SOLVERS[solver], int(max_batch), 0
        cfg.dt, cfg.alpha, cfg.mu, cfg.fmin, cfg.fmax = float(dt), float(alpha), float(mu), float(fmin), float(fmax)
        sw = np.asarray(state_weights, dtype=np.float64).reshape(12)
        for i in range(12):
            cfg.state_weights[i] = sw[i]
        cfg.cond_N = int(condensing_N) if condensing_N else 0
        cfg.warm_start = int(warm_start)
        for k, v in opts.items():
            setattr(cfg, k, v)
        self.N, self.max_batch, self.devices = int(horizon), int(max_batch), tuple(int(d) for d in devices)
        devs = (C.c_int32 * len(self.devices))(*self.devices)
        self._h = C.c_void_p()
        _lib.check(self._lib.b200qp_mpc_create_sharded(C.byref(cfg), devs, len(self.devices), C.byref(self._h)), "b200qp_mpc_create_sharded")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200qp_mpc_destroy_sharded(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve_records(self, rec, contact, forces=None, xpred=None):
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        contact = np.ascontiguousarray(contact, dtype=np.uint8)
        n = rec.shape[0]
        if rec.shape != (n, in_doubles(self.N)) or contact.shape != (n, self.N, 4):
            raise ValueError("record / contact shape mismatch")
        forces = np.zeros((n, self.N, 12)) if forces is None else forces
        xpred = np.zeros((n, self.N + 1, 13)) if xpred is None else xpred
        info = (_lib.SolverInfo * max(n, 1))()
        rc = self._lib.b200qp_mpc_solve_batch_sharded(self._h, n, rec.ctypes.data, contact.ctypes.data, forces.ctypes.data,
                                                      xpred.ctypes.data, C.addressof(info))
        _lib.check(rc, "b200qp_mpc_solve_batch_sharded")
        arr = np.frombuffer(info, dtype=np.dtype([("status", "i4"), ("iterations", "i4"), ("polish_rounds", "i4"), ("n_free", "i4"),
                                                  ("kkt", "f8", (4,))]), count=n).copy()
        return forces, xpred, arr
