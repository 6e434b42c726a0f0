"""ctypes binding of oracle/_ref/libdfki_ref_h<N>.so - the REFERENCE'S OWN sources (gait.cpp, mpc.cpp, the trajectory
planner, the Raibert planner, both gait sequencers), compiled where they lie under /root/reference by
oracle/build_ref.sh behind the C entry points of oracle/ref_harness.cpp.   TEST INFRASTRUCTURE ONLY.

This is what pins the oracle (and through it the CUDA path): contact masks / swing times / bounds come out of the
reference's `Gait::update_sequence` and `MPC::GetUBounds` byte for byte, the QP stage data (A, B_k, q_k, x0, D, lg, ug)
out of the reference's `MPC::GetWrenchSequence`, the sequencer records out of `SimpleGaitSequencer::GetGaitSequence`.
Third-party arithmetic that is NOT the reference's: Eigen (oracle/ref_shim/mini_eigen.hpp) and the QP solve behind the
acados call (oracle/mpc_ref.c).  `/root/reference` is not needed at run time, only the built library."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs: dict[int, C.CDLL] = {}

# acados plan names of the stand-in header (oracle/ref_shim/acados_c/ocp_qp_interface.h)
SOLVERS = dict(PARTIAL_CONDENSING_HPIPM=0, PARTIAL_CONDENSING_OSQP=1, PARTIAL_CONDENSING_QPDUNES=2, FULL_CONDENSING_HPIPM=3,
               FULL_CONDENSING_QPOASES=4, FULL_CONDENSING_DAQP=5)
# GaitDatabase::GaitType order (gait.hpp)
GAIT_TYPES = ["STAND", "STATIC_WALK", "WALKING_TROT", "TROT", "FLYING_TROT", "PACE", "BOUND", "ROTARY_GALLOP", "TRAVERSE_GALLOP",
              "PRONK"]
TARGET_FLAGS = ["x", "y", "z", "roll", "pitch", "yaw", "full_orientation", "x_dot", "y_dot", "z_dot", "hybrid_x_dot",
                "hybrid_y_dot", "hybrid_z_dot", "wx", "wy", "wz"]  # Target::activateTarget, target.hpp:28-45
TARGET_VALUES = ["x", "y", "z", "roll", "pitch", "yaw", "qx", "qy", "qz", "qw", "x_dot", "y_dot", "z_dot", "hybrid_x_dot",
                 "hybrid_y_dot", "hybrid_z_dot", "wx", "wy", "wz"]


def lib_path(N: int) -> str:
    return os.path.join(_HERE, "_ref", f"libdfki_ref_h{N}.so")


def available(N: int = 10) -> bool:
    return os.path.exists(lib_path(N))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def load(N: int = 10) -> C.CDLL:
    if N not in _libs:
        lib = C.CDLL(lib_path(N))
        lib.dfkiref_horizon.restype = C.c_int
        lib.dfkiref_last_error.restype = C.c_char_p
        lib.dfkiref_mpc_create.restype = C.c_void_p
        lib.dfkiref_mpc_create.argtypes = [C.c_double, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_char_p,
                                           C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        lib.dfkiref_mpc_destroy.argtypes = [C.c_void_p]
        lib.dfkiref_mpc_solve.restype = C.c_int
        lib.dfkiref_mpc_solve.argtypes = [C.c_void_p] * 6 + [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.dfkiref_mpc_last_qp.argtypes = [C.c_void_p] * 13
        lib.dfkiref_mpc_set_state_weights.argtypes = [C.c_void_p, C.c_void_p]
        lib.dfkiref_mpc_set_input_weights.argtypes = [C.c_void_p, C.c_double]
        lib.dfkiref_mpc_set_fmax.argtypes = [C.c_void_p, C.c_double]
        lib.dfkiref_mpc_set_mu.argtypes = [C.c_void_p, C.c_double]
        lib.dfkiref_set_solver_options.argtypes = [C.c_double, C.c_int]
        lib.dfkiref_gait_database.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p]
        lib.dfkiref_gait_update_sequence.argtypes = [C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 9
        lib.dfkiref_gait_contact_sequence.argtypes = [C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 6
        lib.dfkiref_seq_create.restype = C.c_void_p
        lib.dfkiref_seq_create.argtypes = [C.c_void_p] * 4
        lib.dfkiref_seq_destroy.argtypes = [C.c_void_p]
        lib.dfkiref_seq_update_state.argtypes = [C.c_void_p] * 4 + [C.c_double]
        lib.dfkiref_seq_update_target.argtypes = [C.c_void_p] * 3
        lib.dfkiref_seq_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 11
        assert lib.dfkiref_horizon() == N
        _libs[N] = lib
    return _libs[N]


# ---------------------------------------------------------------------------------------------------------- gait
def gait_database(gait: str | int, dt: float = 0.05, N: int = 10):
    """GaitDatabase::getGait (gait.cpp:732-767) -> (period, duty[4], offset[4])."""
    lib = load(N)
    t = GAIT_TYPES.index(gait) if isinstance(gait, str) else int(gait)
    period = C.c_double()
    duty = np.zeros(4)
    off = np.zeros(4)
    lib.dfkiref_gait_database(t, dt, C.byref(period), _p(duty), _p(off), None)
    return period.value, duty, off


def gait_update_sequence(period, duty, offset, phase, dt=0.05, steps=100, measured=None, N: int = 10):
    """Gait::update_sequence (gait.cpp:67-112) for n (gait, phase) pairs -> contact uint8 [n][steps][4], swing_time, t_swing,
    keep_mode."""
    lib = load(N)
    period = np.ascontiguousarray(np.atleast_1d(period), dtype=np.float64)
    n = period.shape[0]
    duty = np.ascontiguousarray(np.broadcast_to(duty, (n, 4)), dtype=np.float64)
    offset = np.ascontiguousarray(np.broadcast_to(offset, (n, 4)), dtype=np.float64)
    phase = np.ascontiguousarray(np.broadcast_to(phase, (n,)), dtype=np.float64)
    meas = None if measured is None else np.ascontiguousarray(np.broadcast_to(measured, (n, 4)), dtype=np.uint8)
    contact = np.zeros((n, steps, 4), dtype=np.uint8)
    swing = np.zeros((n, steps, 4))
    t_swing = np.zeros((n, 4))
    keep = np.zeros(n, dtype=np.int32)
    lib.dfkiref_gait_update_sequence(n, steps, dt, _p(period), _p(duty), _p(offset), _p(phase), _p(meas), _p(contact), _p(swing),
                                     _p(t_swing), _p(keep))
    return contact, swing, t_swing, keep


def gait_contact_sequence(period, duty, offset, phase, dt=0.05, steps=100, N: int = 10):
    """Gait::get_contact_sequence / get_swing_time_sequence (gait.cpp:30-65)."""
    lib = load(N)
    period = np.ascontiguousarray(np.atleast_1d(period), dtype=np.float64)
    n = period.shape[0]
    duty = np.ascontiguousarray(np.broadcast_to(duty, (n, 4)), dtype=np.float64)
    offset = np.ascontiguousarray(np.broadcast_to(offset, (n, 4)), dtype=np.float64)
    phase = np.ascontiguousarray(np.broadcast_to(phase, (n,)), dtype=np.float64)
    contact = np.zeros((n, steps, 4), dtype=np.uint8)
    swing = np.zeros((n, steps, 4))
    lib.dfkiref_gait_contact_sequence(n, steps, dt, _p(period), _p(duty), _p(offset), _p(phase), _p(contact), _p(swing))
    return contact, swing


# ---------------------------------------------------------------------------------------------------------- MPC
class RefMPC:
    """The reference's `class MPC` (mpc.hpp:13-144) behind its own constructor arguments."""

    def __init__(self, N, params, weights=None, solver="PARTIAL_CONDENSING_OSQP", cond_N=None, hpipm_mode="BALANCE", warm_start=0,
                 mass=15.019, inertia=None, com=(0.0, 0.0, 0.0)):
        self.N = N
        self.lib = load(N)
        w = np.ascontiguousarray(params["weights"] if weights is None else weights, dtype=np.float64)
        inertia = np.eye(3) if inertia is None else np.asarray(inertia, dtype=np.float64)
        icm = np.ascontiguousarray(inertia.T.reshape(-1))
        com = np.ascontiguousarray(com, dtype=np.float64)
        self.h = self.lib.dfkiref_mpc_create(params["alpha"], _p(w), params["mu"], params["fmin"], params["fmax"], SOLVERS[solver],
                                             int(cond_N if cond_N is not None else N), hpipm_mode.encode(), warm_start, mass,
                                             _p(icm), _p(com))

    def close(self):
        if self.h:
            self.lib.dfkiref_mpc_destroy(self.h)
            self.h = None

    __del__ = close

    def set_state_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.lib.dfkiref_mpc_set_state_weights(self.h, _p(w))

    def set_input_weights(self, alpha):
        self.lib.dfkiref_mpc_set_input_weights(self.h, float(alpha))

    def set_fmax(self, fmax):
        self.lib.dfkiref_mpc_set_fmax(self.h, float(fmax))

    def set_mu(self, mu):
        self.lib.dfkiref_mpc_set_mu(self.h, float(mu))

    def solve(self, rec, contact, forces=None, xpred=None):
        """UpdateState / UpdateModel / UpdateGaitSequence / GetWrenchSequence on one record.  `forces` / `xpred` are only
        written on success (mpc.cpp:447-455); returns dict(return_code, success, num_iter, forces, xpred, pose)."""
        N = self.N
        rec = np.ascontiguousarray(rec, dtype=np.float64)
        contact = np.ascontiguousarray(contact, dtype=np.uint8)
        forces = np.zeros((N, 12)) if forces is None else forces
        xpred = np.zeros((N + 1, 13)) if xpred is None else xpred
        pose = np.zeros((N + 1, 13))
        ok, it = C.c_int(), C.c_int()
        rc = self.lib.dfkiref_mpc_solve(self.h, _p(rec), _p(contact), _p(forces), _p(xpred), _p(pose), C.byref(ok), C.byref(it))
        return dict(return_code=rc, success=bool(ok.value), num_iter=it.value, forces=forces, xpred=xpred, pose=pose)

    def last_qp(self):
        """Stage data of the last solve exactly as the reference handed it to acados (numpy 2-D arrays, row = matrix row)."""
        N = self.N
        A = np.zeros((N, 13, 13))
        B = np.zeros((N, 12, 13))
        q = np.zeros((N + 1, 13))
        lbu = np.zeros((N, 12))
        ubu = np.zeros((N, 12))
        x0 = np.zeros(13)
        D = np.zeros((12, 16))
        lg = np.zeros(16)
        ug = np.zeros(16)
        Q = np.zeros((13, 13))
        R = np.zeros((12, 12))
        kkt = np.zeros(4)
        self.lib.dfkiref_mpc_last_qp(self.h, _p(A), _p(B), _p(q), _p(lbu), _p(ubu), _p(x0), _p(D), _p(lg), _p(ug), _p(Q), _p(R), _p(kkt))
        # column-major -> numpy
        return dict(A=A.transpose(0, 2, 1).copy(), B=B.transpose(0, 2, 1).copy(), q=q, lbu=lbu, ubu=ubu, x0=x0, D=D.T.copy(), lg=lg,
                    ug=ug, Q=Q.T.copy(), R=R.T.copy(), kkt=kkt)


def solve_batch(N, recs, contacts, params, weights=None, mass=15.019, inertia=None):
    """Reference MPC over a batch (sequential, one MPC object)."""
    m = RefMPC(N, params, weights=weights, mass=mass, inertia=inertia)
    n = recs.shape[0]
    forces = np.zeros((n, N, 12))
    xpred = np.zeros((n, N + 1, 13))
    ok = np.zeros(n, dtype=bool)
    for p in range(n):
        r = m.solve(recs[p], contacts[p], forces[p], xpred[p])
        ok[p] = r["success"]
    m.close()
    return forces, xpred, ok


# ---------------------------------------------------------------------------------------------------------- sequencers
class _SeqConfig(C.Structure):
    _fields_ = [("adaptive", C.c_int), ("period", C.c_double), ("duty", C.c_double * 4), ("offset", C.c_double * 4), ("dt", C.c_double),
                ("raibert_k", C.c_double), ("shoulders", C.c_double * 12), ("raibert_filtersize", C.c_int),
                ("raibert_z_on_plane", C.c_int), ("fix_standing_position", C.c_int),
                ("fix_position_distance_threshold", C.c_double), ("fix_position_angular_threshold", C.c_double),
                ("fix_position_velocity_threshold", C.c_double), ("early_contact_detection", C.c_int), ("mass", C.c_double),
                ("inertia_cm", C.c_double * 9), ("com", C.c_double * 3), ("max_leg_length", C.c_double),
                ("swing_time", C.c_double), ("zero_velocity_threshold", C.c_double),
                ("standing_foot_position_threshold", C.c_double), ("min_v_cmd_factor", C.c_double),
                ("max_correction_cycles", C.c_double), ("correction_period", C.c_double), ("disturbance_correction", C.c_double),
                ("min_v", C.c_double), ("offset_delay", C.c_double), ("max_stride_length", C.c_double), ("filter_size", C.c_int),
                ("switch_offsets", C.c_int), ("correct_all", C.c_int)]


# defaults: config/mit_controller_sim_go2.yaml (simple sequencer + planner) and mit_controller_node.cpp:250-264 (adaptive gait)
SEQ_DEFAULTS = dict(dt=0.05, raibert_k=0.03, raibert_filtersize=20, raibert_z_on_plane=0, fix_standing_position=1,
                    fix_position_distance_threshold=0.1, fix_position_angular_threshold=0.26, fix_position_velocity_threshold=0.1,
                    early_contact_detection=1, mass=15.019, max_leg_length=0.406, swing_time=0.2, zero_velocity_threshold=0.03,
                    standing_foot_position_threshold=0.08, min_v_cmd_factor=0.5, max_correction_cycles=2.0, correction_period=0.6,
                    disturbance_correction=0.0, min_v=0.0, offset_delay=0.0, max_stride_length=0.0, filter_size=10,
                    switch_offsets=1, correct_all=0)


class RefSequencer:
    """SimpleGaitSequencer / AdaptiveGaitSequencer of the reference, stateful across calls like the node's `gs_`."""

    def __init__(self, N, state13, feet, contacts=None, adaptive=False, period=0.5, duty=(0.5,) * 4, offset=(0.0, 0.5, 0.5, 0.0),
                 shoulders=None, inertia=None, com=(0.0, 0.0, 0.0), **kw):
        self.N = N
        self.lib = load(N)
        cfg = dict(SEQ_DEFAULTS)
        cfg.update(kw)
        c = _SeqConfig()
        c.adaptive = 1 if adaptive else 0
        c.period = period
        for i in range(4):
            c.duty[i] = float(duty[i])
            c.offset[i] = float(offset[i])
        sh = np.asarray(shoulders, dtype=np.float64).reshape(-1)
        for i in range(12):
            c.shoulders[i] = sh[i]
        icm = (np.eye(3) if inertia is None else np.asarray(inertia, dtype=np.float64)).T.reshape(-1)
        for i in range(9):
            c.inertia_cm[i] = icm[i]
        for i in range(3):
            c.com[i] = float(com[i])
        for k, v in cfg.items():
            setattr(c, k, v)
        self._cfg = c
        st = np.ascontiguousarray(state13, dtype=np.float64)
        ft = np.ascontiguousarray(feet, dtype=np.float64).reshape(-1)
        ct = None if contacts is None else np.ascontiguousarray(contacts, dtype=np.uint8)
        self.h = self.lib.dfkiref_seq_create(C.byref(c), _p(st), _p(ft), _p(ct))

    def close(self):
        if self.h:
            self.lib.dfkiref_seq_destroy(self.h)
            self.h = None

    __del__ = close

    def update_state(self, state13, feet, contacts=None, time_s=0.0):
        st = np.ascontiguousarray(state13, dtype=np.float64)
        ft = np.ascontiguousarray(feet, dtype=np.float64).reshape(-1)
        ct = None if contacts is None else np.ascontiguousarray(contacts, dtype=np.uint8)
        self.lib.dfkiref_seq_update_state(self.h, _p(st), _p(ft), _p(ct), float(time_s))

    def update_target(self, **target):
        """Keyword = field of `Target` (target.hpp); every given field is activated, `full_orientation=(x, y, z, w)`."""
        vals = np.zeros(19)
        vals[9] = 1.0
        act = np.zeros(16, dtype=np.uint8)
        for k, v in target.items():
            if k == "full_orientation":
                vals[6:10] = v
            else:
                vals[TARGET_VALUES.index(k)] = v
            act[TARGET_FLAGS.index(k)] = 1
        self.lib.dfkiref_seq_update_target(self.h, _p(vals), _p(act))

    def get(self, steps=None):
        steps = self.N + 1 if steps is None else steps
        out = dict(contact=np.zeros((steps, 4), dtype=np.uint8), swing_time=np.zeros((steps, 4)), foot_pos=np.zeros((steps, 4, 3)),
                   ref_pos=np.zeros((steps, 3)), ref_ori=np.zeros((steps, 4)), des_pos=np.zeros((steps, 3)),
                   des_ori=np.zeros((steps, 4)), ref_vel=np.zeros((steps, 3)), ref_twist=np.zeros((steps, 3)), scalars=np.zeros(8),
                   gait_state=np.zeros(18))
        self.lib.dfkiref_seq_get(self.h, steps, *[_p(out[k]) for k in ("contact", "swing_time", "foot_pos", "ref_pos", "ref_ori",
                                                                        "des_pos", "des_ori", "ref_vel", "ref_twist", "scalars",
                                                                        "gait_state")])
        return out

    def record(self, state13, mass, inertia, com, seq=None):
        """The b200qp input record + contact bytes that MPC::UpdateGaitSequence would see after this GetGaitSequence."""
        N = self.N
        s = self.get(N + 1) if seq is None else seq
        rec = np.zeros(26 + 13 * (N + 1) + 12 * N)
        rec[0:13] = state13
        rec[13] = mass
        rec[14:23] = np.asarray(inertia).reshape(3, 3).T.reshape(-1)
        rec[23:26] = com
        for k in range(N + 1):
            o = 26 + 13 * k
            rec[o:o + 4] = s["des_ori"][k]
            rec[o + 4:o + 7] = s["des_pos"][k]
            rec[o + 7:o + 10] = s["ref_twist"][k]
            rec[o + 10:o + 13] = s["ref_vel"][k]
        rec[26 + 13 * (N + 1):] = s["foot_pos"][:N].reshape(-1)
        return rec, s["contact"][:N].copy()
