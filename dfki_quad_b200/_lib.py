"""ctypes binding of libb200qp.so (include/b200qp.h).  Fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200QP_LIB") or os.path.join(_HERE, "libb200qp.so")  # override: A/B runs of development builds

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3
ACADOS_SUCCESS, ACADOS_NAN_DETECTED, ACADOS_MAXITER, ACADOS_MINSTEP, ACADOS_QP_FAILURE, ACADOS_READY = range(6)
NAN_AFTER_SUCCESS = -7777
SOLVER_CONDENSED_ADMM, SOLVER_SPARSE_RICCATI, SOLVER_AUTO, SOLVER_PARTIAL_CONDENSING = 0, 1, 2, 3
SOLVERS = {"CONDENSED_ADMM": 0, "SPARSE_RICCATI": 1, "AUTO": 2, "PARTIAL_CONDENSING": 3,
           # The reference's acados plan names (mit_controller_node.cpp:502-519).  Its solvers take any problem size, so the
           # condensing plans map onto AUTO (condensed kernel where the reduced problem fits, Riccati kernel otherwise);
           # the HPIPM partial-condensing plan maps onto the block-condensed Riccati interior-point kernel, which takes the
           # reference's `condensing_N` (any value in 1..N, same optimum).
           "PARTIAL_CONDENSING_OSQP": 2, "FULL_CONDENSING_HPIPM": 2, "FULL_CONDENSING_QPOASES": 2,
           "FULL_CONDENSING_DAQP": 2, "PARTIAL_CONDENSING_HPIPM": 3}


class MpcConfig(C.Structure):
    _fields_ = [
        ("horizon", C.c_int32), ("solver", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32),
        ("dt", C.c_double), ("alpha", C.c_double), ("state_weights", C.c_double * 12),
        ("mu", C.c_double), ("fmin", C.c_double), ("fmax", C.c_double),
        ("admm_rho", C.c_double), ("admm_sigma", C.c_double), ("admm_alpha", C.c_double), ("admm_eps", C.c_double),
        ("kkt_tol", C.c_double),
        ("admm_max_iter", C.c_int32), ("admm_check_every", C.c_int32), ("polish_max_rounds", C.c_int32),
        ("ipm_max_iter", C.c_int32), ("warm_start", C.c_int32), ("cond_N", C.c_int32), ("precision", C.c_int32),
    ]


class SolverInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("polish_rounds", C.c_int32), ("n_free", C.c_int32),
                ("kkt", C.c_double * 4)]


class WbcConfig(C.Structure):
    _fields_ = [("nq", C.c_int32), ("neq", C.c_int32), ("nin", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32),
                ("max_iter", C.c_int32), ("kkt_tol", C.c_double)]


class SeqOptions(C.Structure):
    _fields_ = [("mode", C.c_int32), ("raibert_filtersize", C.c_int32), ("fix_standing_position", C.c_int32),
                ("early_contact_detection", C.c_int32), ("fix_position_distance_threshold", C.c_double),
                ("fix_position_angular_threshold", C.c_double), ("fix_position_velocity_threshold", C.c_double),
                ("swing_time", C.c_double), ("zero_velocity_threshold", C.c_double), ("standing_foot_position_threshold", C.c_double),
                ("min_v_cmd_factor", C.c_double), ("max_correction_cycles", C.c_double), ("correction_period", C.c_double),
                ("disturbance_correction", C.c_double), ("min_v", C.c_double), ("offset_delay", C.c_double),
                ("max_stride_length", C.c_double), ("phase_offset", C.c_double * 4), ("control_dt", C.c_double),
                ("filter_size", C.c_int32), ("switch_offsets", C.c_int32), ("correct_all", C.c_int32)]


class WbcSceneConfig(C.Structure):
    _fields_ = [("mu", C.c_double), ("com_pose_weight", C.c_double * 6), ("foot_force_weight", C.c_double * 3),
                ("foot_pose_weight", C.c_double * 3), ("hessian_regularizer", C.c_double), ("tau_max", C.c_double * 12)]


class RolloutOptions(C.Structure):
    _fields_ = [("substeps", C.c_int32), ("plant_inertia", C.c_int32), ("wbc_per_mpc", C.c_int32), ("reserved", C.c_int32),
                ("com_pose_Kp", C.c_double * 6), ("com_pose_Kd", C.c_double * 6), ("feet_pose_Kp", C.c_double * 3),
                ("feet_pose_Kd", C.c_double * 3), ("control_dt", C.c_double)]


class SeqModel(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 9), ("com", C.c_double * 3), ("shoulders", C.c_double * 12),
                ("raibert_k", C.c_double), ("raibert_g", C.c_double), ("max_leg_length", C.c_double)]


EXPORTS = [
    "b200qp_version", "b200qp_device_count", "b200qp_last_error", "b200qp_fp64_peak_tflops", "b200qp_mpc_create", "b200qp_mpc_destroy",
    "b200qp_mpc_set_state_weights", "b200qp_mpc_set_input_weights", "b200qp_mpc_set_fmax", "b200qp_mpc_set_mu", "b200qp_mpc_set_warm_start",
    "b200qp_mpc_solve_batch", "b200qp_mpc_solve_batch_device", "b200qp_mpc_get_duals", "b200qp_mpc_get_condensed",
    "b200qp_mpc_last_timing", "b200qp_mpc_stream", "b200qp_mpc_set_model", "b200qp_mpc_sequence_batch",
    "b200qp_mpc_sequence_solve_batch", "b200qp_mpc_sequence_solve_batch_device", "b200qp_gait_schedule", "b200qp_gait_schedule_device",
    "b200qp_mpc_sequencer_configure", "b200qp_mpc_sequencer_reset", "b200qp_mpc_sequencer_step", "b200qp_mpc_sequencer_step_solve",
    "b200qp_wbc_scene_configure", "b200qp_wbc_scene_solve_batch", "b200qp_wbc_scene_assemble", "b200qp_wbc_scene_solve_batch_device",
    "b200qp_mpc_rollout",
    "b200qp_mpc_create_sharded", "b200qp_mpc_destroy_sharded", "b200qp_mpc_solve_batch_sharded", "b200qp_mpc_sharded_handle",
    "b200qp_mpc_sharded_devices", "b200qp_shard_range",
    "b200qp_wbc_create", "b200qp_wbc_destroy", "b200qp_wbc_solve_batch", "b200qp_wbc_solve_batch_device", "b200qp_wbc_stream",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C dfki_quad_b200/csrc` (sm_100a CUDA library; this package has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    p, i32, dbl, vp = C.POINTER, C.c_int32, C.c_double, C.c_void_p
    lib.b200qp_version.restype = C.c_char_p
    lib.b200qp_last_error.restype = C.c_char_p
    lib.b200qp_device_count.restype = i32
    lib.b200qp_fp64_peak_tflops.argtypes = [i32, p(dbl)]
    lib.b200qp_mpc_create.argtypes = [p(MpcConfig), p(vp)]
    lib.b200qp_mpc_create.restype = i32
    lib.b200qp_mpc_destroy.argtypes = [vp]
    lib.b200qp_mpc_destroy.restype = None
    lib.b200qp_mpc_set_state_weights.argtypes = [vp, p(dbl)]
    lib.b200qp_mpc_set_input_weights.argtypes = [vp, dbl]
    lib.b200qp_mpc_set_fmax.argtypes = [vp, dbl]
    lib.b200qp_mpc_set_mu.argtypes = [vp, dbl]
    lib.b200qp_mpc_set_warm_start.argtypes = [vp, i32]
    lib.b200qp_mpc_solve_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b200qp_mpc_solve_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32]
    lib.b200qp_mpc_get_duals.argtypes = [vp, i32, vp, vp]
    lib.b200qp_mpc_get_condensed.argtypes = [vp, i32, vp, i32, vp]
    lib.b200qp_mpc_last_timing.argtypes = [vp, p(C.c_float), p(i32)]
    lib.b200qp_mpc_stream.argtypes = [vp]
    lib.b200qp_mpc_set_model.argtypes = [vp, p(SeqModel)]
    lib.b200qp_mpc_sequence_batch.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.b200qp_mpc_sequence_solve_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b200qp_mpc_sequence_solve_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32]
    lib.b200qp_mpc_stream.restype = vp
    lib.b200qp_mpc_sequencer_configure.argtypes = [vp, p(SeqOptions)]
    lib.b200qp_mpc_sequencer_reset.argtypes = [vp]
    lib.b200qp_mpc_sequencer_step.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.b200qp_mpc_sequencer_step_solve.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b200qp_wbc_scene_configure.argtypes = [vp, p(WbcSceneConfig)]
    lib.b200qp_wbc_scene_solve_batch.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.b200qp_wbc_scene_assemble.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.b200qp_wbc_scene_solve_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    lib.b200qp_mpc_rollout.argtypes = [vp, vp, i32, i32, vp, p(RolloutOptions), vp]
    lib.b200qp_mpc_create_sharded.argtypes = [p(MpcConfig), p(i32), i32, p(vp)]
    lib.b200qp_mpc_destroy_sharded.argtypes = [vp]
    lib.b200qp_mpc_destroy_sharded.restype = None
    lib.b200qp_mpc_solve_batch_sharded.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    lib.b200qp_mpc_sharded_handle.argtypes = [vp, i32]
    lib.b200qp_mpc_sharded_handle.restype = vp
    lib.b200qp_mpc_sharded_devices.argtypes = [vp]
    lib.b200qp_shard_range.argtypes = [i32, i32, i32, p(i32), p(i32)]
    lib.b200qp_gait_schedule.argtypes = [i32, i32, i32, dbl, vp, vp, vp, vp, vp, dbl, dbl, vp, vp, vp, vp]
    lib.b200qp_gait_schedule_device.argtypes = [i32, i32, dbl, vp, vp, vp, vp, vp, dbl, dbl, vp, vp, vp, vp, vp]
    lib.b200qp_wbc_create.argtypes = [p(WbcConfig), p(vp)]
    lib.b200qp_wbc_destroy.argtypes = [vp]
    lib.b200qp_wbc_destroy.restype = None
    lib.b200qp_wbc_solve_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.b200qp_wbc_solve_batch_device.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32]
    lib.b200qp_wbc_stream.argtypes = [vp]
    lib.b200qp_wbc_stream.restype = vp
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:  # default restype
            fn.restype = i32
    _lib = lib
    return lib


def fp64_peak_tflops(device=0) -> float:
    v = C.c_double()
    check(load().b200qp_fp64_peak_tflops(device, C.byref(v)), "b200qp_fp64_peak_tflops")
    return v.value


def last_error() -> str:
    return load().b200qp_last_error().decode()


class B200QPError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != OK:
        raise B200QPError(f"{what} failed with code {rc}: {last_error()}")
