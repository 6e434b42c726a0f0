"""Host build of the one-thread-per-robot sequencer kernel (dfki_quad_b200/csrc/sequencer_full.cu) - TEST INFRASTRUCTURE.
g++ compiles the kernel source as plain C++ (emu_preamble.h blanks the CUDA qualifiers) so that the kernel's logic can be compared
with the compiled reference in the container that has no GPU.  The product never loads this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libseqfull_host.so")
CUDA_INC = "/usr/local/cuda/include"


def available():
    return os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")) and os.path.exists("/usr/bin/g++")


def build(host="seq_full_host.cpp", kernel="sequencer_full.cu", so=None):
    so = so or SO
    src = [os.path.join(HERE, host), os.path.join(HERE, "emu_preamble.h"),
           os.path.join(ROOT, "dfki_quad_b200", "csrc", kernel), os.path.join(ROOT, "dfki_quad_b200", "csrc", "common.cuh")]
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in src):
        return so
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-w", "-I" + CUDA_INC, "-o", so, src[0]])
    return so


class _Model(C.Structure):  # b200qp::SeqModel (common.cuh)
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 9), ("com", C.c_double * 3), ("shoulders", C.c_double * 12),
                ("raibert_k", C.c_double), ("raibert_g", C.c_double), ("max_leg_length", C.c_double), ("dt", C.c_double), ("N", C.c_int)]


class HostSequencer:
    """Same call sequence as BatchedMPC.sequencer_configure / sequencer_step, on the host build of the kernel."""

    def __init__(self, N, dt, mass, inertia, com, shoulders, raibert_k, raibert_g, max_leg_length, mode="simple", **options):
        from dfki_quad_b200 import _lib
        from dfki_quad_b200.mpc import SEQ_OPTION_DEFAULTS, in_doubles

        self.lib = C.CDLL(build())
        self.N, self.stride = N, in_doubles(N)
        m = _Model()
        m.mass = mass
        m.inertia[:] = list(np.asarray(inertia, float).reshape(9))
        m.com[:] = list(np.asarray(com, float).reshape(3))
        m.shoulders[:] = list(np.asarray(shoulders, float).reshape(12))
        m.raibert_k, m.raibert_g, m.max_leg_length, m.dt, m.N = raibert_k, raibert_g, max_leg_length, dt, N
        self.m = m
        o = _lib.SeqOptions()
        vals = dict(SEQ_OPTION_DEFAULTS)
        vals.update(options)
        o.mode = {"simple": 0, "adaptive": 1}[mode]
        for k, v in vals.items():
            if k == "phase_offset":
                o.phase_offset[:] = [float(x) for x in v]
            elif isinstance(getattr(o, k), int):
                setattr(o, k, int(v))
            else:
                setattr(o, k, float(v))
        self.o = o
        self.state = None

    def sequencer_reset(self):
        self.state = None

    def sequencer_step(self, seq2_in, measured=None):
        seq2_in = np.ascontiguousarray(seq2_in, dtype=np.float64)
        n = seq2_in.shape[0]
        if self.state is None:
            self.state = np.zeros((n, 192))
        rec = np.zeros((n, self.stride))
        ct = np.zeros((n, self.N, 4), dtype=np.uint8)
        mc = None if measured is None else np.ascontiguousarray(measured, dtype=np.uint8)
        self.lib.seqfull_host_step(n, C.byref(self.m), C.byref(self.o), seq2_in.ctypes.data_as(C.c_void_p),
                                   None if mc is None else mc.ctypes.data_as(C.c_void_p), self.state.ctypes.data_as(C.c_void_p),
                                   rec.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p), self.stride)
        return rec, ct


class _SceneConfig(C.Structure):  # b200qp::WbcSceneConfig (wbc_scene.cu)
    _fields_ = [("mu", C.c_double), ("com_pose_weight", C.c_double * 6), ("foot_force_weight", C.c_double * 3),
                ("foot_pose_weight", C.c_double * 3), ("hessian_regularizer", C.c_double), ("tau_max", C.c_double * 12)]


def host_wbc_scene(scene_in, cfg, tau_max):
    """Host build of wbc_scene_kernel: scene_in [n,72] -> (H [n,900], g [n,30], A [n,1380], lo [n,46], hi [n,46], feet [n,24])."""
    lib = C.CDLL(build("wbc_scene_host.cpp", "wbc_scene.cu", os.path.join(HERE, "_build", "libwbc_scene_host.so")))
    c = _SceneConfig()
    c.mu = float(cfg["mu"])
    c.com_pose_weight[:] = [float(v) for v in cfg["com_pose_weight"]]
    c.foot_force_weight[:] = [float(v) for v in cfg["foot_force_weight"]]
    c.foot_pose_weight[:] = [float(v) for v in cfg["foot_pose_weight"]]
    c.hessian_regularizer = float(cfg["hessian_regularizer"])
    c.tau_max[:] = [float(v) for v in tau_max]
    scene_in = np.ascontiguousarray(scene_in, dtype=np.float64)
    n = scene_in.shape[0]
    out = [np.zeros((n, k)) for k in (900, 30, 1380, 46, 46, 24)]
    lib.wbc_scene_host(n, C.byref(c), scene_in.ctypes.data_as(C.c_void_p), *[o.ctypes.data_as(C.c_void_p) for o in out])
    return out
