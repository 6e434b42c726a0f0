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


def build():
    src = [os.path.join(HERE, "seq_full_host.cpp"), os.path.join(HERE, "emu_preamble.h"),
           os.path.join(ROOT, "dfki_quad_b200", "csrc", "sequencer_full.cu"), os.path.join(ROOT, "dfki_quad_b200", "csrc", "common.cuh")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in src):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-w", "-I" + CUDA_INC, "-o", SO, src[0]])
    return SO


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
