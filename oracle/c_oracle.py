"""ctypes binding of oracle/_build/libmpc_oracle.so (C restatement of the reference CPU path).
TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Pinning status: see the mpc_ref.c header (QP data pinned on oracle/_ref, solver restated)."""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libmpc_oracle.so")


class RefParams(C.Structure):
    _fields_ = [("N", C.c_int), ("dt", C.c_double), ("alpha", C.c_double), ("w", C.c_double * 12), ("mu", C.c_double),
                ("fmin", C.c_double), ("fmax", C.c_double), ("rho", C.c_double), ("sigma", C.c_double),
                ("relax", C.c_double), ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("kkt_tol", C.c_double),
                ("max_iter", C.c_int), ("check_every", C.c_int), ("polish", C.c_int), ("adaptive_rho", C.c_int)]


class RefInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("n_fact", C.c_int), ("polished", C.c_int), ("kkt", C.c_double * 4)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.ref_solve_batch.restype = C.c_int
        _lib.ref_solve_batch.argtypes = [C.POINTER(RefParams), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int]
        _lib.ref_condense.restype = C.c_int
        _lib.ref_condense.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_max_threads.restype = C.c_int
        _lib.ref_kkt_batch.restype = C.c_int
        _lib.ref_riccati_batch.restype = C.c_int
        _lib.ref_riccati_batch.argtypes = [C.POINTER(RefParams), C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_int]
        _lib.ref_kkt_batch.argtypes = [C.POINTER(RefParams), C.c_int] + [C.c_void_p] * 7 + [C.c_int]
    return _lib


def make_params(N, params, eps=1e-3, polish=True, rho=3e-4, max_iter=4000, kkt_tol=1e-7):
    """OSQP 0.6.3 defaults (sigma 1e-6, alpha 1.6, eps 1e-3, max_iter 4000, check every 25, polish on) except the initial
    penalty: rho = 3e-4 instead of OSQP's 0.1, which on this problem family halves the iterations and saves the
    adaptive-rho refactorisation (measured ~1.8x faster) - the CPU baseline gets the same tuning as the GPU kernel."""
    P = RefParams()
    P.N, P.dt, P.alpha, P.mu, P.fmin, P.fmax = N, params["dt"], params["alpha"], params["mu"], params["fmin"], params["fmax"]
    for i in range(12):
        P.w[i] = float(params["weights"][i])
    P.rho, P.sigma, P.relax, P.eps_abs, P.eps_rel, P.kkt_tol = rho, 1e-6, 1.6, eps, eps, kkt_tol
    P.max_iter, P.check_every, P.polish, P.adaptive_rho = max_iter, 25, 1 if polish else 0, 1
    return P


def solve_batch(N, rec, contact, params, threads=0, **kw):
    lib = load()
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    n = rec.shape[0]
    forces = np.zeros((n, N, 12))
    xpred = np.zeros((n, N + 1, 13))
    info = (RefInfo * n)()
    P = make_params(N, params, **kw)
    solved = lib.ref_solve_batch(C.byref(P), n, rec.ctypes.data, contact.ctypes.data, forces.ctypes.data, xpred.ctypes.data,
                                 C.addressof(info), threads)
    arr = np.frombuffer(info, dtype=np.dtype([("status", "i4"), ("iters", "i4"), ("n_fact", "i4"), ("polished", "i4"),
                                              ("kkt", "f8", (4,))]), count=n).copy()
    return forces, xpred, arr, solved


class WbcInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("kkt", C.c_double * 4)]


def wbc_batch(H, g, A, lo, hi, neq, tol=1e-6, max_iter=40, threads=0):
    """C dense primal-dual IPM (wbc_ref.c) over a batch of WBC QPs in the C-ABI layout (column-major H [n,nq*nq], A [n,(neq+nin)*nq])."""
    lib = load()
    lib.wbc_ref_batch.restype = C.c_int
    lib.wbc_ref_batch.argtypes = [C.c_int] * 4 + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    H, g, A, lo, hi = (np.ascontiguousarray(a, dtype=np.float64) for a in (H, g, A, lo, hi))
    n, nq = g.shape[0], g.shape[1]
    nc = lo.shape[1]
    x = np.zeros((n, nq))
    info = (WbcInfo * n)()
    solved = lib.wbc_ref_batch(n, nq, neq, nc - neq, H.ctypes.data, g.ctypes.data, A.ctypes.data, lo.ctypes.data, hi.ctypes.data, tol,
                               max_iter, x.ctypes.data, C.addressof(info), threads)
    arr = np.frombuffer(info, dtype=np.dtype([("status", "i4"), ("iters", "i4"), ("kkt", "f8", (4,))]), count=n).copy()
    return x, arr, solved


RICCATI_MODES = {"SPEED": (1e-6, 15), "BALANCE": (1e-8, 30), "parity": (1e-11, 60)}  # (residual target, iteration cap)


def riccati_batch(N, rec, contact, params, mode="parity", threads=0):
    """C Riccati primal-dual IPM on the sparse OCP-QP (mpc_ref.c::riccati_ipm_one): the restatement of the reference's default
    PARTIAL_CONDENSING_HPIPM plan.  mode: SPEED-like / BALANCE-like / parity (same tolerance as the GPU kernels)."""
    lib = load()
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    n = rec.shape[0]
    forces = np.zeros((n, N, 12))
    xpred = np.zeros((n, N + 1, 13))
    info = (RefInfo * n)()
    P = make_params(N, params)
    tol, cap = RICCATI_MODES[mode]
    solved = lib.ref_riccati_batch(C.byref(P), n, rec.ctypes.data, contact.ctypes.data, forces.ctypes.data, xpred.ctypes.data,
                                   C.addressof(info), tol, cap, threads)
    arr = np.frombuffer(info, dtype=np.dtype([("status", "i4"), ("iters", "i4"), ("n_fact", "i4"), ("polished", "i4"),
                                              ("kkt", "f8", (4,))]), count=n).copy()
    return forces, xpred, arr, solved


def kkt_batch(N, rec, contact, params, forces, xpred=None, lam_box=None, lam_g=None, threads=0):
    """Independent C KKT checker (mpc_ref.c::kkt_one) over a whole batch -> [n][4] = stationarity, dynamics, inequality,
    complementarity inf-norms in the sparse OCP form."""
    lib = load()
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    n = rec.shape[0]
    P = make_params(N, params)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None for a in (forces, xpred, lam_box, lam_g)]
    kkt = np.zeros((n, 4))
    ptr = [a.ctypes.data if a is not None else None for a in arrs]
    rc = lib.ref_kkt_batch(C.byref(P), n, rec.ctypes.data, contact.ctypes.data, ptr[0], ptr[1], ptr[2], ptr[3], kkt.ctypes.data,
                           threads)
    assert rc == 0
    return kkt


def condense(N, rec, contact, params):
    lib = load()
    n = 12 * N
    H = np.zeros((n, n))
    g = np.zeros(n)
    P = make_params(N, params)
    rec = np.ascontiguousarray(rec, dtype=np.float64)
    contact = np.ascontiguousarray(contact, dtype=np.uint8)
    lib.ref_condense(C.byref(P), rec.ctypes.data, contact.ctypes.data, H.ctypes.data, g.ctypes.data)
    return H, g


def time_sample(N, rec, contact, budget_s, threads=None):
    """Bounded sample: solve chunks of the batch with all host threads until `budget_s` is used up."""
    from . import mpc_oracle as mo

    lib = load()
    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    cores = int(threads or os.cpu_count() or lib.ref_max_threads())
    n = rec.shape[0]
    chunk = max(64, 32 * cores)
    solve_batch(N, rec[: min(n, 2 * cores)], contact[: min(n, 2 * cores)], params, threads=cores)  # warm-up
    t0 = time.perf_counter()
    solved = done = 0
    while done < n and time.perf_counter() - t0 < budget_s:
        hi = min(n, done + chunk)
        _, _, _, s = solve_batch(N, rec[done:hi], contact[done:hi], params, threads=cores)
        solved += s
        done = hi
    dt = time.perf_counter() - t0
    return dict(solved=solved, seconds=dt, cores=cores, kind="port",
                solver="C port: full condensing + OSQP-style ADMM (eps 1e-3, rho0 3e-4) + polish to KKT<=1e-7, OpenMP",
                sample=f"first {done} of {n} problems of the same batch, {cores} threads")
