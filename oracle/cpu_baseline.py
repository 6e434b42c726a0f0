"""Timing harness for the restated reference CPU path (bench.py `cpu_baseline` leg and `--impl reference`).
TEST / MEASUREMENT INFRASTRUCTURE ONLY - never on the product path.  Pinning status: see mpc_oracle.py.

Uses the C port (oracle/_build/libmpc_oracle.so: OSQP-style ADMM + polish on the condensed QP, OpenMP over the
batch) when it has been built, else the numpy port (one core)."""
from __future__ import annotations

import os
import time

import numpy as np

from . import mpc_oracle as mo


def make_batch_cpu(n, N, seed, gaits=("TROT",)):
    """Same synthetic distribution as dfki_quad_b200.sequencer.random_batch, with the gait schedule from the oracle's
    own restatement (no GPU involved)."""
    from dfki_quad_b200 import sequencer

    def sched(period, duty, offset, phase, steps, dt):
        c = np.zeros((len(period), steps, 4), dtype=np.uint8)
        s = np.zeros((len(period), steps, 4))
        for p in range(len(period)):
            c[p], s[p] = mo.gait_update_sequence(period[p], duty[p], offset[p], dt, float(phase[p]), steps, None)
        return c, s

    return sequencer.random_batch(n, N, seed, gaits=gaits, schedule_fn=sched)


def _params():
    p = dict(mo.GO2_PARAMS)
    p["weights"] = p["w_move"]
    return p


def time_sample(N, rec, contact, budget_s, threads=None):
    try:
        from . import c_oracle

        if c_oracle.available():
            return c_oracle.time_sample(N, rec, contact, budget_s, threads)
    except ImportError:
        pass
    params = _params()
    t0 = time.perf_counter()
    solved = 0
    for p in range(rec.shape[0]):
        qp = mo.build_ocp_qp(N, rec[p], contact[p], params)
        dq = mo.condense(qp)
        x, y, info = mo.solve_admm_dense(dq["H"], dq["g"], dq["A"], dq["l"], dq["u"], return_info=True)
        solved += 1 if info["polished"] else 0
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(solved=solved, seconds=dt, cores=1, kind="port", solver="numpy OSQP-style ADMM + polish (condensed)",
                sample=f"first {p + 1} problems of the batch, numpy port, 1 thread")
