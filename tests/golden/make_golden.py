"""Generates tests/golden/mpc_golden.npz: seeded input records + oracle solutions (committed fixture).

The reference (/root/reference) cannot be imported or compiled here (Eigen/acados/ROS missing) and ships no golden
vectors, so these are ORACLE outputs (numpy IPM + active-set polish, cross-checked against the C ADMM port to 1e-8
relative), not reference outputs - "parity unpinned" in the sense of the task statement.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import c_oracle, cpu_baseline  # noqa: E402
from oracle import mpc_oracle as mo  # noqa: E402
from dfki_quad_b200 import sequencer  # noqa: E402  (host-side input generator only)


def oracle_schedule(period, duty, offset, phase, steps, dt):
    c = np.zeros((len(period), steps, 4), dtype=np.uint8)
    s = np.zeros((len(period), steps, 4))
    for p in range(len(period)):
        c[p], s[p] = mo.gait_update_sequence(period[p], duty[p], offset[p], dt, float(phase[p]), steps, None)
    return c, s


def polished_reference(N, rec, contact, params):
    U, X, qp, dq = mo.solve_reference_qp(N, rec, contact, params)
    lo = np.where(dq["l"] <= -mo.INF_THRESHOLD, -np.inf, dq["l"])
    hi = np.where(dq["u"] >= mo.INF_THRESHOLD, np.inf, dq["u"])
    u = U.reshape(-1)
    Ax = dq["A"] @ u
    tol = 1e-6 * max(1.0, np.abs(u).max())
    y = np.zeros(len(lo))
    y[(np.abs(Ax - hi) <= tol)] = 1.0
    y[(np.abs(Ax - lo) <= tol)] = -1.0
    x2, _, ok = mo.polish_active_set(dq["H"], dq["g"], dq["A"], lo, hi, u, y)
    assert ok, "oracle polish failed"
    U2 = x2.reshape(N, 12)
    return U2, mo.rollout(qp, U2), qp


def main():
    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    out = {}
    for tag, N, gaits, n, seed in (("h10_trot", 10, ("TROT",), 8, 4242), ("h10_mixed", 10, ("STAND", "PACE", "BOUND", "PRONK"), 8, 4243),
                                   ("h16_mixed", 16, ("TROT", "PACE", "BOUND"), 6, 4244)):
        b = sequencer.random_batch(n, N, seed, gaits=gaits, schedule_fn=oracle_schedule)
        F = np.zeros((n, N, 12))
        X = np.zeros((n, N + 1, 13))
        fc, xc, info, solved = c_oracle.solve_batch(N, b["rec"], b["contact"], params)
        assert solved == n
        for p in range(n):
            F[p], X[p], qp = polished_reference(N, b["rec"][p], b["contact"][p], params)
            assert max(mo.kkt_residuals(qp, F[p])) < 1e-9
            assert np.abs(F[p] - fc[p]).max() <= 1e-8 * max(1.0, np.abs(F[p]).max()), (tag, p, np.abs(F[p] - fc[p]).max())
        out[tag + "_rec"], out[tag + "_contact"], out[tag + "_forces"], out[tag + "_xpred"] = b["rec"], b["contact"], F, X
        out[tag + "_phase"], out[tag + "_gait_ids"] = b["phase"], b["gait_ids"]
    # gait schedule golden: all database gaits at boundary and random phases, 100 steps
    rng = np.random.default_rng(77)
    names = list(mo.GAITS.keys())
    gid = np.repeat(np.arange(len(names)), 6)
    phase = np.concatenate([[0.0, 0.25, 0.5, 0.75, rng.uniform(), rng.uniform()] for _ in names])
    meas = rng.integers(0, 2, (len(gid), 4)).astype(np.uint8)
    C = np.zeros((len(gid), 100, 4), dtype=np.uint8)
    S = np.zeros((len(gid), 100, 4))
    for i, (g, ph) in enumerate(zip(gid, phase)):
        T, d, off = mo.GAITS[names[g]]
        C[i], S[i] = mo.gait_update_sequence(T, d, off, 0.05, float(ph), 100, meas[i])
    out.update(gait_ids=gid, gait_phase=phase, gait_meas=meas, gait_contact=C, gait_swing=S)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mpc_golden.npz"), **out)
    print("wrote mpc_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
