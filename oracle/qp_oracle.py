"""CPU oracle for the WBC task-space QP (generic dense QP with equalities and two-sided inequality rows).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: in the reference this QP is built by ARC-OPT's scene and solved by a
`wbc::QPSolver` plugin (wbc_arc_opt.cpp:51-74,292-297); ARC-OPT and its solver libraries are un-vendored third-party
code (docker/Dockerfile:129-138, branch `refactoring`, no commit pin) and there are no fixtures.  Correctness is anchored
on uniqueness of the optimum of the strictly convex QP, two independent solvers and a KKT checker.

    min 1/2 x'Hx + g'x   s.t.  Aeq x = beq,   lo <= Ain x <= hi        (|bound| >= 1e20 means absent)
"""
from __future__ import annotations

import numpy as np

INF = 1e20


def solve_ipm(H, g, Aeq, beq, Ain, lo, hi, tol=1e-11, max_iter=60):
    """Mehrotra primal-dual interior point on the full KKT system (dense LU)."""
    n, me = H.shape[0], Aeq.shape[0]
    G = np.vstack([Ain[hi < INF], -Ain[lo > -INF]])
    h = np.concatenate([hi[hi < INF], -lo[lo > -INF]])
    m = G.shape[0]
    x = np.zeros(n)
    nu = np.zeros(me)
    s = np.maximum(h - G @ x, 1.0)
    z = np.ones(m)
    it = 0
    for it in range(max_iter):
        rd = H @ x + g + Aeq.T @ nu + G.T @ z
        re = Aeq @ x - beq
        rp = G @ x + s - h
        mu = s @ z / max(m, 1)
        if max(np.abs(rd).max(), np.abs(re).max(initial=0.0), np.abs(rp).max(initial=0.0), (s * z).max(initial=0.0)) < tol:
            break
        W = z / s
        K = np.block([[H + G.T @ (W[:, None] * G), Aeq.T], [Aeq, -1e-14 * np.eye(me)]])
        import scipy.linalg as sla

        lu = sla.lu_factor(K)

        def step(rc):
            t = (z * rp - rc) / s
            sol = sla.lu_solve(lu, np.concatenate([-(rd + G.T @ t), -re]))
            dx, dnu = sol[:n], sol[n:]
            ds = -rp - G @ dx
            dz = (-rc - z * ds) / s
            return dx, dnu, ds, dz

        def maxstep(ds, dz):
            a = 1.0
            if (ds < 0).any():
                a = min(a, np.min(-s[ds < 0] / ds[ds < 0]))
            if (dz < 0).any():
                a = min(a, np.min(-z[dz < 0] / dz[dz < 0]))
            return a

        dx, dnu, ds, dz = step(s * z)
        a = maxstep(ds, dz)
        sig = (((s + a * ds) @ (z + a * dz) / max(m, 1)) / max(mu, 1e-300)) ** 3
        dx, dnu, ds, dz = step(s * z + ds * dz - sig * mu)
        a = min(1.0, 0.995 * maxstep(ds, dz))
        x, nu, s, z = x + a * dx, nu + a * dnu, s + a * ds, z + a * dz
    return x, it


def solve_active_set(H, g, Aeq, beq, Ain, lo, hi, x0, tol=1e-9, rounds=20):
    """Polish: equality-constrained solves on the active rows with primal-dual active-set repair."""
    n = H.shape[0]
    y = Ain @ x0
    sc = max(1.0, np.abs(y).max())
    al = (lo > -INF) & (y - lo <= 1e-6 * sc)
    au = (hi < INF) & (hi - y <= 1e-6 * sc) & ~al
    x = x0
    for _ in range(rounds):
        rows = np.where(al | au)[0]
        Aa = np.vstack([Aeq, Ain[rows]])
        b = np.concatenate([beq, np.where(al[rows], lo[rows], hi[rows])])
        K = np.block([[H, Aa.T], [Aa, -1e-13 * np.eye(len(b))]])
        sol = np.linalg.lstsq(K, np.concatenate([-g, b]), rcond=None)[0]
        x = sol[:n]
        lam = sol[n + Aeq.shape[0]:]
        y = Ain @ x
        wrong = np.zeros(len(lo), bool)
        wrong[rows] = np.where(al[rows], lam > tol, lam < -tol)
        viol_l = (y < lo - tol) & ~al
        viol_u = (y > hi + tol) & ~au
        if not (wrong.any() or viol_l.any() or viol_u.any()):
            return x, True
        al = (al & ~wrong) | viol_l
        au = (au & ~wrong) | viol_u
    return x, False


def solve(H, g, Aeq, beq, Ain, lo, hi):
    x, _ = solve_ipm(H, g, Aeq, beq, Ain, lo, hi)
    xp, ok = solve_active_set(H, g, Aeq, beq, Ain, lo, hi, x)
    return xp if ok else x


def kkt_residuals(H, g, Aeq, beq, Ain, lo, hi, x):
    """(stationarity after least-squares sign-feasible multipliers, equality, inequality) inf-norms."""
    from scipy.optimize import nnls

    y = Ain @ x
    sc = max(1.0, np.abs(y).max())
    cols, = [[]]
    M = [Aeq.T, -Aeq.T]
    for r in range(len(lo)):
        if hi[r] < INF and hi[r] - y[r] <= 1e-7 * sc:
            M.append(Ain[r][:, None])
        if lo[r] > -INF and y[r] - lo[r] <= 1e-7 * sc:
            M.append(-Ain[r][:, None])
    Mm = np.hstack(M)
    grad = H @ x + g
    coef, res = nnls(Mm, -grad)
    stat = np.abs(grad + Mm @ coef).max()
    eq = np.abs(Aeq @ x - beq).max(initial=0.0)
    ineq = max(np.max(np.maximum(lo - y, 0), initial=0.0), np.max(np.maximum(y - hi, 0), initial=0.0))
    return float(stat), float(eq), float(ineq)
