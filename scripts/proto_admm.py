"""Development prototype (numpy) of the algorithm the CUDA ADMM kernel implements: reduced (stance-only) condensed
QP assembled in closed form, K^-1 ADMM, null-space active-set polish with per-foot NNLS dual recovery.
Not shipped, not imported by the product; kept as a readable model of dfki_quad_b200/csrc/mpc_admm.cu."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo")
from oracle import mpc_oracle as mo  # noqa: E402
from oracle import sequencer_oracle as so  # noqa: E402


def assemble_reduced(N, rec, contacts, params):
    """Closed-form reduced Hessian/gradient. Returns H (n,n), g (n), idx list of (k, f), plus stuff for rollout."""
    qp = mo.build_ocp_qp(N, rec, contacts, params)  # stage data (prototype reuses oracle for B_k, xref)
    dt = params["dt"]
    Q = qp["Q"]
    A = qp["A"]
    Rt = A[0:3, 6:9] / dt
    m = rec[13]
    fs = [(k, f) for k in range(N) for f in range(4) if contacts[k][f]]
    n = 3 * len(fs)
    QT, Qp, Qw, Qv = np.diag(Q[0:3]), Q[3:6], np.diag(Q[6:9]), Q[9:12]
    G = Rt.T @ QT @ Rt
    H = np.zeros((n, n))
    for a, (j, fa) in enumerate(fs):
        Bw_a = qp["B"][j][6:9, 3 * fa : 3 * fa + 3]
        for b, (l, fb) in enumerate(fs):
            Bw_b = qp["B"][l][6:9, 3 * fb : 3 * fb + 3]
            i0 = max(j, l) + 1
            cnt = N - i0 + 1
            S2 = sum((i - 1 - j) * (i - 1 - l) for i in range(i0, N + 1))
            W = cnt * Qw + S2 * dt * dt * G
            blk = Bw_a.T @ W @ Bw_b + (dt / m) ** 2 * np.diag(cnt * Qv + S2 * dt * dt * Qp)
            H[3 * a : 3 * a + 3, 3 * b : 3 * b + 3] = blk
    H += params["alpha"] * np.eye(n)
    # gradient via free response + adjoint
    X = [qp["x0"]]
    for k in range(N):
        X.append(A @ X[-1])
    lam = np.zeros(13)
    g = np.zeros(n)
    lams = [None] * (N + 2)
    for k in range(N, 0, -1):
        lam = Q * (X[k] - qp["xref"][k]) + (A.T @ lam if k < N else 0)
        lams[k] = lam
    for a, (j, fa) in enumerate(fs):
        g[3 * a : 3 * a + 3] = qp["B"][j][:, 3 * fa : 3 * fa + 3].T @ lams[j + 1]
    return H, g, fs, qp


def foot_rows(mu):
    # rows: x, y, z boxes, x-mu z<=0, x+mu z>=0, y-mu z<=0, y+mu z>=0
    return np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, -mu], [1, 0, mu], [0, 1, -mu], [0, 1, mu]], float)


def admm_reduced(H, g, params, rho=1e-3, sigma=1e-6, alpha=1.6, max_iter=400, check=10, eps=1e-3, adapt=True):
    n = len(g)
    nf = n // 3
    mu, fmin, fmax = params["mu"], params["fmin"], params["fmax"]
    Af = foot_rows(mu)
    lo = np.tile([-fmax, -fmax, fmin, -np.inf, 0, -np.inf, 0], nf)
    hi = np.tile([fmax, fmax, fmax, 0, np.inf, 0, np.inf], nf)
    E = np.tile([3.0, 3.0, 1 + 4 * mu * mu], nf)

    def Aop(x):
        return (x.reshape(nf, 3) @ Af.T).reshape(-1)

    def ATop(y):
        return (y.reshape(nf, 7) @ Af).reshape(-1)

    Kinv = np.linalg.inv(H + np.diag(sigma + rho * E))
    nfact = 1
    x = np.zeros(n)
    z = np.zeros(7 * nf)
    y = np.zeros(7 * nf)
    for it in range(1, max_iter + 1):
        xt = Kinv @ (sigma * x - g + ATop(rho * z - y))
        zt = Aop(xt)
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        zn = np.clip(zr + y / rho, lo, hi)
        y = y + rho * (zr - zn)
        z = zn
        if it % check == 0:
            Ax = Aop(x)
            Hx = H @ x
            Aty = ATop(y)
            rp = np.abs(Ax - z).max()
            rd = np.abs(Hx + g + Aty).max()
            np_ = max(np.abs(Ax).max(), np.abs(z).max(), 1e-30)
            nd = max(np.abs(Hx).max(), np.abs(Aty).max(), np.abs(g).max(), 1e-30)
            if rp <= eps * (1 + np_) and rd <= eps * (1 + nd):
                break
            if adapt:
                new = rho * math.sqrt((rp / np_) / max(rd / nd, 1e-30))
                new = min(max(new, 1e-6), 1e6)
                if new > 5 * rho or new < rho / 5:
                    rho = new
                    Kinv = np.linalg.inv(H + np.diag(sigma + rho * E))
                    nfact += 1
    return x, z, y, it, nfact, rho


def nnls_small(M, b, iters=30):
    from scipy.optimize import nnls
    return nnls(M, b)[0]


def polish(H, g, x, z, y, params, rounds=6, tol=1e-9):
    n = len(g)
    nf = n // 3
    mu, fmin, fmax = params["mu"], params["fmin"], params["fmax"]
    Af = foot_rows(mu)
    lo = np.array([-fmax, -fmax, fmin, -np.inf, 0, -np.inf, 0])
    hi = np.array([fmax, fmax, fmax, 0, np.inf, 0, np.inf])
    z = z.reshape(nf, 7)
    y = y.reshape(nf, 7)
    actl = (z - lo) < -y
    actu = (hi - z) < y
    for rnd in range(rounds):
        # per-foot rref -> c_f, Z_f
        Zcols = []
        c = np.zeros(n)
        for f in range(nf):
            rows = []
            rhs = []
            for r in range(7):
                if actl[f, r]:
                    rows.append(Af[r]); rhs.append(lo[r])
                elif actu[f, r]:
                    rows.append(Af[r]); rhs.append(hi[r])
            basis = []
            M = np.zeros((0, 4))
            if rows:
                M = np.column_stack([np.array(rows), np.array(rhs)])
            # gaussian elimination with column pivoting on 3 cols
            piv = []
            Mr = M.copy()
            rr = 0
            for col in range(3):
                if rr >= len(Mr):
                    break
                p = rr + np.argmax(np.abs(Mr[rr:, col]))
                if abs(Mr[p, col]) < 1e-12:
                    continue
                Mr[[rr, p]] = Mr[[p, rr]]
                Mr[rr] /= Mr[rr, col]
                for q in range(len(Mr)):
                    if q != rr:
                        Mr[q] -= Mr[q, col] * Mr[rr]
                piv.append(col)
                rr += 1
            cf = np.zeros(3)
            for i, col in enumerate(piv):
                cf[col] = Mr[i, 3]
            c[3 * f : 3 * f + 3] = cf
            for col in range(3):
                if col not in piv:
                    v = np.zeros(3)
                    v[col] = 1.0
                    for i, pc in enumerate(piv):
                        v[pc] = -Mr[i, col]
                    zc = np.zeros(n)
                    zc[3 * f : 3 * f + 3] = v
                    Zcols.append(zc)
        if Zcols:
            Z = np.array(Zcols).T
            Hr = Z.T @ H @ Z
            gr = Z.T @ (g + H @ c)
            w = np.linalg.solve(Hr, -gr)
            xn = c + Z @ w
        else:
            xn = c
        r = H @ xn + g
        # per foot: feasibility and NNLS duals over geometrically tight rows
        ok = True
        stat = 0.0
        viol = 0.0
        ynew = np.zeros((nf, 7))
        for f in range(nf):
            xf = xn[3 * f : 3 * f + 3]
            ax = Af @ xf
            sc = max(1.0, np.abs(xf).max())
            viol_l = ax < lo - 1e-9 * sc
            viol_u = ax > hi + 1e-9 * sc
            viol = max(viol, np.max(np.maximum(lo - ax, 0)), np.max(np.maximum(ax - hi, 0)))
            tight_l = np.abs(ax - lo) <= 1e-9 * sc
            tight_u = np.abs(ax - hi) <= 1e-9 * sc
            cols = []
            tags = []
            for rr_ in range(7):
                if tight_u[rr_]:
                    cols.append(Af[rr_]); tags.append((rr_, 1))
                if tight_l[rr_]:
                    cols.append(-Af[rr_]); tags.append((rr_, -1))
            rf = r[3 * f : 3 * f + 3]
            if cols:
                Mm = np.array(cols).T
                coef = nnls_small(Mm, -rf)
                res = rf + Mm @ coef
                for cc, (rr_, sg) in zip(coef, tags):
                    ynew[f, rr_] += sg * cc
            else:
                res = rf
                coef = []
            stat = max(stat, np.abs(res).max())
            # new active set: tight rows with positive multiplier + violated rows
            na_l = np.zeros(7, bool)
            na_u = np.zeros(7, bool)
            for cc, (rr_, sg) in zip(coef, tags):
                if cc > 0:
                    (na_u if sg > 0 else na_l)[rr_] = True
            na_l |= viol_l
            na_u |= viol_u
            if (na_l != actl[f]).any() or (na_u != actu[f]).any():
                pass
            actl[f] = na_l
            actu[f] = na_u
        gs = max(1.0, np.abs(r).max())
        if stat <= 1e-9 * gs and viol <= 1e-9 * 400:
            return xn, ynew, True, rnd + 1, stat, viol
    return xn, ynew, False, rounds, stat, viol


def main():
    seed = 20261021
    rng = np.random.default_rng(seed)
    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    gaits = sys.argv[2].split(",") if len(sys.argv) > 2 else ["TROT"]
    ntest = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    rho0 = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
    eps = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-3
    its, oks, rnds, errs, nfs, fails = [], [], [], [], [], 0
    for t in range(ntest):
        gait = gaits[t % len(gaits)]
        (rec, contact), meta = so.random_problem(rng, N, gait, params)
        H, g, fs, qp = assemble_reduced(N, rec, contact, params)
        if t == 0:
            dq = mo.condense(qp)
            idx = np.concatenate([[12 * k + 3 * f + i for i in range(3)] for (k, f) in fs])
            print("assembly check H", np.abs(dq["H"][np.ix_(idx, idx)] - H).max() / np.abs(H).max(), "g", np.abs(dq["g"][idx] - g).max() / np.abs(g).max())
        x, z, y, it, nfact, rho = admm_reduced(H, g, params, rho=rho0, eps=eps)
        xp, yp, ok, rnd, stat, viol = polish(H, g, x, z, y, params)
        tries = 0
        while not ok and tries < 4:
            tries += 1
            eps2 = eps * 0.1 ** tries
            x, z, y, it2, nf2, rho = admm_reduced(H, g, params, rho=rho, eps=eps2, max_iter=2000)
            it += it2
            nfact += nf2
            xp, yp, ok, rnd, stat, viol = polish(H, g, x, z, y, params)
        U = np.zeros((N, 12))
        for a, (k, f) in enumerate(fs):
            U[k, 3 * f : 3 * f + 3] = xp[3 * a : 3 * a + 3]
        kkt = mo.kkt_residuals(qp, U)
        its.append(it); oks.append(ok); rnds.append(rnd); nfs.append(nfact)
        if not ok or max(kkt) > 1e-6:
            fails += 1
            print("FAIL", t, gait, it, ok, rnd, stat, viol, kkt)
    print(f"N={N} gaits={gaits} n={ntest} rho0={rho0} eps={eps}: iters mean {np.mean(its):.1f} med {np.median(its)} max {np.max(its)}; "
          f"polish ok {np.mean(oks):.3f} rounds mean {np.mean(rnds):.2f}; nfact mean {np.mean(nfs):.2f}; fails {fails}")


if __name__ == "__main__":
    main()
