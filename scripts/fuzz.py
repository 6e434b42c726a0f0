"""Randomised soak of the MPC path: random solver options, weights, friction, force limits, horizons and gaits; every solved
problem is checked by the kernel's own KKT record, a sample by the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dfki_quad_b200 import BatchedMPC, sequencer, gait as gaitmod
from oracle import mpc_oracle as mo
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
rng2 = np.random.default_rng(12345)  # oracle sample picks: separate stream, so that ONLY=<trial> reproduces a trial
cfg = sequencer.GO2_MPC
names = gaitmod.GAIT_NAMES
bad = 0
for trial in range(int(os.environ.get("TRIALS", "24"))):
    N = int(rng.choice([1, 2, 5, 8, 10, 13, 16, 20]))
    n = 256
    gaits = tuple(rng.choice(names, size=rng.integers(1, 4), replace=False))
    w = np.array(cfg["state_weights_move"], float) * np.exp(rng.normal(0, 1.0, 12))
    alpha = float(10 ** rng.uniform(-6, -3))
    mu = float(rng.choice([0.2, 0.4, 0.6, 0.9, 1.2]))
    fmin = float(rng.choice([0.0, 0.0, 5.0]))
    fmax = float(rng.choice([120.0, 250.0, 400.0]))
    solver = str(rng.choice(["CONDENSED_ADMM", "SPARSE_RICCATI", "AUTO"]))
    ws = int(rng.choice([0, 0, 2]))
    if os.environ.get("ONLY") and trial not in {int(t) for t in os.environ["ONLY"].split(",")}:
        continue  # (the random stream above is consumed either way, so ONLY=<trial> reproduces that trial)
    b = sequencer.random_batch(n, N, seed=1000 + trial, gaits=gaits)
    mpc = BatchedMPC(alpha, w, mu, fmin, fmax, solver, warm_start=ws, horizon=N, dt=cfg["dt"], max_batch=n)
    f, x, info = mpc.solve_records(b["rec"], b["contact"])
    if ws:
        f, x, info = mpc.solve_records(b["rec"], b["contact"])
    nfree = 3 * b["contact"].reshape(n, -1).sum(axis=1)
    fits = (nfree <= 156) | (solver != "CONDENSED_ADMM")
    ok = info.success
    kk = info.kkt.max(axis=1)
    params = dict(mo.GO2_PARAMS, weights=w, alpha=alpha, mu=mu, fmin=fmin, fmax=fmax)
    worst = 0.0
    for p in rng2.choice(np.nonzero(ok)[0], size=min(3, int(ok.sum())), replace=False) if ok.any() else []:
        U, X, qp, dq = mo.solve_reference_qp(N, b["rec"][p], b["contact"][p], params)
        worst = max(worst, np.abs(f[p].reshape(N, 12) - U).max() / max(1.0, np.abs(U).max()))
    F = f.reshape(n, N, 4, 3)[ok]
    feas = (np.abs(F[..., 0]) <= mu * F[..., 2] + 1e-6).all() and (np.abs(F[..., 1]) <= mu * F[..., 2] + 1e-6).all() and (F[..., 2] <= fmax + 1e-6).all()
    swing0 = np.all(F[b["contact"][ok] == 0] == 0.0)
    flag = "" if (ok[fits].all() and (kk[ok] <= 1e-6).all() and worst < 1e-4 and feas and swing0) else "   <-- CHECK"
    bad += bool(flag)
    if flag:
        badp = np.nonzero(~ok & fits)[0]
        if os.environ.get("ONLY"):
            q = int(badp[0])
            U, X, qp, dq = mo.solve_reference_qp(N, b["rec"][q], b["contact"][q], params)
            print("    oracle forces:", np.round(U.reshape(N, 12), 4).tolist(), "contact", b["contact"][q].tolist())
            print("    oracle kkt:", mo.kkt_residuals(qp, U.reshape(N, 12)))
            for s2 in ("CONDENSED_ADMM",):
                m2 = BatchedMPC(alpha, w, mu, fmin, fmax, s2, horizon=N, dt=cfg["dt"], max_batch=n)
                f2, x2, i2 = m2.solve_records(b["rec"], b["contact"])
                print("    ", s2, "status", int(i2.return_code[q]), "forces", np.round(f2[q], 4).tolist(), "kkt", i2.kkt[q].tolist())
        print("    failed:", [(int(q), int(info.return_code[q]), int(info.acados_num_iter[q]), info.kkt[q].tolist(), int(nfree[q])) for q in badp[:4]])
    print(f"{trial:2d} N={N:2d} {solver:15s} ws={ws} mu={mu} fmin={fmin} fmax={fmax} alpha={alpha:.1e} gaits={','.join(gaits):40s} ok(fits)={ok[fits].mean():.4f} "
          f"kkt_max={kk[ok].max() if ok.any() else float('nan'):.1e} oracle_rel={worst:.1e} feas={feas} swing0={swing0} iters={info.acados_num_iter.mean():.1f} max={int(info.acados_num_iter[ok].max()) if ok.any() else -1}{flag}", flush=True)
    mpc.close()
print("trials needing a look:", bad)
