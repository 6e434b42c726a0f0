"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libb200qp.so via
ctypes); the oracle (oracle/) is the checker only."""
import math

import numpy as np
import pytest

from oracle import mpc_oracle as mo

pytestmark = pytest.mark.gpu


def _mpc(N, max_batch, solver="CONDENSED_ADMM", weights="state_weights_move", **opts):
    from dfki_quad_b200 import BatchedMPC, sequencer

    cfg = sequencer.GO2_MPC
    return BatchedMPC(cfg["alpha"], cfg[weights], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=N, dt=cfg["dt"],
                      max_batch=max_batch, **opts)


# ---------------------------------------------------------------------------------------------------- K4 gait masks
def test_gait_schedule_bit_exact_random():
    from dfki_quad_b200 import gait

    rng = np.random.default_rng(20261019)
    n, steps, dt = 1500, 21, 0.05
    ids = rng.integers(0, len(gait.GAIT_NAMES), n)
    per, duty, off = gait.gait_arrays(ids)
    phase = rng.uniform(0, 1, n)
    # exact-boundary phases: multiples of dt/T and of the offsets
    phase[:200] = rng.integers(0, 20, 200) * 0.05
    phase[200:300] = 0.0
    meas = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    contact, swing, lbu, ubu = gait.contact_schedule(per, duty, off, phase, steps, dt, measured_contact=meas, fmin=0.0,
                                                     fmax=400.0, want_bounds=True)
    for p in range(n):
        c_ref, s_ref = mo.gait_update_sequence(per[p], duty[p], off[p], dt, float(phase[p]), steps, meas[p])
        assert np.array_equal(contact[p], c_ref), (p, gait.GAIT_NAMES[ids[p]], phase[p])
        assert np.array_equal(swing[p].view(np.uint64), s_ref.view(np.uint64)), (p, "swing time not bit exact")
        for k in range(0, steps, 5):
            lb, ub = mo.get_u_bounds(c_ref[k], 0.0, 400.0)
            assert np.array_equal(lbu[p, k], lb) and np.array_equal(ubu[p, k], ub)


def test_gait_schedule_known_answer_trot_phase0():
    """TROT at phase 0, dt = 0.05: rows 0-4 [1,0,0,1], rows 5-9 [0,1,1,0], row 10 [1,0,0,1]
    (the 0.5 < 0.5 comparison at i = 5 must come out false)."""
    from dfki_quad_b200 import gait

    per, duty, off = gait.gait_arrays(np.array([gait.GAIT_NAMES.index("TROT")]))
    contact, _ = gait.contact_schedule(per, duty, off, np.array([0.0]), 11, 0.05)
    exp = np.array([[1, 0, 0, 1]] * 5 + [[0, 1, 1, 0]] * 5 + [[1, 0, 0, 1]], dtype=np.uint8)
    assert np.array_equal(contact[0], exp)


def test_gait_schedule_no_measured_contacts_and_empty():
    from dfki_quad_b200 import gait

    per, duty, off = gait.gait_arrays(np.arange(len(gait.GAIT_NAMES)))
    phase = np.linspace(0, 0.95, len(per))
    contact, swing = gait.contact_schedule(per, duty, off, phase, 100, 0.05)
    for p in range(len(per)):
        c_ref, s_ref = mo.gait_update_sequence(per[p], duty[p], off[p], 0.05, float(phase[p]), 100, None)
        assert np.array_equal(contact[p], c_ref) and np.array_equal(swing[p], s_ref)
    c0, s0 = gait.contact_schedule(np.zeros(0), np.zeros((0, 4)), np.zeros((0, 4)), np.zeros(0), 10, 0.05)
    assert c0.shape == (0, 10, 4)


# ---------------------------------------------------------------------------------------------------- K1 assembly
@pytest.mark.parametrize("N,gaits", [(10, ("TROT",)), (10, ("STAND", "PACE", "STATIC_WALK")), (16, ("TROT", "BOUND"))])
def test_condensed_hessian_gradient_parity(N, gaits, go2_params):
    from dfki_quad_b200 import sequencer

    n = 6
    batch = sequencer.random_batch(n, N, seed=11 + N, gaits=gaits)
    mpc = _mpc(N, n)
    _, _, info = mpc.solve_records(batch["rec"], batch["contact"])
    for p in range(n):
        if info.n_free[p] > 156:
            continue
        H, g = mpc.get_condensed(p)
        qp = mo.build_ocp_qp(N, batch["rec"][p], batch["contact"][p], go2_params)
        dq = mo.condense(qp)
        idx = np.array([12 * k + 3 * f + c for k in range(N) for f in range(4) if batch["contact"][p, k, f] for c in range(3)],
                       dtype=int)
        assert H.shape[0] == len(idx) == info.n_free[p]
        Href, gref = dq["H"][np.ix_(idx, idx)], dq["g"][idx]
        assert np.abs(H - Href).max() <= 1e-12 * np.abs(Href).max()
        assert np.abs(g - gref).max() <= 1e-12 * max(np.abs(gref).max(), 1.0)


# ---------------------------------------------------------------------------------------------------- K2a ADMM solve
def _check_against_oracle(N, batch, forces, xpred, info, params, idx, grf_tol=1e-5, kkt_tol=1e-6):
    for p in idx:
        assert info.return_code[p] == 0, (p, info.return_code[p])
        U, X, qp, dq = mo.solve_reference_qp(N, batch["rec"][p], batch["contact"][p], params)
        Ug = forces[p].reshape(N, 12)
        # oracle IPM is an interior solution: polish it on the active set it identified, for a 1e-10 reference
        x2, y2, ok = mo.polish_active_set(dq["H"], dq["g"], dq["A"], np.where(dq["l"] <= -mo.INF_THRESHOLD, -np.inf, dq["l"]),
                                          np.where(dq["u"] >= mo.INF_THRESHOLD, np.inf, dq["u"]), U.reshape(-1),
                                          _ipm_duals(dq, U.reshape(-1)))
        Uref = x2.reshape(N, 12) if ok else U
        scale = max(1.0, np.abs(Uref).max())
        assert np.abs(Ug - Uref).max() / scale <= grf_tol, (p, np.abs(Ug - Uref).max() / scale)
        kkt = mo.kkt_residuals(qp, Ug)  # FP64 KKT 4-tuple by the independent checker
        assert max(kkt) <= kkt_tol, (p, kkt)
        Xr = mo.rollout(qp, Ug)
        assert np.abs(xpred[p] - Xr).max() <= 1e-9 * max(1.0, np.abs(Xr).max())
        # swing legs are exactly zero (constraint indexing bit-exact)
        mask = np.repeat(batch["contact"][p].reshape(N, 4), 3, axis=1) == 0
        assert np.all(Ug[mask] == 0.0)


def _ipm_duals(dq, u):
    """Sign pattern of the duals at the IPM solution from the gradient (for the oracle-side polish)."""
    r = dq["H"] @ u + dq["g"]
    A, l, uu = dq["A"], dq["l"], dq["u"]
    Ax = A @ u
    y = np.zeros(len(l))
    tol = 1e-6 * max(1.0, np.abs(u).max())
    y[(np.abs(Ax - uu) <= tol) & (uu < mo.INF_THRESHOLD)] = 1.0
    y[(np.abs(Ax - l) <= tol) & (l > -mo.INF_THRESHOLD)] = -1.0
    return y


@pytest.mark.parametrize("gaits,N", [(("TROT",), 10), (("STAND", "PACE", "BOUND", "WALKING_TROT", "PRONK"), 10),
                                      (("TROT", "PACE"), 16)])
def test_admm_parity_small(gaits, N, go2_params):
    from dfki_quad_b200 import sequencer

    n = 24
    batch = sequencer.random_batch(n, N, seed=100 + N + len(gaits), gaits=gaits)
    mpc = _mpc(N, n)
    forces, xpred, info = mpc.solve_records(batch["rec"], batch["contact"])
    ok = [p for p in range(n) if info.n_free[p] <= 156]
    assert len(ok) > 0
    _check_against_oracle(N, batch, forces, xpred, info, go2_params, ok)
    assert np.all(info.kkt[ok].max(axis=1) <= 1e-6)


def test_config2_full_batch_converges(go2_params):
    """BASELINE.json configs[1]: 4096 randomised trot problems, N = 10, condensed ADMM."""
    from dfki_quad_b200 import sequencer

    N, n = 10, 4096
    batch = sequencer.random_batch(n, N, seed=20261019 + 2, gaits=("TROT",))
    mpc = _mpc(N, n)
    forces, xpred, info = mpc.solve_records(batch["rec"], batch["contact"])
    assert info.success.mean() >= 0.999, (info.success.mean(), np.bincount(info.return_code[info.return_code >= 0]))
    assert np.all(info.kkt[info.success].max(axis=1) <= 1e-6)
    rng = np.random.default_rng(3)
    idx = rng.choice(np.nonzero(info.success)[0], 24, replace=False)
    _check_against_oracle(N, batch, forces, xpred, info, go2_params, idx)
    # size-independent properties on the whole batch: friction pyramid, boxes, swing legs zero
    F = forces.reshape(n, N, 4, 3)[info.success]
    C = batch["contact"][info.success]
    assert np.all(F[C == 0] == 0.0)
    mu, fmax = 0.6, 400.0
    assert np.all(np.abs(F[..., 0]) <= mu * F[..., 2] + 1e-7) and np.all(np.abs(F[..., 1]) <= mu * F[..., 2] + 1e-7)
    assert np.all(F[..., 2] >= -1e-7) and np.all(F[..., 2] <= fmax + 1e-7)


def test_known_answer_standing(go2_params):
    """level stand, 4 symmetric stance feet, zero velocity, reference = current state =>
    sum f_z = m g (the finite horizon lets the last stages sag, so the early stages are checked), f_xy ~ 0, equal
    split over the symmetric feet at every stage."""
    from dfki_quad_b200 import pack_records

    N = 10
    quat = np.array([[0.0, 0.0, 0.0, 1.0]])
    pos = np.array([[0.0, 0.0, 0.3]])
    z3 = np.zeros((1, 3))
    feet = np.array([[[0.18, 0.13, 0.0], [0.18, -0.13, 0.0], [-0.18, 0.13, 0.0], [-0.18, -0.13, 0.0]]])
    rec = pack_records(N, quat, pos, z3, z3, np.array([mo.GO2_MASS]), mo.GO2_INERTIA, z3,
                       np.repeat(quat[:, None], N + 1, 1), np.repeat(pos[:, None], N + 1, 1), np.zeros((1, N + 1, 3)),
                       np.zeros((1, N + 1, 3)), np.repeat(feet[:, None], N, 1))
    contact = np.ones((1, N, 4), dtype=np.uint8)
    mpc = _mpc(N, 1, weights="state_weights_stand")
    forces, xpred, info = mpc.solve_records(rec, contact)
    assert info.success[0] and info.n_free[0] == 120
    F = forces.reshape(N, 4, 3)
    mg = mo.GO2_MASS * mo.GRAVITY_CONSTANT
    assert np.allclose(F[:6, :, 2].sum(axis=1), mg, rtol=1e-3)
    assert np.abs(F[:, :, :2]).max() < 1e-6 * mg
    assert np.allclose(F[:, :, 2], F[:, :1, 2], rtol=1e-9)
    U, _, qp, _ = mo.solve_reference_qp(N, rec[0], contact[0], dict(go2_params, weights=go2_params["w_stand"]))
    assert np.abs(forces[0].reshape(N, 12) - U).max() <= 1e-5 * np.abs(U).max()


def test_all_swing_is_ballistic(go2_params):
    """Appendix A (ii): no stance foot at any stage => U = 0 and X is the free roll-out."""
    from dfki_quad_b200 import sequencer

    N, n = 10, 3
    batch = sequencer.random_batch(n, N, seed=5)
    contact = np.zeros((n, N, 4), dtype=np.uint8)
    mpc = _mpc(N, n)
    forces, xpred, info = mpc.solve_records(batch["rec"], contact)
    assert info.success.all() and np.all(forces == 0.0) and np.all(info.n_free == 0)
    for p in range(n):
        qp = mo.build_ocp_qp(N, batch["rec"][p], contact[p], go2_params)
        assert np.abs(xpred[p] - mo.rollout(qp, np.zeros((N, 12)))).max() <= 1e-12


def test_setters_and_failure_semantics(go2_params):
    from dfki_quad_b200 import sequencer

    N, n = 10, 8
    batch = sequencer.random_batch(n, N, seed=9)
    mpc = _mpc(N, n)
    f0, _, _ = mpc.solve_records(batch["rec"], batch["contact"])
    mpc.SetMu(0.3)
    mpc.SetFmax(150.0)
    f1, _, info = mpc.solve_records(batch["rec"], batch["contact"])
    assert info.success.all()
    F = f1.reshape(n, N, 4, 3)
    assert np.all(np.abs(F[..., 0]) <= 0.3 * F[..., 2] + 1e-7) and np.all(F[..., 2] <= 150.0 + 1e-7)
    p2 = dict(go2_params, mu=0.3, fmax=150.0)
    _check_against_oracle(N, batch, f1, mpc.solve_records(batch["rec"], batch["contact"])[1], info, p2, range(3))
    # NaN input => that problem fails and leaves its output slot untouched; the others are unaffected
    rec = batch["rec"].copy()
    rec[2, 5] = np.nan
    forces = np.full((n, N, 12), 123.0)
    xpred = np.full((n, N + 1, 13), 321.0)
    _, _, info = mpc.solve_records(rec, batch["contact"], forces, xpred)
    assert not info.success[2] and info.return_code[2] in (1, -7777)
    assert np.all(forces[2] == 123.0) and np.all(xpred[2] == 321.0)
    assert info.success[[0, 1, 3, 4, 5, 6, 7]].all()
    assert np.array_equal(forces[0], f1[0])


@pytest.mark.parametrize("solver", ["CONDENSED_ADMM", "SPARSE_RICCATI"])
def test_page_locked_outputs_are_written_in_place_with_failure_semantics(solver):
    """Pinned caller buffers are written by the kernels directly (no staging copy): results identical to the staged
    path, failed problems leave their slots untouched."""
    import torch
    from dfki_quad_b200 import sequencer

    N, n = 10, 512
    batch = sequencer.random_batch(n, N, seed=31, gaits=("TROT", "PACE", "STAND"))
    mpc = _mpc(N, n, solver=solver)
    f_ref, x_ref, info_ref = mpc.solve_records(batch["rec"], batch["contact"])  # pageable numpy -> staged path
    rec = torch.from_numpy(batch["rec"].copy()).pin_memory()
    ct = torch.from_numpy(batch["contact"]).pin_memory()
    rec[7, 5] = float("nan")
    rec[300, 30] = float("nan")
    f_p = torch.full((n, N, 12), 123.0, dtype=torch.float64).pin_memory()
    x_p = torch.full((n, N + 1, 13), 321.0, dtype=torch.float64).pin_memory()
    _, _, info = mpc.solve_records(rec.numpy(), ct.numpy(), f_p.numpy(), x_p.numpy())
    bad = np.zeros(n, bool)
    bad[[7, 300]] = True
    assert not info.success[bad].any() and info.success[~bad].all()
    assert (f_p.numpy()[bad] == 123.0).all() and (x_p.numpy()[bad] == 321.0).all()
    assert np.array_equal(f_p.numpy()[~bad], f_ref[~bad]) and np.array_equal(x_p.numpy()[~bad], x_ref[~bad])


def test_duals_reported_in_reference_row_order(go2_params):
    from dfki_quad_b200 import sequencer

    N, n = 10, 6
    batch = sequencer.random_batch(n, N, seed=21, gaits=("TROT", "STAND"))
    mpc = _mpc(N, n)
    forces, _, info = mpc.solve_records(batch["rec"], batch["contact"])
    lam_box, lam_g = mpc.get_duals(n)
    for p in range(n):
        qp = mo.build_ocp_qp(N, batch["rec"][p], batch["contact"][p], go2_params)
        kkt = mo.kkt_residuals(qp, forces[p].reshape(N, 12), lam_box[p], lam_g[p])
        assert max(kkt) <= 1e-6, (p, kkt)


# ---------------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("tag,N", [("h10_trot", 10), ("h10_mixed", 10), ("h16_mixed", 16)])
def test_golden_fixtures(tag, N):
    """Committed oracle fixtures (tests/golden/make_golden.py): GRFs within 1e-5 relative, predicted states 1e-9."""
    import os

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_golden.npz"))
    rec, contact, F, X = G[tag + "_rec"], G[tag + "_contact"], G[tag + "_forces"], G[tag + "_xpred"]
    mpc = _mpc(N, rec.shape[0])
    forces, xpred, info = mpc.solve_records(rec, contact)
    ok = info.n_free <= 156
    assert info.success[ok].all() and ok.sum() >= rec.shape[0] - 2
    assert np.abs(forces[ok] - F[ok]).max() <= 1e-5 * max(1.0, np.abs(F).max())
    assert np.abs(xpred[ok] - X[ok]).max() <= 1e-8
    assert (info.return_code[~ok] == 4).all()  # too large for the condensed kernel -> ACADOS_QP_FAILURE, outputs untouched
    assert not forces[~ok].any()


def test_gait_golden_bit_exact():
    import os

    from dfki_quad_b200 import gait

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_golden.npz"))
    per, duty, off = gait.gait_arrays(G["gait_ids"])
    c, s = gait.contact_schedule(per, duty, off, G["gait_phase"], 100, 0.05, measured_contact=G["gait_meas"])
    assert np.array_equal(c, G["gait_contact"]) and np.array_equal(s.view(np.uint64), G["gait_swing"].view(np.uint64))


# ---------------------------------------------------------------------------------------------------- K2b Riccati IPM
def _check_riccati(N, batch, mpc, forces, xpred, info, params, idx, grf_tol=1e-5):
    lam_box, lam_g = mpc.get_duals(len(batch["rec"]))
    for p in idx:
        assert info.return_code[p] == 0, (p, info.return_code[p], info.kkt[p])
        U, X, qp, dq = mo.solve_reference_qp(N, batch["rec"][p], batch["contact"][p], params)
        x2, _, ok = mo.polish_active_set(dq["H"], dq["g"], dq["A"], np.where(dq["l"] <= -mo.INF_THRESHOLD, -np.inf, dq["l"]),
                                         np.where(dq["u"] >= mo.INF_THRESHOLD, np.inf, dq["u"]), U.reshape(-1),
                                         _ipm_duals(dq, U.reshape(-1)))
        Uref = x2.reshape(N, 12) if ok else U
        Ug = forces[p].reshape(N, 12)
        scale = max(1.0, np.abs(Uref).max())
        assert np.abs(Ug - Uref).max() / scale <= grf_tol, (p, np.abs(Ug - Uref).max() / scale)
        kkt = mo.kkt_residuals(qp, Ug, lam_box[p], lam_g[p])  # interior point: check with the solver's own duals
        assert max(kkt) <= 1e-6, (p, kkt)
        assert np.abs(xpred[p] - mo.rollout(qp, Ug)).max() <= 1e-9 * max(1.0, np.abs(xpred[p]).max())
        mask = np.repeat(batch["contact"][p].reshape(N, 4), 3, axis=1) == 0
        assert np.all(Ug[mask] == 0.0)


@pytest.mark.parametrize("gaits,N", [(("TROT", "PACE", "BOUND"), 16), (("STAND", "WALKING_TROT", "STATIC_WALK"), 10),
                                      (("TROT", "STAND"), 20)])
def test_riccati_parity_small(gaits, N, go2_params):
    from dfki_quad_b200 import sequencer

    n = 16
    batch = sequencer.random_batch(n, N, seed=300 + N, gaits=gaits)
    mpc = _mpc(N, n, solver="SPARSE_RICCATI")
    forces, xpred, info = mpc.solve_records(batch["rec"], batch["contact"])
    assert info.success.all(), (info.return_code, info.kkt)
    _check_riccati(N, batch, mpc, forces, xpred, info, go2_params, range(n))


def test_riccati_golden_h16():
    import os

    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_golden.npz"))
    for tag, N in (("h16_mixed", 16), ("h10_mixed", 10)):
        rec, contact, F, X = G[tag + "_rec"], G[tag + "_contact"], G[tag + "_forces"], G[tag + "_xpred"]
        mpc = _mpc(N, rec.shape[0], solver="PARTIAL_CONDENSING_HPIPM")
        forces, xpred, info = mpc.solve_records(rec, contact)
        assert info.success.all()
        assert np.abs(forces - F).max() <= 1e-5 * max(1.0, np.abs(F).max())
        assert np.abs(xpred - X).max() <= 1e-5  # interior-point solution vs polished vertex


def test_riccati_matches_admm(go2_params):
    """The two kernels solve the same QP: their GRFs agree to 1e-5 relative on a mixed-gait batch."""
    from dfki_quad_b200 import sequencer

    N, n = 10, 256
    batch = sequencer.random_batch(n, N, seed=77, gaits=("TROT", "PACE", "BOUND", "STAND"))
    fa, xa, ia = _mpc(N, n).solve_records(batch["rec"], batch["contact"])
    fr, xr, ir = _mpc(N, n, solver="SPARSE_RICCATI").solve_records(batch["rec"], batch["contact"])
    assert ia.success.all() and ir.success.all()
    scale = np.maximum(1.0, np.abs(fa).reshape(n, -1).max(axis=1))
    assert (np.abs(fa - fr).reshape(n, -1).max(axis=1) / scale).max() <= 1e-5
    assert np.abs(xa - xr).max() <= 1e-5


def test_auto_solver_dispatches_by_reduced_size(go2_params):
    """Solver AUTO: problems whose reduced size fits go to the condensed kernel, the rest (N = 16 all-stance, n = 192) to the
    Riccati kernel; every problem is solved and agrees with the single-kernel handles."""
    from dfki_quad_b200 import sequencer

    N, n = 16, 512
    batch = sequencer.random_batch(n, N, seed=78, gaits=("TROT", "STAND", "STATIC_WALK", "BOUND"))
    nfree = 3 * batch["contact"].reshape(n, -1).sum(axis=1)
    assert (nfree > 156).any() and (nfree <= 128).any()
    auto = _mpc(N, n, solver="AUTO")
    fa, xa, ia = auto.solve_records(batch["rec"], batch["contact"])
    lam_box, lam_g = auto.get_duals(n)
    assert ia.success.all(), np.unique(ia.return_code, return_counts=True)
    assert np.all(ia.kkt.max(axis=1) <= 1e-6)
    fr, xr, ir = _mpc(N, n, solver="SPARSE_RICCATI").solve_records(batch["rec"], batch["contact"])
    fc, xc, ic = _mpc(N, n).solve_records(batch["rec"], batch["contact"])
    assert ir.success.all()
    assert (ic.return_code[nfree > 156] == 4).all() and ic.success[nfree <= 156].all()  # condensed alone cannot take them
    scale = np.maximum(1.0, np.abs(fr).reshape(n, -1).max(axis=1))
    assert (np.abs(fa - fr).reshape(n, -1).max(axis=1) / scale).max() <= 1e-5
    small = nfree <= 156
    assert np.array_equal(fa[small], fc[small])  # same kernel, same bits
    # polish rounds are only reported by the condensed kernel: the dispatch is visible in the info records
    assert (ia.polish_rounds[~small] == 0).all() and (ia.polish_rounds[small & (nfree > 0)] >= 1).all()
    params = dict(go2_params, weights=go2_params["w_move"])
    for p in list(np.nonzero(~small)[0][:2]) + list(np.nonzero(small)[0][:2]):
        U, X, qp, _ = mo.solve_reference_qp(N, batch["rec"][p], batch["contact"][p], params)
        assert np.abs(fa[p].reshape(N, 12) - U).max() / max(1.0, np.abs(U).max()) <= 1e-5
        assert max(mo.kkt_residuals(qp, fa[p].reshape(N, 12), lam_box[p], lam_g[p])) <= 1e-6  # with the solver's own duals


def test_config3_full_batch_riccati(go2_params):
    """BASELINE.json configs[2]: 16384 mixed trot/pace/bound problems, N = 16, sparse Riccati kernel."""
    from dfki_quad_b200 import sequencer

    N, n = 16, 16384
    batch = sequencer.random_batch(n, N, seed=20261019 + 3, gaits=("TROT", "PACE", "BOUND"))
    mpc = _mpc(N, n, solver="SPARSE_RICCATI")
    forces, xpred, info = mpc.solve_records(batch["rec"], batch["contact"])
    assert info.success.mean() >= 0.999, (info.success.mean(), np.unique(info.return_code, return_counts=True))
    assert np.all(info.kkt[info.success].max(axis=1) <= 1e-6)
    rng = np.random.default_rng(4)
    idx = rng.choice(np.nonzero(info.success)[0], 12, replace=False)
    _check_riccati(N, batch, mpc, forces, xpred, info, go2_params, idx)
    F = forces.reshape(n, N, 4, 3)[info.success]
    assert np.all(F[batch["contact"][info.success] == 0] == 0.0)
    assert np.all(np.abs(F[..., 0]) <= 0.6 * F[..., 2] + 1e-6) and np.all(F[..., 2] >= -1e-6)


# ---------------------------------------------------------------------------------------------------- edge cases
def test_edge_cases_sizes_and_arguments(go2_params):
    from dfki_quad_b200 import BatchedMPC, _lib, sequencer

    cfg = sequencer.GO2_MPC
    # empty batch is a no-op
    mpc = _mpc(10, 4)
    f, x, info = mpc.solve_records(np.zeros((0, 289)), np.zeros((0, 10, 4), dtype=np.uint8))
    assert f.shape == (0, 10, 12) and len(info.success) == 0
    # more problems than max_batch is an argument error, not a crash
    b = sequencer.random_batch(6, 10, seed=1)
    with pytest.raises(_lib.B200QPError):
        mpc.solve_records(b["rec"], b["contact"])
    # horizon limits: 1 and B200QP_MAX_HORIZON = 20 (the reference's compile-time value), both kernels
    for N in (1, 20):
        bb = sequencer.random_batch(5, N, seed=40 + N, gaits=("TROT", "STAND"))
        for solver in ("CONDENSED_ADMM", "SPARSE_RICCATI"):
            m = _mpc(N, 5, solver=solver)
            ff, xx, ii = m.solve_records(bb["rec"], bb["contact"])
            for p in range(5):
                if solver == "CONDENSED_ADMM" and ii.n_free[p] > 156:
                    assert ii.return_code[p] == 4
                    continue
                assert ii.return_code[p] == 0, (N, solver, p, ii.return_code[p])
                U, X, qp, _ = mo.solve_reference_qp(N, bb["rec"][p], bb["contact"][p], go2_params)
                assert np.abs(ff[p].reshape(N, 12) - U).max() <= 1e-5 * max(1.0, np.abs(U).max())
    with pytest.raises(_lib.B200QPError):
        BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], horizon=21, max_batch=4)


def test_mixed_sizes_in_one_call_and_determinism(go2_params):
    """One call with problems of all three size bins (and one too large for the condensed kernel); results do not
    depend on the batch composition or on the run."""
    from dfki_quad_b200 import sequencer

    N, n = 12, 40
    b = sequencer.random_batch(n, N, seed=8, gaits=("TROT", "STAND", "STATIC_WALK", "BOUND"))
    mpc = _mpc(N, n)
    f1, x1, i1 = mpc.solve_records(b["rec"], b["contact"])
    f2, x2, i2 = mpc.solve_records(b["rec"], b["contact"])
    assert np.array_equal(f1, f2) and np.array_equal(i1.return_code, i2.return_code)
    assert set(np.unique(np.digitize(i1.n_free, [65, 129, 157]))) >= {0, 1, 2}
    perm = np.random.default_rng(0).permutation(n)
    f3, _, i3 = mpc.solve_records(b["rec"][perm], b["contact"][perm])
    assert np.array_equal(f3, f1[perm])
    ok = i1.return_code == 0
    assert ok.sum() >= n - 12 and np.all(i1.return_code[~ok] == 4)
    _check_against_oracle(N, b, f1, x1, i1, go2_params, np.nonzero(ok)[0][:8])


# ------------------------------------------------------------------ on-device gait sequencer
def _seq_case(rng, N, gait, early=False):
    from oracle import sequencer_oracle as so
    (rec, contact), raw = so.random_problem(rng, N, gait)
    return rec, contact, raw


@pytest.mark.parametrize("N", [10, 20])
def test_sequencer_kernel_matches_scalar_oracle(N):
    """Sequencer kernel vs the scalar restatement of SimpleGaitSequencer / MPCTrajectoryPlanner / Raibert planner:
    contact masks bit-exact, records to 1e-12 (libm vs CUDA trig differ in the last ulp)."""
    from oracle import mpc_oracle as mo
    from dfki_quad_b200 import sequencer as sq
    rng = np.random.Generator(np.random.PCG64(77 + N))
    gaits = list(mo.GAITS)
    recs, cts, seqs = [], [], []
    for i in range(96):
        g = gaits[i % len(gaits)]
        rec, contact, raw = _seq_case(rng, N, g)
        T, duty, off = mo.GAITS[g]
        st = raw["state"]
        min_h = raw["feet"][:, 2].min()
        tgt = {k: np.array([v]) for k, v in raw["tgt"].items()}
        seqs.append(sq.pack_sequencer_inputs(st["quat"][None], st["pos"][None], st["omega"][None], st["vel"][None],
                                             raw["feet"][None], st["vel"][None], np.full((1, 4), min_h), tgt,
                                             np.array([raw["phase"]]), np.array([T]), np.full((1, 4), duty),
                                             np.asarray(off, float)[None])[0])
        recs.append(rec)
        cts.append(contact)
    mpc = _mpc(N, 128)
    mpc.set_model(**sq.go2_seq_model())
    rec_g, ct_g = mpc.sequence_records(np.array(seqs))
    assert np.array_equal(ct_g, np.array(cts).astype(np.uint8))
    np.testing.assert_allclose(rec_g, np.array(recs), rtol=0, atol=1e-12)


def test_sequencer_early_contact_rule():
    """Measured contact late in a planned swing phase latches the leg into stance (gait.cpp:78-90)."""
    from oracle import mpc_oracle as mo
    from oracle import sequencer_oracle as so
    from dfki_quad_b200 import sequencer as sq
    N = 10
    rng = np.random.Generator(np.random.PCG64(5))
    T, duty, off = mo.GAITS["TROT"]
    seqs, cts, meas = [], [], []
    for i in range(64):
        (_, _), raw = so.random_problem(rng, N, "TROT")
        st = raw["state"]
        m = rng.integers(0, 2, 4).astype(np.uint8)
        rec, contact = so.make_problem(N, "TROT", raw["phase"], st, raw["tgt"], raw["feet"], mo.GO2_PARAMS, measured_contacts=list(m))
        min_h = raw["feet"][:, 2].min()
        tgt = {k: np.array([v]) for k, v in raw["tgt"].items()}
        seqs.append(sq.pack_sequencer_inputs(st["quat"][None], st["pos"][None], st["omega"][None], st["vel"][None],
                                             raw["feet"][None], st["vel"][None], np.full((1, 4), min_h), tgt,
                                             np.array([raw["phase"]]), np.array([T]), np.full((1, 4), duty),
                                             np.asarray(off, float)[None])[0])
        cts.append(contact)
        meas.append(m)
    mpc = _mpc(N, 64)
    mpc.set_model(**sq.go2_seq_model())
    _, ct_g = mpc.sequence_records(np.array(seqs), np.array(meas))
    ref = np.array(cts).astype(np.uint8)
    assert np.array_equal(ct_g, ref)
    # the rule must actually have fired somewhere in the sample
    _, ct_plain = mpc.sequence_records(np.array(seqs))
    assert (ct_g != ct_plain).any()


def test_sequence_and_solve_equals_two_step_path():
    """Fused sequencer+solve call gives the same forces as host sequencer -> solve_records (records agree to 1e-12,
    so the QP solutions agree to the solver tolerance)."""
    from dfki_quad_b200 import sequencer as sq
    N, n = 10, 2048
    batch = sq.random_batch(n, N, seed=11, gaits=("TROT", "WALKING_TROT", "STAND"))
    mpc = _mpc(N, n)
    mpc.set_model(**sq.go2_seq_model())
    rec_g, ct_g = mpc.sequence_records(batch["seq_in"])
    assert np.array_equal(ct_g, batch["contact"])
    np.testing.assert_allclose(rec_g, batch["rec"], rtol=0, atol=1e-12)
    f1, x1, i1 = mpc.solve_records(batch["rec"], batch["contact"])
    f2, x2, i2 = mpc.sequence_and_solve(batch["seq_in"])
    assert i1.success.all() and i2.success.all()
    scale = np.abs(f1).max(axis=(1, 2), keepdims=True)
    assert (np.abs(f1 - f2) / scale).max() < 1e-5
    with pytest.raises(Exception):
        _mpc(N, 4).sequence_and_solve(batch["seq_in"][:4])  # no model set


# ------------------------------------------------------------------ warm / hot start (MPC::MPC warm_start, mpc.hpp:127)
@pytest.mark.parametrize("mode", [1, 2])
def test_warm_and_hot_start_reach_the_same_optimum(mode, go2_params):
    """Second solve of slightly perturbed problems: warm (1) and hot (2) start return the cold-start optimum to 1e-5 with far
    fewer ADMM iterations; hot start needs none when the active set did not change."""
    from dfki_quad_b200 import sequencer

    N, n = 10, 1024
    batch = sequencer.random_batch(n, N, seed=91, gaits=("TROT", "WALKING_TROT", "STAND"))
    rec2 = batch["rec"].copy()
    rng = np.random.default_rng(5)
    rec2[:, 10:13] += rng.uniform(-0.02, 0.02, (n, 3))  # linear velocity
    rec2[:, 7:10] += rng.uniform(-0.02, 0.02, (n, 3))  # angular velocity
    cold = _mpc(N, n)
    fc, xc, ic = cold.solve_records(rec2, batch["contact"])
    warm = _mpc(N, n, warm_start=mode)
    f1, _, i1 = warm.solve_records(batch["rec"], batch["contact"])  # nothing stored yet: identical to a cold solve
    f1c, _, i1c = cold.solve_records(batch["rec"], batch["contact"])
    assert np.array_equal(f1, f1c) and np.array_equal(i1.acados_num_iter, i1c.acados_num_iter)
    f2, x2, i2 = warm.solve_records(rec2, batch["contact"])
    assert ic.success.all() and i2.success.all()
    assert np.all(i2.kkt.max(axis=1) <= 1e-6)
    scale = np.maximum(1.0, np.abs(fc).reshape(n, -1).max(axis=1))
    assert (np.abs(f2 - fc).reshape(n, -1).max(axis=1) / scale).max() <= 1e-5
    assert i2.acados_num_iter.mean() < (0.75 if mode == 1 else 0.3) * ic.acados_num_iter.mean(), (i2.acados_num_iter.mean(), ic.acados_num_iter.mean())
    if mode == 2:
        assert (i2.acados_num_iter == 0).mean() > 0.7, (i2.acados_num_iter == 0).mean()
    _check_against_oracle(N, dict(batch, rec=rec2), f2, x2, i2, go2_params, range(3))
    # a different contact pattern in the same slots (another gait phase) must still converge to the right answer
    other = sequencer.random_batch(n, N, seed=92, gaits=("TROT", "PACE"))
    f3, x3, i3 = warm.solve_records(other["rec"], other["contact"])
    f3c, _, i3c = cold.solve_records(other["rec"], other["contact"])
    assert i3.success.all() and i3c.success.all()
    scale = np.maximum(1.0, np.abs(f3c).reshape(n, -1).max(axis=1))
    assert (np.abs(f3 - f3c).reshape(n, -1).max(axis=1) / scale).max() <= 1e-5
    # switching the mode forgets the stored solutions
    warm.SetWarmStart(mode)
    f4, _, i4 = warm.solve_records(other["rec"], other["contact"])
    assert np.array_equal(i4.acados_num_iter, i3c.acados_num_iter)


def test_sequencer_warp_kernel_equals_the_serial_statement(monkeypatch):
    """The warp-per-robot sequencer performs the serial kernel's floating-point operations in the same order: identical bits."""
    from dfki_quad_b200 import sequencer as sq

    N, n = 20, 1500
    batch = sq.random_batch(n, N, seed=123, gaits=tuple(sq.gaitmod.GAIT_NAMES))
    rng = np.random.default_rng(1)
    meas = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    mpc = _mpc(N, n)
    mpc.set_model(**sq.go2_seq_model())
    rec_w, ct_w = mpc.sequence_records(batch["seq_in"], meas)
    monkeypatch.setenv("B200QP_SERIAL_SEQUENCER", "1")
    rec_s, ct_s = mpc.sequence_records(batch["seq_in"], meas)
    assert np.array_equal(ct_w, ct_s)
    assert np.array_equal(rec_w.view(np.uint64), rec_s.view(np.uint64))


def test_randomised_options_soak():
    """Random horizons, gaits, weights, alpha, friction (incl. mu > 1: the x/y box rows can bind, NNLS dual path), force limits,
    solver and warm-start mode: every problem that fits its solver converges to KKT <= 1e-6, is feasible, has exact zeros on
    swing legs, and a sample agrees with the (polished) oracle to 1e-5."""
    from dfki_quad_b200 import BatchedMPC, sequencer
    from dfki_quad_b200 import gait as gaitmod

    rng = np.random.default_rng(7)
    cfg = sequencer.GO2_MPC
    for trial in range(12):
        N = int(rng.choice([1, 2, 5, 8, 10, 13, 16, 20]))
        n = 192
        gaits = tuple(rng.choice(gaitmod.GAIT_NAMES, size=rng.integers(1, 4), replace=False))
        w = np.array(cfg["state_weights_move"], float) * np.exp(rng.normal(0, 1.0, 12))
        alpha = float(10 ** rng.uniform(-5.5, -3))
        mu = float(rng.choice([0.2, 0.4, 0.6, 0.9, 1.2]))
        fmin = float(rng.choice([0.0, 0.0, 5.0]))
        fmax = float(rng.choice([120.0, 250.0, 400.0]))
        solver = str(rng.choice(["CONDENSED_ADMM", "SPARSE_RICCATI", "AUTO"]))
        ws = int(rng.choice([0, 0, 2]))
        b = sequencer.random_batch(n, N, seed=2000 + trial, gaits=gaits)
        mpc = BatchedMPC(alpha, w, mu, fmin, fmax, solver, warm_start=ws, horizon=N, dt=cfg["dt"], max_batch=n)
        f, x, info = mpc.solve_records(b["rec"], b["contact"])
        if ws:
            f, x, info = mpc.solve_records(b["rec"], b["contact"])
        tag = (trial, N, solver, ws, mu, fmin, fmax, alpha, gaits)
        nfree = 3 * b["contact"].reshape(n, -1).sum(axis=1)
        fits = (nfree <= 156) | (solver != "CONDENSED_ADMM")
        ok = info.success
        assert ok[fits].all(), (tag, np.unique(info.return_code[fits], return_counts=True))
        assert (info.kkt[ok].max(axis=1) <= 1e-6).all(), tag
        F = f.reshape(n, N, 4, 3)[ok]
        assert (np.abs(F[..., 0]) <= mu * F[..., 2] + 1e-6).all() and (np.abs(F[..., 1]) <= mu * F[..., 2] + 1e-6).all(), tag
        assert (F[..., 2] <= fmax + 1e-6).all() and np.all(F[b["contact"][ok] == 0] == 0.0), tag
        assert (F[..., 2][b["contact"][ok] != 0] >= fmin - 1e-6).all(), tag
        params = dict(mo.GO2_PARAMS, weights=w, alpha=alpha, mu=mu, fmin=fmin, fmax=fmax)
        p = int(np.nonzero(ok)[0][trial % int(ok.sum())])
        U, X, qp, dq = mo.solve_reference_qp(N, b["rec"][p], b["contact"][p], params)
        x2, _, okp = mo.polish_active_set(dq["H"], dq["g"], dq["A"], np.where(dq["l"] <= -mo.INF_THRESHOLD, -np.inf, dq["l"]),
                                          np.where(dq["u"] >= mo.INF_THRESHOLD, np.inf, dq["u"]), U.reshape(-1),
                                          _ipm_duals(dq, U.reshape(-1)))
        Uref = x2.reshape(N, 12) if okp else U
        assert np.abs(f[p].reshape(N, 12) - Uref).max() / max(1.0, np.abs(Uref).max()) <= 1e-5, tag
        mpc.close()


def test_plugin_interface_calls_like_the_reference(go2_params):
    """The MPCInterface call sequence of mit_controller_node.cpp:764-794 - UpdateState, UpdateModel, UpdateGaitSequence,
    GetWrenchSequence - on the batched mirror: same answer as the record-level entry point, WrenchSequence / MPCPrediction /
    SolverInformation shaped as in the reference (wrench_sequence.hpp:9, mpc_prediction.hpp:7-13, mpc_interface.hpp:10-20)."""
    from dfki_quad_b200 import BatchedMPC, sequencer

    N, n = 10, 64
    batch = sequencer.random_batch(n, N, seed=17, gaits=("TROT", "STAND"))
    cfg = sequencer.GO2_MPC
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "PARTIAL_CONDENSING_OSQP",
                     horizon=N, dt=cfg["dt"], max_batch=n)  # the reference's plan string -> AUTO (condensed kernel here: n <= 156)
    with pytest.raises(RuntimeError):
        mpc.GetWrenchSequence()  # nothing updated yet
    mpc.UpdateState(batch["quat"], batch["pos"], batch["omega"], batch["vel"])
    mpc.UpdateModel(sequencer.GO2_MASS, sequencer.GO2_INERTIA, sequencer.GO2_COM)
    traj = batch["traj"]
    mpc.UpdateGaitSequence(dict(contact_sequence=np.pad(batch["contact"], ((0, 0), (0, 5), (0, 0))),  # longer than N, like GAIT_SEQUENCE_SIZE
                                foot_position_sequence=batch["foot_pos"],
                                desired_reference_trajectory_orientation=traj["des_quat"],
                                desired_reference_trajectory_position=traj["des_pos"],
                                reference_trajectory_twist=traj["ref_twist"], reference_trajectory_velocity=traj["ref_vel"]))
    forces, pred, info = mpc.GetWrenchSequence()
    assert forces.shape == (n, N, 4, 3) and pred.raw_data.shape == (n, N + 1, 13)
    assert pred.orientation.shape == (n, N + 1, 4) and pred.position.shape == (n, N + 1, 3)
    assert info.success.all() and (info.return_code == 0).all() and info.acados_num_iter.min() >= 0 and info.total_solver_time > 0
    f2, x2, i2 = mpc.solve_records(batch["rec"], batch["contact"])
    assert np.array_equal(forces.reshape(n, N, 12), f2) and np.array_equal(pred.raw_data, x2)
    # prediction row 0 is the measured state (mpc.cpp:392-396, 460-466): position without the COM offset, quaternion up to sign
    assert np.allclose(pred.position[:, 0], batch["pos"], atol=1e-12)
    assert np.allclose(np.abs(np.sum(pred.orientation[:, 0] * batch["quat"], axis=1)), 1.0, atol=1e-12)
    assert np.allclose(pred.linear_velocity[:, 0], batch["vel"], atol=1e-12)
    _check_against_oracle(N, batch, f2, x2, i2, go2_params, range(2))


def test_streamed_inputs_give_the_resident_answer_every_time():
    """Host-buffer call with streamed record chunks (kernel polls per-chunk ready flags) vs the device-resident solve of the same
    batch: identical bits, 25 times in a row (pinned and pageable callers)."""
    import torch
    from dfki_quad_b200 import sequencer

    N, n = 10, 4096
    batch = sequencer.random_batch(n, N, seed=314, gaits=("TROT", "PACE", "WALKING_TROT"))
    mpc = _mpc(N, n)
    dev = torch.device("cuda", 0)
    ext = torch.cuda.ExternalStream(mpc.stream(), device=dev)
    with torch.cuda.stream(ext):
        d_in, d_ct = torch.from_numpy(batch["rec"]).to(dev), torch.from_numpy(batch["contact"]).to(dev)
        d_f = torch.zeros((n, N, 12), dtype=torch.float64, device=dev)
        d_x = torch.zeros((n, N + 1, 13), dtype=torch.float64, device=dev)
        d_i = torch.zeros((n, 48), dtype=torch.uint8, device=dev)
    ext.synchronize()
    mpc.solve_device(n, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_i.data_ptr(), sync=True)
    f_ref, x_ref = d_f.cpu().numpy(), d_x.cpu().numpy()
    rec_p, ct_p = torch.from_numpy(batch["rec"]).pin_memory().numpy(), torch.from_numpy(batch["contact"]).pin_memory().numpy()
    f_p = torch.zeros((n, N, 12), dtype=torch.float64).pin_memory().numpy()
    x_p = torch.zeros((n, N + 1, 13), dtype=torch.float64).pin_memory().numpy()
    for rep in range(25):
        f_p[:] = 0.0
        if rep % 2 == 0:
            f, x, info = mpc.solve_records(rec_p, ct_p, f_p, x_p)
        else:
            f, x, info = mpc.solve_records(batch["rec"], batch["contact"])
        assert info.success.all(), rep
        assert np.array_equal(f, f_ref) and np.array_equal(x, x_ref), rep


def test_bench_own_arm_prints_the_contract_line():
    """bench.py on a small batch: one JSON line with every key the measurement contract names."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--batch", "512", "--steps", "3", "--warmup", "3",
                          "--cpu-budget", "1", "--no-closed-loop"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 3 and line["dtype"] == "f64" and line["scaling"] == "weak"
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and line["config"]["converged_fraction"] == 1.0 and "workload" in line["config"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    assert line["cpu_baseline"]["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_auto_solver_gives_condensed_failures_a_second_opinion(go2_params):
    """Solver AUTO hands problems the condensed kernel gives up on (iteration cap) to the Riccati kernel.  With the ADMM cap set
    absurdly low the condensed kernel alone fails most problems; AUTO still solves all of them, to the same optimum."""
    from dfki_quad_b200 import sequencer

    N, n = 10, 512
    batch = sequencer.random_batch(n, N, seed=55, gaits=("TROT", "WALKING_TROT"))
    f_ref, x_ref, i_ref = _mpc(N, n).solve_records(batch["rec"], batch["contact"])
    fc, _, ic = _mpc(N, n, admm_max_iter=10, polish_max_rounds=1).solve_records(batch["rec"], batch["contact"])
    assert (~ic.success).mean() > 0.2 and set(np.unique(ic.return_code[~ic.success])) <= {2, 3}
    fa, xa, ia = _mpc(N, n, solver="AUTO", admm_max_iter=10, polish_max_rounds=1).solve_records(batch["rec"], batch["contact"])
    assert ia.success.all() and np.all(ia.kkt.max(axis=1) <= 1e-6)
    rescued = ~ic.success
    assert (ia.polish_rounds[rescued] == 0).all()  # those came from the interior-point kernel
    scale = np.maximum(1.0, np.abs(f_ref).reshape(n, -1).max(axis=1))
    assert (np.abs(fa - f_ref).reshape(n, -1).max(axis=1) / scale).max() <= 1e-5


def test_riccati_two_cycle_regression():
    """A rotary-gallop problem with single-foot stance stages at alpha = 1.1e-6 on which Mehrotra's method used to fall into a
    two-cycle (gap 0.3 <-> 0.65 for 200 iterations; found by scripts/fuzz.py).  The monotone-gap safeguard must solve it and its
    neighbourhood, to the condensed kernel's answer."""
    import os
    from dfki_quad_b200 import BatchedMPC, sequencer

    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "riccati_two_cycle.npz"))
    N, cfg = int(d["N"]), sequencer.GO2_MPC
    rng = np.random.default_rng(0)
    recs = np.repeat(d["rec"][None], 64, axis=0)
    recs[1:, 4:13] += rng.normal(0, 1e-3, (63, 9))
    cts = np.repeat(d["contact"][None], 64, axis=0)
    args = (float(d["alpha"]), d["w"], float(d["mu"]), float(d["fmin"]), float(d["fmax"]))
    fr, _, ir = BatchedMPC(*args, "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=64).solve_records(recs, cts)
    fa, _, ia = BatchedMPC(*args, "CONDENSED_ADMM", horizon=N, dt=cfg["dt"], max_batch=64).solve_records(recs, cts)
    assert ir.success.all() and ia.success.all(), (np.unique(ir.return_code, return_counts=True))
    assert ir.acados_num_iter.max() <= 25
    scale = np.maximum(1.0, np.abs(fa).reshape(64, -1).max(axis=1))
    assert (np.abs(fr - fa).reshape(64, -1).max(axis=1) / scale).max() <= 1e-5


def test_interior_point_warm_start_saves_iterations_and_keeps_the_optimum():
    """warm_start = 1 with solver SPARSE_RICCATI (MPC::MPC(..., warm_start), mpc.cpp:236-237): the warp kernel starts from 0.9 x the slot's
    previous forces + 0.1 x the cold interior point with complementarity 0.1; same optimum as the cold solve, fewer iterations."""
    from dfki_quad_b200 import BatchedMPC, sequencer

    N, n = 16, 2048
    c = sequencer.GO2_MPC
    b = sequencer.random_batch(n, N, seed=5, gaits=("TROT", "PACE", "BOUND"))
    rng = np.random.default_rng(1)
    rec2 = b["rec"].copy()
    rec2[:, 4:7] += rng.normal(0, 3e-3, (n, 3))
    rec2[:, 7:13] += rng.normal(0, 2e-2, (n, 6))
    args = (c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], "SPARSE_RICCATI")
    cold = BatchedMPC(*args, horizon=N, dt=c["dt"], max_batch=n)
    fc, xc, ic = cold.solve_records(rec2, b["contact"])
    warm = BatchedMPC(*args, horizon=N, dt=c["dt"], max_batch=n, warm_start=1)
    f0, _, i0 = warm.solve_records(b["rec"], b["contact"])       # nothing stored yet: a cold solve
    assert abs(i0.acados_num_iter.mean() - ic.acados_num_iter.mean()) < 0.5
    f1, x1, i1 = warm.solve_records(rec2, b["contact"])
    ok = i1.success & ic.success
    assert ok.mean() > 0.999
    assert i1.acados_num_iter[ok].mean() < ic.acados_num_iter[ok].mean() - 1.0
    sc = np.maximum(1.0, np.abs(fc).reshape(n, -1).max(axis=1))
    err = np.abs(f1 - fc).reshape(n, -1).max(axis=1) / sc
    assert err[ok].max() <= 5e-5 and np.quantile(err[ok], 0.999) <= 1e-5
    warm.SetWarmStart(0)
    f2, _, i2 = warm.solve_records(rec2, b["contact"])
    assert np.array_equal(f2[ok], fc[ok])  # warm start off again: the cold solve, bit for bit
    cold.close()
    warm.close()
