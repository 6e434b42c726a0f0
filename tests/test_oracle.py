"""CPU tests of the oracle itself: golden fixtures, known answers, two independent solvers, KKT checker."""
import os

import numpy as np
import pytest

from oracle import mpc_oracle as mo
from oracle import sequencer_oracle as so

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mpc_golden.npz"))


@pytest.fixture(scope="module")
def c_oracle():
    import __graft_entry__ as ge
    from oracle import c_oracle as co

    if not co.available():
        ge.build()
    return co


def test_gait_known_answer_trot_phase0():
    """SURVEY.md Appendix A (iii)."""
    T, d, off = mo.GAITS["TROT"]
    c, s = mo.gait_update_sequence(T, d, off, 0.05, 0.0, 11)
    exp = np.array([[1, 0, 0, 1]] * 5 + [[0, 1, 1, 0]] * 5 + [[1, 0, 0, 1]], dtype=np.uint8)
    assert np.array_equal(c, exp)
    assert s[0, 1] == (1 - 0.5) * 0.5 and s[0, 0] == 0.0


def test_gait_golden_and_early_contact():
    names = list(mo.GAITS.keys())
    for i in range(len(GOLD["gait_ids"])):
        T, d, off = mo.GAITS[names[GOLD["gait_ids"][i]]]
        c, s = mo.gait_update_sequence(T, d, off, 0.05, float(GOLD["gait_phase"][i]), 100, GOLD["gait_meas"][i])
        assert np.array_equal(c, GOLD["gait_contact"][i]) and np.array_equal(s, GOLD["gait_swing"][i])
    # early contact: leg in the second half of swing with a measured contact is held in contact until its planned stance
    T, d, off = mo.GAITS["TROT"]
    c_no, _ = mo.gait_update_sequence(T, d, off, 0.05, 0.4, 10, None)
    c_ec, _ = mo.gait_update_sequence(T, d, off, 0.05, 0.4, 10, [1, 1, 1, 1])
    assert c_no[0].tolist() == [1, 0, 0, 1] and c_ec[0].tolist() == [1, 1, 1, 1]
    assert np.array_equal(c_no[1:], c_ec[1:]) or c_ec[:2, 1].all()


def test_stand_gait_is_always_contact():
    T, d, off = mo.GAITS["STAND"]
    c, s = mo.gait_update_sequence(T, d, off, 0.05, 0.3, 100)
    assert c.all() and not s.any()


@pytest.mark.parametrize("tag,N", [("h10_trot", 10), ("h10_mixed", 10), ("h16_mixed", 16)])
def test_oracle_reproduces_golden(tag, N, go2_params):
    rec, contact, F = GOLD[tag + "_rec"], GOLD[tag + "_contact"], GOLD[tag + "_forces"]
    for p in range(0, rec.shape[0], 3):
        U, X, qp, dq = mo.solve_reference_qp(N, rec[p], contact[p], go2_params)
        assert np.abs(U - F[p]).max() <= 2e-6 * max(1.0, np.abs(F[p]).max())  # interior IPM vs polished vertex
        assert max(mo.kkt_residuals(qp, F[p])) <= 1e-9
        assert np.abs(mo.rollout(qp, F[p]) - GOLD[tag + "_xpred"][p]).max() <= 1e-12


@pytest.mark.parametrize("tag,N", [("h10_trot", 10), ("h10_mixed", 10), ("h16_mixed", 16)])
def test_c_port_matches_golden(tag, N, go2_params, c_oracle):
    rec, contact, F = GOLD[tag + "_rec"], GOLD[tag + "_contact"], GOLD[tag + "_forces"]
    f, x, info, solved = c_oracle.solve_batch(N, rec, contact, go2_params, threads=2)
    assert solved == rec.shape[0] and np.all(info["kkt"].max(axis=1) <= 1e-7)
    assert np.abs(f - F).max() <= 1e-8 * max(1.0, np.abs(F).max())
    assert np.abs(x - GOLD[tag + "_xpred"]).max() <= 1e-9
    H, g = c_oracle.condense(N, rec[0], contact[0], go2_params)
    dq = mo.condense(mo.build_ocp_qp(N, rec[0], contact[0], go2_params))
    assert np.abs(H - dq["H"]).max() <= 1e-13 * np.abs(H).max() and np.abs(g - dq["g"]).max() <= 1e-12 * np.abs(g).max()


def test_two_numpy_solvers_agree(go2_params):
    """Dense IPM and OSQP-style ADMM + polish reach the same (unique) optimum."""
    rec, contact = GOLD["h10_trot_rec"][1], GOLD["h10_trot_contact"][1]
    qp = mo.build_ocp_qp(10, rec, contact, go2_params)
    dq = mo.condense(qp)
    u1, _ = mo.solve_ipm_dense(dq["H"], dq["g"], dq["A"], dq["l"], dq["u"])
    u2, y2, inf = mo.solve_admm_dense(dq["H"], dq["g"], dq["A"], dq["l"], dq["u"], return_info=True)
    assert inf["polished"]
    assert np.abs(u1 - u2).max() <= 2e-6 * np.abs(u2).max()
    assert np.abs(u2.reshape(10, 12) - GOLD["h10_trot_forces"][1]).max() <= 1e-8 * np.abs(u2).max()


def test_known_answer_standing(go2_params):
    """Appendix A (i): level stand on 4 symmetric feet: sum f_z = m g in the early stages, f_xy = 0, equal split."""
    N = 10
    quat = np.array([0.0, 0.0, 0.0, 1.0])
    pos = np.array([0.0, 0.0, 0.3])
    feet = np.array([[0.18, 0.13, 0.0], [0.18, -0.13, 0.0], [-0.18, 0.13, 0.0], [-0.18, -0.13, 0.0]])
    rec = mo.pack_problem(N, quat, pos, np.zeros(3), np.zeros(3), mo.GO2_MASS, mo.GO2_INERTIA, np.zeros(3), [quat] * (N + 1),
                          [pos] * (N + 1), np.zeros((N + 1, 3)), np.zeros((N + 1, 3)), np.array([feet] * N))
    contact = np.ones((N, 4), dtype=np.uint8)
    U, X, qp, _ = mo.solve_reference_qp(N, rec, contact, dict(go2_params, weights=go2_params["w_stand"]))
    F = U.reshape(N, 4, 3)
    mg = mo.GO2_MASS * mo.GRAVITY_CONSTANT
    assert np.allclose(F[:6, :, 2].sum(axis=1), mg, rtol=1e-3)
    assert np.abs(F[:, :, :2]).max() < 1e-6 * mg and np.allclose(F[:, :, 2], F[:, :1, 2], rtol=1e-6)


def test_known_answer_all_swing_is_ballistic(go2_params):
    """Appendix A (ii)."""
    N = 10
    rec = GOLD["h10_trot_rec"][0]
    contact = np.zeros((N, 4), dtype=np.uint8)
    U, X, qp, _ = mo.solve_reference_qp(N, rec, contact, go2_params)
    assert np.all(U == 0.0)
    x = qp["x0"].copy()
    for k in range(N):
        x = qp["A"] @ x
        assert np.allclose(X[k + 1], x, atol=1e-14)
    assert np.isclose(X[N][11], qp["x0"][11] - N * 0.05 * mo.GRAVITY_CONSTANT)


def test_qp_data_quirks(go2_params):
    """Quirks restated from mpc.cpp: A uses the CURRENT yaw at every stage, B the unwrapped DESIRED yaw; R is Rz^T;
    force-to-velocity block (dt/m) I is present for swing feet as well; bounds pin swing feet to 0."""
    rec, contact = GOLD["h10_trot_rec"][2], GOLD["h10_trot_contact"][2]
    qp = mo.build_ocp_qp(10, rec, contact, go2_params)
    psi = mo.yaw_from_quaternion(rec[0:4])
    assert np.allclose(qp["A"][0:3, 6:9], 0.05 * np.array([[np.cos(psi), np.sin(psi), 0], [-np.sin(psi), np.cos(psi), 0], [0, 0, 1]]))
    assert qp["A"][11, 12] == -0.05 and np.allclose(qp["A"][3:6, 9:12], 0.05 * np.eye(3))
    for k in range(10):
        for f in range(4):
            assert np.allclose(qp["B"][k][9:12, 3 * f:3 * f + 3], np.eye(3) * 0.05 / mo.GO2_MASS)
            if not contact[k, f]:
                assert not qp["B"][k][6:9, 3 * f:3 * f + 3].any()
                assert not qp["lbu"][k][3 * f:3 * f + 3].any() and not qp["ubu"][k][3 * f:3 * f + 3].any()
    assert qp["x0"][12] == 9.8067 and qp["Q"][12] == 0.0
    yaws = [t[2] for t in qp["xref"]]
    assert np.all(np.abs(np.diff([psi] + yaws)) <= np.pi)


def test_kkt_checker_flags_wrong_solutions(go2_params):
    rec, contact, F = GOLD["h10_trot_rec"][0], GOLD["h10_trot_contact"][0], GOLD["h10_trot_forces"][0]
    qp = mo.build_ocp_qp(10, rec, contact, go2_params)
    assert max(mo.kkt_residuals(qp, F)) <= 1e-9
    bad = F.copy()
    k, f = np.argwhere(contact == 1)[0]
    bad[k, 3 * f + 2] += 1.0
    assert max(mo.kkt_residuals(qp, bad)) > 1e-4
    bad = F.copy()
    bad[k, 3 * f + 2] = -5.0
    assert mo.kkt_residuals(qp, bad)[2] >= 5.0 - 1e-9


def test_scalar_sequencer_invariants(go2_params):
    rng = np.random.default_rng(5)
    (rec, contact), meta = so.random_problem(rng, 10, "TROT", go2_params)
    d = mo.unpack_problem(10, rec)
    assert np.all(d["foot_pos"][contact == 0] == 0.0)  # swing feet get the zero vector (raibert_foot_step_planner.cpp:125)
    assert np.allclose(np.linalg.norm(d["des_quat"], axis=1), 1.0)
    assert np.allclose(d["des_pos"][1:, 2], 0.30)


def test_c_wbc_restatement_solves_the_hard_fixture():
    """oracle/wbc_ref.c (same method as the device kernel) on tests/golden/wbc_hard.npz against the active-set oracle's solutions."""
    import os

    from oracle import c_oracle

    if not c_oracle.available():
        pytest.skip("C oracle not built")
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wbc_hard.npz"))
    x, info, solved = c_oracle.wbc_batch(d["H"], d["g"], d["A"], d["lo"], d["hi"], 18)
    assert solved == d["g"].shape[0]
    assert np.abs(x - d["x"]).max() <= 1e-8 * np.abs(d["x"]).max()
