"""CPU tests that PIN the oracle: the numpy / C restatements under oracle/ are compared with outputs of the reference's
own sources - either live through oracle/_ref (built here from /root/reference by oracle/build_ref.sh) or through the
committed fixture tests/golden/ref_golden.npz that tests/golden/make_ref_golden.py wrote from it."""
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import mpc_oracle as mo
from oracle import ref_binding as rb

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(HERE, "golden", "ref_golden.npz"))


@pytest.fixture(scope="module")
def ref_built():
    """oracle/_ref is built on demand where the reference is present; elsewhere the prebuilt libraries are used."""
    if not all(rb.available(N) for N in (10, 16, 20)) and os.path.isdir("/root/reference/ws/src"):
        import subprocess

        subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    if not rb.available(10):
        pytest.skip("oracle/_ref not built and /root/reference not present")
    return True


def _gait_params(ids):
    per = np.array([mo.GAITS[rb.GAIT_TYPES[i]][0] for i in ids])
    duty = np.array([np.broadcast_to(mo.GAITS[rb.GAIT_TYPES[i]][1], 4) for i in ids], dtype=float)
    off = np.array([mo.GAITS[rb.GAIT_TYPES[i]][2] for i in ids], dtype=float)
    return per, duty, off


# ------------------------------------------------------------------------------------------------ fixture-based (always run)
def test_oracle_gait_database_equals_reference(G):
    for g, name in enumerate(rb.GAIT_TYPES):
        T, duty, off = mo.GAITS[name]
        assert np.array_equal(G[f"gaitdb_{g}"], np.concatenate([[T], np.broadcast_to(duty, 4), off])), name


def test_oracle_gait_schedule_bit_exact_with_reference(G):
    """Gait::update_sequence of the reference (gait.cpp:67-112) vs the oracle's restatement: masks and swing times as bits."""
    ids, phase, meas = G["gait_ids"], G["gait_phase"], G["gait_meas"]
    per, duty, off = _gait_params(ids)
    for p in range(len(ids)):
        c, s = mo.gait_update_sequence(per[p], duty[p], off[p], 0.05, float(phase[p]), 25, meas[p])[:2]
        assert np.array_equal(np.asarray(c, dtype=np.uint8), G["gait_contact"][p]), (p, rb.GAIT_TYPES[ids[p]], phase[p])
        assert np.array_equal(np.asarray(s).view(np.uint64), G["gait_swing"][p].view(np.uint64)), p
        if p % 5 == 0:
            c, s = mo.gait_update_sequence(per[p], duty[p], off[p], 0.05, float(phase[p]), 25, None)[:2]
            assert np.array_equal(np.asarray(c, dtype=np.uint8), G["gait_contact_nomeas"][p])
            assert np.array_equal(np.asarray(s).view(np.uint64), G["gait_swing_nomeas"][p].view(np.uint64))


@pytest.mark.parametrize("N", [10, 16, 20])
def test_oracle_qp_data_equals_reference(G, N, go2_params):
    """The stage data MPC::GetWrenchSequence hands to acados (mpc.cpp:332-398) vs oracle build_ocp_qp: bounds, D, lg, ug byte-
    equal; A, B_k, q_k, x0 to round-off (the stand-in Eigen may order a 3-term sum differently from the numpy code)."""
    rec, ct = G[f"mpc{N}_rec"], G[f"mpc{N}_contact"]
    for p in range(rec.shape[0]):
        qp = mo.build_ocp_qp(N, rec[p], ct[p], go2_params)
        assert np.array_equal(np.array(qp["lbu"]), G[f"mpc{N}_lbu"][p]) and np.array_equal(np.array(qp["ubu"]), G[f"mpc{N}_ubu"][p])
        assert np.array_equal(qp["D"], G[f"mpc{N}_D"][p]) and np.array_equal(qp["lg"], G[f"mpc{N}_lg"][p])
        assert np.array_equal(qp["ug"], G[f"mpc{N}_ug"][p])
        assert np.array_equal(np.diag(qp["Q"]), G[f"mpc{N}_Q"][p]) and np.array_equal(qp["R"] * np.eye(12), G[f"mpc{N}_R"][p])
        for k in range(N):
            assert np.abs(qp["A"] - G[f"mpc{N}_A"][p, k]).max() <= 1e-16
            Bk = G[f"mpc{N}_B"][p, k]
            assert np.abs(qp["B"][k] - Bk).max() <= 4e-16 * max(1.0, np.abs(Bk).max())
        assert np.abs(np.array(qp["q"]) - G[f"mpc{N}_q"][p]).max() <= 1e-13
        assert np.abs(qp["x0"] - G[f"mpc{N}_x0"][p]).max() <= 1e-15


@pytest.mark.parametrize("N", [10, 16, 20])
def test_c_port_and_checkers_on_reference_solutions(G, N, go2_params):
    """Solutions that came back through the reference's GetWrenchSequence: the C port reproduces the GRFs (1e-6 relative),
    both KKT checkers accept them, the predicted states are the roll-out."""
    if not c_oracle.available():
        pytest.skip("C oracle not built")
    rec, ct, F, X = G[f"mpc{N}_rec"], G[f"mpc{N}_contact"], G[f"mpc{N}_forces"], G[f"mpc{N}_xpred"]
    f2, x2, info, solved = c_oracle.solve_batch(N, rec, ct, go2_params)
    assert solved == rec.shape[0]
    assert np.abs(f2 - F).max() <= 1e-6 * max(1.0, np.abs(F).max())
    assert np.abs(x2 - X).max() <= 1e-8
    kkt = c_oracle.kkt_batch(N, rec, ct, go2_params, F, X)
    assert kkt.max() <= 1e-9, kkt.max(axis=0)
    for p in range(min(3, rec.shape[0])):
        qp = mo.build_ocp_qp(N, rec[p], ct[p], go2_params)
        assert max(mo.kkt_residuals(qp, F[p])) <= 1e-9
        assert np.abs(mo.rollout(qp, F[p]) - X[p]).max() <= 1e-10
        # MPCPrediction: orientation = euler_to_quaternion(raw[0:3]), position = raw[3:6] - com (mpc.cpp:461-466)
        pose = G[f"mpc{N}_pose"][p]
        for k in (0, N):
            assert np.abs(np.abs(pose[k, 0:4] @ mo.euler_to_quaternion(X[p, k, 0:3])) - 1.0) <= 1e-12
            assert np.array_equal(pose[k, 4:7], X[p, k, 3:6]) and np.array_equal(pose[k, 10:13], X[p, k, 9:12])


def test_c_kkt_checker_rejects_wrong_answers(G, go2_params):
    if not c_oracle.available():
        pytest.skip("C oracle not built")
    N = 10
    rec, ct, F, X = G["mpc10_rec"], G["mpc10_contact"], G["mpc10_forces"].copy(), G["mpc10_xpred"].copy()
    k0 = c_oracle.kkt_batch(N, rec, ct, go2_params, F, X)
    stance = np.argwhere(ct[0] == 1)[0]
    F[0, stance[0], 3 * stance[1] + 2] += 1e-3          # 1 mN more on one stance foot: stationarity + dynamics break
    X[1, 3, 4] += 1e-6                                   # one corrupted predicted state: dynamics breaks
    k1 = c_oracle.kkt_batch(N, rec, ct, go2_params, F, X)
    assert k0.max() <= 1e-9 and k1[0, 0] > 1e-6 and k1[0, 1] > 1e-7 and k1[1, 1] > 5e-7
    assert np.array_equal(k0[2:], k1[2:])
    p = int(np.argmax((ct == 0).any(axis=(1, 2))))        # a problem with a swing leg somewhere
    swing = np.argwhere(ct[p] == 0)[0]
    F[p, swing[0], 3 * swing[1]] = 1e-4                  # force on a swing leg: inequality (lbu = ubu = 0) breaks
    assert c_oracle.kkt_batch(N, rec, ct, go2_params, F, X)[p, 2] >= 1e-4


# ------------------------------------------------------------------------------------------------ live against oracle/_ref
def test_reference_library_reproduces_the_fixture(G, ref_built):
    ids, phase, meas = G["gait_ids"], G["gait_phase"], G["gait_meas"]
    per, duty, off = _gait_params(ids)
    c, s, ts, keep = rb.gait_update_sequence(per, duty, off, phase, dt=0.05, steps=25, measured=meas)
    assert np.array_equal(c, G["gait_contact"]) and np.array_equal(s.view(np.uint64), G["gait_swing"].view(np.uint64))
    c, s = rb.gait_contact_sequence(per, duty, off, phase, dt=0.05, steps=25)
    assert np.array_equal(c, G["gait_contact_nomeas"])


def test_oracle_gait_schedule_vs_reference_library_many_pairs(ref_built):
    """Random (gait parameters, phase, measured contacts) beyond the database: arbitrary periods, duty factors, offsets."""
    rng = np.random.default_rng(5)
    n = 1500
    per = rng.choice([0.3, 0.35, 0.4, 0.5, 0.6, 1.25], n) * rng.choice([1.0, 1.0, 1.1], n)
    duty = rng.choice([0.2, 0.4, 0.5, 0.6, 0.75, 0.8, 1.0, 0.0], (n, 4))
    off = rng.choice([0.0, 0.25, 0.3571, 0.5, 0.75, 0.8571], (n, 4))
    phase = rng.uniform(0, 1, n)
    phase[:300] = rng.integers(0, 40, 300) * 0.025
    meas = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    c, s, _, keep = rb.gait_update_sequence(per, duty, off, phase, dt=0.05, steps=21, measured=meas)
    for p in range(n):
        co, so = mo.gait_update_sequence(per[p], duty[p], off[p], 0.05, float(phase[p]), 21, meas[p])[:2]
        assert np.array_equal(np.asarray(co, dtype=np.uint8), c[p]) and np.array_equal(np.asarray(so).view(np.uint64), s[p].view(np.uint64))


@pytest.mark.parametrize("N,gaits", [(10, ("TROT", "STAND", "BOUND")), (16, ("TROT", "PACE"))])
def test_reference_mpc_vs_oracle_live(ref_built, N, gaits, go2_params):
    """New random problems through the reference's MPC class: oracle QP data and C-port GRFs agree; the reference's
    setters (SetMu / SetFmax / SetInputWeights / SetStateWeights, mpc.cpp:294-326) change the QP like the oracle's params."""
    from oracle import cpu_baseline

    b = cpu_baseline.make_batch_cpu(6, N, 4321 + N, gaits=gaits)
    m = rb.RefMPC(N, go2_params, mass=mo.GO2_MASS, inertia=mo.GO2_INERTIA)
    params = dict(go2_params)
    for step in range(2):
        for p in range(6):
            r = m.solve(b["rec"][p], b["contact"][p])
            assert r["success"] and r["return_code"] == 0
            qp, oq = m.last_qp(), mo.build_ocp_qp(N, b["rec"][p], b["contact"][p], params)
            assert np.array_equal(qp["lbu"], np.array(oq["lbu"])) and np.array_equal(qp["D"], oq["D"])
            assert np.array_equal(np.diag(qp["Q"]), oq["Q"]) and qp["R"][0, 0] == oq["R"]
            assert np.abs(qp["B"] - np.array(oq["B"])).max() <= 1e-15 * max(1.0, np.abs(qp["B"]).max())
            if c_oracle.available():
                f2, _, _, solved = c_oracle.solve_batch(N, b["rec"][p:p + 1], b["contact"][p:p + 1], params)
                assert solved == 1 and np.abs(f2[0] - r["forces"]).max() <= 1e-6 * max(1.0, np.abs(f2).max())
        m.set_mu(0.35)
        m.set_fmax(180.0)
        m.set_input_weights(3e-5)
        m.set_state_weights(go2_params["w_stand"])
        params = dict(go2_params, mu=0.35, fmax=180.0, alpha=3e-5, weights=go2_params["w_stand"])
    m.close()


def test_reference_failure_semantics(ref_built, go2_params):
    """NaN in the state: the reference reports failure and leaves the caller's wrench / prediction untouched (mpc.cpp:445-455)."""
    from oracle import cpu_baseline

    N = 10
    b = cpu_baseline.make_batch_cpu(1, N, 99)
    rec = b["rec"][0].copy()
    rec[5] = np.nan
    m = rb.RefMPC(N, go2_params, mass=mo.GO2_MASS, inertia=mo.GO2_INERTIA)
    forces = np.full((N, 12), 123.0)
    xpred = np.full((N + 1, 13), 321.0)
    r = m.solve(rec, b["contact"][0], forces, xpred)
    assert not r["success"] and r["return_code"] != 0
    assert np.all(forces == 123.0) and np.all(xpred == 321.0)
    m.close()
