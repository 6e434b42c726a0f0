"""Closed loop on the device (b200qp_mpc_rollout, kernels K7 + K5b + K1/K2 [+ K6 + K3]) against the same loop stepped on the host in
numpy with one library call per solve (rollout.closed_loop_stateful): plant integration, foot bookkeeping, leg inverse kinematics and
the MPC -> WBC coupling are restated independently there."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MOVE = dict(hybrid_x_dot=0.4, hybrid_y_dot=0.1, wz=0.3, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)


def _setup(n, gait="TROT", seed=0, solver="CONDENSED_ADMM", target=MOVE, yaw_range=0.3):
    from dfki_quad_b200 import BatchedMPC, gait as gaitmod, pack_seq2, sequencer as sq
    from dfki_quad_b200.mpc import SEQ_OPTION_DEFAULTS

    rng = np.random.default_rng(seed)
    c = sq.GO2_MPC
    m = BatchedMPC(c["alpha"], c["state_weights_move"], c["mu"], c["fmin"], c["fmax"], solver, horizon=10, dt=c["dt"], max_batch=n)
    m.set_model(sq.GO2_MASS, sq.GO2_INERTIA, sq.GO2_COM, sq.GO2_SHOULDERS, sq.RAIBERT_K, sq.RAIBERT_G, sq.MAX_LEG_LENGTH)
    m.sequencer_configure("simple", **{k: v for k, v in SEQ_OPTION_DEFAULTS.items() if k in ("raibert_filtersize", "fix_standing_position")},
                          early_contact_detection=0)
    yaw = rng.uniform(-yaw_range, yaw_range, n)
    quat = sq._yaw_quat(yaw)
    pos = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.full(n, 0.30)], axis=-1)
    feet = pos[:, None, :] + np.einsum("nij,lj->nli", sq.quat_to_rot(quat), sq.GO2_SHOULDERS)
    feet[:, :, 2] = 0.0
    state = np.concatenate([quat, pos, np.zeros((n, 6))], axis=1)
    gid = np.full(n, gaitmod.GAIT_NAMES.index(gait))
    period, duty, offset = gaitmod.gait_arrays(gid)
    t0 = rng.integers(0, 500_000_000, n).astype(np.float64)
    return m, pack_seq2(state, feet, t0, target, (period, duty, offset))


@pytest.mark.parametrize("plant_inertia,yaw_range", [("reference_model", np.pi), ("physical", 0.03)])
def test_device_loop_equals_the_host_stepped_loop(plant_inertia, yaw_range):
    """A physical plant (Iw = R Ib R') is only consistent with the reference MPC's model near heading zero: MPC::GetIWorld rotates the
    body inertia with GetR(psi) = Rz(psi)' (mpc.cpp:46-64), reproduced in the kernels for parity; "reference_model" uses that inertia."""
    from dfki_quad_b200 import rollout

    n, periods = 96, 16
    m1, seq = _setup(n, seed=1, yaw_range=yaw_range)
    m2, _ = _setup(n, seed=1, yaw_range=yaw_range)
    s1, s2 = seq.copy(), seq.copy()
    h_dev = m1.rollout(s1, periods, plant_inertia=plant_inertia)
    h_host = rollout.closed_loop_stateful(m2, s2, periods, plant_inertia=plant_inertia)
    assert (h_dev[:, :, 11] == 0).mean() > 0.99
    assert np.array_equal(h_dev[:, :, 11], h_host[:, :, 11])
    np.testing.assert_allclose(h_dev[:, :, 0:10], h_host[:, :, 0:10], rtol=0, atol=1e-7)
    np.testing.assert_allclose(s1[:, :26], s2[:, :26], rtol=0, atol=1e-7)
    assert np.array_equal(s1[:, 55], s2[:, 55]) and np.array_equal(s1[:, 25], s2[:, 25])
    # the robots walk: forward progress in the heading direction, height held, gravity carried
    assert np.abs(h_dev[-1, :, 2] - 0.30).max() < 0.03
    assert np.abs(h_dev[4:, :, 10].mean() - 15.019 * 9.8067) < 5.0
    m1.close()
    m2.close()


def test_device_loop_with_the_wbc_in_it_equals_the_host_stepped_loop():
    from dfki_quad_b200 import rollout, wbc

    n, periods = 48, 12
    m1, seq = _setup(n, seed=2, yaw_range=0.03)
    m2, _ = _setup(n, seed=2, yaw_range=0.03)
    w1, w2 = wbc.BatchedWBCSolver(max_batch=n), wbc.BatchedWBCSolver(max_batch=n)
    w1.scene_configure()
    w2.scene_configure()
    s1, s2 = seq.copy(), seq.copy()
    # the reference's rates: MPC every 10 ms (MPC_CONTROL_DT), WBC every 2 ms
    h_dev = m1.rollout(s1, periods, wbc=w1, substeps=5, wbc_per_mpc=5, control_dt=0.01)
    h_host = rollout.closed_loop_stateful(m2, s2, periods, wbc_solver=w2, substeps=5, wbc_per_mpc=5, control_dt=0.01)
    assert (h_dev[:, :, 14] == 0).mean() > 0.98 and (h_dev[:, :, 11] == 0).mean() > 0.98
    assert np.array_equal(h_dev[:, :, 11], h_host[:, :, 11]) and np.array_equal(h_dev[:, :, 14], h_host[:, :, 14])
    np.testing.assert_allclose(h_dev[:, :, 0:10], h_host[:, :, 0:10], rtol=0, atol=1e-6)
    assert np.abs(h_dev[-1, :, 2] - 0.30).max() < 0.03
    for o in (m1, m2, w1, w2):
        o.close()


def test_four_thousand_robots_one_second_of_trot_with_mpc_and_wbc_on_the_device():
    """SURVEY 8(f)-4 at scale: 4096 robots, 20 MPC periods (1 s) x 5 WBC solves each, no host work inside the loop."""
    from dfki_quad_b200 import wbc

    n, periods = 4096, 20
    m, seq = _setup(n, seed=3, yaw_range=np.pi)
    w = wbc.BatchedWBCSolver(max_batch=n)
    w.scene_configure()
    hist = m.rollout(seq, periods, wbc=w, plant_inertia="reference_model")
    assert (hist[:, :, 11] == 0).mean() > 0.995 and (hist[:, :, 14] == 0).mean() > 0.995
    assert np.abs(hist[-1, :, 2] - 0.30).max() < 0.12 and np.abs(hist[-1, :, 2] - 0.30).mean() < 0.03
    up = 1.0 - 2.0 * (hist[-1, :, 6] ** 2 + hist[-1, :, 7] ** 2)  # z-axis of the body stays close to vertical
    assert up.min() > 0.97
    assert np.isfinite(seq).all()
    m.close()
    w.close()


def test_rollout_argument_checks():
    from dfki_quad_b200 import wbc

    m, seq = _setup(8, seed=4)
    with pytest.raises(Exception):
        m.rollout(seq, 2, substeps=0)
    w = wbc.BatchedWBCSolver(max_batch=8)
    with pytest.raises(Exception):
        m.rollout(seq, 2, wbc=w)  # scene not configured
    w.scene_configure()
    with pytest.raises(Exception):
        m.rollout(seq, 2, wbc=w, substeps=10, wbc_per_mpc=3)  # does not divide
    with pytest.raises(ValueError):
        m.rollout(seq[:, :50], 2)
    m.rollout(seq, 0)
    m.close()
    w.close()
