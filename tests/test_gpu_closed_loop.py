"""Closed-loop sanity of the whole path: sequencer kernel -> QP assembly -> solve -> first-stage GRFs applied
to a single-rigid-body plant.  This is not an oracle comparison: it checks that the signs, frames and units of everything the
path produces are physically right - the robots stand, trot and track the commanded velocity."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mpc(n, solver="CONDENSED_ADMM"):
    from dfki_quad_b200 import BatchedMPC, sequencer

    cfg = sequencer.GO2_MPC
    mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], solver, horizon=10,
                     dt=cfg["dt"], max_batch=n)
    mpc.set_model(**sequencer.go2_seq_model())
    return mpc


def _yaw(q):
    from dfki_quad_b200 import sequencer

    return sequencer.yaw_from_quaternion(q)


# Physical plant for headings near zero; for random headings / turning the plant uses the world inertia of the reference's own
# MPC model (MPC::GetIWorld rotates the body inertia by -psi, see rollout.srbd_step) so that plant and prediction stay consistent.
@pytest.mark.parametrize("gait,vx,wz,yaw0,plant", [("STAND", 0.0, 0.0, 0.0, "physical"), ("TROT", 0.4, 0.0, 0.0, "physical"),
                                                  ("WALKING_TROT", -0.3, 0.0, 0.0, "physical"),
                                                  ("TROT", 0.2, 0.5, None, "reference_model"),
                                                  ("STAND", 0.0, 0.0, None, "reference_model")])
def test_closed_loop_tracks_command(gait, vx, wz, yaw0, plant):
    from dfki_quad_b200 import rollout, sequencer

    n, steps = 64, 60  # 3 s at 20 Hz
    tgt = dict(hybrid_x_dot=vx, hybrid_y_dot=0.0, wz=wz, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    h = rollout.closed_loop(_mpc(n), n, steps, tgt, gait=gait, seed=3, yaw0=yaw0, plant_inertia=plant)
    assert h["ok"].mean() >= 0.995
    pos, vel, quat = h["pos"], h["vel"], h["quat"]
    # upright and at the commanded height over the last second
    R = sequencer.quat_to_rot(quat[-20:].reshape(-1, 4))
    tilt = np.arccos(np.clip(R[:, 2, 2], -1, 1))
    assert tilt.max() < 0.15, tilt.max()
    assert np.abs(pos[-20:, :, 2] - 0.30).max() < 0.04
    # body-frame velocity over the last second follows the command; yaw rate follows wz
    yaw = _yaw(quat.reshape(-1, 4)).reshape(quat.shape[:2])
    c, s = np.cos(yaw[-20:]), np.sin(yaw[-20:])
    vbx = c * vel[-20:, :, 0] + s * vel[-20:, :, 1]
    vby = -s * vel[-20:, :, 0] + c * vel[-20:, :, 1]
    assert np.abs(vbx.mean(axis=0) - vx).max() < 0.08, (vbx.mean(axis=0).min(), vbx.mean(axis=0).max())
    assert np.abs(vby.mean(axis=0)).max() < 0.08
    dyaw = np.unwrap(yaw, axis=0)
    rate = (dyaw[-1] - dyaw[-21]) / (20 * 0.05)
    assert np.abs(rate - wz).max() < 0.12, (rate.min(), rate.max())
    # the stance feet carry the weight on average
    assert abs(h["fz"][-20:].mean() - sequencer.GO2_MASS * 9.8067) < 0.1 * sequencer.GO2_MASS * 9.8067


def test_closed_loop_hot_start_follows_the_cold_trajectory():
    """The reference runs with mpc_warm_start = 1 (mit_controller_sim_go2.yaml:25).  In closed loop consecutive problems are
    close: hot start must follow the cold-start trajectory and skip most of the ADMM work."""
    from dfki_quad_b200 import BatchedMPC, rollout, sequencer

    n, steps = 64, 40
    tgt = dict(hybrid_x_dot=0.3, hybrid_y_dot=0.0, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    cfg = sequencer.GO2_MPC
    hist = {}
    for mode in (0, 2):
        mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM",
                         warm_start=mode, horizon=10, dt=cfg["dt"], max_batch=n)
        mpc.set_model(**sequencer.go2_seq_model())
        hist[mode] = rollout.closed_loop(mpc, n, steps, tgt, gait="TROT", seed=3, yaw0=0.0)
    assert hist[0]["ok"].all() and hist[2]["ok"].all()
    assert np.abs(hist[0]["pos"] - hist[2]["pos"]).max() < 1e-4
    it0, it2 = hist[0]["iters"][5:].mean(), hist[2]["iters"][5:].mean()
    assert it2 < 0.5 * it0, (it0, it2)
