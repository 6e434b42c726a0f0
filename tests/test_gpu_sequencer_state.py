"""Kernel K5b on the GPU: the gait sequencers WITH their state across calls (b200qp_mpc_sequencer_*), one state slot per robot,
against (a) the committed multi-call sequences recorded from the compiled reference and (b) live SimpleGaitSequencer /
AdaptiveGaitSequencer objects of oracle/_ref driven over the same streams (the library travels to the GPU box)."""
import os

import numpy as np
import pytest

import seq_cases as sc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")


def _mpc(N=10, solver="CONDENSED_ADMM", max_batch=256):
    from dfki_quad_b200 import BatchedMPC
    from oracle import mpc_oracle as mo

    P = dict(mo.GO2_PARAMS)
    m = BatchedMPC(P["alpha"], P["w_move"], P["mu"], P["fmin"], P["fmax"], solver=solver, horizon=N, max_batch=max_batch)
    m.set_model(**sc.model_kwargs())
    return m


def test_committed_reference_sequences_call_after_call():
    from test_seq_kernel_logic import golden_streams

    gold = np.load(GOLD)
    m = _mpc()
    for name in ("trot_move_stop", "stand_hold", "pace_turn", "walk_world_frame"):
        rows, ms = golden_streams(gold, name)
        m.sequencer_configure("simple", **sc.OPTIONS)
        for i, row in enumerate(rows):
            rec, ct = m.sequencer_step(row, ms[i][None])
            assert np.array_equal(ct[0], gold[f"seq_{name}_contact"][i]), (name, i)
            np.testing.assert_allclose(rec[0], gold[f"seq_{name}_rec"][i], rtol=0, atol=1e-12, err_msg=f"{name} call {i}")
    m.close()


@pytest.mark.parametrize("adaptive", [False, True])
@pytest.mark.parametrize("variant", ["default", "fast_long_holds", "delayed_offsets_stride_limit", "disturbance_correct_all",
                                     "free_position_no_latch"])
def test_batched_state_slots_follow_live_reference_objects(adaptive, variant):
    from oracle import ref_binding as rb
    from test_seq_kernel_logic import VARIANTS

    if not rb.available(10):
        pytest.skip("oracle/_ref not built")
    N = 10
    stream_kw, opt, aopt = VARIANTS.get(variant, ({}, {}, {}))
    po = (0.0, 0.25, 0.5, 0.75) if variant == "fast_long_holds" else (0.0, 0.5, 0.5, 0.0)
    robots = sc.make_streams(96, 80, seed=300 + adaptive, adaptive=adaptive, phase_offset=po, **stream_kw)
    want_rec, want_ct = sc.reference_records(robots, N, adaptive=adaptive, options=opt, adaptive_options=aopt)
    m = _mpc(N)
    kw = {**sc.OPTIONS, **sc.ADAPTIVE, **opt, **aopt}
    m.sequencer_configure("adaptive" if adaptive else "simple", phase_offset=po, **kw)
    worst = 0.0
    for i in range(len(robots[0]["calls"])):
        rows, meas = sc.pack_call(robots, i)
        rec, ct = m.sequencer_step(rows, meas)
        bad = np.nonzero((ct != want_ct[i]).reshape(len(robots), -1).any(axis=1))[0]
        assert bad.size == 0, (i, bad[:5], [robots[b]["calls"][i]["cmd"] for b in bad[:2]])
        worst = max(worst, float(np.abs(rec - want_rec[i]).max()))
    assert worst <= 1e-11, worst
    # a reset forgets everything: the first call of a new run equals the first call of the old one
    m.sequencer_reset()
    rows, meas = sc.pack_call(robots, 0)
    rec, ct = m.sequencer_step(rows, meas)
    assert np.array_equal(ct, want_ct[0])
    np.testing.assert_allclose(rec, want_rec[0], rtol=0, atol=1e-11)
    m.close()


@pytest.mark.parametrize("N", [16, 20])
@pytest.mark.parametrize("adaptive", [False, True])
def test_other_horizons_follow_live_reference_objects(adaptive, N):
    from oracle import ref_binding as rb

    if not rb.available(N):
        pytest.skip("oracle/_ref not built for this horizon")
    robots = sc.make_streams(48, 40, seed=900 + adaptive + N, adaptive=adaptive, phase_offset=(0.0, 0.5, 0.5, 0.0), speed=2.0, hold=(4, 12))
    want_rec, want_ct = sc.reference_records(robots, N, adaptive=adaptive)
    m = _mpc(N, solver="SPARSE_RICCATI")
    m.sequencer_configure("adaptive" if adaptive else "simple", phase_offset=(0.0, 0.5, 0.5, 0.0), **sc.OPTIONS, **sc.ADAPTIVE)
    for i in range(len(robots[0]["calls"])):
        rows, meas = sc.pack_call(robots, i)
        rec, ct = m.sequencer_step(rows, meas)
        assert np.array_equal(ct, want_ct[i]), i
        np.testing.assert_allclose(rec, want_rec[i], rtol=0, atol=1e-11)
    m.close()


def test_step_solve_is_step_then_solve():
    N = 10
    robots = sc.make_streams(128, 6, seed=9)
    a, b = _mpc(N), _mpc(N)
    a.sequencer_configure("simple", **sc.OPTIONS)
    b.sequencer_configure("simple", **sc.OPTIONS)
    for i in range(6):
        rows, meas = sc.pack_call(robots, i)
        rec, ct = a.sequencer_step(rows, meas)
        f1, x1, i1 = a.solve_records(rec, ct)
        f2, x2, i2 = b.sequencer_step_solve(rows, meas)
        assert np.array_equal(f1, f2) and np.array_equal(x1, x2)
        assert np.array_equal(i1.return_code, i2.return_code)
        assert (i2.return_code == 0).mean() > 0.95
    a.close()
    b.close()


def test_argument_checks():

    m = _mpc(max_batch=8)
    with pytest.raises(Exception):
        m.sequencer_step(np.zeros((1, 64)))  # not configured
    m.sequencer_configure("simple", **sc.OPTIONS)
    with pytest.raises(Exception):
        m.sequencer_step(np.zeros((9, 64)))  # more robots than slots
    with pytest.raises(Exception):
        m.sequencer_configure("simple", raibert_filtersize=33)
    with pytest.raises(Exception):
        m.sequencer_configure("adaptive", filter_size=25)
    m.close()
