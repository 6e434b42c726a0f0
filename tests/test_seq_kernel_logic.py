"""Logic of the stateful sequencer kernel (K5b, dfki_quad_b200/csrc/sequencer_full.cu) against the COMPILED REFERENCE, on the CPU:
the kernel source is built as host C++ (tests/host_emu) and driven call after call next to SimpleGaitSequencer /
AdaptiveGaitSequencer objects of oracle/_ref.  The GPU run of the same kernel is tests/test_gpu_sequencer_state.py."""
import os

import numpy as np
import pytest

import seq_cases as sc
from host_emu import emu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")
needs_emu = pytest.mark.skipif(not emu.available(), reason="no CUDA headers / g++ for the host build of the kernel")


def _have_ref():
    from oracle import ref_binding as rb
    return rb.available(10)


def golden_streams(gold, name):
    """The committed multi-call sequences of make_ref_golden.py as seq2 input rows."""
    from dfki_quad_b200.mpc import TARGET_FLAGS, TARGET_VALUES, pack_seq2

    st, ft, ms, tm = gold[f"seq_{name}_state"], gold[f"seq_{name}_feet"], gold[f"seq_{name}_meas"], gold[f"seq_{name}_time"]
    tv, ta, g = gold[f"seq_{name}_tvals"], gold[f"seq_{name}_tact"], gold[f"seq_{name}_gait"]
    rows = []
    for i in range(st.shape[0]):
        cmd = {}
        for b, flag in enumerate(TARGET_FLAGS):
            if ta[i][b]:
                cmd[flag] = tuple(tv[i][6:10]) if flag == "full_orientation" else tv[i][TARGET_VALUES.index(flag)]
        rows.append(pack_seq2(st[i][None], ft[i][None], [int(tm[i] * 1e9)], cmd, (g[0], g[1:5], g[5:9])))
    return rows, ms


@needs_emu
@pytest.mark.parametrize("name", ["trot_move_stop", "stand_hold", "pace_turn", "walk_world_frame"])
def test_host_build_of_the_kernel_reproduces_the_committed_reference_sequences(name):
    gold = np.load(GOLD)
    rows, ms = golden_streams(gold, name)
    seq = emu.HostSequencer(10, 0.05, mode="simple", **sc.model_kwargs(), **sc.OPTIONS)
    for i, row in enumerate(rows):
        rec, ct = seq.sequencer_step(row, ms[i][None])
        assert np.array_equal(ct[0], gold[f"seq_{name}_contact"][i]), (name, i)
        np.testing.assert_allclose(rec[0], gold[f"seq_{name}_rec"][i], rtol=0, atol=1e-12, err_msg=f"{name} call {i}")


@needs_emu
@pytest.mark.parametrize("N", [10, 16, 20])
@pytest.mark.parametrize("adaptive", [False, True])
def test_host_build_of_the_kernel_follows_live_reference_objects(adaptive, N):
    from oracle import ref_binding as rb

    if not rb.available(N):
        pytest.skip("oracle/_ref not built for this horizon")
    robots = sc.make_streams(24 if N == 10 else 10, 40, seed=5 + adaptive + N, adaptive=adaptive)
    want_rec, want_ct = sc.reference_records(robots, N, adaptive=adaptive)
    # one sequencer object per robot in the reference <-> one state slot per robot here; phase offsets are per object
    for r, rob in enumerate(robots):
        seq = emu.HostSequencer(N, 0.05, mode="adaptive" if adaptive else "simple", **sc.model_kwargs(), **sc.OPTIONS, **sc.ADAPTIVE,
                                phase_offset=rob["phase_offset"])
        for i in range(len(rob["calls"])):
            rows, meas = sc.pack_call([rob], i)
            rec, ct = seq.sequencer_step(rows, meas)
            assert np.array_equal(ct[0], want_ct[i][r]), (r, i, rob["calls"][i]["cmd"])
            np.testing.assert_allclose(rec[0], want_rec[i][r], rtol=0, atol=1e-11, err_msg=f"robot {r} call {i} {rob['calls'][i]['cmd']}")


VARIANTS = {
    "fast_long_holds": (dict(speed=3.0, hold=(8, 30), p_stop=0.15), {}, {}),
    "delayed_offsets_stride_limit": (dict(speed=2.5, hold=(5, 20), p_stop=0.2),
                                     {}, dict(offset_delay=0.15, max_stride_length=0.25, min_v=0.2, switch_offsets=1)),
    "disturbance_correct_all": (dict(speed=2.0, hold=(5, 20), p_stop=0.2), dict(early_contact_detection=0),
                                dict(disturbance_correction=0.5, correct_all=1, max_correction_cycles=3.0, correction_period=0.4,
                                     filter_size=4, switch_offsets=0)),
    "free_position_no_latch": (dict(speed=1.0, hold=(2, 8), p_stop=0.4), dict(fix_standing_position=0, raibert_filtersize=3), {}),
}


@needs_emu
@pytest.mark.parametrize("variant", sorted(VARIANTS))
@pytest.mark.parametrize("adaptive", [False, True])
def test_host_build_of_the_kernel_option_variants(adaptive, variant):
    if not _have_ref():
        pytest.skip("oracle/_ref not built")
    N = 10
    stream_kw, opt, aopt = VARIANTS[variant]
    robots = sc.make_streams(12, 120, seed=77 + adaptive, adaptive=adaptive, **stream_kw)
    want_rec, want_ct = sc.reference_records(robots, N, adaptive=adaptive, options=opt, adaptive_options=aopt)
    if adaptive:
        assert 0.3 < want_ct.mean() < 0.97  # the gait is really stepping
    for r, rob in enumerate(robots):
        kw = {**sc.OPTIONS, **sc.ADAPTIVE, **opt, **aopt}
        seq = emu.HostSequencer(N, 0.05, mode="adaptive" if adaptive else "simple", **sc.model_kwargs(), **kw, phase_offset=rob["phase_offset"])
        for i in range(len(rob["calls"])):
            rows, meas = sc.pack_call([rob], i)
            rec, ct = seq.sequencer_step(rows, meas)
            assert np.array_equal(ct[0], want_ct[i][r]), (r, i, rob["calls"][i]["cmd"])
            np.testing.assert_allclose(rec[0], want_rec[i][r], rtol=0, atol=1e-11, err_msg=f"robot {r} call {i} {rob['calls'][i]['cmd']}")


@needs_emu
def test_host_build_of_the_wbc_scene_kernel_equals_the_numpy_assembly():
    """Kernel K6 (wbc_scene.cu: Go2 rigid-body terms + reduced-TSID scene) against go2_model.dynamics + wbc.reduced_tsid_qp."""
    from dfki_quad_b200 import go2_model as gm
    from dfki_quad_b200 import wbc
    from test_host import _oracle_schedule

    n, N = 64, 10
    zf = np.zeros((n, N, 12))
    zf[:, :, 2::3] = 40.0
    xp = np.zeros((n, N + 1, 13))
    xp[:, :, 5] = 0.3
    b = wbc.random_wbc_batch(n, seed=11, schedule_fn=_oracle_schedule, mpc_solution=(zf, xp))
    mb = b["mpc_batch"]
    scene = wbc.pack_scene(mb["pos"], mb["quat"], mb["vel"], mb["omega"], b["q"], b["qd"], b["contacts"], b["a_body"], b["a_foot"], b["f_ref"])
    H, g, A, lo, hi, feet = emu.host_wbc_scene(scene, wbc.GO2_WBC, gm.TAU_MAX)
    qp = b["qp"]
    Hn = np.transpose(qp.H, (0, 2, 1)).reshape(n, -1)
    An = np.transpose(qp.A, (0, 2, 1)).reshape(n, -1)
    assert len({tuple(c) for c in b["contacts"]}) >= 2
    np.testing.assert_allclose(H, Hn, rtol=0, atol=1e-10)
    np.testing.assert_allclose(g, qp.g, rtol=0, atol=1e-9)
    np.testing.assert_allclose(A, An, rtol=0, atol=1e-11)
    np.testing.assert_allclose(lo, np.clip(qp.lower_y, -1e20, 1e20), rtol=0, atol=1e-10)
    np.testing.assert_allclose(hi, np.clip(qp.upper_y, -1e20, 1e20), rtol=0, atol=1e-10)
    np.testing.assert_allclose(feet[:, :12], b["dyn"]["foot_pos"].reshape(n, 12), rtol=0, atol=1e-13)
    np.testing.assert_allclose(feet[:, 12:], b["dyn"]["foot_vel"].reshape(n, 12), rtol=0, atol=1e-12)
