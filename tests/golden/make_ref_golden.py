"""Generates tests/golden/ref_golden.npz from oracle/_ref - i.e. from the REFERENCE'S OWN sources (gait.cpp, mpc.cpp,
simple_gait_sequencer.cpp, mpc_trajectory_planner.cpp, raibert_foot_step_planner.cpp) compiled where they lie under
/root/reference by oracle/build_ref.sh.  These are reference outputs, not oracle outputs:
  gait_*      Gait::update_sequence (gait.cpp:67-112): contact masks + swing times for random and boundary phases
  gaitdb_*    GaitDatabase::getGait (gait.cpp:732-767)
  mpcN_*      MPC::GetWrenchSequence (mpc.cpp:329-470): the stage data it hands to acados (A, B_k, q_k, x0, lbu, ubu, D, lg,
              ug) and the unpacked solution; the QP SOLVE behind the acados call is oracle/mpc_ref.c (third party in the
              reference), everything else is reference code
  seq_*       SimpleGaitSequencer over several consecutive calls (planner latching, Raibert state, phase from time)
Run from the repo root in the build container (needs /root/reference):  python tests/golden/make_ref_golden.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_baseline  # noqa: E402
from oracle import mpc_oracle as mo  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402


def seq_scenarios():
    """(name, gait, list of per-call dicts: time, target kwargs, velocity) - commands that exercise the stand-still latch
    (mpc_trajectory_planner.cpp:158-188): move, stop, stand, turn, stop."""
    move = dict(hybrid_x_dot=0.4, hybrid_y_dot=0.1, wz=0.3, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    stop = dict(hybrid_x_dot=0.0, hybrid_y_dot=0.0, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    turn = dict(hybrid_x_dot=0.0, hybrid_y_dot=0.0, wz=0.5, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=0.30)
    world = dict(x_dot=0.2, y_dot=-0.1, z_dot=0.0, wz=0.0, yaw=0.3)
    return [("trot_move_stop", "TROT", [move, move, stop, stop, stop, move, stop, stop]),
            ("stand_hold", "STAND", [stop, stop, stop, turn, stop, stop]),
            ("pace_turn", "PACE", [turn, turn, stop, stop, move, move]),
            ("walk_world_frame", "STATIC_WALK", [world, world, dict(world, x_dot=0.0, y_dot=0.0), dict(world, x_dot=0.0, y_dot=0.0)])]


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    out = {}
    rng = np.random.default_rng(20261020)
    # ---- gait schedule
    n = 1500
    ids = rng.integers(0, 10, n)
    per = np.zeros(n)
    duty = np.zeros((n, 4))
    off = np.zeros((n, 4))
    for g in range(10):
        p, d, o = rb.gait_database(g)
        out[f"gaitdb_{g}"] = np.concatenate([[p], d, o])
        per[ids == g], duty[ids == g], off[ids == g] = p, d, o
    phase = rng.uniform(0, 1, n)
    phase[:200] = rng.integers(0, 20, 200) * 0.05      # exact multiples of dt / T for T = 0.5
    phase[200:300] = 0.0
    phase[300:400] = rng.integers(0, 8, 100) * 0.125
    phase[400:450] = np.nextafter(rng.integers(1, 20, 50) * 0.05, 0.0)
    meas = rng.integers(0, 2, (n, 4)).astype(np.uint8)
    c, s, ts, keep = rb.gait_update_sequence(per, duty, off, phase, dt=0.05, steps=25, measured=meas)
    c2, s2, _, _ = rb.gait_update_sequence(per, duty, off, phase, dt=0.05, steps=25, measured=None)
    out.update(gait_ids=ids, gait_phase=phase, gait_meas=meas, gait_contact=c, gait_swing=s, gait_t_swing=ts, gait_keep=keep,
               gait_contact_nomeas=c2, gait_swing_nomeas=s2)
    # ---- MPC stage data + solutions
    params = dict(mo.GO2_PARAMS)
    params["weights"] = params["w_move"]
    for N, gaits, nprob in ((10, ("TROT", "STAND", "PACE", "PRONK"), 8), (16, ("TROT", "PACE", "BOUND"), 6), (20, ("TROT", "STATIC_WALK"), 4)):
        b = cpu_baseline.make_batch_cpu(nprob, N, 777 + N, gaits=gaits)
        m = rb.RefMPC(N, params, mass=mo.GO2_MASS, inertia=mo.GO2_INERTIA)
        keys = ("A", "B", "q", "lbu", "ubu", "x0", "D", "lg", "ug", "Q", "R")
        acc = {k: [] for k in keys}
        F = np.zeros((nprob, N, 12))
        X = np.zeros((nprob, N + 1, 13))
        pose = np.zeros((nprob, N + 1, 13))
        for p in range(nprob):
            r = m.solve(b["rec"][p], b["contact"][p], F[p], X[p])
            assert r["success"], m.lib.dfkiref_last_error()
            pose[p] = r["pose"]
            qp = m.last_qp()
            assert qp["kkt"].max() < 1e-9
            for k in keys:
                acc[k].append(qp[k])
        out[f"mpc{N}_rec"], out[f"mpc{N}_contact"], out[f"mpc{N}_forces"], out[f"mpc{N}_xpred"] = b["rec"], b["contact"], F, X
        out[f"mpc{N}_pose"] = pose
        for k in keys:
            out[f"mpc{N}_{k}"] = np.array(acc[k])
        m.close()
    # ---- SimpleGaitSequencer over consecutive calls
    N = 10
    for name, gait, cmds in seq_scenarios():
        T, d, o = rb.gait_database(gait)
        yaw0 = rng.uniform(-1.0, 1.0)
        quat = mo.euler_to_quaternion([0.02, -0.03, yaw0])
        pos = np.array([0.3, -0.2, 0.31])
        vel = np.array([0.05, 0.02, 0.0])
        omega = np.array([0.0, 0.0, 0.05])
        cy, sy = np.cos(yaw0), np.sin(yaw0)
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        feet = np.array([pos + Rz @ mo.GO2_SHOULDERS[l] for l in range(4)])
        feet[:, 2] = rng.uniform(0.0, 0.01, 4)
        state = np.concatenate([quat, pos, omega, vel])
        seq = rb.RefSequencer(N, state, feet, contacts=np.ones(4, dtype=np.uint8), period=T, duty=d, offset=o,
                              shoulders=mo.GO2_SHOULDERS, inertia=mo.GO2_INERTIA)
        states, feets, meass, times, tvals, tact, recs, cts, gstates = [], [], [], [], [], [], [], [], []
        t = 0.0
        for i, cmd in enumerate(cmds):
            # a crude plant: drift with the commanded velocity so that consecutive calls see different states
            vx, vy = cmd.get("hybrid_x_dot", cmd.get("x_dot", 0.0)), cmd.get("hybrid_y_dot", cmd.get("y_dot", 0.0))
            pos = pos + 0.01 * np.array([vx * cy - vy * sy, vx * sy + vy * cy, 0.0])
            vel = 0.8 * vel + 0.2 * np.array([vx * cy - vy * sy, vx * sy + vy * cy, 0.0])
            state = np.concatenate([quat, pos, omega, vel])
            meas = rng.integers(0, 2, 4).astype(np.uint8) if i % 3 == 2 else np.ones(4, dtype=np.uint8)
            seq.update_state(state, feet, meas, t)
            seq.update_target(**cmd)
            g = seq.get(N + 1)
            rec, ct = seq.record(state, mo.GO2_MASS, mo.GO2_INERTIA, mo.GO2_COM, g)
            vals = np.zeros(19)
            vals[9] = 1.0
            act = np.zeros(16, dtype=np.uint8)
            for k, v in cmd.items():
                vals[rb.TARGET_VALUES.index(k)] = v
                act[rb.TARGET_FLAGS.index(k)] = 1
            states.append(state); feets.append(feet.copy()); meass.append(meas); times.append(t); tvals.append(vals); tact.append(act)
            recs.append(rec); cts.append(ct); gstates.append(g["gait_state"])
            t += 0.01 * (7 if i % 2 else 13)
        seq.close()
        for k, v in (("state", states), ("feet", feets), ("meas", meass), ("time", times), ("tvals", tvals), ("tact", tact),
                     ("rec", recs), ("contact", cts), ("gait_state", gstates)):
            out[f"seq_{name}_{k}"] = np.array(v)
        out[f"seq_{name}_gait"] = np.concatenate([[T], d, o])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
