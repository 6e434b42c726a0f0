"""Multi-call scenarios for the gait sequencers with state (kernel K5b): randomised per-robot command / state streams, and the
expected MPC records from the COMPILED REFERENCE (oracle/_ref, SimpleGaitSequencer / AdaptiveGaitSequencer objects that live
across the calls).  Shared by the CPU logic test (host build of the kernel) and the GPU parity test."""
import numpy as np

from oracle import mpc_oracle as mo
from oracle import ref_binding as rb

MAX_LEG_LENGTH = 0.213 + 0.213 - 0.02
# options common to both implementations (names of b200qp_seq_options); the reference's own defaults
OPTIONS = dict(raibert_filtersize=20, fix_standing_position=1, early_contact_detection=1, fix_position_distance_threshold=0.1,
               fix_position_angular_threshold=0.26, fix_position_velocity_threshold=0.1)
ADAPTIVE = dict(swing_time=0.2, zero_velocity_threshold=0.03, standing_foot_position_threshold=0.08, min_v_cmd_factor=0.5,
                max_correction_cycles=2.0, correction_period=0.6, disturbance_correction=0.0, min_v=0.0, offset_delay=0.0,
                max_stride_length=0.0, filter_size=10, switch_offsets=1, correct_all=0)


def _command(rng, kind, speed=1.0):
    """A Target as keyword arguments; kinds cover every field / activation flag of target.hpp."""
    z = 0.30
    if kind == "move":
        return dict(hybrid_x_dot=speed * rng.uniform(-0.5, 0.5), hybrid_y_dot=speed * rng.uniform(-0.3, 0.3), wz=rng.uniform(-0.6, 0.6),
                    wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=z)
    if kind == "stop":
        return dict(hybrid_x_dot=0.0, hybrid_y_dot=0.0, wz=0.0, wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=z)
    if kind == "turn":
        return dict(hybrid_x_dot=0.0, hybrid_y_dot=0.0, wz=rng.uniform(-0.6, 0.6), wx=0.0, wy=0.0, roll=0.0, pitch=0.0, z=z)
    if kind == "world":
        return dict(x_dot=speed * rng.uniform(-0.3, 0.3), y_dot=speed * rng.uniform(-0.3, 0.3), z_dot=0.0, wz=0.0, yaw=rng.uniform(-0.5, 0.5))
    if kind == "world_stop":
        return dict(x_dot=0.0, y_dot=0.0, z_dot=0.0, wz=0.0, wx=0.0, wy=0.0)
    if kind == "pose":
        return dict(x=rng.uniform(-0.5, 0.5), y=rng.uniform(-0.5, 0.5), z=z, roll=rng.uniform(-0.1, 0.1), pitch=rng.uniform(-0.1, 0.1),
                    yaw=rng.uniform(-1, 1))
    if kind == "full":
        q = mo.euler_to_quaternion([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-1, 1)])
        return dict(full_orientation=tuple(q), hybrid_x_dot=rng.uniform(-0.3, 0.3), hybrid_z_dot=rng.uniform(-0.05, 0.05))
    if kind == "climb":
        return dict(hybrid_x_dot=0.2, hybrid_y_dot=0.0, hybrid_z_dot=rng.uniform(-0.1, 0.1), wz=0.1)
    if kind == "zdot":
        return dict(x_dot=0.1, z_dot=rng.uniform(-0.1, 0.1), wx=rng.uniform(-0.1, 0.1), wy=rng.uniform(-0.1, 0.1))
    raise KeyError(kind)


KINDS = ("move", "stop", "turn", "world", "world_stop", "pose", "full", "climb", "zdot")


def make_streams(n_robots, n_calls, seed, adaptive=False, speed=1.0, hold=(1, 5), p_stop=0.3, phase_offset=None):
    """Per robot: gait, and per call a state / feet / measured contacts / clock / Target.  Commands persist for a few calls so that the
    stand-still latch, the moving averages and the adaptive gait's corrections all see history."""
    rng = np.random.default_rng(seed)
    robots = []
    for r in range(n_robots):
        gait = ("TROT", "STAND", "PACE", "STATIC_WALK", "BOUND", "PRONK", "WALKING_TROT")[rng.integers(0, 7)]
        T, d, o = rb.gait_database(gait)
        yaw = rng.uniform(-np.pi, np.pi)
        eul = np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), yaw])
        pos = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), 0.30 + rng.uniform(-0.02, 0.02)])
        vel = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), 0.0])
        omega = np.array([0.0, 0.0, rng.uniform(-0.1, 0.1)])
        cy, sy = np.cos(yaw), np.sin(yaw)
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        feet = np.array([pos + Rz @ mo.GO2_SHOULDERS[l] for l in range(4)])
        feet[:, 2] = rng.uniform(0.0, 0.01, 4)
        t_ns = int(rng.integers(0, 10 ** 9))
        calls = []
        cmd, left = None, 0
        for i in range(n_calls):
            if left == 0:
                cmd = _command(rng, KINDS[rng.integers(0, len(KINDS))] if rng.uniform() > p_stop else "stop", speed)
                left = int(rng.integers(hold[0], hold[1]))
            left -= 1
            vx = cmd.get("hybrid_x_dot", cmd.get("x_dot", 0.0))
            vy = cmd.get("hybrid_y_dot", cmd.get("y_dot", 0.0))
            wz = cmd.get("wz", 0.0)
            eul = eul + np.array([0.0, 0.0, 0.02 * wz]) + rng.normal(0, 1e-3, 3)
            cy, sy = np.cos(eul[2]), np.sin(eul[2])
            vw = np.array([vx * cy - vy * sy, vx * sy + vy * cy, 0.0])
            pos = pos + 0.02 * vw + rng.normal(0, 1e-3, 3)
            vel = 0.7 * vel + 0.3 * vw + rng.normal(0, 5e-3, 3)
            omega = np.array([0.0, 0.0, 0.5 * wz]) + rng.normal(0, 5e-3, 3)
            feet = feet + np.concatenate([np.tile(0.02 * vw[:2], (4, 1)), rng.uniform(-0.002, 0.002, (4, 1))], axis=1)
            meas = rng.integers(0, 2, 4).astype(np.uint8) if rng.uniform() < 0.4 else np.ones(4, dtype=np.uint8)
            state = np.concatenate([mo.euler_to_quaternion(eul), pos, omega, vel])
            calls.append(dict(state=state, feet=feet.copy(), meas=meas, t_ns=t_ns, cmd=dict(cmd)))
            t_ns += int(rng.integers(5_000_000, 30_000_000))
        po = (0.0, 0.5, 0.5, 0.0) if rng.uniform() < 0.7 else (0.0, 0.25, 0.5, 0.75)
        po = po if phase_offset is None else tuple(phase_offset)
        robots.append(dict(gait=(T, np.array(d), np.array(o)), phase_offset=po, calls=calls))
    return robots


def reference_records(robots, N, adaptive=False, options=None, adaptive_options=None):
    """Runs one reference sequencer object per robot over its calls -> rec [calls][robots][stride], contact [calls][robots][N][4]."""
    opt = dict(OPTIONS)
    opt.update(options or {})
    aopt = dict(ADAPTIVE)
    aopt.update(adaptive_options or {})
    n_calls = len(robots[0]["calls"])
    recs = [[None] * len(robots) for _ in range(n_calls)]
    cts = [[None] * len(robots) for _ in range(n_calls)]
    for r, rob in enumerate(robots):
        c0 = rob["calls"][0]
        T, d, o = rob["gait"]
        t0 = c0["t_ns"]
        seq = rb.RefSequencer(N, c0["state"], c0["feet"], contacts=c0["meas"], adaptive=adaptive, period=T, duty=d,
                              offset=rob["phase_offset"] if adaptive else o, shoulders=mo.GO2_SHOULDERS, inertia=mo.GO2_INERTIA,
                              max_leg_length=MAX_LEG_LENGTH, **opt, **aopt)
        for i, c in enumerate(rob["calls"]):
            seq.update_state(c["state"], c["feet"], c["meas"], (c["t_ns"] - t0) * 1e-9)
            seq.update_target(**c["cmd"])
            recs[i][r], cts[i][r] = seq.record(c["state"], mo.GO2_MASS, mo.GO2_INERTIA, mo.GO2_COM)
        seq.close()
    return np.array(recs), np.array(cts)


def ref_ns(dt_ns):
    """Nanoseconds the reference's clock shows for an offset handed to the harness in seconds (duration_cast truncates)."""
    return int((dt_ns * 1e-9) * 1e9)


def pack_call(robots, i):
    """-> (seq2_in [robots][64], measured [robots][4]) of call i (b200qp.h, B200QP_SEQ2_IN_DOUBLES)."""
    from dfki_quad_b200.mpc import pack_seq2

    n = len(robots)
    rows = np.zeros((n, 64))
    meas = np.zeros((n, 4), dtype=np.uint8)
    for r, rob in enumerate(robots):
        c = rob["calls"][i]
        T, d, o = rob["gait"]
        rows[r] = pack_seq2(c["state"][None], c["feet"][None], [ref_ns(c["t_ns"] - rob["calls"][0]["t_ns"])], c["cmd"], (T, d, o))[0]
        meas[r] = c["meas"]
    return rows, meas


def model_kwargs():
    return dict(mass=mo.GO2_MASS, inertia=mo.GO2_INERTIA, com=mo.GO2_COM, shoulders=mo.GO2_SHOULDERS, raibert_k=0.03, raibert_g=9.81,
                max_leg_length=MAX_LEG_LENGTH)
