"""CPU-side tests: C-ABI surface, host logic (sequencer, sharding), failure without a GPU.  No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_schedule(period, duty, offset, phase, steps, dt):
    from oracle import mpc_oracle as mo

    n = len(period)
    c = np.zeros((n, steps, 4), dtype=np.uint8)
    s = np.zeros((n, steps, 4))
    for p in range(n):
        c[p], s[p] = mo.gait_update_sequence(period[p], duty[p], offset[p], dt, float(phase[p]), steps, None)
    return c, s


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from dfki_quad_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "b200qp.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(b200qp_[a-z_0-9]+)\s*\(", header)) - {"b200qp_mpc_in_doubles"}
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200qp.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert b"sm_100a" in _lib.load().b200qp_version()


def test_struct_layout_matches_header():
    from dfki_quad_b200 import _lib

    assert ctypes.sizeof(_lib.SolverInfo) == 48
    assert ctypes.sizeof(_lib.MpcConfig) == 16 + 8 * (2 + 12 + 3 + 5) + 32  # 7 trailing int32 (incl. warm_start, cond_N, precision) + padding


def test_no_gpu_fails_loudly():
    from dfki_quad_b200 import BatchedMPC, _lib, sequencer

    lib = _lib.load()
    if lib.b200qp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cfg = sequencer.GO2_MPC
    with pytest.raises(_lib.B200QPError):
        BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], horizon=10, max_batch=4)
    with pytest.raises(_lib.B200QPError):
        from dfki_quad_b200 import gait

        gait.contact_schedule(np.array([0.5]), np.full((1, 4), 0.5), np.zeros((1, 4)), np.array([0.0]), 5, 0.05)


def test_bad_arguments_rejected():
    from dfki_quad_b200 import _lib

    lib = _lib.load()
    cfg = _lib.MpcConfig()
    h = ctypes.c_void_p()
    cfg.horizon = 99
    assert lib.b200qp_mpc_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_ARG
    assert lib.b200qp_mpc_create(None, ctypes.byref(h)) == _lib.ERR_ARG
    assert lib.b200qp_gait_schedule(0, 1, 0, 0.05, None, None, None, None, None, 0.0, 1.0, None, None, None, None) == _lib.ERR_ARG


def test_product_never_imports_oracle():
    """The shipped path must not route through the CPU oracle (or /root/reference)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dfki_quad_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in src


def test_batched_sequencer_matches_scalar_oracle():
    from dfki_quad_b200 import sequencer
    from oracle import mpc_oracle as mo
    from oracle import sequencer_oracle as so

    N, n = 10, 12
    batch = sequencer.random_batch(n, N, seed=3, gaits=("TROT", "PACE", "BOUND", "STATIC_WALK"), schedule_fn=_oracle_schedule)
    params = dict(mo.GO2_PARAMS)
    for p in range(n):
        from dfki_quad_b200.gait import GAIT_NAMES

        tgt = {k: float(v[p]) for k, v in batch["target"].items()}
        state = dict(quat=batch["quat"][p], pos=batch["pos"][p], omega=batch["omega"][p], vel=batch["vel"][p])
        rec, contact = so.make_problem(N, GAIT_NAMES[batch["gait_ids"][p]], float(batch["phase"][p]), state, tgt,
                                       batch["feet"][p], params)
        assert np.array_equal(contact, batch["contact"][p])
        assert np.allclose(rec, batch["rec"][p], rtol=0, atol=1e-12), np.abs(rec - batch["rec"][p]).max()


def test_pack_records_layout_matches_oracle():
    from dfki_quad_b200 import pack_records
    from oracle import mpc_oracle as mo

    rng = np.random.default_rng(0)
    N = 7
    a = lambda *s: rng.normal(size=s)  # noqa: E731
    I = a(3, 3)
    rec = pack_records(N, a(1, 4), a(1, 3), a(1, 3), a(1, 3), np.array([2.0]), I, a(1, 3), a(1, N + 1, 4), a(1, N + 1, 3),
                       a(1, N + 1, 3), a(1, N + 1, 3), a(1, N, 4, 3))
    d = mo.unpack_problem(N, rec[0])
    assert np.array_equal(d["inertia"], I)
    rec2 = mo.pack_problem(N, d["quat"], d["pos"], d["omega"], d["vel"], d["mass"], d["inertia"], d["com"], d["des_quat"],
                           d["des_pos"], d["ref_twist"], d["ref_vel"], d["foot_pos"])
    assert np.array_equal(rec[0], rec2)
    assert rec.shape[1] == mo.in_doubles(N) == 26 + 13 * (N + 1) + 12 * N


def test_shard_range_partitions():
    from dfki_quad_b200.shard import shard_range

    for n in (0, 1, 7, 4096, 1048576 + 3):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 host path: each rank takes its slice, statistics are reduced off the clock."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import torch.distributed as dist\n"
        "from dfki_quad_b200.shard import shard_range, gather_stats\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = shard_range(1001, r, w)\n"
        "s = gather_stats({'solved': hi - lo, 'iters_max': 10 + r})\n"
        "assert s['solved'] == 1001 and s['iters_max'] == 11, s\n"
        "dist.destroy_process_group()\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]


def test_go2_dynamics_consistency():
    """Host-side rigid-body terms: total mass (= the MPC's 15.019 kg), symmetric positive definite M, foot Jacobians
    against finite differences of the forward kinematics."""
    from dfki_quad_b200 import go2_model as gm

    rng = np.random.default_rng(0)
    n = 2
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    pos, q = rng.normal(size=(n, 3)), gm.nominal_joint_angles() + rng.uniform(-0.3, 0.3, (n, 12))
    v, w, qd = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.normal(size=(n, 12))
    D = gm.dynamics(pos, quat, q, v, w, qd)
    assert np.allclose(D["M"][:, 0, 0], 15.019, atol=1e-9) and np.linalg.eigvalsh(D["M"]).min() > 1e-4
    eps = 1e-6
    for k in list(range(3)) + list(range(6, 18)):  # translations and joints (rotations need the quaternion exponential)
        dp, dq = np.zeros((n, 3)), np.zeros((n, 12))
        if k < 3:
            dp[:, k] = eps
        else:
            dq[:, k - 6] = eps
        fp = gm.dynamics(pos + dp, quat, q + dq, v, w, qd)["foot_pos"]
        fm = gm.dynamics(pos - dp, quat, q - dq, v, w, qd)["foot_pos"]
        assert np.abs((fp - fm) / (2 * eps) - D["foot_J"][:, :, :, k]).max() < 1e-7
    assert np.allclose(D["foot_vel"], np.einsum("nfij,nj->nfi", D["foot_J"], np.concatenate([v, w, qd], axis=1)))


def test_tsid_scene_against_oracle():
    """Reduced-TSID assembly is a well-posed strictly convex QP; the two oracle solvers agree on it."""
    from dfki_quad_b200 import go2_model as gm
    from dfki_quad_b200 import wbc
    from oracle import qp_oracle as qo

    N, n = 10, 4
    zf = np.zeros((n, N, 12))
    zf[:, :, 2::3] = 40.0
    xp = np.zeros((n, N + 1, 13))
    xp[:, :, 5] = 0.3
    b = wbc.random_wbc_batch(n, seed=5, schedule_fn=_oracle_schedule, mpc_solution=(zf, xp))
    qp = b["qp"]
    assert qp.A.shape == (n, 46, 30) and qp.neq == 18
    for p in range(n):
        args = (qp.H[p], qp.g[p], qp.A[p, :18], qp.lower_y[p, :18], qp.A[p, 18:], qp.lower_y[p, 18:], qp.upper_y[p, 18:])
        x, it = qo.solve_ipm(*args)
        x2, ok = qo.solve_active_set(*args, x)
        assert ok and np.abs(x - x2).max() <= 1e-5 * max(1.0, np.abs(x2).max())
        stat, eq, ineq = qo.kkt_residuals(*args, x2)
        assert max(stat, eq, ineq) <= 1e-7
        f = x2[18:].reshape(4, 3)
        assert np.abs(f[~b["contacts"][p]]).max() <= 1e-9
        assert np.abs(wbc.joint_torques({k: v[p:p + 1] for k, v in b["dyn"].items()}, x2[None])).max() <= gm.TAU_MAX.max() + 1e-6


def test_cpp_header_compiles_and_links(tmp_path):
    """include/b200qp_mpc.hpp (header-only C++ mirror of the MPC plugin) compiles with g++ -std=c++17 and links against
    libb200qp.so; without a GPU construction throws with the library's error string (no CPU fallback)."""
    import __graft_entry__ as ge

    ge.build()
    src = tmp_path / "t.cpp"
    src.write_text(
        '#include "b200qp_mpc.hpp"\n#include <cstdio>\n'
        "int main() {\n"
        "  std::array<double, 12> w{}; w.fill(1.0);\n"
        "  try { b200qp::BatchedMPC mpc(1e-5, w, 0.6, 0.0, 400.0, B200QP_SOLVER_CONDENSED_ADMM, 10, 0.05, 2);\n"
        "        std::printf(\"created\\n\"); }\n"
        "  catch (const std::exception &e) { std::printf(\"error: %s\\n\", e.what()); return b200qp_device_count() > 0 ? 1 : 0; }\n"
        "  return 0; }\n")
    libdir = os.path.join(ROOT, "dfki_quad_b200")
    exe = tmp_path / "t"
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    r = subprocess.run([gxx, "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lb200qp",
                        "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "created" in out.stdout or "no CPU fallback" in out.stdout or "error" in out.stdout


def test_ctypes_structs_match_the_c_header_field_by_field(tmp_path):
    """offsetof / sizeof of every ABI struct as the C compiler sees include/b200qp.h == the ctypes mirrors in _lib.py."""
    import ctypes

    from dfki_quad_b200 import _lib

    structs = {"b200qp_mpc_config": _lib.MpcConfig, "b200qp_solver_info": _lib.SolverInfo, "b200qp_wbc_config": _lib.WbcConfig,
               "b200qp_seq_model": _lib.SeqModel, "b200qp_seq_options": _lib.SeqOptions, "b200qp_wbc_scene_config": _lib.WbcSceneConfig,
               "b200qp_rollout_options": _lib.RolloutOptions}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "b200qp.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += [f'  printf("seq_in_doubles %d\\n", B200QP_SEQ_IN_DOUBLES);', "  return 0; }"]
    src = tmp_path / "o.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "o"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split("\n")
    seen = {}
    for ln in out:
        parts = ln.split()
        if len(parts) == 3:
            seen[(parts[0], parts[1])] = int(parts[2])
        elif len(parts) == 2:
            seen[(parts[0],)] = int(parts[1])
    for cname, cls in structs.items():
        assert seen[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert seen[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    from dfki_quad_b200 import sequencer

    assert seen[("seq_in_doubles",)] == sequencer.SEQ_IN_DOUBLES == 50
    # the kernel-side mirrors of the option structs are plain copies field by field (api.cu); their sizes are pinned here through the header
    assert ctypes.sizeof(_lib.SeqOptions) == 4 * 4 + 3 * 8 + 10 * 8 + 4 * 8 + 8 + 3 * 4 + 4  # 4 ints, 3 + 10 doubles, phase_offset[4], control_dt, 3 ints + tail padding
    assert ctypes.sizeof(_lib.RolloutOptions) == 4 * 4 + 18 * 8 + 8


def test_rollout_plant_conserves_what_it_should():
    """Host-side SRBD plant of the closed-loop harness: free fall, torque-free rotation (|L| constant), weight balance."""
    from dfki_quad_b200 import rollout, sequencer as sq

    n = 4
    rng = np.random.default_rng(0)
    quat = sq._yaw_quat(rng.uniform(-1, 1, n))
    pos = np.tile([0.0, 0.0, 1.0], (n, 1))
    omega = rng.uniform(-1, 1, (n, 3))
    vel = rng.uniform(-1, 1, (n, 3))
    feet = np.zeros((n, 4, 3))
    zero = np.zeros((n, 4, 3))
    L0 = np.einsum("nij,nj->ni", sq.quat_to_rot(quat) @ sq.GO2_INERTIA @ np.swapaxes(sq.quat_to_rot(quat), 1, 2), omega)
    q1, p1, w1, v1 = rollout.srbd_step(quat, pos, omega, vel, feet, zero, 0.2, substeps=400)
    assert np.allclose(v1, vel + 0.2 * rollout.GRAVITY, atol=1e-12)
    assert np.allclose(p1, pos + 0.2 * vel + 0.5 * 0.04 * rollout.GRAVITY, atol=2e-3)
    R1 = sq.quat_to_rot(q1)
    L1 = np.einsum("nij,nj->ni", R1 @ sq.GO2_INERTIA @ np.swapaxes(R1, 1, 2), w1)
    assert np.allclose(L1, L0, atol=2e-3)  # angular momentum in the world frame
    assert np.allclose(np.linalg.norm(q1, axis=1), 1.0)
    # four feet under the hips carrying m g / 4 each: no acceleration
    yaw0 = np.zeros(n)
    quat = sq._yaw_quat(yaw0)
    pos = np.tile([0.0, 0.0, 0.3], (n, 1))
    feet = pos[:, None, :] + sq.GO2_SHOULDERS[None]
    feet[:, :, 2] = 0.0
    f = np.zeros((n, 4, 3))
    xf, xr = sq.GO2_SHOULDERS[0, 0], -sq.GO2_SHOULDERS[2, 0]  # hips are not symmetric about the COM: moment balance front / rear
    w = sq.GO2_MASS * 9.8067
    f[:, :2, 2] = 0.5 * w * xr / (xf + xr)
    f[:, 2:, 2] = 0.5 * w * xf / (xf + xr)
    q2, p2, w2, v2 = rollout.srbd_step(quat, pos, np.zeros((n, 3)), np.zeros((n, 3)), feet, f, 0.05)
    assert np.abs(v2).max() < 1e-12 and np.abs(w2).max() < 1e-12


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU (it times the CPU restatement): one JSON line with the contract keys."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "64", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "solves/s" and line["vs_baseline"] is None
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"]


def test_shard_range_of_the_c_abi_equals_the_python_rule():
    """b200qp_shard_range (pure arithmetic, no CUDA) - the slicing rule of b200qp_mpc_solve_batch_sharded - equals
    dfki_quad_b200.shard.shard_range, covers [0, n) exactly once and keeps slices contiguous."""
    from dfki_quad_b200 import _lib
    from dfki_quad_b200.shard import shard_range

    lib = _lib.load()
    lo, hi = ctypes.c_int32(), ctypes.c_int32()
    for n in (0, 1, 7, 8, 9, 4096, 1048576, 1048577, 2**31 - 1):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                assert lib.b200qp_shard_range(n, rank, world, ctypes.byref(lo), ctypes.byref(hi)) == 0
                assert (lo.value, hi.value) == shard_range(n, rank, world)
                assert lo.value == prev and hi.value >= lo.value
                prev = hi.value
            assert prev == n
    assert lib.b200qp_shard_range(10, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) == _lib.ERR_ARG
