"""CPU oracle (numpy, FP64) for the dfki-quad MPC QP hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product path (dfki_quad_b200/) never does.

PINNING: the reference ships no golden vectors, unit tests or fixtures for this path, so the
anchor is the reference's own code compiled here: oracle/_ref (oracle/build_ref.sh builds gait.cpp, mpc.cpp, the planner
and the sequencers from /root/reference against stand-in Eigen / ROS / acados-interface headers).  tests/test_ref.py
holds this file to it: (a) the gait schedule bit-exact, (b) the QP *data* (A, x0, bounds, D byte-equal, B <= 4e-17, q)
equal to what the compiled MPC::GetWrenchSequence hands to acados; tests/golden/ref_golden.npz carries the same
comparison to the GPU box.  What stays "parity unpinned" is the QP SOLVE: it lives in un-vendored third-party code
(acados v0.3.6 -> HPIPM / OSQP 0.6.3; docker/Dockerfile:74-76); this file solves the strictly convex QP to 1e-10 with two
independent methods, anchored on uniqueness of the optimum plus an independent KKT checker and the hand-derived known
answers in tests/test_oracle.py.

Conventions: quaternions are stored (x, y, z, w) = Eigen's coeffs() order; matrices handed to the C ABI are
column-major like the reference's acados pointers (mpc.hpp:39-47); in numpy they are ordinary 2-D arrays.
"""
from __future__ import annotations

import math
import numpy as np

NX = 13
NU = 12
NLEGS = 4
NG = 16
GRAVITY_CONSTANT = 9.8067  # mpc.hpp:22
POS_UNBOUND = 99999999.0  # mpc.hpp:19
INF_THRESHOLD = 9.0e7  # bounds with |b| >= this are treated as absent by the solvers

# Go2 constants: ws/src/common/src/quad_model_symbolic.cpp:15-30
GO2_MASS = 15.019
GO2_INERTIA = np.array(
    [[0.202755, 0.00012166, -0.0143742], [0.00012166, 0.490926, -3.12e-05], [-0.0143742, -3.12e-05, 0.52697]]
)
GO2_COM = np.zeros(3)
# ws/src/controllers/config/mit_controller_sim_go2.yaml:23-38
GO2_PARAMS = dict(
    alpha=1e-5,
    mu=0.6,
    fmin=0.0,
    fmax=400.0,
    w_move=np.array([20.0, 200.0, 50.0, 0.0001, 0.0001, 100.0, 0.2, 0.2, 0.8, 20.0, 20.0, 0.3]),
    w_stand=np.array([30.0, 30.0, 30.0, 30.0, 30.0, 30.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]),
    dt=0.05,
)
GO2_SHOULDERS = np.array([[0.167, 0.1738, 0.0], [0.167, -0.1738, 0.0], [-0.197, 0.1738, 0.0], [-0.197, -0.1738, 0.0]])


# ------------------------------------------------------------------ record layout (shared with include/b200qp.h)
def in_doubles(N: int) -> int:
    return 26 + 13 * (N + 1) + 12 * N


def pack_problem(N, quat, pos, omega, vel, mass, inertia, com, des_quat, des_pos, ref_twist, ref_vel, foot_pos):
    """Pack one problem into the flat FP64 input record (see include/b200qp.h)."""
    rec = np.zeros(in_doubles(N))
    rec[0:4] = quat
    rec[4:7] = pos
    rec[7:10] = omega
    rec[10:13] = vel
    rec[13] = mass
    rec[14:23] = np.asarray(inertia).reshape(3, 3).T.reshape(-1)  # column-major
    rec[23:26] = com
    o = 26
    for k in range(N + 1):
        rec[o : o + 4] = des_quat[k]
        rec[o + 4 : o + 7] = des_pos[k]
        rec[o + 7 : o + 10] = ref_twist[k]
        rec[o + 10 : o + 13] = ref_vel[k]
        o += 13
    rec[o : o + 12 * N] = np.asarray(foot_pos)[:N].reshape(-1)
    return rec


def unpack_problem(N, rec):
    d = dict(
        quat=rec[0:4],
        pos=rec[4:7],
        omega=rec[7:10],
        vel=rec[10:13],
        mass=rec[13],
        inertia=rec[14:23].reshape(3, 3).T,
        com=rec[23:26],
    )
    st = rec[26 : 26 + 13 * (N + 1)].reshape(N + 1, 13)
    d["des_quat"] = st[:, 0:4]
    d["des_pos"] = st[:, 4:7]
    d["ref_twist"] = st[:, 7:10]
    d["ref_vel"] = st[:, 10:13]
    d["foot_pos"] = rec[26 + 13 * (N + 1) :].reshape(N, 4, 3)
    return d


# ------------------------------------------------------------------ quaternion / euler helpers
def quat_to_rot(q):
    """Eigen::Quaternion::toRotationMatrix, q = (x, y, z, w)."""
    x, y, z, w = (float(v) for v in q)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array(
        [
            [1.0 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1.0 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
        ]
    )


def yaw_from_quaternion(q):
    """quaternion_operations.hpp:136-141."""
    R = quat_to_rot(q)
    return math.atan2(R[1, 0], R[0, 0])


def matrix_to_euler(R):
    """quaternion_operations.hpp:106-124 (including its gimbal branches; the second branch's first operand is a
    non-zero constant in the source, i.e. always true)."""
    eps = np.finfo(float).eps
    m20 = R[2, 0]
    pitch = -math.asin(m20) if abs(m20) <= 1.0 else float("nan")
    if (1.0 - eps <= m20) and (m20 <= 1.0 + eps):
        yaw = 0.0
        roll = math.atan2(-R[0, 1], -R[0, 2])
    elif m20 <= -1.0 + eps:
        yaw = 0.0
        roll = math.atan2(R[0, 1], R[0, 2])
    else:
        yaw = math.atan2(R[1, 0], R[0, 0])
        roll = math.atan2(R[2, 1], R[2, 2])
    return np.array([roll, pitch, yaw])


def quaternion_to_euler(q):
    """quaternion_operations.hpp:126-134."""
    return matrix_to_euler(quat_to_rot(q))


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ]
    )


def euler_to_quaternion(e):
    """quaternion_operations.hpp:19-25: AngleAxis(yaw,Z) * AngleAxis(pitch,Y) * AngleAxis(roll,X)."""
    r, p, y = (float(v) for v in e)
    qz = np.array([0.0, 0.0, math.sin(0.5 * y), math.cos(0.5 * y)])
    qy = np.array([0.0, math.sin(0.5 * p), 0.0, math.cos(0.5 * p)])
    qx = np.array([math.sin(0.5 * r), 0.0, 0.0, math.cos(0.5 * r)])
    return quat_mul(quat_mul(qz, qy), qx)


def skew(v):
    """eigen_util.hpp:6."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


# ------------------------------------------------------------------ gait schedule (gait.cpp)
GAITS = {  # GaitDatabase::getGait, gait.cpp:732-767: (period, duty, offsets)
    "STAND": (0.5, 1.0, (0.0, 0.0, 0.0, 0.0)),
    "STATIC_WALK": (1.25, 0.8, (0.0, 0.5, 0.75, 0.25)),
    "WALKING_TROT": (0.5, 0.6, (0.0, 0.5, 0.5, 0.0)),
    "TROT": (0.5, 0.5, (0.0, 0.5, 0.5, 0.0)),
    "FLYING_TROT": (0.4, 0.4, (0.0, 0.5, 0.5, 0.0)),
    "PACE": (0.35, 0.5, (0.0, 0.5, 0.0, 0.5)),
    "BOUND": (0.4, 0.4, (0.0, 0.0, 0.5, 0.5)),
    "ROTARY_GALLOP": (0.4, 0.2, (0.0, 0.8571, 0.3571, 0.5)),
    "TRAVERSE_GALLOP": (0.5, 0.2, (0.0, 0.8571, 0.3571, 0.5)),
    "PRONK": (0.5, 0.5, (0.0, 0.0, 0.0, 0.0)),
}
GAIT_IDS = list(GAITS.keys())


def gait_update_sequence(period, duty, offset, dt, phase, steps, contact_state=None):
    """Gait::update_sequence, gait.cpp:67-112, strict IEEE-754 double evaluation of the source expression
    `fmod((phase + 1.0 - phase_offset_[leg]) + i * dt_ / T_, 1.0)` with i unsigned -> double and
    left-to-right (i * dt) / T.  Returns (contact uint8 [steps,4], swing_time float64 [steps,4])."""
    duty = [float(duty)] * 4 if np.isscalar(duty) else [float(d) for d in duty]
    contact = np.zeros((steps, NLEGS), dtype=np.uint8)
    swing = np.zeros((steps, NLEGS))
    early = [False] * NLEGS
    for i in range(steps):
        for leg in range(NLEGS):
            leg_phase = math.fmod((phase + 1.0 - offset[leg]) + float(i) * dt / period, 1.0)
            c = leg_phase < duty[leg]
            if i == 0 and contact_state is not None:
                if contact_state[leg] and (not c) and (leg_phase > ((duty[leg] + 1.0) / 2.0)):
                    early[leg] = True
            if c:
                early[leg] = False
            elif early[leg]:
                c = True
            t = (1 - leg_phase) * period if ((not c) and duty[leg] > 0) else 0.0
            contact[i, leg] = 1 if c else 0
            swing[i, leg] = t
    return contact, swing


def sequencer_phase(time_in_phase, period):
    """SimpleGaitSequencer::UpdateState, simple_gait_sequencer.cpp:71-75."""
    return math.fmod(time_in_phase / period, 1.0)


# ------------------------------------------------------------------ QP data exactly as mpc.cpp builds it
def get_R(psi, s=1.0):
    """MPC::GetR, mpc.cpp:46-58 (note: this is Rz(psi)^T scaled by s)."""
    c, sn = math.cos(psi), math.sin(psi)
    return np.array([[c * s, sn * s, 0.0], [-sn * s, c * s, 0.0], [0.0, 0.0, s]])


def get_A(psi, dt):
    """MPC::GetA, mpc.cpp:11-18."""
    A = np.eye(NX)
    A[3:6, 9:12] = np.eye(3) * dt
    A[11, 12] = -dt
    A[0:3, 6:9] = get_R(psi, dt)
    return A


def get_B(psi, m, inertia, r, contacts, dt):
    """MPC::GetB, mpc.cpp:20-44."""
    B = np.zeros((NX, NU))
    for f in range(NLEGS):
        B[9:12, 3 * f : 3 * f + 3] = np.eye(3) * (dt / m)
    R = get_R(psi, 1.0)
    I_rot = R @ inertia @ R.T  # MPC::GetIWorld, mpc.cpp:60-64
    I_inv = np.linalg.inv(I_rot)
    for f in range(NLEGS):
        if contacts[f]:
            B[6:9, 3 * f : 3 * f + 3] = I_inv @ skew(r[f]) * dt
    return B


def get_u_bounds(contact, fmin, fmax):
    """MPC::GetUBounds, mpc.cpp:94-115."""
    lb = np.zeros(NU)
    ub = np.zeros(NU)
    for f in range(NLEGS):
        if contact[f]:
            lb[3 * f : 3 * f + 3] = [-fmax, -fmax, fmin]
            ub[3 * f : 3 * f + 3] = [fmax, fmax, fmax]
    return lb, ub


def get_C(mu):
    """MPC::GetC, mpc.cpp:66-71 (acados 'D')."""
    D = np.zeros((NG, NU))
    for f in range(NLEGS):
        D[4 * f : 4 * f + 4, 3 * f : 3 * f + 3] = [[1, 0, -mu], [1, 0, mu], [0, 1, -mu], [0, 1, mu]]
    return D


def get_C_bounds():
    """MPC::GetCbounds, mpc.cpp:73-92."""
    lg = np.tile([-POS_UNBOUND, 0.0, -POS_UNBOUND, 0.0], NLEGS)
    ug = np.tile([0.0, POS_UNBOUND, 0.0, POS_UNBOUND], NLEGS)
    return lg, ug


def build_ocp_qp(N, rec, contacts, params):
    """MPC::GetWrenchSequence up to the ocp_qp_solve call, mpc.cpp:332-398 (+ constants of the ctor :160-213).
    Returns the stage-wise OCP-QP: A (one, identical for all k), B[k], Q diag, R scalar, q[k], x0, lbu/ubu[k], D, lg, ug."""
    d = unpack_problem(N, rec)
    dt = params["dt"]
    w = np.asarray(params["weights"], dtype=float)
    Q = np.concatenate([w, [0.0]])
    psi_avg = yaw_from_quaternion(d["quat"])
    yaw_before = psi_avg
    com = d["com"]
    A = get_A(psi_avg, dt)
    Bs, lbu, ubu, qs, xref = [], [], [], [], []
    for k in range(N + 1):
        desired_yaw = yaw_from_quaternion(d["des_quat"][k])
        diff = desired_yaw - yaw_before
        if diff < -math.pi:
            desired_yaw += 2 * math.pi
        elif diff > math.pi:
            desired_yaw -= 2 * math.pi
        yaw_before = desired_yaw
        if k < N:
            r = [d["foot_pos"][k][f] - (d["des_pos"][k] + com) for f in range(NLEGS)]
            Bs.append(get_B(desired_yaw, d["mass"], d["inertia"], r, contacts[k], dt))
            lb, ub = get_u_bounds(contacts[k], params["fmin"], params["fmax"])
            lbu.append(lb)
            ubu.append(ub)
        t = np.zeros(NX)
        t[0:3] = quaternion_to_euler(d["des_quat"][k])
        t[2] = desired_yaw
        t[3:6] = d["des_pos"][k] + com
        t[6:9] = d["ref_twist"][k]
        t[9:12] = d["ref_vel"][k]
        xref.append(t)
        qs.append(-(Q * t))
    x0 = np.zeros(NX)
    x0[0:3] = quaternion_to_euler(d["quat"])
    x0[3:6] = d["pos"] + com
    x0[6:9] = d["omega"]
    x0[9:12] = d["vel"]
    x0[12] = GRAVITY_CONSTANT
    lg, ug = get_C_bounds()
    return dict(
        N=N, A=A, B=Bs, Q=Q, R=params["alpha"], q=qs, xref=xref, x0=x0, lbu=lbu, ubu=ubu,
        D=get_C(params["mu"]), lg=lg, ug=ug, contacts=np.asarray(contacts)[:N],
    )


def condense(qp):
    """Exact dense condensing (what acados/HPIPM full condensing produces up to ordering):
    X = Phi x0 + Gamma U;  H = Gamma^T Qbar Gamma + alpha I;  g = Gamma^T Qbar (Phi x0 - Xref).
    Dense layout cross-check: quad_linear_matrices.cpp:78-114."""
    N, A = qp["N"], qp["A"]
    Gamma = np.zeros((NX * N, NU * N))
    Phi = np.zeros((NX * N, NX))
    Ap = np.eye(NX)
    for i in range(1, N + 1):
        Ap = A @ Ap
        Phi[(i - 1) * NX : i * NX] = Ap
    for j in range(N):
        blk = qp["B"][j]
        for i in range(j + 1, N + 1):
            Gamma[(i - 1) * NX : i * NX, j * NU : (j + 1) * NU] = blk
            blk = A @ blk
    Qbar = np.tile(qp["Q"], N)
    xref = np.concatenate(qp["xref"][1:])
    H = Gamma.T @ (Qbar[:, None] * Gamma) + qp["R"] * np.eye(NU * N)
    g = Gamma.T @ (Qbar * (Phi @ qp["x0"] - xref))
    # inequality rows: box (12N) then pyramids (16N), both block diagonal over stages
    Aineq = np.zeros((28 * N, NU * N))
    l = np.zeros(28 * N)
    u = np.zeros(28 * N)
    for k in range(N):
        Aineq[12 * k : 12 * k + 12, 12 * k : 12 * k + 12] = np.eye(12)
        l[12 * k : 12 * k + 12] = qp["lbu"][k]
        u[12 * k : 12 * k + 12] = qp["ubu"][k]
        r0 = 12 * N + 16 * k
        Aineq[r0 : r0 + 16, 12 * k : 12 * k + 12] = qp["D"]
        l[r0 : r0 + 16] = qp["lg"]
        u[r0 : r0 + 16] = qp["ug"]
    return dict(H=H, g=g, A=Aineq, l=l, u=u, Gamma=Gamma, Phi=Phi)


def rollout(qp, U):
    """x_{k+1} = A x_k + B_k u_k from x0; U is (N,12)."""
    X = np.zeros((qp["N"] + 1, NX))
    X[0] = qp["x0"]
    for k in range(qp["N"]):
        X[k + 1] = qp["A"] @ X[k] + qp["B"][k] @ U[k]
    return X


# ------------------------------------------------------------------ solver 1: dense primal-dual IPM (Mehrotra)
def solve_ipm_dense(H, g, A, l, u, tol=1e-11, max_iter=100):
    """Mehrotra predictor-corrector on  min 1/2 x'Hx + g'x  s.t. l <= Ax <= u.
    Rows with l == u are eliminated as equalities when they are simple pins (a single 1 in the row), which is
    the only equality form the MPC QP contains (swing-leg boxes, mpc.cpp:110-113)."""
    n = H.shape[0]
    eq = np.where((u - l) == 0.0)[0]
    pinned = np.zeros(n, dtype=bool)
    xpin = np.zeros(n)
    for r in eq:
        nz = np.nonzero(A[r])[0]
        assert len(nz) == 1 and A[r, nz[0]] == 1.0
        pinned[nz[0]] = True
        xpin[nz[0]] = l[r]
    free = ~pinned
    Hf = H[np.ix_(free, free)]
    gf = g[free] + H[np.ix_(free, pinned)] @ xpin[pinned]
    rows_u = np.where((u < INF_THRESHOLD) & ((u - l) != 0.0))[0]
    rows_l = np.where((l > -INF_THRESHOLD) & ((u - l) != 0.0))[0]
    G = np.vstack([A[rows_u][:, free], -A[rows_l][:, free]])
    h = np.concatenate([u[rows_u] - A[rows_u][:, pinned] @ xpin[pinned], -(l[rows_l] - A[rows_l][:, pinned] @ xpin[pinned])])
    keep = np.abs(G).sum(axis=1) > 0  # rows that only touch pinned variables are constant (0 <= 0)
    G, h = G[keep], h[keep]
    m = G.shape[0]
    nf = Hf.shape[0]
    x = np.zeros(nf)
    if nf == 0:
        xfull = xpin.copy()
        return xfull, 0
    s = np.maximum(h - G @ x, 1.0)
    z = np.ones(m)
    it = 0
    for it in range(1, max_iter + 1):
        rd = Hf @ x + gf + G.T @ z
        rp = G @ x + s - h
        mu = s @ z / max(m, 1)
        if max(np.abs(rd).max(), np.abs(rp).max() if m else 0.0, mu) < tol:
            break
        W = z / s
        K = Hf + G.T @ (W[:, None] * G)
        Lc = None
        reg = 0.0
        for _try in range(6):  # degenerate vertices (cone apex) make K numerically indefinite late in the solve
            try:
                Lc = np.linalg.cholesky(K + reg * np.eye(nf))
                break
            except np.linalg.LinAlgError:
                reg = max(10.0 * reg, 1e-12 * np.abs(np.diag(K)).max())
        if Lc is None or not np.all(np.isfinite(Lc)):
            break

        def kkt(rc):
            # eliminate ds = -rp - G dx ; dz = (-rc - z ds)/s
            t = (-rc + z * rp) / s
            rhs = -rd - G.T @ t
            dx = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
            ds = -rp - G @ dx
            dz = (-rc - z * ds) / s
            return dx, ds, dz

        dx, ds, dz = kkt(s * z)
        a_aff = _step(s, ds, z, dz, 1.0)
        mu_aff = (s + a_aff * ds) @ (z + a_aff * dz) / max(m, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        dx, ds, dz = kkt(s * z + ds * dz - sigma * mu)
        a = _step(s, ds, z, dz, 0.995)
        x += a * dx
        s += a * ds
        z += a * dz
    xfull = xpin.copy()
    xfull[free] = x
    return xfull, it


def _step(s, ds, z, dz, frac):
    a = 1.0
    neg = ds < 0
    if neg.any():
        a = min(a, frac * np.min(-s[neg] / ds[neg]))
    neg = dz < 0
    if neg.any():
        a = min(a, frac * np.min(-z[neg] / dz[neg]))
    return a


# ------------------------------------------------------------------ solver 2: OSQP-style ADMM + active-set polish
def solve_admm_dense(H, g, A, l, u, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3, max_iter=4000,
                     adaptive_rho=True, check_every=25, polish=True, polish_refine=3, return_info=False):
    """Restated OSQP 0.6.3 algorithm (Stellato et al. 2020, Alg. 1) without problem scaling:
    K = H + sigma I + A' diag(rho_vec) A, rho_vec = 1e3*rho on equality rows; relaxed x/z updates; residual-ratio
    rho adaptation; polish on the active set identified from the duals."""
    n, m = H.shape[0], A.shape[0]
    lo = np.where(l <= -INF_THRESHOLD, -np.inf, l)
    hi = np.where(u >= INF_THRESHOLD, np.inf, u)
    is_eq = (hi - lo) == 0.0

    def rho_vec(r):
        v = np.full(m, r)
        v[is_eq] = 1e3 * r
        v[np.isinf(lo) & np.isinf(hi)] = 1e-6
        return v

    rv = rho_vec(rho)
    x = np.zeros(n)
    z = np.zeros(m)
    y = np.zeros(m)
    Lc = np.linalg.cholesky(H + sigma * np.eye(n) + A.T @ (rv[:, None] * A))
    n_fact = 1
    it = 0
    status = "max_iter"
    for it in range(1, max_iter + 1):
        rhs = sigma * x - g + A.T @ (rv * z - y)
        xt = np.linalg.solve(Lc.T, np.linalg.solve(Lc, rhs))
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        zr = alpha * zt + (1 - alpha) * z
        znew = np.clip(zr + y / rv, lo, hi)
        y = y + rv * (zr - znew)
        z = znew
        if it % check_every == 0 or it == max_iter:
            Ax = A @ x
            Hx = H @ x
            Aty = A.T @ y
            rp = np.abs(Ax - z).max()
            rd = np.abs(Hx + g + Aty).max()
            ep = eps_abs + eps_rel * max(np.abs(Ax).max(), np.abs(z).max())
            ed = eps_abs + eps_rel * max(np.abs(Hx).max(), np.abs(Aty).max(), np.abs(g).max())
            if rp <= ep and rd <= ed:
                status = "solved"
                break
            if adaptive_rho:
                num = rp / max(np.abs(Ax).max(), np.abs(z).max(), 1e-30)
                den = rd / max(np.abs(Hx).max(), np.abs(Aty).max(), np.abs(g).max(), 1e-30)
                new_rho = float(np.clip(rho * math.sqrt(num / max(den, 1e-30)), 1e-6, 1e6))
                if new_rho > 5 * rho or new_rho < rho / 5:
                    rho = new_rho
                    rv = rho_vec(rho)
                    Lc = np.linalg.cholesky(H + sigma * np.eye(n) + A.T @ (rv[:, None] * A))
                    n_fact += 1
    polished = False
    if polish:
        xp, yp, ok = polish_active_set(H, g, A, lo, hi, x, y, refine=polish_refine)
        if ok:
            x, y, polished = xp, yp, True
    if return_info:
        return x, y, dict(iters=it, status=status, polished=polished, n_fact=n_fact, rho=rho)
    return x, y


def polish_active_set(H, g, A, lo, hi, x, y, delta=1e-9, refine=5, rounds=8):
    """OSQP-style polish: solve the equality-constrained QP on the active rows (y<0 -> lower, y>0 -> upper) with a
    regularised KKT system + iterative refinement; then repair the active set (primal-dual active-set sweeps)
    until primal feasible and duals sign-correct."""
    n, m = H.shape[0], A.shape[0]
    Ax = A @ x
    act_l = (y < 0) & np.isfinite(lo)
    act_u = (y > 0) & np.isfinite(hi)
    eqr = (hi - lo) == 0.0
    act_l = act_l | eqr
    act_u = act_u & ~eqr
    for _ in range(rounds):
        rows = np.where(act_l | act_u)[0]
        Aa = A[rows]
        b = np.where(act_l[rows], lo[rows], hi[rows])
        na = len(rows)
        K = np.block([[H, Aa.T], [Aa, np.zeros((na, na))]])
        Kr = K + np.diag(np.concatenate([np.full(n, delta), np.full(na, -delta)]))
        rhs = np.concatenate([-g, b])
        try:
            import scipy.linalg as sla

            lu = sla.lu_factor(Kr)
            sol = sla.lu_solve(lu, rhs)
            for _r in range(refine):
                sol = sol + sla.lu_solve(lu, rhs - K @ sol)
        except Exception:
            return x, y, False
        xn = sol[:n]
        yn = np.zeros(m)
        yn[rows] = sol[n:]
        Axn = A @ xn
        viol_l = (Axn < lo - 1e-9) & ~act_l
        viol_u = (Axn > hi + 1e-9) & ~act_u
        wrong_l = act_l & ~eqr & (yn > 1e-9)
        wrong_u = act_u & (yn < -1e-9)
        if not (viol_l.any() or viol_u.any() or wrong_l.any() or wrong_u.any()):
            return xn, yn, True
        act_l = (act_l | viol_l) & ~wrong_l
        act_u = (act_u | viol_u) & ~wrong_u
    return x, y, False


# ------------------------------------------------------------------ KKT checker (acados 4-tuple analogue)
def kkt_residuals(qp, U, lam_box=None, lam_g=None):
    """Independent KKT check in the *sparse* OCP form (the quantity ocp_qp_inf_norm_residuals would print,
    mpc.cpp:430-433): returns (stat, dyn, ineq, comp) inf-norms.  X is re-rolled from U, the dynamics
    multipliers are the exact adjoint, so `dyn` measures only round-off; if duals are not given, the minimal
    sign-feasible multipliers are reconstructed per foot from the reduced gradient (duals are non-unique for
    swing legs / at the cone apex)."""
    N = qp["N"]
    U = np.asarray(U).reshape(N, NU)
    X = rollout(qp, U)
    Q, A = qp["Q"], qp["A"]
    lamx = np.zeros((N + 2, NX))
    for k in range(N, 0, -1):
        lamx[k] = Q * X[k] + qp["q"][k] + (A.T @ lamx[k + 1] if k < N else 0.0)
    grad = np.array([qp["R"] * U[k] + qp["B"][k].T @ lamx[k + 1] for k in range(N)])  # dL/du without ineq duals
    D, lg, ug = qp["D"], qp["lg"], qp["ug"]
    stat = 0.0
    ineq = 0.0
    comp = 0.0
    for k in range(N):
        lb, ub = qp["lbu"][k], qp["ubu"][k]
        gk = D @ U[k]
        ineq = max(ineq, np.max(np.maximum(lb - U[k], 0)), np.max(np.maximum(U[k] - ub, 0)))
        fin_l, fin_u = lg > -INF_THRESHOLD, ug < INF_THRESHOLD
        ineq = max(ineq, np.max(np.maximum((lg - gk)[fin_l], 0)), np.max(np.maximum((gk - ug)[fin_u], 0)))
        if lam_box is not None:
            yb = np.asarray(lam_box).reshape(N, NU)[k]
            yg = np.asarray(lam_g).reshape(N, NG)[k]
        else:
            yb, yg = _min_duals(grad[k], U[k], lb, ub, D, lg, ug, qp["contacts"][k])
        r = grad[k] + yb + D.T @ yg
        pinned = (ub - lb) == 0.0
        stat = max(stat, np.abs(r[~pinned]).max() if (~pinned).any() else 0.0)
        # complementarity: y+ * (ub - u), y- * (u - lb)
        ypos, yneg = np.maximum(yb, 0), np.maximum(-yb, 0)
        comp = max(comp, np.abs(ypos * (ub - U[k]))[~pinned].max(initial=0.0), np.abs(yneg * (U[k] - lb))[~pinned].max(initial=0.0))
        gpos, gneg = np.maximum(yg, 0), np.maximum(-yg, 0)
        comp = max(comp, np.abs(gpos * (ug - gk))[fin_u].max(initial=0.0), np.abs(gneg * (gk - lg))[fin_l].max(initial=0.0))
        # sign feasibility of duals on one-sided rows counts as stationarity error
        stat = max(stat, np.abs(gpos[~fin_u]).max(initial=0.0), np.abs(gneg[~fin_l]).max(initial=0.0))
    Xr = rollout(qp, U)
    dyn = float(np.abs(Xr - X).max())
    return float(stat), dyn, float(ineq), float(comp)


def _min_duals(grad, u, lb, ub, D, lg, ug, contact, tol=1e-7):
    """Reconstruct multipliers for one stage from the gradient: non-negative least squares over the
    constraints that are active within `tol` (scaled)."""
    from scipy.optimize import nnls

    yb = np.zeros(NU)
    yg = np.zeros(NG)
    for f in range(NLEGS):
        sl = slice(3 * f, 3 * f + 3)
        if not contact[f]:
            yb[sl] = -grad[sl]
            continue
        cols = []
        tags = []
        uf = u[sl]
        scale = max(1.0, np.abs(uf).max())
        for i in range(3):
            e = np.zeros(3)
            e[i] = 1.0
            if ub[3 * f + i] - uf[i] <= tol * scale:
                cols.append(e)
                tags.append(("b", i, +1))
            if uf[i] - lb[3 * f + i] <= tol * scale:
                cols.append(-e)
                tags.append(("b", i, -1))
        Df = D[4 * f : 4 * f + 4, sl]
        gv = Df @ uf
        for r in range(4):
            if ug[4 * f + r] < INF_THRESHOLD and ug[4 * f + r] - gv[r] <= tol * scale:
                cols.append(Df[r])
                tags.append(("g", r, +1))
            if lg[4 * f + r] > -INF_THRESHOLD and gv[r] - lg[4 * f + r] <= tol * scale:
                cols.append(-Df[r])
                tags.append(("g", r, -1))
        if cols:
            M = np.array(cols).T
            coef, _ = nnls(M, -grad[sl])
            for c, (kind, idx, sg) in zip(coef, tags):
                if kind == "b":
                    yb[3 * f + idx] += sg * c
                else:
                    yg[4 * f + idx] += sg * c
    return yb, yg


def solve_reference_qp(N, rec, contacts, params, method="ipm"):
    """Full oracle pipeline for one problem: build (mpc.cpp) -> condense -> solve -> unpack (mpc.cpp:457-467)."""
    qp = build_ocp_qp(N, rec, contacts, params)
    dq = condense(qp)
    if method == "ipm":
        u, _ = solve_ipm_dense(dq["H"], dq["g"], dq["A"], dq["l"], dq["u"])
    else:
        u, _ = solve_admm_dense(dq["H"], dq["g"], dq["A"], dq["l"], dq["u"])
    U = u.reshape(N, NU)
    X = rollout(qp, U)
    return U, X, qp, dq


# ------------------------------------------------------------------ solver 3: sparse OCP-QP, Riccati-based Mehrotra IPM
def _foot_G(mu):
    return np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1, 0, -mu], [-1, 0, -mu],
                     [0, 1, -mu], [0, -1, -mu]], float)


def _chol_solve(L, b):
    import scipy.linalg as sla

    return sla.solve_triangular(L.T, sla.solve_triangular(L, b, lower=True), lower=False)


def solve_riccati_ipm(qp, params, tol=1e-10, max_iter=60, return_info=False):
    """HPIPM-style solve of the stage-wise QP of build_ocp_qp: Mehrotra predictor-corrector, every Newton system solved
    by a Riccati recursion on (A, B_k, Q, R~_k = alpha I + G'WG).  States are the full 13-vector (the g state included),
    so this is independent of the structure exploitation in the CUDA kernel."""
    N, A, Q = qp["N"], qp["A"], np.diag(qp["Q"])
    mu_f, fmin, fmax, alpha = params["mu"], params["fmin"], params["fmax"], qp["R"]
    Gf = _foot_G(mu_f)
    hf = np.array([fmax, fmax, fmax, fmax, fmax, -fmin, 0, 0, 0, 0], float)
    on = np.asarray(qp["contacts"]).astype(bool)
    G = np.zeros((N, 40, 12))
    for f in range(4):
        G[:, 10 * f:10 * f + 10, 3 * f:3 * f + 3] = Gf
    act = np.repeat(on, 10, axis=1)  # [N,40] rows that exist
    h = np.tile(hf, 4)
    U = np.zeros((N, 12))
    s = np.where(act, np.maximum(h, 1.0), 1.0)
    z = np.where(act, 1.0, 0.0)
    m = max(int(act.sum()), 1)
    pin = ~np.repeat(on, 3, axis=1)  # swing inputs pinned to zero
    it = 0
    for it in range(max_iter + 1):
        X = rollout(qp, U)
        lam = np.zeros((N + 2, NX))
        for k in range(N, 0, -1):
            lam[k] = qp["Q"] * X[k] + qp["q"][k] + (A.T @ lam[k + 1] if k < N else 0.0)
        rd = np.array([alpha * U[k] + qp["B"][k].T @ lam[k + 1] + G[k].T @ z[k] for k in range(N)])
        rd[pin] = 0.0
        rp = np.where(act, np.einsum("kij,kj->ki", G, U) + s - h, 0.0)
        mu_c = float((s * z)[act].sum()) / m
        res = max(np.abs(rd).max(), np.abs(rp).max(), float((s * z)[act].max(initial=0.0)))
        if res <= tol or not act.any():
            break
        W = np.where(act, z / s, 0.0)
        # backward Riccati (matrices)
        Pm = Q.copy()
        Ks, Hinv = [None] * N, [None] * N
        for k in range(N - 1, -1, -1):
            B = qp["B"][k]
            Rt = alpha * np.eye(12) + G[k].T @ (W[k][:, None] * G[k])
            Huu = Rt + B.T @ Pm @ B
            Hux = B.T @ Pm @ A
            p_ = pin[k]
            Huu[p_, :] = 0.0
            Huu[:, p_] = 0.0
            Huu[p_, p_] = 1.0
            Hux[p_, :] = 0.0
            Hinv[k] = np.linalg.cholesky(Huu)  # Cholesky factor; explicit inverses lose the endgame accuracy
            Ks[k] = -_chol_solve(Hinv[k], Hux)
            Pm = Q + A.T @ Pm @ A + Hux.T @ Ks[k]
            Pm = 0.5 * (Pm + Pm.T)

        def solve(rc):
            t = np.where(act, (z * rp - rc) / s, 0.0)
            rt = rd + np.einsum("kij,ki->kj", G, t)
            rt[pin] = 0.0
            p = np.zeros(NX)
            kk = np.zeros((N, 12))
            for k in range(N - 1, -1, -1):
                hu = rt[k] + qp["B"][k].T @ p
                hu[pin[k]] = 0.0
                kk[k] = -_chol_solve(Hinv[k], hu)
                p = A.T @ p + Ks[k].T @ hu
            dx = np.zeros(NX)
            dU = np.zeros((N, 12))
            for k in range(N):
                dU[k] = Ks[k] @ dx + kk[k]
                dU[k][pin[k]] = 0.0
                dx = A @ dx + qp["B"][k] @ dU[k]
            ds = np.where(act, -rp - np.einsum("kij,kj->ki", G, dU), 0.0)
            dz = np.where(act, (-rc - z * ds) / s, 0.0)
            return dU, ds, dz

        def maxstep(ds, dz):
            a = 1.0
            msk = act & (ds < 0)
            if msk.any():
                a = min(a, float(np.min(-s[msk] / ds[msk])))
            msk = act & (dz < 0)
            if msk.any():
                a = min(a, float(np.min(-z[msk] / dz[msk])))
            return a

        dU, ds, dz = solve(s * z)
        a = maxstep(ds, dz)
        mu_aff = float(((s + a * ds) * (z + a * dz))[act].sum()) / m
        sig = (mu_aff / max(mu_c, 1e-300)) ** 3
        dU, ds, dz = solve(s * z + ds * dz - sig * mu_c)
        a = min(1.0, 0.995 * maxstep(ds, dz))
        U = U + a * dU
        s = np.where(act, s + a * ds, s)
        z = np.where(act, z + a * dz, z)
    U[pin] = 0.0
    if return_info:
        return U, dict(iters=it, res=res, z=z)
    return U
