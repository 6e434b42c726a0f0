"""Batched floating-base rigid-body dynamics of the Go2 (numpy, host side) - the input generator of the WBC QP.

In the reference these terms come from Pinocchio through ARC-OPT's `RobotModelPinocchio`
(ws/src/controllers/src/mit_controller/wbc_arc_opt.cpp:30-47,236-241); the WBC QP kernel takes them as inputs
("QP solve only; RBD terms are inputs").  Physical parameters below are the numbers of
ws/src/common/model/urdf/go2/urdf/go2_description.urdf (link inertials, joint origins/axes, effort limits); all joint
frames have rpy = 0, abad joints rotate about x, shoulder and knee joints about y.

Generalised velocity convention used throughout this package: nu = [v_base (world), omega_base (world), qdot(12)],
leg order fl, fr, bl, br (quad_model_symbolic.hpp:163), joint order abad, shoulder, knee.
Equation of motion:  M(q) nu_dot + h(q, nu) = S^T tau + sum_f J_f^T f_f,   foot velocity = J_f nu.
"""
from __future__ import annotations

import numpy as np

GRAVITY = 9.81
NV = 18
LEGS = ("fl", "fr", "bl", "br")
SX = np.array([1.0, 1.0, -1.0, -1.0])  # front / back
SY = np.array([1.0, -1.0, 1.0, -1.0])  # left / right
TAU_MAX = np.array([23.7, 23.7, 45.43] * 4)  # <limit effort=...> of *_abad, *_shoulder, *_knee

# base_link (+ the two 1 g head links, fixed)
BASE_MASS = 6.921
BASE_COM = np.array([0.021112, 0.0, -0.005366])
BASE_INERTIA = np.array([[0.02448, 0.00012166, 0.0014849], [0.00012166, 0.098077, -3.12e-05], [0.0014849, -3.12e-05, 0.107]])
HEADS = [(0.001, np.array([0.285, 0.0, 0.01]), 9.6e-06), (0.001, np.array([0.293, 0.0, -0.06]), 9.6e-06)]

# joint origins in the parent frame (front-left values; x mirrors with SX on the abad joint, y with SY)
ABAD_ORIGIN = np.array([0.1934, 0.0465, 0.0])
SHOULDER_ORIGIN = np.array([0.0, 0.0955, 0.0])
KNEE_ORIGIN = np.array([0.0, 0.0, -0.213])
FOOT_ORIGIN = np.array([0.0, 0.0, -0.213])  # *_contact_frame
FOOT_MASS, FOOT_INERTIA = 0.04, 9.6e-06


def _sym(ixx, ixy, ixz, iyy, iyz, izz):
    return np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])


def _leg_links(leg):
    """(mass, com, inertia) of the HAA, HFE, KFE links of leg `leg` with the URDF's mirrored signs."""
    sx, sy = SX[leg], SY[leg]
    haa = (0.678, np.array([-0.0054 * sx, 0.00194 * sy, -0.000105]),
           _sym(0.00048, -3.01e-06 * sx * sy, 1.11e-06 * sx, 0.000884, -1.42e-06 * sy, 0.000596))
    hfe = (1.152, np.array([-0.00374, -0.0223 * sy, -0.0327]),
           _sym(0.00584, 8.72e-05 * sy, -0.000289, 0.0058, 0.000808 * sy, 0.00103))
    kfe = (0.154, np.array([0.00548, -0.000975 * sy, -0.115]),
           _sym(0.00108, 3.4e-07 * sy, 1.72e-05, 0.0011, 8.28e-06 * sy, 3.29e-05))
    return haa, hfe, kfe


def quat_to_rot(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _rot_axis(axis, ang):
    c, s = np.cos(ang), np.sin(ang)
    R = np.zeros(ang.shape + (3, 3))
    if axis == 0:
        R[..., 0, 0] = 1; R[..., 1, 1] = c; R[..., 1, 2] = -s; R[..., 2, 1] = s; R[..., 2, 2] = c
    else:
        R[..., 1, 1] = 1; R[..., 0, 0] = c; R[..., 0, 2] = s; R[..., 2, 0] = -s; R[..., 2, 2] = c
    return R


def _skew(v):
    S = np.zeros(v.shape[:-1] + (3, 3))
    S[..., 0, 1] = -v[..., 2]; S[..., 0, 2] = v[..., 1]; S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]; S[..., 2, 0] = -v[..., 1]; S[..., 2, 1] = v[..., 0]
    return S


def dynamics(pos, quat, q, v_lin, omega, qd):
    """Batched M [n,18,18], h [n,18], foot positions [n,4,3], foot Jacobians J [n,4,3,18], Jdot*nu [n,4,3],
    foot velocities [n,4,3].  Inputs: pos [n,3], quat [n,4] (x,y,z,w), q [n,12], v_lin/omega [n,3] world, qd [n,12]."""
    n = pos.shape[0]
    Rb = quat_to_rot(quat)
    M = np.zeros((n, NV, NV))
    h = np.zeros((n, NV))
    gvec = np.array([0.0, 0.0, GRAVITY])

    def add_body(mass, com_w, Iw, Jv, Jw, w_body, a_bias, al_bias):
        """Accumulate one rigid body: com position (world), inertia (world), com Jacobians, angular velocity, bias accels."""
        nonlocal M, h
        M += mass * np.einsum("nij,nik->njk", Jv, Jv) + np.einsum("nij,nil,nlk->njk", Jw, Iw, Jw)
        Iw_w = np.einsum("nij,nj->ni", Iw, w_body)
        wrench_f = mass * (a_bias + gvec)
        wrench_t = np.einsum("nij,nj->ni", Iw, al_bias) + np.cross(w_body, Iw_w)
        h += np.einsum("nij,ni->nj", Jv, wrench_f) + np.einsum("nij,ni->nj", Jw, wrench_t)

    def point_kin(r, Jw_link, origin_Jv, w_link, origin_abias, al_bias):
        """Jacobian and bias acceleration of a point at offset r (world) from a link origin."""
        Jv = origin_Jv - np.einsum("nij,njk->nik", _skew(r), Jw_link)
        ab = origin_abias + np.cross(al_bias, r) + np.cross(w_link, np.cross(w_link, r))
        return Jv, ab

    # base
    Jv0 = np.zeros((n, 3, NV)); Jv0[:, :, 0:3] = np.eye(3)
    Jw0 = np.zeros((n, 3, NV)); Jw0[:, :, 3:6] = np.eye(3)
    zero3 = np.zeros((n, 3))
    for mass, com, inertia in [(BASE_MASS, BASE_COM, BASE_INERTIA)] + [(m, c, np.eye(3) * i) for m, c, i in HEADS]:
        r = np.einsum("nij,j->ni", Rb, com)
        Jv, ab = point_kin(r, Jw0, Jv0, omega, zero3, zero3)
        Iw = np.einsum("nij,jk,nlk->nil", Rb, inertia, Rb)
        add_body(mass, pos + r, Iw, Jv, Jw0, omega, ab, zero3)

    foot_pos = np.zeros((n, 4, 3)); foot_J = np.zeros((n, 4, 3, NV)); foot_ab = np.zeros((n, 4, 3)); foot_vel = np.zeros((n, 4, 3))
    nu = np.concatenate([v_lin, omega, qd], axis=1)
    for leg in range(4):
        links = _leg_links(leg)
        origins = [ABAD_ORIGIN * np.array([SX[leg], SY[leg], 1.0]), SHOULDER_ORIGIN * np.array([1.0, SY[leg], 1.0]), KNEE_ORIGIN]
        axes = [0, 1, 1]
        # running state of the parent link frame
        R = Rb; p_o = pos; Jv_o = Jv0; Jw = Jw0; w = omega; ab_o = zero3; al = zero3
        for j in range(3):
            col = 6 + 3 * leg + j
            r = np.einsum("nij,j->ni", R, origins[j])                     # parent origin -> joint origin
            Jv_o, ab_o = point_kin(r, Jw, Jv_o, w, ab_o, al)
            p_o = p_o + r
            axis_w = R[:, :, axes[j]]                                       # joint axis in world (parent frame axis)
            Jw = Jw.copy(); Jw[:, :, col] = axis_w
            al = al + np.cross(w, axis_w * qd[:, 3 * leg + j][:, None])     # bias angular acceleration
            w = w + axis_w * qd[:, 3 * leg + j][:, None]
            R = np.einsum("nij,njk->nik", R, _rot_axis(axes[j], q[:, 3 * leg + j]))
            mass, com, inertia = links[j]
            bodies = [(mass, com, inertia)]
            if j == 2:
                bodies.append((FOOT_MASS, FOOT_ORIGIN, np.eye(3) * FOOT_INERTIA))
            for m_b, c_b, I_b in bodies:
                rc = np.einsum("nij,j->ni", R, c_b)
                Jv, ab = point_kin(rc, Jw, Jv_o, w, ab_o, al)
                Iw = np.einsum("nij,jk,nlk->nil", R, I_b, R)
                add_body(m_b, p_o + rc, Iw, Jv, Jw, w, ab, al)
        rf = np.einsum("nij,j->ni", R, FOOT_ORIGIN)
        Jf, abf = point_kin(rf, Jw, Jv_o, w, ab_o, al)
        foot_pos[:, leg] = p_o + rf
        foot_J[:, leg] = Jf
        foot_ab[:, leg] = abf
        foot_vel[:, leg] = np.einsum("nij,nj->ni", Jf, nu)
    M = 0.5 * (M + np.transpose(M, (0, 2, 1)))
    return dict(M=M, h=h, foot_pos=foot_pos, foot_J=foot_J, foot_Jdnu=foot_ab, foot_vel=foot_vel)


def nominal_joint_angles():
    """Stand pose of the experiment script (solver_experiments.py:126-139): (0.126, 0.61, -1.22), abad mirrored."""
    return np.array([0.126 * SY[l] if j == 0 else (0.61 if j == 1 else -1.22) for l in range(4) for j in range(3)])
