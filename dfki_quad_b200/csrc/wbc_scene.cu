// Kernel K6: the whole-body-control QP assembled ON THE DEVICE from (base state, joint state, contacts, task references), one thread
// per robot - what `scene->update()` does on the host in the reference before `scene->solve()` (wbc_arc_opt.cpp:292-297):
//   rigid-body terms of the Go2 (Pinocchio in the reference, wbc_arc_opt.cpp:30-47,236-241): M(q), h(q, nu), foot Jacobians, Jdot nu
//   reduced TSID scene (ARC-OPT AccelerationSceneReducedTSID, task set and weights wbc_arc_opt.cpp:78-123,
//   mit_controller_sim_go2.yaml:62-72): x = [nu_dot(18), f(4x3)], 18 equalities, 28 two-sided rows
// so that the host -> device traffic of a WBC solve drops from the assembled QP (19.5 KB) to the robot's state (576 B).
// Physical parameters: ws/src/common/model/urdf/go2/urdf/go2_description.urdf (link inertials, joint origins and axes, effort
// limits); all joint frames have rpy = 0, abad joints rotate about x, hip and knee joints about y; fixed topology, so every body's
// Jacobian only has the 6 base columns and the 3 columns of its own leg.
// Conventions: nu = [v_base (world), omega_base (world), qdot(12)], leg order fl, fr, bl, br, joint order abad, hip, knee;
// M nu_dot + h = S' tau + sum_f J_f' f_f.  Output layouts are those of b200qp_wbc_solve_batch (column-major H and A).
#include "common.cuh"

namespace b200qp {

namespace {
constexpr int NV = 18, NQP = 30, NEQ = 18, NIN = 28, NC = NEQ + NIN;
constexpr double GRAV = 9.81;
constexpr double WINF_ = 1.0e20;

struct V3 { double x, y, z; };
__device__ inline V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
__device__ inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ inline V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ inline double comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M3 { double m[3][3]; };
__device__ inline V3 mul(const M3 &A, V3 v) {
  return v3(A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z);
}
__device__ inline V3 col(const M3 &A, int j) { return v3(A.m[0][j], A.m[1][j], A.m[2][j]); }
__device__ inline M3 mul(const M3 &A, const M3 &B) {
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
  return C;
}
// R I R'
__device__ inline M3 rot_inertia(const M3 &R, const M3 &I) {
  M3 T = mul(R, I), C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C.m[i][j] = T.m[i][0] * R.m[j][0] + T.m[i][1] * R.m[j][1] + T.m[i][2] * R.m[j][2];
  return C;
}
__device__ inline M3 sym(double ixx, double ixy, double ixz, double iyy, double iyz, double izz) {
  M3 I = {{{ixx, ixy, ixz}, {ixy, iyy, iyz}, {ixz, iyz, izz}}};
  return I;
}
__device__ inline M3 diag3(double d) { return sym(d, 0.0, 0.0, d, 0.0, d); }
__device__ inline M3 rot_axis(int axis, double ang) {  // joint rotation about x (axis 0) or y (axis 1)
  const double c = cos(ang), s = sin(ang);
  M3 R = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
  if (axis == 0) {
    R.m[0][0] = 1; R.m[1][1] = c; R.m[1][2] = -s; R.m[2][1] = s; R.m[2][2] = c;
  } else {
    R.m[1][1] = 1; R.m[0][0] = c; R.m[0][2] = s; R.m[2][0] = -s; R.m[2][2] = c;
  }
  return R;
}

// Jacobians of a point / a link in the 9 columns that can be non-zero for leg `leg`: [v_base(3), omega_base(3), qdot_leg(3)]
struct Jac9 { double v[3][9]; };

// Jacobian and bias acceleration of a point at offset r (world) from a link origin
__device__ inline void point_kin(V3 r, const Jac9 &Jw, const Jac9 &Jo, V3 w, V3 abo, V3 al, Jac9 &Jv, V3 &ab) {
  for (int c = 0; c < 9; ++c) {  // Jv = Jo - skew(r) Jw
    const V3 jw = v3(Jw.v[0][c], Jw.v[1][c], Jw.v[2][c]);
    const V3 t = cross(r, jw);
    Jv.v[0][c] = Jo.v[0][c] - t.x;
    Jv.v[1][c] = Jo.v[1][c] - t.y;
    Jv.v[2][c] = Jo.v[2][c] - t.z;
  }
  ab = abo + cross(al, r) + cross(w, cross(w, r));
}

__device__ inline int gcol(int leg, int c) { return c < 6 ? c : 6 + 3 * leg + (c - 6); }

// accumulate one rigid body into M and h
__device__ inline void add_body(double *M, double *h, int leg, double mass, const M3 &Iw, const Jac9 &Jv, const Jac9 &Jw, V3 w, V3 ab, V3 al) {
  for (int a = 0; a < 9; ++a) {
    const V3 ja = v3(Jw.v[0][a], Jw.v[1][a], Jw.v[2][a]);
    const V3 Ija = mul(Iw, ja);  // Iw symmetric: Jw' Iw Jw
    const int ga = gcol(leg, a);
    for (int b = 0; b < 9; ++b) {
      const double lin = Jv.v[0][a] * Jv.v[0][b] + Jv.v[1][a] * Jv.v[1][b] + Jv.v[2][a] * Jv.v[2][b];
      const double ang = Ija.x * Jw.v[0][b] + Ija.y * Jw.v[1][b] + Ija.z * Jw.v[2][b];
      M[ga * NV + gcol(leg, b)] += mass * lin + ang;
    }
  }
  const V3 wf = mass * (ab + v3(0.0, 0.0, GRAV));
  const V3 wt = mul(Iw, al) + cross(w, mul(Iw, w));
  for (int a = 0; a < 9; ++a)
    h[gcol(leg, a)] += Jv.v[0][a] * wf.x + Jv.v[1][a] * wf.y + Jv.v[2][a] * wf.z + Jw.v[0][a] * wt.x + Jw.v[1][a] * wt.y + Jw.v[2][a] * wt.z;
}
}  // namespace

// mirror of b200qp_wbc_scene_config (include/b200qp.h)
struct WbcSceneConfig {
  double mu, com_pose_weight[6], foot_force_weight[3], foot_pose_weight[3], hessian_regularizer, tau_max[12];
};

// scene_in [n][72]: quat xyzw 0:4, pos 4:7, v 7:10, omega 10:13, q 13:25, qd 25:37, a_body_des 37:43, a_foot_des 43:55, f_ref 55:67,
// contact bit mask 67.  Outputs as b200qp_wbc_solve_batch takes them; feet [n][24] = foot positions and velocities (optional).
__global__ void __launch_bounds__(64) wbc_scene_kernel(int n, WbcSceneConfig C, const double *__restrict__ scene_in, double *__restrict__ Hout,
                                                       double *__restrict__ gout, double *__restrict__ Aout, double *__restrict__ loout,
                                                       double *__restrict__ hiout, double *__restrict__ feet) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double *S = scene_in + (size_t)p * 72;
  M3 Rb;
  quat_to_rot(S, Rb.m);
  const V3 pos = v3(S[4], S[5], S[6]), vlin = v3(S[7], S[8], S[9]), omega = v3(S[10], S[11], S[12]);
  const double *q = S + 13, *qd = S + 25;
  const unsigned cmask = (unsigned)S[67];
  double M[NV * NV], h[NV];
  for (int i = 0; i < NV * NV; ++i) M[i] = 0.0;
  for (int i = 0; i < NV; ++i) h[i] = 0.0;
  const V3 zero = v3(0.0, 0.0, 0.0);
  Jac9 Jv0, Jw0;  // base origin: Jv = [I 0 0], Jw = [0 I 0]
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 9; ++c) {
      Jv0.v[r][c] = (c == r) ? 1.0 : 0.0;
      Jw0.v[r][c] = (c == 3 + r) ? 1.0 : 0.0;
    }
  // base_link and the two 1 g head links (fixed to the base); the leg columns of these bodies are zero, any `leg` index serves
  {
    const double bm[3] = {6.921, 0.001, 0.001};
    const V3 bc[3] = {v3(0.021112, 0.0, -0.005366), v3(0.285, 0.0, 0.01), v3(0.293, 0.0, -0.06)};
    const M3 bi[3] = {sym(0.02448, 0.00012166, 0.0014849, 0.098077, -3.12e-05, 0.107), diag3(9.6e-06), diag3(9.6e-06)};
    for (int b = 0; b < 3; ++b) {
      const V3 r = mul(Rb, bc[b]);
      Jac9 Jv;
      V3 ab;
      point_kin(r, Jw0, Jv0, omega, zero, zero, Jv, ab);
      add_body(M, h, 0, bm[b], rot_inertia(Rb, bi[b]), Jv, Jw0, omega, ab, zero);
    }
  }
  double fJ[4][3][NV], fab[4][3], fpos[4][3], fvel[4][3];
  for (int leg = 0; leg < 4; ++leg) {
    const double sx = (leg < 2) ? 1.0 : -1.0, sy = (leg % 2 == 0) ? 1.0 : -1.0;
    const double lm[3] = {0.678, 1.152, 0.154};
    const V3 lc[3] = {v3(-0.0054 * sx, 0.00194 * sy, -0.000105), v3(-0.00374, -0.0223 * sy, -0.0327), v3(0.00548, -0.000975 * sy, -0.115)};
    const M3 li[3] = {sym(0.00048, -3.01e-06 * sx * sy, 1.11e-06 * sx, 0.000884, -1.42e-06 * sy, 0.000596),
                      sym(0.00584, 8.72e-05 * sy, -0.000289, 0.0058, 0.000808 * sy, 0.00103),
                      sym(0.00108, 3.4e-07 * sy, 1.72e-05, 0.0011, 8.28e-06 * sy, 3.29e-05)};
    const V3 org[3] = {v3(0.1934 * sx, 0.0465 * sy, 0.0), v3(0.0, 0.0955 * sy, 0.0), v3(0.0, 0.0, -0.213)};
    const V3 foot_org = v3(0.0, 0.0, -0.213);
    M3 R = Rb;
    V3 po = pos, w = omega, abo = zero, al = zero;
    Jac9 Jo = Jv0, Jw = Jw0;
    for (int j = 0; j < 3; ++j) {
      const int axis = (j == 0) ? 0 : 1;
      const double qdj = qd[3 * leg + j];
      const V3 r = mul(R, org[j]);  // parent origin -> joint origin
      Jac9 Jn;
      V3 abn;
      point_kin(r, Jw, Jo, w, abo, al, Jn, abn);
      Jo = Jn;
      abo = abn;
      po = po + r;
      const V3 ax = col(R, axis);  // joint axis in world
      Jw.v[0][6 + j] = ax.x;
      Jw.v[1][6 + j] = ax.y;
      Jw.v[2][6 + j] = ax.z;
      al = al + cross(w, qdj * ax);
      w = w + qdj * ax;
      R = mul(R, rot_axis(axis, q[3 * leg + j]));
      for (int b = 0; b < (j == 2 ? 2 : 1); ++b) {  // the knee link also carries the foot (0.04 kg sphere at the contact frame)
        const double mb = (b == 0) ? lm[j] : 0.04;
        const V3 cb = (b == 0) ? lc[j] : foot_org;
        const M3 Ib = (b == 0) ? li[j] : diag3(9.6e-06);
        const V3 rc = mul(R, cb);
        Jac9 Jv;
        V3 ab;
        point_kin(rc, Jw, Jo, w, abo, al, Jv, ab);
        add_body(M, h, leg, mb, rot_inertia(R, Ib), Jv, Jw, w, ab, al);
      }
    }
    const V3 rf = mul(R, foot_org);
    Jac9 Jf;
    V3 abf;
    point_kin(rf, Jw, Jo, w, abo, al, Jf, abf);
    const V3 pf = po + rf;
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < NV; ++c) fJ[leg][r][c] = 0.0;
      for (int c = 0; c < 9; ++c) fJ[leg][r][gcol(leg, c)] = Jf.v[r][c];
      fab[leg][r] = comp(abf, r);
      fpos[leg][r] = comp(pf, r);
      double vel = Jf.v[r][0] * vlin.x + Jf.v[r][1] * vlin.y + Jf.v[r][2] * vlin.z + Jf.v[r][3] * omega.x + Jf.v[r][4] * omega.y + Jf.v[r][5] * omega.z;
      for (int c = 0; c < 3; ++c) vel += Jf.v[r][6 + c] * qd[3 * leg + c];
      fvel[leg][r] = vel;
    }
  }
  for (int i = 0; i < NV; ++i)  // M = (M + M') / 2
    for (int j = 0; j < i; ++j) {
      const double s = 0.5 * (M[i * NV + j] + M[j * NV + i]);
      M[i * NV + j] = s;
      M[j * NV + i] = s;
    }
  if (feet != nullptr)
    for (int l = 0; l < 4; ++l)
      for (int r = 0; r < 3; ++r) {
        feet[(size_t)p * 24 + 3 * l + r] = fpos[l][r];
        feet[(size_t)p * 24 + 12 + 3 * l + r] = fvel[l][r];
      }

  // ---- reduced TSID scene ------------------------------------------------------------------------------------------------
  double *H = Hout + (size_t)p * NQP * NQP, *g = gout + (size_t)p * NQP, *A = Aout + (size_t)p * NC * NQP;
  double *lo = loout + (size_t)p * NC, *hi = hiout + (size_t)p * NC;
  const double *abody = S + 37, *afoot = S + 43, *fref = S + 55;
  for (int e = 0; e < NQP * NQP; ++e) H[e] = 0.0;
  for (int e = 0; e < NC * NQP; ++e) A[e] = 0.0;
  for (int i = 0; i < NQP; ++i) {
    H[i * NQP + i] = C.hessian_regularizer;
    g[i] = 0.0;
  }
  // body_pose task on base_link in world (wbc_arc_opt.cpp:82-91): J = [I6 0], Jdot nu = 0
  for (int i = 0; i < 6; ++i) {
    H[i * NQP + i] += C.com_pose_weight[i];
    g[i] -= C.com_pose_weight[i] * abody[i];
  }
  for (int f = 0; f < 4; ++f) {
    const bool on = (cmask >> f) & 1u;
    if (!on) {  // foot_pose task, active for feet NOT in contact (wbc_arc_opt.cpp:108-117)
      for (int a = 0; a < NV; ++a) {
        double ga = 0.0;
        for (int r = 0; r < 3; ++r) ga += fJ[f][r][a] * C.foot_pose_weight[r] * (fab[f][r] - afoot[3 * f + r]);
        g[a] += ga;
        for (int b = 0; b < NV; ++b) {
          double s = 0.0;
          for (int r = 0; r < 3; ++r) s += fJ[f][r][a] * C.foot_pose_weight[r] * fJ[f][r][b];
          H[b * NQP + a] += s;
        }
      }
    } else {  // foot_contact wrench-forward task, active for feet in contact (wbc_arc_opt.cpp:95-104)
      for (int r = 0; r < 3; ++r) {
        const int s3 = 18 + 3 * f + r;
        H[s3 * NQP + s3] += C.foot_force_weight[r];
        g[s3] -= C.foot_force_weight[r] * fref[3 * f + r];
      }
    }
  }
#define AE(row, colm) A[(size_t)(colm) * NC + (row)]
  // floating-base rows of the equation of motion: M_b nud - sum_f J_f[:, :6]' f_f = -h_b
  for (int r = 0; r < 6; ++r) {
    for (int c = 0; c < NV; ++c) AE(r, c) = M[r * NV + c];
    for (int f = 0; f < 4; ++f)
      for (int k = 0; k < 3; ++k) AE(r, 18 + 3 * f + k) = -fJ[f][k][r];
    lo[r] = -h[r];
    hi[r] = -h[r];
  }
  for (int f = 0; f < 4; ++f) {
    const bool on = (cmask >> f) & 1u;
    const int r0 = 6 + 3 * f;
    for (int k = 0; k < 3; ++k) {
      if (on) {  // rigid contact: J_f nud = -Jdot nu
        for (int c = 0; c < NV; ++c) AE(r0 + k, c) = fJ[f][k][c];
        lo[r0 + k] = -fab[f][k];
        hi[r0 + k] = -fab[f][k];
      } else {  // swing: the contact wrench is removed (f_f = 0)
        AE(r0 + k, 18 + 3 * f + k) = 1.0;
        lo[r0 + k] = -0.0;
        hi[r0 + k] = -0.0;
      }
    }
    // friction pyramid rows (absent for swing feet)
    const int p0 = NEQ + 12 + 4 * f;
    const double pyr[4][3] = {{1, 0, -C.mu}, {1, 0, C.mu}, {0, 1, -C.mu}, {0, 1, C.mu}};
    for (int k = 0; k < 4; ++k) {
      for (int c = 0; c < 3; ++c) AE(p0 + k, 18 + 3 * f + c) = pyr[k][c];
      lo[p0 + k] = (on && (k & 1)) ? 0.0 : -WINF_;
      hi[p0 + k] = (on && !(k & 1)) ? 0.0 : WINF_;
    }
  }
  // joint torque limits: tau = M_j nud + h_j - sum_f J_f[:, 6:]' f_f in [-tau_max, tau_max]
  for (int j = 0; j < 12; ++j) {
    for (int c = 0; c < NV; ++c) AE(NEQ + j, c) = M[(6 + j) * NV + c];
    for (int f = 0; f < 4; ++f)
      for (int k = 0; k < 3; ++k) AE(NEQ + j, 18 + 3 * f + k) = -fJ[f][k][6 + j];
    lo[NEQ + j] = -C.tau_max[j] - h[6 + j];
    hi[NEQ + j] = C.tau_max[j] - h[6 + j];
  }
#undef AE
}

// joint torques of a solved scene (the scene's solver_output -> effort mapping): tau = M_j nud + h_j - sum_f J_f[:, 6:]' f_f, read back
// from the torque-limit rows of A (= [M_j, -J_j']) and their bounds (h_j = -(lo + hi) / 2)
__global__ void wbc_torque_kernel(int n, const double *__restrict__ A, const double *__restrict__ lo, const double *__restrict__ hi,
                                  const double *__restrict__ x, double *__restrict__ tau) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * 12) return;
  const int p = e / 12, j = e - 12 * p;
  const double *Ap = A + (size_t)p * NC * NQP, *xp = x + (size_t)p * NQP;
  double t = -0.5 * (lo[(size_t)p * NC + NEQ + j] + hi[(size_t)p * NC + NEQ + j]);
  for (int c = 0; c < NQP; ++c) t += Ap[(size_t)c * NC + NEQ + j] * xp[c];
  tau[e] = t;
}

#ifndef B200QP_HOST_EMULATION
cudaError_t launch_wbc_scene(int n, const WbcSceneConfig &C, const double *scene_in, double *H, double *g, double *A, double *lo, double *hi,
                             double *feet, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  wbc_scene_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, C, scene_in, H, g, A, lo, hi, feet);
  return cudaGetLastError();
}
cudaError_t launch_wbc_torque(int n, const double *A, const double *lo, const double *hi, const double *x, double *tau, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  wbc_torque_kernel<<<(n * 12 + 127) / 128, 128, 0, st>>>(n, A, lo, hi, x, tau);
  return cudaGetLastError();
}
#endif
}  // namespace b200qp
