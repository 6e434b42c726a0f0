// Kernels K7: the closed loop of SURVEY.md 8(f)-4 kept on the device - a batch of single-rigid-body robots (the "potato" plant the
// reference drives through Drake, potato_sim.cpp:92-140: first-stage GRFs applied at the foot positions, gravity, nothing else)
// stepped between MPC solves, and the coupling of the MPC to the whole-body controller (mit_controller_node.cpp:1000-1006: stage-1
// prediction -> WBC target, stage-0 GRFs -> wrench references, stage-0 contacts -> foot contact flags).  One thread per robot; the
// robot's state lives in its row of the stateful sequencer's input (b200qp.h, B200QP_SEQ2_IN_DOUBLES), so
//   sequencer (K5b) -> QP solve (K1/K2) -> [WBC target (here) -> scene (K6) -> WBC QP (K3)] -> plant (here)
// runs period after period without the host touching a number.
#include "common.cuh"

namespace b200qp {

// mirror of b200qp_rollout_options (include/b200qp.h)
struct RolloutOptions {
  int substeps, plant_inertia, wbc_per_mpc, reserved;
  double com_pose_Kp[6], com_pose_Kd[6], feet_pose_Kp[3], feet_pose_Kd[3];
};

namespace {
constexpr int S_FEET = 13, S_TIME = 25, S_PREV = 55;  // columns of the seq2 row; S_PREV: contact mask of the previous period + 16 (0: none yet)
constexpr double ABAD_X = 0.1934, ABAD_Y = 0.0465, HIP_Y = 0.0955, L1 = 0.213, L2 = 0.213;

__device__ inline void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline void solve3(const double A[3][3], const double *b, double *x) {  // Cramer (well-conditioned 3x3 only)
  const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02, id = 1.0 / det;
  x[0] = id * (c00 * b[0] + (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * b[1] + (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * b[2]);
  x[1] = id * (c01 * b[0] + (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * b[1] + (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * b[2]);
  x[2] = id * (c02 * b[0] + (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * b[1] + (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * b[2]);
}
}  // namespace

// After an MPC solve: touch-down / swing-foot bookkeeping of the massless feet and the forces the plant will see.
//   stance foot (planned contact at stage 0)  -> sits on the foothold planned for stage 0
//   swing foot                                -> travels with the plan: the foothold of its first stance stage (kept if it has none)
//   f0 = stage-0 GRFs of the robots whose solve succeeded, zero otherwise and for swing feet
__global__ void rollout_feet_kernel(int n, int N, int in_stride, const double *__restrict__ rec, const uint8_t *__restrict__ contact,
                                    const double *__restrict__ forces, const b200qp_solver_info *__restrict__ info,
                                    double *__restrict__ seq2, uint8_t *__restrict__ measured, double *__restrict__ f0) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint8_t *ct = contact + (size_t)p * N * 4;
  const double *fp = rec + (size_t)p * in_stride + 26 + 13 * (N + 1);
  double *S = seq2 + (size_t)p * 64;
  const bool ok = info[p].status == B200QP_ACADOS_SUCCESS;
  unsigned mask = 0;
  for (int leg = 0; leg < 4; ++leg) {
    const bool c0 = ct[leg] != 0;
    int first = -1;
    if (c0) first = 0;
    else
      for (int k = 1; k < N && first < 0; ++k)
        if (ct[4 * k + leg]) first = k;
    if (first >= 0)
      for (int c = 0; c < 3; ++c) S[S_FEET + 3 * leg + c] = fp[12 * first + 3 * leg + c];
    for (int c = 0; c < 3; ++c) f0[(size_t)p * 12 + 3 * leg + c] = (ok && c0) ? forces[(size_t)p * N * 12 + 3 * leg + c] : 0.0;
    measured[4 * p + leg] = c0 ? 1 : 0;
    if (c0) mask |= 1u << leg;
  }
  S[S_PREV] = (double)(16u + mask);
}

// Single rigid body under the applied forces for `h_total` seconds (semi-implicit Euler, `substeps` steps):
//   m v' = sum f + m g,   Iw w' = sum (r_f - p) x f - w x Iw w,   q' = 1/2 [w, 0] q
// plant_inertia 0: Iw = R Ib R' (physical);  1: the world inertia of the reference MPC's own model, built from GetR(psi) (mpc.cpp:46-64)
__global__ void rollout_plant_kernel(int n, SeqModel M, int plant_inertia, double h_total, int substeps, const double *__restrict__ f_app,
                                     int f_stride, int f_off, const b200qp_solver_info *__restrict__ f_info, const double *__restrict__ f_fallback,
                                     double *__restrict__ seq2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double *S = seq2 + (size_t)p * 64;
  double q[4] = {S[0], S[1], S[2], S[3]}, pos[3] = {S[4], S[5], S[6]}, w[3] = {S[7], S[8], S[9]}, v[3] = {S[10], S[11], S[12]};
  double f[12];
  const bool use_app = f_info == nullptr || f_info[p].status == B200QP_ACADOS_SUCCESS;
  for (int c = 0; c < 12; ++c) f[c] = use_app ? f_app[(size_t)p * f_stride + f_off + c] : f_fallback[(size_t)p * 12 + c];
  const double h = h_total / (double)substeps;
  for (int s = 0; s < substeps; ++s) {
    double R[3][3], Iw[3][3];
    quat_to_rot(q, R);
    if (plant_inertia == 1) {
      const double psi = atan2(R[1][0], R[0][0]), cs = cos(psi), sn = sin(psi);
      const double Rz[3][3] = {{cs, -sn, 0.0}, {sn, cs, 0.0}, {0.0, 0.0, 1.0}};
      double T[3][3];  // Iw = Rz' Ib Rz
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = Rz[0][i] * M.inertia[0 * 3 + j] + Rz[1][i] * M.inertia[1 * 3 + j] + Rz[2][i] * M.inertia[2 * 3 + j];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Iw[i][j] = T[i][0] * Rz[0][j] + T[i][1] * Rz[1][j] + T[i][2] * Rz[2][j];
    } else {
      double T[3][3];  // Iw = R Ib R'
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i][j] = R[i][0] * M.inertia[0 * 3 + j] + R[i][1] * M.inertia[1 * 3 + j] + R[i][2] * M.inertia[2 * 3 + j];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Iw[i][j] = T[i][0] * R[j][0] + T[i][1] * R[j][1] + T[i][2] * R[j][2];
    }
    double tau[3] = {0.0, 0.0, 0.0}, fs[3] = {0.0, 0.0, 0.0};
    for (int leg = 0; leg < 4; ++leg) {
      const double r[3] = {S[S_FEET + 3 * leg] - pos[0], S[S_FEET + 3 * leg + 1] - pos[1], S[S_FEET + 3 * leg + 2] - pos[2]};
      double t[3];
      cross3(r, f + 3 * leg, t);
      for (int c = 0; c < 3; ++c) {
        tau[c] += t[c];
        fs[c] += f[3 * leg + c];
      }
    }
    double Iww[3], wxI[3], rhs[3], wd[3];
    for (int i = 0; i < 3; ++i) Iww[i] = Iw[i][0] * w[0] + Iw[i][1] * w[1] + Iw[i][2] * w[2];
    cross3(w, Iww, wxI);
    for (int c = 0; c < 3; ++c) rhs[c] = tau[c] - wxI[c];
    solve3(Iw, rhs, wd);
    for (int c = 0; c < 3; ++c) w[c] += h * wd[c];
    v[0] += h * (fs[0] / M.mass);
    v[1] += h * (fs[1] / M.mass);
    v[2] += h * (fs[2] / M.mass - GRAVITY_CONSTANT);
    for (int c = 0; c < 3; ++c) pos[c] += h * v[c];
    // dq = [w, 0] * q (Hamilton product, xyzw)
    const double dq[4] = {w[0] * q[3] + w[1] * q[2] - w[2] * q[1], w[1] * q[3] - w[0] * q[2] + w[2] * q[0],
                          w[2] * q[3] + w[0] * q[1] - w[1] * q[0], -w[0] * q[0] - w[1] * q[1] - w[2] * q[2]};
    double nn = 0.0;
    for (int c = 0; c < 4; ++c) {
      q[c] += 0.5 * h * dq[c];
      nn += q[c] * q[c];
    }
    nn = 1.0 / sqrt(nn);
    for (int c = 0; c < 4; ++c) q[c] *= nn;
  }
  for (int c = 0; c < 4; ++c) S[c] = q[c];
  for (int c = 0; c < 3; ++c) {
    S[4 + c] = pos[c];
    S[7 + c] = w[c];
    S[10 + c] = v[c];
  }
}

// End of a control period: the clock advances by dt, one history row per robot.
// history row (16 doubles): pos 0:3, vel 3:6, quat 6:10, sum f_z applied 10, MPC status 11, iterations 12, polish rounds 13, WBC status of the
// last WBC solve of the period 14 (-1: no WBC), spare 15
__global__ void rollout_finish_kernel(int n, double dt_ns, const double *__restrict__ f0, const b200qp_solver_info *__restrict__ info,
                                      const b200qp_solver_info *__restrict__ winfo, double *__restrict__ seq2, double *__restrict__ hist) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double *S = seq2 + (size_t)p * 64;
  S[S_TIME] += dt_ns;
  if (hist == nullptr) return;
  double *Hh = hist + (size_t)p * 16;
  for (int c = 0; c < 3; ++c) {
    Hh[c] = S[4 + c];
    Hh[3 + c] = S[10 + c];
  }
  for (int c = 0; c < 4; ++c) Hh[6 + c] = S[c];
  Hh[10] = f0[(size_t)p * 12 + 2] + f0[(size_t)p * 12 + 5] + f0[(size_t)p * 12 + 8] + f0[(size_t)p * 12 + 11];
  Hh[11] = (double)info[p].status;
  Hh[12] = (double)info[p].iterations;
  Hh[13] = (double)info[p].polish_rounds;
  Hh[14] = winfo != nullptr ? (double)winfo[p].status : -1.0;
  Hh[15] = 0.0;
}

// MPC -> WBC coupling (mit_controller_node.cpp:1000-1006, wbc_arc_opt.cpp:252-289): the scene input of kernel K6 from the plant state.
//   joints: closed-form leg inverse kinematics of the massless legs (foot position in the body frame -> abad, hip, knee), joint
//           velocities from the leg Jacobian with the feet at rest in the world
//   body task: Cartesian PD on the MPC's stage-1 prediction (orientation error as a rotation vector), feet task: PD towards rest
//   wrench references: f0, contacts: the stage-0 mask
__global__ void rollout_wbc_target_kernel(int n, int N, RolloutOptions O, const double *__restrict__ seq2, const double *__restrict__ xpred,
                                          const double *__restrict__ f0, double *__restrict__ scene) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double *S = seq2 + (size_t)p * 64;
  double *Z = scene + (size_t)p * 72;
  double R[3][3];
  quat_to_rot(S, R);
  const double pos[3] = {S[4], S[5], S[6]}, w[3] = {S[7], S[8], S[9]}, v[3] = {S[10], S[11], S[12]};
  for (int c = 0; c < 4; ++c) Z[c] = S[c];
  for (int c = 0; c < 3; ++c) {
    Z[4 + c] = pos[c];
    Z[7 + c] = v[c];
    Z[10 + c] = w[c];
  }
  const unsigned mask = ((unsigned)S[S_PREV]) & 15u;
  for (int leg = 0; leg < 4; ++leg) {
    const double sx = (leg < 2) ? 1.0 : -1.0, sy = (leg % 2 == 0) ? 1.0 : -1.0;
    const double rw[3] = {S[S_FEET + 3 * leg] - pos[0], S[S_FEET + 3 * leg + 1] - pos[1], S[S_FEET + 3 * leg + 2] - pos[2]};
    // foot in the body frame, relative to the abad joint
    const double bx = R[0][0] * rw[0] + R[1][0] * rw[1] + R[2][0] * rw[2] - sx * ABAD_X;
    const double by = R[0][1] * rw[0] + R[1][1] * rw[1] + R[2][1] * rw[2] - sy * ABAD_Y;
    const double bz = R[0][2] * rw[0] + R[1][2] * rw[1] + R[2][2] * rw[2];
    const double l0 = sy * HIP_Y;
    const double hh = sqrt(fmax(by * by + bz * bz - l0 * l0, 1e-12));
    const double q0 = atan2(hh * by + l0 * bz, l0 * by - hh * bz);
    double d2 = bx * bx + hh * hh;
    const double dmax = (L1 + L2) * (L1 + L2) * (1.0 - 1e-9), dmin = 1e-6;
    d2 = fmin(fmax(d2, dmin), dmax);  // out of reach: the leg is stretched towards the foot
    const double ck = fmin(fmax((d2 - L1 * L1 - L2 * L2) / (2.0 * L1 * L2), -1.0), 1.0);
    const double q2 = -acos(ck);
    const double q1 = atan2(-bx, hh) - atan2(L2 * sin(q2), L1 + L2 * cos(q2));
    Z[13 + 3 * leg] = q0;
    Z[13 + 3 * leg + 1] = q1;
    Z[13 + 3 * leg + 2] = q2;
    // leg Jacobian (world): columns = joint axis x (foot - joint origin)
    const double c0 = cos(q0), s0 = sin(q0), c1 = cos(q1), s1 = sin(q1), c12 = cos(q1 + q2), s12 = sin(q1 + q2);
    // body-frame vectors: hip origin -> foot, knee origin -> foot (after the abad rotation Rx(q0))
    const double tx = -L1 * s1 - L2 * s12, tz = -L1 * c1 - L2 * c12;      // in the leg plane (x, z')
    const double kx = -L2 * s12, kz = -L2 * c12;
    const double a_foot[3] = {tx, l0 * c0 - tz * s0, l0 * s0 + tz * c0};   // abad origin -> foot
    const double h_foot[3] = {tx, -tz * s0, tz * c0};                      // hip origin -> foot
    const double k_foot[3] = {kx, -kz * s0, kz * c0};                      // knee origin -> foot
    const double ax0[3] = {1.0, 0.0, 0.0}, ax1[3] = {0.0, c0, s0};         // abad axis, hip / knee axis (body frame)
    double jb[3][3], col[3];
    cross3(ax0, a_foot, col);
    for (int r = 0; r < 3; ++r) jb[r][0] = col[r];
    cross3(ax1, h_foot, col);
    for (int r = 0; r < 3; ++r) jb[r][1] = col[r];
    cross3(ax1, k_foot, col);
    for (int r = 0; r < 3; ++r) jb[r][2] = col[r];
    double J[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) J[r][c] = R[r][0] * jb[0][c] + R[r][1] * jb[1][c] + R[r][2] * jb[2][c];
    // foot at rest in the world: v + w x r + J qd = 0
    double wr[3], rhs[3], qd[3];
    cross3(w, rw, wr);
    for (int c = 0; c < 3; ++c) rhs[c] = -(v[c] + wr[c]);
    solve3(J, rhs, qd);
    for (int c = 0; c < 3; ++c) {
      Z[25 + 3 * leg + c] = qd[c];
      Z[43 + 3 * leg + c] = 0.0;  // feet task: PD towards the foot's own position and zero velocity
      Z[55 + 3 * leg + c] = f0[(size_t)p * 12 + 3 * leg + c];
    }
  }
  // body task: PD on the stage-1 prediction (euler -> quaternion as mpc.cpp:463)
  const double *x1 = xpred + (size_t)p * (N + 1) * 13 + 13;
  const double hr = 0.5 * x1[0], hp = 0.5 * x1[1], hy = 0.5 * x1[2];
  const double cr = cos(hr), sr = sin(hr), cp = cos(hp), sp = sin(hp), cy = cos(hy), sy_ = sin(hy);
  const double qr[4] = {cy * cp * sr - sy_ * sp * cr, cy * sp * cr + sy_ * cp * sr, sy_ * cp * cr - cy * sp * sr, cy * cp * cr + sy_ * sp * sr};
  // rotation-vector error of q_ref * conj(q)
  const double *q = S;
  double e[4] = {qr[3] * -q[0] + qr[0] * q[3] + qr[1] * -q[2] - qr[2] * -q[1], qr[3] * -q[1] - qr[0] * -q[2] + qr[1] * q[3] + qr[2] * -q[0],
                 qr[3] * -q[2] + qr[0] * -q[1] - qr[1] * -q[0] + qr[2] * q[3], qr[3] * q[3] + qr[0] * q[0] + qr[1] * q[1] + qr[2] * q[2]};
  if (e[3] < 0.0)
    for (int c = 0; c < 4; ++c) e[c] = -e[c];
  const double vn = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
  const double ang = 2.0 * atan2(vn, e[3]);
  const double sc = vn > 1e-12 ? ang / vn : 2.0;
  for (int c = 0; c < 3; ++c) {
    Z[37 + c] = O.com_pose_Kd[c] * (x1[9 + c] - v[c]) + O.com_pose_Kp[c] * (x1[3 + c] - pos[c]);
    Z[40 + c] = O.com_pose_Kd[3 + c] * (x1[6 + c] - w[c]) + O.com_pose_Kp[3 + c] * (sc * e[c]);
  }
  Z[67] = (double)mask;
  for (int c = 68; c < 72; ++c) Z[c] = 0.0;
}

#ifndef B200QP_HOST_EMULATION
static inline int nblk(int n) { return (n + 63) / 64; }
cudaError_t launch_rollout_feet(int n, int N, int in_stride, const double *rec, const uint8_t *contact, const double *forces,
                                const b200qp_solver_info *info, double *seq2, uint8_t *measured, double *f0, cudaStream_t st) {
  rollout_feet_kernel<<<nblk(n), 64, 0, st>>>(n, N, in_stride, rec, contact, forces, info, seq2, measured, f0);
  return cudaGetLastError();
}
cudaError_t launch_rollout_plant(int n, const SeqModel &M, int plant_inertia, double h_total, int substeps, const double *f_app, int f_stride,
                                 int f_off, const b200qp_solver_info *f_info, const double *f_fallback, double *seq2, cudaStream_t st) {
  rollout_plant_kernel<<<nblk(n), 64, 0, st>>>(n, M, plant_inertia, h_total, substeps, f_app, f_stride, f_off, f_info, f_fallback, seq2);
  return cudaGetLastError();
}
cudaError_t launch_rollout_finish(int n, double dt_ns, const double *f0, const b200qp_solver_info *info, const b200qp_solver_info *winfo,
                                  double *seq2, double *hist, cudaStream_t st) {
  rollout_finish_kernel<<<nblk(n), 64, 0, st>>>(n, dt_ns, f0, info, winfo, seq2, hist);
  return cudaGetLastError();
}
cudaError_t launch_rollout_wbc_target(int n, int N, const RolloutOptions &O, const double *seq2, const double *xpred, const double *f0,
                                      double *scene, cudaStream_t st) {
  rollout_wbc_target_kernel<<<nblk(n), 64, 0, st>>>(n, N, O, seq2, xpred, f0, scene);
  return cudaGetLastError();
}
#endif
}  // namespace b200qp
