// Kernel K5b: the reference's gait sequencers WITH THEIR STATE ACROSS CALLS, one thread per robot (compiled -fmad=false).
//   SimpleGaitSequencer::UpdateState / GetGaitSequence      simple_gait_sequencer.cpp:43-80  (phase from the clock, standing rule)
//   AdaptiveGaitSequencer::GetGaitSequence                  adaptive_gait_sequencer.cpp:42-59
//   MPCTrajectoryPlanner::plan_trajectory                   mpc_trajectory_planner.cpp:28-140   (every Target field and flag)
//   MPCTrajectoryPlanner::plan_desired_trajectory           mpc_trajectory_planner.cpp:142-264  (stand-still latch :158-188)
//   Gait::update_sequence                                   gait.cpp:67-112                     (early-contact rule)
//   AdaptiveGait::update_sequence / update_params / nextPhase / disturbance_factor / set_offset   gait.cpp:411-700
//   RaibertFootStepPlanner::get_foot_position_sequence / get_foothold   raibert_foot_step_planner.cpp:43-169 (z_on_plane = false)
//   MovingAverage<T>::update                                common/filters.hpp:105-132
// What the reference keeps in its objects between calls lives in a per-robot state record in device memory (layout below);
// a call reads the robot's measured state, feet, contacts, clock and Target, advances that record exactly as the reference's
// objects would, and writes the MPC input record + contact bytes the solve kernels consume.  The stage-parallel first-call
// kernel of gait_schedule.cu stays the fast path for stateless use; this kernel is the faithful one.
#include "common.cuh"

namespace b200qp {



namespace {
constexpr double EPS = 2.220446049250313e-16;
constexpr int ROWS = MAXH + 1;
// state record (B200QP_SEQ_STATE_DOUBLES = 192)
constexpr int ST_INIT = 0, ST_DPOS = 1, ST_DQ = 4, ST_PFH = 8, ST_RINIT = 12, ST_RFH = 13, ST_FCNT = 17, ST_FIDX = 18, ST_FSUM = 19,
              ST_FBUF = 22 /* [32][3] */, ST_TINIT = 118, ST_T0 = 119, ST_APH = 120, ST_AT = 124, ST_ADF = 125, ST_APO = 129,
              ST_ANPO = 133, ST_ADEL = 137, ST_ASW = 138, ST_VCNT = 139, ST_VIDX = 140, ST_VSUM = 141, ST_VBUF = 143 /* [24][2] */,
              ST_AINIT = 191;
// input record (B200QP_SEQ2_IN_DOUBLES = 64)
constexpr int IN_FEET = 13, IN_TIME = 25, IN_TGT = 26, IN_ACT = 45, IN_PER = 46, IN_DUTY = 47, IN_OFF = 51;
// Target value order: x y z roll pitch yaw | qx qy qz qw | x_dot y_dot z_dot | hybrid_x_dot hybrid_y_dot hybrid_z_dot | wx wy wz
// Target flag order (target.hpp:28-45): x y z roll pitch yaw full_orientation x_dot y_dot z_dot hybrid_x_dot hybrid_y_dot hybrid_z_dot wx wy wz
enum { A_X, A_Y, A_Z, A_ROLL, A_PITCH, A_YAW, A_FULL, A_XD, A_YD, A_ZD, A_HXD, A_HYD, A_HZD, A_WX, A_WY, A_WZ };

__device__ inline double q_yaw(const double *q) {
  double R[3][3];
  quat_to_rot(q, R);
  return atan2(R[1][0], R[0][0]);
}
__device__ inline void q_euler(const double *q, double *e) {
  double R[3][3];
  quat_to_rot(q, R);
  rot_to_euler(R, e);
}
__device__ inline void e_to_q(const double *e, double *q) {  // euler_to_quaternion, quaternion_operations.hpp:19-25
  const double r = 0.5 * e[0], p = 0.5 * e[1], y = 0.5 * e[2];
  const double cr = cos(r), sr = sin(r), cp = cos(p), sp = sin(p), cy = cos(y), sy = sin(y);
  q[3] = cy * cp * cr + sy * sp * sr;
  q[0] = cy * cp * sr - sy * sp * cr;
  q[1] = cy * sp * cr + sy * cp * sr;
  q[2] = sy * cp * cr - cy * sp * sr;
}
__device__ inline void yaw_q(double yaw, double *q) {
  q[0] = 0.0; q[1] = 0.0; q[2] = sin(0.5 * yaw); q[3] = cos(0.5 * yaw);
}
__device__ inline void slerp_half(const double *a, const double *b, double *o) {  // Eigen slerp(0.5, b)
  const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  const double ad = fabs(d);
  double s0 = 0.5, s1 = 0.5;
  if (ad < 1.0 - EPS) {
    const double th = acos(ad), st = sin(th);
    s0 = sin(0.5 * th) / st;
    s1 = s0;
  }
  if (d < 0) s1 = -s1;
  for (int c = 0; c < 4; ++c) o[c] = s0 * a[c] + s1 * b[c];
}
__device__ inline void rot_v(const double *q, const double *v, double *o) {
  double R[3][3];
  quat_to_rot(q, R);
  for (int r = 0; r < 3; ++r) o[r] = R[r][0] * v[0] + R[r][1] * v[1] + R[r][2] * v[2];
}
__device__ inline void rot_vt(const double *q, const double *v, double *o) {  // conjugate rotation
  double R[3][3];
  quat_to_rot(q, R);
  for (int r = 0; r < 3; ++r) o[r] = R[0][r] * v[0] + R[1][r] * v[1] + R[2][r] * v[2];
}
__device__ inline double min4(const double *h) { return fmin(fmin(h[0], h[1]), fmin(h[2], h[3])); }

// MovingAverage<VectorDd>::update (filters.hpp:116-132): cnt / idx / sum / buffer live in the state record
template <int D>
__device__ inline void moving_average(double *cnt, double *idx, double *sum, double *buf, int size, const double *value, double *out) {
  const int n = (int)*cnt;
  if (n == size) {
    const int i = (int)*idx;
    for (int c = 0; c < D; ++c) {
      sum[c] = sum[c] - buf[D * i + c];
      buf[D * i + c] = value[c];
      sum[c] = sum[c] + value[c];
      out[c] = sum[c] / (double)(unsigned)size;
    }
    *idx = (double)((i + 1) % size);
  } else if (n > 0) {
    for (int c = 0; c < D; ++c) {
      buf[D * n + c] = value[c];
      sum[c] = sum[c] + value[c];
      out[c] = sum[c] / (double)(n + 1);
    }
    *cnt = (double)(n + 1);
  } else {
    for (int c = 0; c < D; ++c) {
      buf[c] = value[c];
      sum[c] = value[c];
      out[c] = value[c];
    }
    *cnt = 1.0;
  }
}

// RaibertFootStepPlanner::get_foothold (raibert_foot_step_planner.cpp:133-169) without the reachability clip
__device__ inline void raibert_foothold(const SeqModel &M, int leg, const double *pos, const double *yawq, double t_stance, const double *v,
                                        const double *v_cmd, const double *w_cmd, const double *heights, double *out, double *psh) {
  const double sh[3] = {M.shoulders[3 * leg], M.shoulders[3 * leg + 1], M.shoulders[3 * leg + 2]};
  rot_v(yawq, sh, psh);
  for (int c = 0; c < 3; ++c) psh[c] = pos[c] + psh[c];
  double h = pos[2] - min4(heights);
  h = fmin(fmax(h, 0.0), M.max_leg_length);
  const double cr[3] = {v[1] * w_cmd[2] - v[2] * w_cmd[1], v[2] * w_cmd[0] - v[0] * w_cmd[2], v[0] * w_cmd[1] - v[1] * w_cmd[0]};
  const double kc = 0.5 * sqrt(h / M.raibert_g);
  for (int c = 0; c < 3; ++c) out[c] = psh[c] + kc * cr[c] + ((t_stance / 2.0) * v[c] + M.raibert_k * (v[c] - v_cmd[c]));
}

__device__ inline double froude(double v, double h) { return fabs((v * v) / (GRAVITY_CONSTANT * h)); }

// AdaptiveGait::set_offset (gait.cpp:698-707) on the state record
__device__ inline void adaptive_set_offset(double *St, const SeqOptions &O, const double *po) {
  St[ST_ASW] = 1.0;
  if (O.offset_delay <= EPS) {
    for (int l = 0; l < 4; ++l) {
      St[ST_APO + l] = po[l];
      St[ST_ANPO + l] = po[l];
    }
  } else {
    bool same = true;
    for (int l = 0; l < 4; ++l) same &= (po[l] == St[ST_ANPO + l]);
    if (!same) {
      for (int l = 0; l < 4; ++l) St[ST_ANPO + l] = po[l];
      St[ST_ADEL] = O.offset_delay;
    }
  }
}

// AdaptiveGait::nextPhase (gait.cpp:579-679)
__device__ inline void adaptive_next_phase(double *phase, double dt, double T, const double *phase_offset, const double *df_old,
                                           const double *df_new, const uint8_t *contact_state, double max_correction_cycles) {
  for (int i = 0; i < 4; ++i) {
    if (T <= EPS || df_old[i] <= EPS || df_old[i] >= 1.0 - EPS) {
      bool all = true;
      for (int j = 0; j < 4; ++j) {
        if (T >= EPS && (df_old[j] <= EPS || df_old[j] >= 1.0 - EPS)) {
          all = false;
          break;
        }
      }
      if (!all) phase[i] = fmod(phase[i] + dt / T, 1.0);
    } else {
      phase[i] = fmod(phase[i] + dt / T, 1.0);
      if (phase[i] < df_old[i]) phase[i] = fmod(phase[i] * df_new[i] / df_old[i], 1.0);
      else phase[i] = fmod((phase[i] - df_old[i]) * (1.0 - df_new[i]) / (1.0 - df_old[i]) + df_new[i], 1.0);
    }
  }
  if (contact_state != nullptr) {
    for (int i = 0; i < 4; ++i)
      if (contact_state[i] && (phase[i] > df_new[i]) && (phase[i] > ((df_new[i] + 1.0) / 2.0))) phase[i] = 0.0;
  }
  int ground_idx = 0;
  double min_dif = 1.0;
  bool swing_found = false;
  for (int i = 0; i < 4; ++i) {
    const double dif = phase[i] - df_new[i];
    if (swing_found && dif >= 0.0 && dif < min_dif) {
      ground_idx = i;
      min_dif = dif;
    } else if (!swing_found && dif >= 0) {
      ground_idx = i;
      min_dif = dif;
      swing_found = true;
    } else if (!swing_found && (double)abs((int)dif) < min_dif) {  // the reference calls the integer abs() here (gait.cpp:648)
      ground_idx = i;
      min_dif = (double)abs((int)dif);
    }
  }
  double err[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 4; ++i) {
    if (i != ground_idx) {
      const double cur = fmod((phase[ground_idx] + phase_offset[ground_idx]) - phase[i], 1.0);
      err[i] = fmod((cur - phase_offset[i]) + 0.5, 1.0) - 0.5;
    }
  }
  const double mx = dt / (max_correction_cycles * T);
  for (int i = 0; i < 4; ++i) {
    if (phase[i] < df_new[i] - mx) {
      if (err[i] >= 0.0) phase[i] = fmin(phase[i] + fmin(err[i], mx), df_new[i]);
      else phase[i] = fmax(phase[i] + fmax(err[i], -mx), 0.0);
    }
  }
}
}  // namespace

// in: [n][64], measured: [n][4] or nullptr (= all feet in contact), state: [n][192]
__global__ void sequencer_full_kernel(int n, SeqModel M, SeqOptions O, const double *__restrict__ seq_in,
                                      const uint8_t *__restrict__ measured, double *__restrict__ state, double *__restrict__ rec_out,
                                      uint8_t *__restrict__ contact_out, int in_stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double *S = seq_in + (size_t)p * 64;
  double *St = state + (size_t)p * 192;
  const int N = M.N, steps = N + 1;
  const double dt = M.dt;
  double *rec = rec_out + (size_t)p * in_stride;
  uint8_t *ct = contact_out + (size_t)p * N * 4;
  uint8_t meas[4];
  for (int l = 0; l < 4; ++l) meas[l] = measured != nullptr ? (measured[4 * p + l] != 0) : 1;
  const unsigned act = (unsigned)S[IN_ACT];
#define ACT(bit) (((act >> (bit)) & 1u) != 0u)
  const double *T_ = S + IN_TGT;
  const double t_x = T_[0], t_y = T_[1], t_z = T_[2], t_roll = T_[3], t_pitch = T_[4], t_yaw = T_[5];
  const double *t_q = T_ + 6;
  const double t_xd = T_[10], t_yd = T_[11], t_zd = T_[12], t_hx = T_[13], t_hy = T_[14], t_hz = T_[15];
  const double t_w[3] = {T_[16], T_[17], T_[18]};

  // ---- header: state + model --------------------------------------------------------------------------------------------
  for (int c = 0; c < 13; ++c) rec[c] = S[c];
  rec[13] = M.mass;
  for (int c = 0; c < 9; ++c) rec[14 + c] = M.inertia[c];
  for (int c = 0; c < 3; ++c) rec[23 + c] = M.com[c];

  // ---- MPCTrajectoryPlanner::plan_trajectory (mpc_trajectory_planner.cpp:28-140) ---------------------------------------------
  for (int l = 0; l < 4; ++l)
    if (meas[l]) St[ST_PFH + l] = S[IN_FEET + 3 * l + 2];
  const double min_foot_height = min4(St + ST_PFH);
  double ref_pos[ROWS][3], ref_q[ROWS][4], ref_vel[ROWS][3], ref_tw[ROWS][3];
  {
    double position[3] = {S[4], S[5], S[6]}, orientation[4] = {S[0], S[1], S[2], S[3]}, velocity[3] = {S[10], S[11], S[12]};
    double twist[3] = {S[7], S[8], S[9]}, eul[3];
    q_euler(orientation, eul);
    for (int c = 0; c < 3; ++c) { ref_pos[0][c] = position[c]; ref_vel[0][c] = velocity[c]; ref_tw[0][c] = twist[c]; }
    for (int c = 0; c < 4; ++c) ref_q[0][c] = orientation[c];
    for (int i = 1; i < steps; ++i) {
      if (ACT(A_WX)) twist[0] = t_w[0];
      if (ACT(A_WY)) twist[1] = t_w[1];
      if (ACT(A_WZ)) twist[2] = t_w[2];
      for (int c = 0; c < 3; ++c) eul[c] = eul[c] + dt * twist[c];
      if (ACT(A_FULL)) {
        for (int c = 0; c < 4; ++c) orientation[c] = t_q[c];
      } else {
        if (ACT(A_YAW)) eul[2] = t_yaw;
        if (ACT(A_PITCH)) eul[1] = t_pitch;
        if (ACT(A_ROLL)) eul[0] = t_roll;
        e_to_q(eul, orientation);
      }
      if (ACT(A_HXD) || ACT(A_HYD)) {
        const double hv[3] = {ACT(A_HXD) ? t_hx : 0.0, ACT(A_HYD) ? t_hy : 0.0, 0.0};
        rot_v(orientation, hv, velocity);
        double yq_prev[4], yq[4], mid[4], stp[3];
        const double dv[3] = {dt * hv[0], dt * hv[1], dt * hv[2]};
        yaw_q(q_yaw(ref_q[i - 1]), yq_prev);
        yaw_q(q_yaw(orientation), yq);
        slerp_half(yq_prev, yq, mid);
        rot_v(mid, dv, stp);
        for (int c = 0; c < 3; ++c) position[c] = position[c] + stp[c];
      } else {
        if (ACT(A_XD)) { velocity[0] = t_xd; position[0] = position[0] + dt * t_xd; }
        if (ACT(A_YD)) { velocity[1] = t_yd; position[1] = position[1] + dt * t_yd; }
      }
      if (ACT(A_ZD)) { velocity[2] = t_zd; position[2] = position[2] + dt * dt * t_zd; }
      else if (ACT(A_HZD)) { velocity[2] = t_hz; position[2] = position[2] + dt * velocity[2]; }
      if (ACT(A_X)) position[0] = t_x;
      if (ACT(A_Y)) position[1] = t_y;
      if (ACT(A_Z)) position[2] = t_z + min_foot_height;
      for (int c = 0; c < 3; ++c) { ref_pos[i][c] = position[c]; ref_vel[i][c] = velocity[c]; ref_tw[i][c] = twist[c]; }
      for (int c = 0; c < 4; ++c) ref_q[i][c] = orientation[c];
    }
  }
  // ---- plan_desired_trajectory (:142-264) -> rows of the MPC record ----------------------------------------------------------
  double des_q0[4];
  {
    double position[3], orientation[4], velocity[3], eul[3];
    for (int c = 0; c < 3; ++c) { position[c] = ref_pos[0][c]; velocity[c] = ref_vel[0][c]; }
    for (int c = 0; c < 4; ++c) orientation[c] = ref_q[0][c];
    q_euler(orientation, eul);
    const bool fix = O.fix_standing_position != 0;
    // row 0 is the measured state in both trajectories
    double *st0 = rec + 26;
    for (int c = 0; c < 4; ++c) { st0[c] = orientation[c]; des_q0[c] = orientation[c]; }
    for (int c = 0; c < 3; ++c) { st0[4 + c] = position[c]; st0[7 + c] = ref_tw[0][c]; st0[10 + c] = ref_vel[0][c]; }
    if (fix && St[ST_INIT] != 0.0) {  // stand-still latch (:158-188) on the previous call's desired row 1
      const double *dp1 = St + ST_DPOS, *dq1 = St + ST_DQ;
      const double dx = position[0] - dp1[0], dy = position[1] - dp1[1];
      if (!(ACT(A_XD) && fabs(t_xd) >= EPS) && !(ACT(A_HXD) && fabs(t_hx) >= EPS) && !(ACT(A_YD) && fabs(t_yd) >= EPS) &&
          !(ACT(A_HYD) && fabs(t_hy) >= EPS) && !(ACT(A_WZ) && fabs(t_w[2]) >= EPS) &&
          (ACT(A_XD) || ACT(A_HXD) || ACT(A_YD) || ACT(A_HYD)) && (sqrt(dx * dx + dy * dy) < O.dist_thr) &&
          (sqrt(velocity[0] * velocity[0] + velocity[1] * velocity[1] + velocity[2] * velocity[2]) < O.vel_thr)) {
        position[0] = dp1[0];
        position[1] = dp1[1];
      }
      if (!(ACT(A_ZD) && fabs(t_zd) >= EPS) && !(ACT(A_HZD) && fabs(t_hz) >= EPS) && (ACT(A_ZD) || ACT(A_HZD))) position[2] = dp1[2];
      double old[3];
      q_euler(dq1, old);
      if (ACT(A_WX) && fabs(t_w[0]) <= EPS) eul[0] = old[0];
      if (ACT(A_WY) && fabs(t_w[1]) <= EPS) eul[1] = old[1];
      if (ACT(A_WZ) && fabs(t_w[2]) <= EPS && (fabs(eul[2] - old[2]) < O.ang_thr)) eul[2] = old[2];
    }
    double prev_q[4] = {orientation[0], orientation[1], orientation[2], orientation[3]};
    for (int i = 1; i < steps; ++i) {
      double *st = rec + 26 + 13 * i;
      if (fix) {
        for (int c = 0; c < 3; ++c) eul[c] = eul[c] + dt * ref_tw[i][c];
        if (ACT(A_FULL)) {
          for (int c = 0; c < 4; ++c) orientation[c] = t_q[c];
        } else {
          if (ACT(A_YAW)) eul[2] = t_yaw;
          if (ACT(A_PITCH)) eul[1] = t_pitch;
          if (ACT(A_ROLL)) eul[0] = t_roll;
          e_to_q(eul, orientation);
        }
        if (ACT(A_HXD) || ACT(A_HYD)) {
          const double hv[3] = {ACT(A_HXD) ? t_hx : 0.0, ACT(A_HYD) ? t_hy : 0.0, 0.0};
          double yq_prev[4], yq[4], mid[4], stp[3];
          const double dv[3] = {dt * hv[0], dt * hv[1], dt * hv[2]};
          yaw_q(q_yaw(prev_q), yq_prev);
          yaw_q(q_yaw(orientation), yq);
          slerp_half(yq_prev, yq, mid);
          rot_v(mid, dv, stp);
          for (int c = 0; c < 3; ++c) position[c] = position[c] + stp[c];
        } else {
          if (ACT(A_XD)) position[0] = position[0] + dt * t_xd;
          if (ACT(A_YD)) position[1] = position[1] + dt * t_yd;
        }
        if (ACT(A_ZD)) position[2] = position[2] + dt * dt * t_zd;
        else if (ACT(A_HZD)) position[2] = position[2] + dt * ref_vel[i][2];
        if (ACT(A_X)) position[0] = t_x;
        if (ACT(A_Y)) position[1] = t_y;
        if (ACT(A_Z)) position[2] = t_z + min_foot_height;
      } else {  // desired = reference (:143-147)
        for (int c = 0; c < 3; ++c) position[c] = ref_pos[i][c];
        for (int c = 0; c < 4; ++c) orientation[c] = ref_q[i][c];
      }
      for (int c = 0; c < 4; ++c) { st[c] = orientation[c]; prev_q[c] = orientation[c]; }
      for (int c = 0; c < 3; ++c) { st[4 + c] = position[c]; st[7 + c] = ref_tw[i][c]; st[10 + c] = ref_vel[i][c]; }
      if (i == 1 && fix) {
        for (int c = 0; c < 3; ++c) St[ST_DPOS + c] = position[c];
        for (int c = 0; c < 4; ++c) St[ST_DQ + c] = orientation[c];
      }
    }
    if (fix) St[ST_INIT] = 1.0;
  }
  // ---- set_current_target (:266-358): what the adaptive gait reads ----------------------------------------------------------------
  double target_velocity[3], target_twist[3];
  const double current_height = ref_pos[0][2] - min_foot_height;
  {
    double orientation[4] = {ref_q[0][0], ref_q[0][1], ref_q[0][2], ref_q[0][3]}, eul[3];
    for (int c = 0; c < 3; ++c) { target_velocity[c] = ref_vel[0][c]; target_twist[c] = ref_tw[0][c]; }
    q_euler(orientation, eul);
    if (ACT(A_WX)) target_twist[0] = t_w[0];
    if (ACT(A_WY)) target_twist[1] = t_w[1];
    if (ACT(A_WZ)) target_twist[2] = t_w[2];
    if (ACT(A_FULL)) {
      for (int c = 0; c < 4; ++c) orientation[c] = t_q[c];
    } else {
      if (ACT(A_YAW)) eul[2] = t_yaw;
      if (ACT(A_PITCH)) eul[1] = t_pitch;
      if (ACT(A_ROLL)) eul[0] = t_roll;
      e_to_q(eul, orientation);
    }
    if (ACT(A_HXD) || ACT(A_HYD)) {
      const double hv[3] = {ACT(A_HXD) ? t_hx : 0.0, ACT(A_HYD) ? t_hy : 0.0, 0.0};
      rot_v(orientation, hv, target_velocity);
    } else {
      if (ACT(A_XD)) target_velocity[0] = t_xd;
      if (ACT(A_YD)) target_velocity[1] = t_yd;
    }
    if (ACT(A_ZD)) target_velocity[2] = t_zd;
    else if (ACT(A_HZD)) target_velocity[2] = t_hz;
  }

  // ---- Raibert planner state: constructor (raibert_foot_step_planner.cpp:25-34) on the first call -----------------------------------
  if (St[ST_RINIT] == 0.0) {
    const double mz = fmin(fmin(S[IN_FEET + 2], S[IN_FEET + 5]), fmin(S[IN_FEET + 8], S[IN_FEET + 11]));
    for (int l = 0; l < 4; ++l) St[ST_RFH + l] = mz;
    St[ST_RINIT] = 1.0;
  }

  // ---- contact schedule, rows 0..N ----------------------------------------------------------------------------------------------
  unsigned cmask[ROWS];
  double t_stance[4];
  if (O.mode == 0) {
    // SimpleGaitSequencer::UpdateState (simple_gait_sequencer.cpp:62-80): phase from the clock; a pure standing gait keeps phase 0
    const double T = S[IN_PER];
    bool moving = false;
    for (int l = 0; l < 4; ++l) moving |= (S[IN_DUTY + l] != 0.0 && S[IN_DUTY + l] != 1.0);
    double phase;
    if (St[ST_TINIT] != 0.0 && moving) {
      const double time_in_phase = (S[IN_TIME] - St[ST_T0]) / 1e9;  // duration_cast<duration<double>>(ns)
      phase = fmod(time_in_phase / T, 1.0);
    } else {
      phase = 0.0;
      St[ST_T0] = S[IN_TIME];
      St[ST_TINIT] = 1.0;
    }
    bool early[4] = {false, false, false, false};
    for (int i = 0; i < steps; ++i) {  // Gait::update_sequence (gait.cpp:67-112)
      unsigned m = 0;
      for (int leg = 0; leg < 4; ++leg) {
        const double d = S[IN_DUTY + leg], off = S[IN_OFF + leg];
        const double leg_phase = fmod(((phase + 1.0) - off) + ((double)(unsigned)i * dt) / T, 1.0);
        bool c = leg_phase < d;
        if (i == 0 && O.early_contact_detection) {
          if (meas[leg] && !c && (leg_phase > ((d + 1.0) / 2.0))) early[leg] = true;
        }
        if (c) early[leg] = false;
        else if (early[leg]) c = true;
        if (c) m |= 1u << leg;
      }
      cmask[i] = m;
    }
    for (int l = 0; l < 4; ++l) t_stance[l] = S[IN_DUTY + l] * T;  // Gait::get_t_stance
  } else {
    // ---- AdaptiveGait::update_sequence (gait.cpp:411-487) ---------------------------------------------------------------------
    if (St[ST_AINIT] == 0.0) {  // AdaptiveGait::AdaptiveGait (gait.cpp:363-409)
      for (int l = 0; l < 4; ++l) {
        St[ST_APH + l] = fmod(1.0 - O.phase_offset[l], 1.0);
        St[ST_ADF + l] = 1.0;
        St[ST_APO + l] = O.phase_offset[l];
        St[ST_ANPO + l] = O.phase_offset[l];
      }
      St[ST_AT] = O.correction_period;
      St[ST_ADEL] = -1.0;
      St[ST_ASW] = O.switch_offsets ? 1.0 : 0.0;
      St[ST_AINIT] = 1.0;
    }
    const double swing_time = O.swing_time, gdt = dt;
    // update_params (gait.cpp:507-577); `member` = the call on the gait's own members with the raibert planner at hand
    auto update_params = [&](double &T, double *df, double *po, double &rem, double dtt, double v, double h, bool sw, bool member) {
      const double Fr = froude(v, h);
      double stride = (2.3 * pow(Fr, 0.3)) * h;
      if (O.max_stride_length > EPS) stride = fmin(stride, O.max_stride_length);
      if (sw && rem < 0.0) {
        const double walk[4] = {0.0, 0.5, 0.75, 0.25}, trot[4] = {0.0, 0.5, 0.5, 0.0};
        adaptive_set_offset(St, O, Fr < 0.02 ? walk : trot);  // acts on the members (and on `po` / `rem` when they are the members)
      }
      if (rem > 0.0) {
        rem -= dtt;
        if (rem <= 0.0)
          for (int l = 0; l < 4; ++l) po[l] = St[ST_ANPO + l];
      }
      if (fabs(v) > (O.zero_velocity_threshold + EPS)) {
        T = stride / fabs(v);
        for (int l = 0; l < 4; ++l) df[l] = (T - swing_time) / T;
      } else {  // standing mode: move a foot that is far from its Raibert foothold
        bool move[4] = {false, false, false, false}, move_any = false;
        if (member) {
          for (int leg = 0; leg < 4; ++leg) {
            const bool in_swing = St[ST_APH + leg] >= df[leg];
            double fh[3], psh[3];
            // calc_p_shoulder rotates with the quaternion it is given: here the FULL desired orientation of row 0
            // get_foothold(leg, ref_pos[0], desired_orientation[0], get_t_stance = T_ - swing_time, ref_vel[0], target_velocity, target_twist)
            const double sh[3] = {M.shoulders[3 * leg], M.shoulders[3 * leg + 1], M.shoulders[3 * leg + 2]};
            rot_v(des_q0, sh, psh);
            for (int c = 0; c < 3; ++c) psh[c] = ref_pos[0][c] + psh[c];
            double hh = ref_pos[0][2] - min4(St + ST_RFH);
            hh = fmin(fmax(hh, 0.0), M.max_leg_length);
            const double *vv = ref_vel[0];
            const double cr[3] = {vv[1] * target_twist[2] - vv[2] * target_twist[1], vv[2] * target_twist[0] - vv[0] * target_twist[2],
                                  vv[0] * target_twist[1] - vv[1] * target_twist[0]};
            const double kc = 0.5 * sqrt(hh / M.raibert_g);
            const double tst = St[ST_AT] - swing_time;
            for (int c = 0; c < 3; ++c) fh[c] = psh[c] + kc * cr[c] + ((tst / 2.0) * vv[c] + M.raibert_k * (vv[c] - target_velocity[c]));
            const double ddx = fh[0] - S[IN_FEET + 3 * leg], ddy = fh[1] - S[IN_FEET + 3 * leg + 1];
            const double dist = sqrt(ddx * ddx + ddy * ddy);
            move[leg] = in_swing || (dist > O.standing_foot_position_threshold && O.standing_foot_position_threshold > EPS);
            move_any = move_any || move[leg];
          }
        }
        T = O.correction_period;
        for (int l = 0; l < 4; ++l) df[l] = (move[l] || (move_any && O.correct_all)) ? (T - swing_time) / T : 1.0;
      }
    };
    double df_old[4];
    for (int l = 0; l < 4; ++l) df_old[l] = St[ST_ADF + l];
    const double wz = ref_tw[0][2];
    double vf[2];
    const double v2[2] = {ref_vel[0][0], ref_vel[0][1]};
    moving_average<2>(St + ST_VCNT, St + ST_VIDX, St + ST_VSUM, St + ST_VBUF, O.filter_size, v2, vf);
    double v = sqrt(vf[0] * vf[0] + vf[1] * vf[1]);
    const double v_cmd = sqrt(target_velocity[0] * target_velocity[0] + target_velocity[1] * target_velocity[1]);
    v = fmax(v, fabs(wz * 0.3));
    v = fmax(v, v_cmd * O.min_v_cmd_factor);
    {  // disturbance_factor (gait.cpp:489-505)
      double vb[3], vc[3], vd[3];
      rot_vt(ref_q[0], ref_vel[0], vb);
      rot_vt(ref_q[0], target_velocity, vc);
      for (int c = 0; c < 3; ++c) vd[c] = fmax(fabs(vb[c]) - fabs(vc[c]), 0.0);
      vd[2] = 0.0;
      v += O.disturbance_correction * sqrt(vd[0] * vd[0] + vd[1] * vd[1] + vd[2] * vd[2]);
    }
    v = fmax(v, O.min_v);
    double h = current_height;
    update_params(St[ST_AT], St + ST_ADF, St + ST_APO, St[ST_ADEL], O.control_dt, v, h, St[ST_ASW] != 0.0, true);
    adaptive_next_phase(St + ST_APH, O.control_dt, St[ST_AT], St + ST_APO, df_old, St + ST_ADF, O.early_contact_detection ? meas : nullptr,
                        O.max_correction_cycles);
    double df[4], ph[4], po[4], T = St[ST_AT], rem = St[ST_ADEL];
    for (int l = 0; l < 4; ++l) { df[l] = St[ST_ADF + l]; ph[l] = St[ST_APH + l]; po[l] = St[ST_APO + l]; }
    double swing_prev[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned prev_m = 0;
    for (int i = 0; i < steps; ++i) {
      if (i > 0) {
        v = sqrt(ref_vel[i][0] * ref_vel[i][0] + ref_vel[i][1] * ref_vel[i][1]);
        v = fmax(v, O.min_v);
        h = current_height;
        for (int l = 0; l < 4; ++l) df_old[l] = df[l];
        update_params(T, df, po, rem, gdt, v, h, false, false);
        adaptive_next_phase(ph, gdt, T, po, df_old, df, nullptr, O.max_correction_cycles);
      }
      unsigned m = 0;
      for (int leg = 0; leg < 4; ++leg) {
        bool c = ph[leg] < df[leg];
        const bool prev_c = (prev_m >> leg) & 1u;
        if (c && i > 0 && !prev_c) c = swing_prev[leg] <= gdt;
        double sw;
        if (c) sw = 0.0;
        else {
          if (i == 0 || prev_c) sw = (1 - ph[leg]) * T;
          else {
            sw = swing_prev[leg] - gdt;
            if (sw < 0.0) {
              c = true;
              sw = 0.0;
            }
          }
        }
        swing_prev[leg] = sw;
        if (c) m |= 1u << leg;
      }
      cmask[i] = m;
      prev_m = m;
    }
    for (int l = 0; l < 4; ++l) t_stance[l] = St[ST_AT] - swing_time;  // AdaptiveGait::get_t_stance
  }
  for (int i = 0; i < N; ++i)
    for (int leg = 0; leg < 4; ++leg) ct[4 * i + leg] = (cmask[i] >> leg) & 1u;

  // ---- RaibertFootStepPlanner::get_foot_position_sequence (:43-129), z_on_plane = false ---------------------------------------
  double vfilt[3];
  {
    const double vin[3] = {S[10], S[11], S[12]};
    moving_average<3>(St + ST_FCNT, St + ST_FIDX, St + ST_FSUM, St + ST_FBUF, O.raibert_filtersize, vin, vfilt);
  }
  double *fp = rec + 26 + 13 * steps;
  bool resume[4] = {true, false, false, false};
  double prevfp[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const double L = M.max_leg_length;
  for (int i = 0; i < steps; ++i) {
    for (int leg = 0; leg < 4; ++leg) {
      double out[3] = {0.0, 0.0, 0.0};
      if ((cmask[i] >> leg) & 1u) {
        if (i == 0) {
          for (int c = 0; c < 3; ++c) out[c] = S[IN_FEET + 3 * leg + c];
          St[ST_RFH + leg] = out[2];
        } else if (resume[leg]) {
          for (int c = 0; c < 3; ++c) out[c] = prevfp[leg][c];
        } else {
          double yq[4], psh[3], pp[3];
          yaw_q(q_yaw(ref_q[i]), yq);
          raibert_foothold(M, leg, ref_pos[i], yq, t_stance[leg], vfilt, ref_vel[i], ref_tw[i], St + ST_RFH, pp, psh);
          pp[2] = St[ST_RFH + leg];
          const double dif[3] = {pp[0] - psh[0], pp[1] - psh[1], pp[2] - psh[2]};
          const double nrm = sqrt(dif[0] * dif[0] + dif[1] * dif[1] + dif[2] * dif[2]);
          if (nrm > L && fabs(dif[2]) <= L) {  // clip to the reachable sphere keeping dx/dy (:97-113)
            double nd[3];
            if (fabs(dif[0]) > fabs(dif[1])) {
              nd[0] = -sqrt(-(dif[2] - L) * (dif[2] + L) / ((dif[1] / dif[0]) * (dif[1] / dif[0]) + 1));
              nd[1] = (dif[1] / dif[0]) * nd[0];
            } else {
              nd[1] = -sqrt(-(dif[2] - L) * (dif[2] + L) / ((dif[0] / dif[1]) * (dif[0] / dif[1]) + 1));
              nd[0] = (dif[0] / dif[1]) * nd[1];
            }
            nd[2] = dif[2];
            for (int c = 0; c < 3; ++c) pp[c] = psh[c] + nd[c];
          }
          for (int c = 0; c < 3; ++c) out[c] = pp[c];
        }
        resume[leg] = true;
      } else {
        resume[leg] = false;
      }
      for (int c = 0; c < 3; ++c) prevfp[leg][c] = out[c];
      if (i < N)
        for (int c = 0; c < 3; ++c) fp[12 * i + 3 * leg + c] = out[c];
    }
  }
#undef ACT
}

#ifndef B200QP_HOST_EMULATION  // tests/host_emu compiles the kernel body as host code for logic checks; the launch is device-only
cudaError_t launch_sequencer_full(int n, const SeqModel &M, const SeqOptions &O, const double *seq_in, const uint8_t *measured,
                                  double *state, double *rec_out, uint8_t *contact_out, int in_stride, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  sequencer_full_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, M, O, seq_in, measured, state, rec_out, contact_out, in_stride);
  return cudaGetLastError();
}
#endif
}  // namespace b200qp
