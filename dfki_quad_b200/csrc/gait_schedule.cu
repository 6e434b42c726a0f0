// Kernel K4: gait schedule -> contact mask -> swing time -> force-box indexing, bit-exact with the strict IEEE-754
// evaluation of the reference expressions.  Compiled with -fmad=false (no contraction) and no fast-math.
//   Gait::update_sequence        ws/src/controllers/src/mit_controller/gait.cpp:67-112
//   MPC::GetUBounds              ws/src/controllers/src/mit_controller/mpc.cpp:94-115
// One thread per (robot, leg); the loop over steps is sequential because the early-contact override
// (gait.cpp:88-99) carries state from step to step.
#include <cstdlib>

#include "common.cuh"

namespace b200qp {

__global__ void gait_schedule_kernel(int n, int steps, double dt, const double *__restrict__ period,
                                     const double *__restrict__ duty, const double *__restrict__ offset,
                                     const double *__restrict__ phase, const uint8_t *__restrict__ measured,
                                     double fmin, double fmax, uint8_t *__restrict__ contact,
                                     double *__restrict__ swing_time, double *__restrict__ lbu,
                                     double *__restrict__ ubu) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * n) return;
  const int p = t >> 2, leg = t & 3;
  const double T = period[p], d = duty[t], off = offset[t], ph = phase[p];
  const bool meas = measured != nullptr && measured[t] != 0;
  const bool has_meas = measured != nullptr;
  bool early = false;
  const double base = (ph + 1.0) - off;  // (phase + 1.0 - phase_offset_[leg]), left to right
  for (int i = 0; i < steps; ++i) {
    const double leg_phase = fmod(base + ((double)(unsigned)i * dt) / T, 1.0);  // gait.cpp:82
    bool c = leg_phase < d;                                                      // gait.cpp:83
    if (i == 0 && has_meas) {                                                    // gait.cpp:85-90
      if (meas && !c && (leg_phase > ((d + 1.0) / 2.0))) early = true;
    }
    if (c) early = false;  // gait.cpp:92-96
    else if (early) c = true;
    const size_t o = ((size_t)p * steps + i) * 4 + leg;
    contact[o] = c ? 1 : 0;
    if (swing_time != nullptr) swing_time[o] = (!c && d > 0) ? (1 - leg_phase) * T : 0.0;  // gait.cpp:98-103
    if (lbu != nullptr) {  // mpc.cpp:103-113
      const size_t b = ((size_t)p * steps + i) * 12 + 3 * leg;
      if (c) {
        lbu[b] = -fmax; lbu[b + 1] = -fmax; lbu[b + 2] = fmin;
        ubu[b] = fmax;  ubu[b + 1] = fmax;  ubu[b + 2] = fmax;
      } else {
        lbu[b] = 0.0; lbu[b + 1] = 0.0; lbu[b + 2] = 0.0;
        ubu[b] = 0.0; ubu[b + 1] = 0.0; ubu[b + 2] = 0.0;
      }
    }
  }
}

cudaError_t launch_gait_schedule(int n, int steps, double dt, const double *period, const double *duty,
                                 const double *offset, const double *phase, const uint8_t *measured, double fmin,
                                 double fmax, uint8_t *contact, double *swing_time, double *lbu, double *ubu,
                                 cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 128;
  gait_schedule_kernel<<<(4 * n + threads - 1) / threads, threads, 0, st>>>(n, steps, dt, period, duty, offset, phase,
                                                                           measured, fmin, fmax, contact, swing_time,
                                                                           lbu, ubu);
  return cudaGetLastError();
}

}  // namespace b200qp

// ---- FP64 FMA peak microbenchmark (roofline denominator for the on-chip solve kernels; MEASURED_PEAKS.json has no
// FP64 entry).  8 independent DFMA chains per thread, 1024 threads per block, 8 blocks per SM.
namespace b200qp {
__global__ void fp64_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
cudaError_t fp64_peak(int num_sms, double *tflops) {
  const int blocks = num_sms * 2, threads = 1024, iters = 20000;
  double *d = nullptr;
  cudaError_t e = cudaMalloc(&d, (size_t)blocks * threads * sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads>>>(d, 100, 1.0);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double t = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (t > best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return e;
}
}  // namespace b200qp

// ---- on-device gait sequencer: one thread per robot turns (state, Target, gait, phase) into
// the MPC input record + contact mask, so the host->device boundary shrinks from 2.3 KB to 0.4 KB per robot.
//   SimpleGaitSequencer::GetGaitSequence            simple_gait_sequencer.cpp:43-60
//   MPCTrajectoryPlanner::plan_trajectory            mpc_trajectory_planner.cpp:28-140  (Target activation of the node,
//       mit_controller_node.cpp:145-165: hybrid_x_dot, hybrid_y_dot, wz, wx, wy, roll, pitch, z active)
//   MPCTrajectoryPlanner::plan_desired_trajectory    mpc_trajectory_planner.cpp:142-264 (first call, no latching)
//   Gait::update_sequence                            gait.cpp:67-112
//   RaibertFootStepPlanner::get_foot_position_sequence / get_foothold   raibert_foot_step_planner.cpp:43-169
namespace b200qp {


namespace {
__device__ inline double q_yaw(const double *q) {
  double R[3][3];
  quat_to_rot(q, R);
  return atan2(R[1][0], R[0][0]);
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
  if (ad < 1.0 - 2.220446049250313e-16) {
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
}  // namespace

// seq_in per robot (50 doubles): quat 0-3, pos 4-6, omega 7-9, vel 10-12, feet 13-24, v_filtered 25-27, last foot heights 28-31,
// target 32-39 (hybrid_x_dot, hybrid_y_dot, wz, wx, wy, roll, pitch, z), phase 40, period 41, duty 42-45, offset 46-49
__global__ void sequencer_kernel(int n, SeqModel M, const double *__restrict__ seq_in, const uint8_t *__restrict__ measured,
                                 double *__restrict__ rec_out, uint8_t *__restrict__ contact_out, int in_stride) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const double *S = seq_in + (size_t)p * 50;
  const int N = M.N, steps = N + 1;
  const double dt = M.dt;
  double *rec = rec_out + (size_t)p * in_stride;
  uint8_t *ct = contact_out + (size_t)p * N * 4;
  // ---- contact schedule, rows 0..N (row N is only needed by the foot-step logic)
  unsigned cmask[MAXH + 1];
  {
    bool early[4] = {false, false, false, false};
    for (int i = 0; i < steps; ++i) {
      unsigned m = 0;
      for (int leg = 0; leg < 4; ++leg) {
        const double T = S[41], d = S[42 + leg], off = S[46 + leg];
        const double leg_phase = fmod(((S[40] + 1.0) - off) + ((double)(unsigned)i * dt) / T, 1.0);
        bool c = leg_phase < d;
        if (i == 0 && measured != nullptr) {
          if (measured[4 * p + leg] != 0 && !c && (leg_phase > ((d + 1.0) / 2.0))) early[leg] = true;
        }
        if (c) early[leg] = false;
        else if (early[leg]) c = true;
        if (c) m |= 1u << leg;
      }
      cmask[i] = m;
      if (i < N)
        for (int leg = 0; leg < 4; ++leg) ct[4 * i + leg] = (m >> leg) & 1u;
    }
  }
  // ---- header: state + model
  for (int c = 0; c < 13; ++c) rec[c] = S[c];
  rec[13] = M.mass;
  for (int c = 0; c < 9; ++c) rec[14 + c] = M.inertia[c];
  for (int c = 0; c < 3; ++c) rec[23 + c] = M.com[c];
  const double hx = S[32], hy = S[33];
  const double twc[3] = {S[35], S[36], S[34]};
  const double troll = S[37], tpitch = S[38], tz = S[39];
  double min_h = fmin(fmin(S[13 + 2], S[16 + 2]), fmin(S[19 + 2], S[22 + 2]));
  // ---- reference ("ref") and desired ("des") trajectories; both recursions coincide on the first call except that the
  // reference twist of row 0 is the measured one.  Per row we also need ref position / yaw-quaternion / velocity / twist for
  // the foot-step planner, kept in small per-row arrays.
  double ref_pos[MAXH + 1][3], ref_yawq[MAXH + 1][4], ref_vel[MAXH + 1][3];
  {
    double R0[3][3], eul[3], position[3], quat_prev[4];
    quat_to_rot(S, R0);
    rot_to_euler(R0, eul);
    for (int c = 0; c < 3; ++c) position[c] = S[4 + c];
    for (int c = 0; c < 4; ++c) quat_prev[c] = S[c];
    // row 0
    double *st0 = rec + 26;
    for (int c = 0; c < 4; ++c) st0[c] = S[c];
    for (int c = 0; c < 3; ++c) { st0[4 + c] = S[4 + c]; st0[7 + c] = S[7 + c]; st0[10 + c] = S[10 + c]; }
    for (int c = 0; c < 3; ++c) { ref_pos[0][c] = S[4 + c]; ref_vel[0][c] = S[10 + c]; }
    yaw_q(q_yaw(S), ref_yawq[0]);
    for (int i = 1; i < steps; ++i) {
      for (int c = 0; c < 3; ++c) eul[c] = eul[c] + dt * twc[c];
      eul[1] = tpitch;
      eul[0] = troll;
      double q[4], yq_prev[4], yq[4], mid[4], step[3], dv[3] = {dt * hx, dt * hy, 0.0};
      e_to_q(eul, q);
      yaw_q(q_yaw(quat_prev), yq_prev);
      yaw_q(q_yaw(q), yq);
      slerp_half(yq_prev, yq, mid);
      rot_v(mid, dv, step);
      for (int c = 0; c < 3; ++c) position[c] = position[c] + step[c];
      position[2] = tz + min_h;
      double hv[3] = {hx, hy, 0.0}, vel[3];
      rot_v(q, hv, vel);
      double *st = rec + 26 + 13 * i;
      for (int c = 0; c < 4; ++c) st[c] = q[c];
      for (int c = 0; c < 3; ++c) { st[4 + c] = position[c]; st[7 + c] = twc[c]; st[10 + c] = vel[c]; }
      for (int c = 0; c < 3; ++c) { ref_pos[i][c] = position[c]; ref_vel[i][c] = vel[c]; }
      for (int c = 0; c < 4; ++c) { ref_yawq[i][c] = yq[c]; quat_prev[c] = q[c]; }
    }
  }
  // ---- Raibert foot positions (z_on_plane = false)
  double *fp = rec + 26 + 13 * steps;
  double heights[4] = {S[28], S[29], S[30], S[31]};
  bool resume[4] = {true, true, true, true};
  double prevfp[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  const double L = M.max_leg_length;
  for (int i = 0; i < steps; ++i) {
    for (int leg = 0; leg < 4; ++leg) {
      double out[3] = {0.0, 0.0, 0.0};
      if ((cmask[i] >> leg) & 1u) {
        if (i == 0) {
          for (int c = 0; c < 3; ++c) out[c] = S[13 + 3 * leg + c];
          heights[leg] = out[2];
        } else if (resume[leg]) {
          for (int c = 0; c < 3; ++c) out[c] = prevfp[leg][c];
        } else {
          double psh[3], sh[3] = {M.shoulders[3 * leg], M.shoulders[3 * leg + 1], M.shoulders[3 * leg + 2]};
          rot_v(ref_yawq[i], sh, psh);
          for (int c = 0; c < 3; ++c) psh[c] = ref_pos[i][c] + psh[c];
          double h = ref_pos[i][2] - fmin(fmin(heights[0], heights[1]), fmin(heights[2], heights[3]));
          h = fmin(fmax(h, 0.0), L);
          const double *v = S + 25;
          const double wcx = (i == 0) ? S[7] : twc[0], wcy = (i == 0) ? S[8] : twc[1], wcz = (i == 0) ? S[9] : twc[2];
          const double cr[3] = {v[1] * wcz - v[2] * wcy, v[2] * wcx - v[0] * wcz, v[0] * wcy - v[1] * wcx};
          const double kc = 0.5 * sqrt(h / M.raibert_g);
          const double tst = S[42 + leg] * S[41];
          double pp[3];
          for (int c = 0; c < 3; ++c) pp[c] = psh[c] + kc * cr[c] + ((tst / 2.0) * v[c] + M.raibert_k * (v[c] - ref_vel[i][c]));
          pp[2] = heights[leg];
          double dif[3] = {pp[0] - psh[0], pp[1] - psh[1], pp[2] - psh[2]};
          const double nrm = sqrt(dif[0] * dif[0] + dif[1] * dif[1] + dif[2] * dif[2]);
          if (nrm > L && fabs(dif[2]) <= L) {  // clip to the reachable sphere keeping dx/dy (raibert_foot_step_planner.cpp:97-113)
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
}

// Same sequencer, one WARP per robot.  The thread-per-robot kernel above is one chain of ~21 stages x ~10 FP64 trigonometric
// calls (0.09 ms for any batch size); here lane i owns stage i: the yaw of stage i is accumulated by the lane itself in the
// reference's order (i additions), its orientation / yaw quaternion / velocity are independent of the other stages, the
// half-way yaw quaternion needs the neighbour's (one shuffle), and the positions are summed per lane in stage order from the
// steps parked in shared memory - the floating-point operations and their order per output are exactly those of the serial
// kernel.  Four lanes (one per leg) then run the stateful parts (early-contact latch, Raibert hold / touch-down logic), and the
// record is written with coalesced stores.  sequencer_kernel stays as the readable statement of the algorithm and as a
// cross-check in the tests of the two kernels against each other.
constexpr int SEQ_WARPS = 4;
__global__ void __launch_bounds__(32 * SEQ_WARPS) sequencer_warp_kernel(int n, SeqModel M, const double *__restrict__ seq_in,
                                                                        const uint8_t *__restrict__ measured,
                                                                        double *__restrict__ rec_out,
                                                                        uint8_t *__restrict__ contact_out, int in_stride) {
  constexpr int RECMAX = 26 + 13 * (MAXH + 1) + 12 * MAXH;
  __shared__ double s_rec[SEQ_WARPS][RECMAX];
  __shared__ double s_in[SEQ_WARPS][52];
  __shared__ double s_step[SEQ_WARPS][MAXH + 1][2];
  __shared__ double s_refpos[SEQ_WARPS][MAXH + 1][3], s_refyq[SEQ_WARPS][MAXH + 1][4], s_refvel[SEQ_WARPS][MAXH + 1][3];
  __shared__ unsigned s_cmask[SEQ_WARPS][MAXH + 1];
  __shared__ double s_h[SEQ_WARPS][4];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int p = blockIdx.x * SEQ_WARPS + w;
  if (p >= n) return;  // whole warps leave together; only __syncwarp below
  const unsigned full = 0xffffffffu;
  const int N = M.N, steps = N + 1;
  const double dt = M.dt;
  double *S = s_in[w];
  for (int e = lane; e < 50; e += 32) S[e] = seq_in[(size_t)p * 50 + e];
  __syncwarp();
  double *rec = s_rec[w];
  // ---- contact schedule: lane = leg (stateful latch, sequential over the stages) -------------------------------------------------
  if (lane < 4) {
    const int leg = lane;
    const double T = S[41], d = S[42 + leg], off = S[46 + leg];
    bool early = false;
    for (int i = 0; i < steps; ++i) {
      const double leg_phase = fmod(((S[40] + 1.0) - off) + ((double)(unsigned)i * dt) / T, 1.0);
      bool c = leg_phase < d;
      if (i == 0 && measured != nullptr) {
        if (measured[4 * p + leg] != 0 && !c && (leg_phase > ((d + 1.0) / 2.0))) early = true;
      }
      if (c) early = false;
      else if (early) c = true;
      const unsigned bits = __ballot_sync(0xfu, c) & 0xfu;
      if (leg == 0) s_cmask[w][i] = bits;
      if (i < N) contact_out[((size_t)p * N + i) * 4 + leg] = c ? 1 : 0;
    }
  }
  // ---- header ------------------------------------------------------------------------------------------------------------------
  if (lane < 13) rec[lane] = S[lane];
  if (lane == 13) rec[13] = M.mass;
  if (lane >= 14 && lane < 23) rec[lane] = M.inertia[lane - 14];
  if (lane >= 23 && lane < 26) rec[lane] = M.com[lane - 23];
  const double hx = S[32], hy = S[33];
  const double twc[3] = {S[35], S[36], S[34]};
  const double troll = S[37], tpitch = S[38], tz = S[39];
  const double min_h = fmin(fmin(S[13 + 2], S[16 + 2]), fmin(S[19 + 2], S[22 + 2]));
  // ---- per-stage orientation, velocity and position step: lane = stage ----------------------------------------------------------------
  double q[4] = {0, 0, 0, 1}, yq[4] = {0, 0, 0, 1}, vel[3] = {0, 0, 0};
  const int i = lane;
  if (i < steps) {
    if (i == 0) {
      for (int c = 0; c < 4; ++c) q[c] = S[c];
      for (int c = 0; c < 3; ++c) vel[c] = S[10 + c];
    } else {
      double R0[3][3], eul[3];
      quat_to_rot(S, R0);
      rot_to_euler(R0, eul);
      for (int t = 0; t < i; ++t) {
        for (int c = 0; c < 3; ++c) eul[c] = eul[c] + dt * twc[c];
        eul[1] = tpitch;
        eul[0] = troll;
      }
      e_to_q(eul, q);
      const double hv[3] = {hx, hy, 0.0};
      rot_v(q, hv, vel);
    }
    yaw_q(q_yaw(q), yq);
  }
  double yp[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) yp[c] = __shfl_up_sync(full, yq[c], 1);
  if (i >= 1 && i < steps) {
    double mid[4], st3[3];
    const double dv[3] = {dt * hx, dt * hy, 0.0};
    slerp_half(yp, yq, mid);
    rot_v(mid, dv, st3);
    s_step[w][i][0] = st3[0];
    s_step[w][i][1] = st3[1];
  }
  __syncwarp();
  if (i < steps) {
    double pos[3] = {S[4], S[5], S[6]};
    for (int t = 1; t <= i; ++t) {  // the reference's summation order
      pos[0] = pos[0] + s_step[w][t][0];
      pos[1] = pos[1] + s_step[w][t][1];
    }
    if (i >= 1) pos[2] = tz + min_h;
    double *st = rec + 26 + 13 * i;
    for (int c = 0; c < 4; ++c) st[c] = q[c];
    for (int c = 0; c < 3; ++c) {
      st[4 + c] = pos[c];
      st[7 + c] = (i == 0) ? S[7 + c] : twc[c];
      st[10 + c] = vel[c];
      s_refpos[w][i][c] = pos[c];
      s_refvel[w][i][c] = vel[c];
    }
    for (int c = 0; c < 4; ++c) s_refyq[w][i][c] = yq[c];
  }
  __syncwarp();
  // ---- Raibert foot positions: lane = leg ---------------------------------------------------------------------------------------------
  if (lane < 4) {
    const int leg = lane;
    double *fp = rec + 26 + 13 * steps;
    // heights only change at stage 0 (stance feet take their current height); every leg needs all four afterwards
    double hleg = S[28 + leg];
    if ((s_cmask[w][0] >> leg) & 1u) hleg = S[13 + 3 * leg + 2];
    s_h[w][leg] = hleg;
    __syncwarp(0xfu);
    const double hmin = fmin(fmin(s_h[w][0], s_h[w][1]), fmin(s_h[w][2], s_h[w][3]));
    bool resume = true;
    double prevfp[3] = {0.0, 0.0, 0.0};
    const double L = M.max_leg_length;
    for (int k = 0; k < steps; ++k) {
      double out[3] = {0.0, 0.0, 0.0};
      if ((s_cmask[w][k] >> leg) & 1u) {
        if (k == 0) {
          for (int c = 0; c < 3; ++c) out[c] = S[13 + 3 * leg + c];
        } else if (resume) {
          for (int c = 0; c < 3; ++c) out[c] = prevfp[c];
        } else {
          double psh[3];
          const double sh[3] = {M.shoulders[3 * leg], M.shoulders[3 * leg + 1], M.shoulders[3 * leg + 2]};
          rot_v(s_refyq[w][k], sh, psh);
          for (int c = 0; c < 3; ++c) psh[c] = s_refpos[w][k][c] + psh[c];
          double h = s_refpos[w][k][2] - hmin;
          h = fmin(fmax(h, 0.0), L);
          const double *v = S + 25;
          const double wcx = twc[0], wcy = twc[1], wcz = twc[2];  // k >= 1 here
          const double cr[3] = {v[1] * wcz - v[2] * wcy, v[2] * wcx - v[0] * wcz, v[0] * wcy - v[1] * wcx};
          const double kc = 0.5 * sqrt(h / M.raibert_g);
          const double tst = S[42 + leg] * S[41];
          double pp[3];
          for (int c = 0; c < 3; ++c) pp[c] = psh[c] + kc * cr[c] + ((tst / 2.0) * v[c] + M.raibert_k * (v[c] - s_refvel[w][k][c]));
          pp[2] = hleg;
          const double dif[3] = {pp[0] - psh[0], pp[1] - psh[1], pp[2] - psh[2]};
          const double nrm = sqrt(dif[0] * dif[0] + dif[1] * dif[1] + dif[2] * dif[2]);
          if (nrm > L && fabs(dif[2]) <= L) {
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
        resume = true;
      } else {
        resume = false;
      }
      for (int c = 0; c < 3; ++c) prevfp[c] = out[c];
      if (k < N)
        for (int c = 0; c < 3; ++c) fp[12 * k + 3 * leg + c] = out[c];
    }
  }
  __syncwarp();
  double *dst = rec_out + (size_t)p * in_stride;
  for (int e = lane; e < in_stride; e += 32) dst[e] = rec[e];
}

cudaError_t launch_sequencer(int n, const SeqModel &M, const double *seq_in, const uint8_t *measured, double *rec_out,
                             uint8_t *contact_out, int in_stride, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (getenv("B200QP_SERIAL_SEQUENCER") != nullptr)  // the thread-per-robot statement of the same algorithm (cross-check)
    sequencer_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, M, seq_in, measured, rec_out, contact_out, in_stride);
  else
    sequencer_warp_kernel<<<(n + SEQ_WARPS - 1) / SEQ_WARPS, 32 * SEQ_WARPS, 0, st>>>(n, M, seq_in, measured, rec_out, contact_out,
                                                                                   in_stride);
  return cudaGetLastError();
}
}  // namespace b200qp
