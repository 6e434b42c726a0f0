// Shared device helpers for the b200qp kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200qp.h"

namespace b200qp {

constexpr int NX = B200QP_NX;
constexpr int NU = B200QP_NU;
constexpr int MAXH = B200QP_MAX_HORIZON;
constexpr double GRAVITY_CONSTANT = 9.8067;  // MPC::GRAVITY_CONSTANT, mpc.hpp:22

// Kernel-side copy of b200qp_mpc_config (passed by value in the launch).
struct MpcParams {
  int N;
  int in_stride;  // doubles per input record
  double dt, alpha, w[12], mu, fmin, fmax;
  double rho0, sigma, relax, eps, kkt_tol;
  int max_iter, check_every, polish_rounds, ipm_max_iter;
  int precision;  // 0 FP64 (product); 1 / 2: FP32 error study of the condensed kernel (b200qp_mpc_config.precision)
  // warm start of the interior-point kernel (b200qp_mpc_config.warm_start, mpc.cpp:236-237): previous forces per problem slot
  double *ipm_warm_x;        // [max_batch][N][12] or nullptr
  uint8_t *ipm_warm_valid;   // [max_batch]
  double ipm_warm_mu;        // complementarity the warm-started slack / multiplier pairs begin at
  double ipm_tau;            // fraction of the step to the boundary that is taken
  double ipm_cold_mu;        // ... and the cold-started ones (z = mu / s at the interior starting point)
  // second opinion for the interior-point path (solver SPARSE_RICCATI): solved problems whose final residual is above ipm_refine_thr are
  // listed here and re-solved by the condensed ADMM + active-set polish kernel, which overwrites the outputs only if it succeeds
  int *ipm_refine_list;
  int *ipm_refine_count;
  double ipm_refine_thr;
  // ... which starts from the interior-point answer: forces [max_batch][N][12] and duals in its own row layout [max_batch][N][4][7]
  // (3 box rows, 4 pyramid rows), so that its polish gets the active set directly (hot start) and only the K^-1 / ADMM-free path runs
  double *ipm_refine_x;
  double *ipm_refine_y;
  uint8_t *ipm_refine_valid;
  int ipm_flags;             // warp Riccati kernel: bit 0 separate primal / dual step lengths, bit 1 centre two steps when the gap did not fall
};

// model constants of the gait sequencer kernels (b200qp_seq_model + the handle's dt / horizon)
struct SeqModel {
  double mass, inertia[9], com[3], shoulders[12], raibert_k, raibert_g, max_leg_length, dt;
  int N;
};
// mirror of b200qp_seq_options (include/b200qp.h)
struct SeqOptions {
  int mode, raibert_filtersize, fix_standing_position, early_contact_detection;
  double dist_thr, ang_thr, vel_thr;
  // AdaptiveGait constructor arguments (gait.hpp:123-139)
  double swing_time, zero_velocity_threshold, standing_foot_position_threshold, min_v_cmd_factor, max_correction_cycles,
      correction_period, disturbance_correction, min_v, offset_delay, max_stride_length, phase_offset[4], control_dt;
  int filter_size, switch_offsets, correct_all;
};

struct MpcBuffers {
  const double *in;        // [n][in_stride]
  const uint8_t *contact;  // [n][N][4]
  double *forces;          // [n][N][12]
  double *xpred;           // [n][N+1][13]
  b200qp_solver_info *info;
  double *lam_box;  // [n][N][12] or nullptr
  double *lam_g;    // [n][N][16] or nullptr
  // work distribution
  const int *bin_list;  // problem indices of this bin
  const int *bin_count;
  int *bin_next;     // atomic work counter
  double *hscratch;  // [gridDim.x][BS*BS] condensed Hessian per resident block
  // solver AUTO: problems the condensed kernel could not solve (iteration cap) are appended here for a second opinion by the
  // Riccati kernel; nullptr otherwise
  int *retry_list;
  int *retry_count;
  int refine;  // 1: this launch re-solves problems that already have a solution (second opinion): a failure leaves info untouched
  // chunked input streaming (host-buffer entry point): ready[c] != 0 once the records of chunk c = prob / ready_chunk have
  // landed in device memory; nullptr = everything is resident
  const int *ready;
  int ready_chunk;
  // warm / hot start (b200qp_mpc_config.warm_start): previous solution per problem slot, forces layout [n][N][12] and ADMM
  // duals [n][N][4][7]; warm_mode 0 = off (buffers may be nullptr)
  double *warm_x;
  double *warm_y;
  uint8_t *warm_valid;
  int warm_mode;
  // debug (parity hook)
  int dbg_index;
  double *dbg_H;
  double *dbg_g;
  long long *dbg_cycles;  // [n][10] per-phase cycles of thread 0, or nullptr
};

// Eigen::Quaternion::toRotationMatrix for q = (x, y, z, w); R is row-major [3][3].
__device__ __forceinline__ void quat_to_rot(const double *q, double R[3][3]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.0 - (tyy + tzz);
  R[0][1] = txy - twz;
  R[0][2] = txz + twy;
  R[1][0] = txy + twz;
  R[1][1] = 1.0 - (txx + tzz);
  R[1][2] = tyz - twx;
  R[2][0] = txz - twy;
  R[2][1] = tyz + twx;
  R[2][2] = 1.0 - (txx + tyy);
}

// matrix_to_euler, quaternion_operations.hpp:106-124 (same gimbal branches), returns roll, pitch, yaw.
__device__ __forceinline__ void rot_to_euler(const double R[3][3], double e[3]) {
  const double eps = 2.220446049250313e-16;
  const double m20 = R[2][0];
  const double pitch = -asin(m20);
  double yaw, roll;
  if ((1.0 - eps <= m20) && (m20 <= 1.0 + eps)) {
    yaw = 0.0;
    roll = atan2(-R[0][1], -R[0][2]);
  } else if (m20 <= -1.0 + eps) {
    yaw = 0.0;
    roll = atan2(R[0][1], R[0][2]);
  } else {
    yaw = atan2(R[1][0], R[0][0]);
    roll = atan2(R[2][1], R[2][2]);
  }
  e[0] = roll;
  e[1] = pitch;
  e[2] = yaw;
}

// max over the block of V values per thread; result valid in all threads. red must hold V * 32 doubles.
template <int V>
__device__ __forceinline__ void block_max(double (&v)[V], double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < V; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = fmax(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
  }
  __syncthreads();  // protect red from a previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) red[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < V; ++k) {
    double m = red[k * 32];
    for (int wv = 1; wv < nw; ++wv) m = fmax(m, red[k * 32 + wv]);
    v[k] = m;
  }
}

// 1/sqrt(x) for x > 0: FP32 seed + two Newton steps in FP64 (relative error ~1e-15); ~10 instructions instead of the
// library rsqrt()/sqrt()+division sequences, which sit on the serial critical path of the small Cholesky factorisations.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r = (double)rsqrtf((float)x);
  r = r * fma(-0.5 * x, r * r, 1.5);
  r = r * fma(-0.5 * x, r * r, 1.5);
  return r;
}

// 1 / x: hardware approximation (MUFU.RCP64H, ~20 bits) + two Newton steps (relative error ~1e-16) - a handful of instructions
// instead of the IEEE division sequence with its slow-path call; x must be finite, non-zero and not denormal.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// NaN-propagating abs-max helper: fmax drops NaNs, so NaN detection is done separately.
__device__ __forceinline__ double absmax(double a, double b) { return fmax(fabs(a), fabs(b)); }

}  // namespace b200qp
