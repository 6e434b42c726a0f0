// b200qp_mpc.hpp - header-only C++17 host mirror of the reference's MPC plugin over the C ABI (include/b200qp.h).
//
// Mirrors `class MPC : public MPCInterface` (ws/src/controllers/include/mit_controller/mpc.hpp:13-144,
// mpc_interface.hpp:22-33) for a BATCH of robots, without Eigen/ROS so it compiles anywhere: the reference's
// StateInterface / ModelInterface / GaitSequence arguments become plain structs holding exactly the fields
// MPC::GetWrenchSequence reads (mpc.cpp:332-396).  A maintainer of the reference tree wraps this class (or the C ABI
// directly) in an `MPCInterface` subclass - see INTEGRATION.md.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200qp.h"

namespace b200qp {

struct BodyState {            // StateInterface getters used at mpc.cpp:332,392-395
  std::array<double, 4> orientation_xyzw;
  std::array<double, 3> position, angular_vel_world, linear_vel_world;
};
struct BodyModel {            // ModelInterface getters used at mpc.cpp:335,357-358
  double mass;
  std::array<double, 9> inertia_colmajor;
  std::array<double, 3> body_to_com;
};
struct GaitSequenceView {     // GaitSequence fields read by the MPC (gait_sequence.hpp:21-29), rows 0..N (N for per-input fields)
  const uint8_t *contact_sequence;                          // [N][4]
  const double *foot_position_sequence;                     // [N][4][3]
  const double *desired_reference_trajectory_orientation;   // [N+1][4] xyzw
  const double *desired_reference_trajectory_position;      // [N+1][3]
  const double *reference_trajectory_twist;                 // [N+1][3]
  const double *reference_trajectory_velocity;              // [N+1][3]
};
struct SolverInformation {    // mpc_interface.hpp:10-20
  bool success;
  int return_code;
  int acados_num_iter;
  double kkt[4];
};

class BatchedMPC {
 public:
  // Arguments follow MPC::MPC (mpc.hpp:117-127); `solver` is B200QP_SOLVER_* (AUTO picks the kernel per problem),
  // `warm_start` as the reference's last constructor argument (0 cold, 1 warm, 2 hot start per batch slot), `condensing_N`
  // as the reference's cond_N (B200QP_SOLVER_PARTIAL_CONDENSING; 0 = library default).
  BatchedMPC(double alpha, const std::array<double, 12> &state_weights, double mu, double fmin, double fmax, int solver,
             int horizon, double dt, int max_batch, int device = 0, int warm_start = 0, int condensing_N = 0)
      : N_(horizon), max_batch_(max_batch) {
    b200qp_mpc_config c{};
    c.horizon = horizon; c.solver = solver; c.max_batch = max_batch; c.device = device; c.warm_start = warm_start;
    c.cond_N = condensing_N;
    c.dt = dt; c.alpha = alpha; c.mu = mu; c.fmin = fmin; c.fmax = fmax;
    for (int i = 0; i < 12; ++i) c.state_weights[i] = state_weights[i];
    check(b200qp_mpc_create(&c, &h_), "b200qp_mpc_create");
    in_.resize((size_t)max_batch * b200qp_mpc_in_doubles(horizon));
    contact_.resize((size_t)max_batch * horizon * 4);
    info_.resize(max_batch);
  }
  ~BatchedMPC() { b200qp_mpc_destroy(h_); }
  void SetWarmStart(int mode) { check(b200qp_mpc_set_warm_start(h_, mode), "b200qp_mpc_set_warm_start"); }
  BatchedMPC(const BatchedMPC &) = delete;
  BatchedMPC &operator=(const BatchedMPC &) = delete;

  // MPCInterface::UpdateState / UpdateModel / UpdateGaitSequence for robot `i` of the batch (deep copies, mpc.cpp:284-292)
  void UpdateState(int i, const BodyState &s) {
    double *r = rec(i);
    for (int k = 0; k < 4; ++k) r[k] = s.orientation_xyzw[k];
    for (int k = 0; k < 3; ++k) { r[4 + k] = s.position[k]; r[7 + k] = s.angular_vel_world[k]; r[10 + k] = s.linear_vel_world[k]; }
  }
  void UpdateModel(int i, const BodyModel &m) {
    double *r = rec(i);
    r[13] = m.mass;
    for (int k = 0; k < 9; ++k) r[14 + k] = m.inertia_colmajor[k];
    for (int k = 0; k < 3; ++k) r[23 + k] = m.body_to_com[k];
  }
  void UpdateGaitSequence(int i, const GaitSequenceView &g) {
    double *r = rec(i) + 26;
    for (int k = 0; k <= N_; ++k, r += 13) {
      for (int c = 0; c < 4; ++c) r[c] = g.desired_reference_trajectory_orientation[4 * k + c];
      for (int c = 0; c < 3; ++c) {
        r[4 + c] = g.desired_reference_trajectory_position[3 * k + c];
        r[7 + c] = g.reference_trajectory_twist[3 * k + c];
        r[10 + c] = g.reference_trajectory_velocity[3 * k + c];
      }
    }
    for (int e = 0; e < 12 * N_; ++e) r[e] = g.foot_position_sequence[e];
    for (int e = 0; e < 4 * N_; ++e) contact_[(size_t)i * N_ * 4 + e] = g.contact_sequence[e];
  }
  // MPCInterface::GetWrenchSequence for robots 0..n-1: forces [n][N][4][3] (WrenchSequence), xpred [n][N+1][13]
  // (MPCPrediction::raw_data).  Failed robots keep their previous output (mpc.cpp:447-455).
  void GetWrenchSequence(int n, double *forces, double *xpred, SolverInformation *info) {
    check(b200qp_mpc_solve_batch(h_, n, in_.data(), contact_.data(), forces, xpred, info_.data()), "b200qp_mpc_solve_batch");
    for (int i = 0; i < n; ++i) {
      info[i].success = info_[i].status == B200QP_ACADOS_SUCCESS;
      info[i].return_code = info_[i].status;
      info[i].acados_num_iter = info_[i].iterations;
      for (int k = 0; k < 4; ++k) info[i].kkt[k] = info_[i].kkt[k];
    }
  }
  // mpc.hpp:138-141
  void SetStateWeights(const std::array<double, 12> &w) { check(b200qp_mpc_set_state_weights(h_, w.data()), "set_state_weights"); }
  void SetInputWeights(double alpha) { check(b200qp_mpc_set_input_weights(h_, alpha), "set_input_weights"); }
  void SetFmax(double fmax) { check(b200qp_mpc_set_fmax(h_, fmax), "set_fmax"); }
  void SetMu(double mu) { check(b200qp_mpc_set_mu(h_, mu), "set_mu"); }
  b200qp_mpc *handle() { return h_; }

 private:
  double *rec(int i) {
    if (i < 0 || i >= max_batch_) throw std::out_of_range("robot index");
    return in_.data() + (size_t)i * b200qp_mpc_in_doubles(N_);
  }
  static void check(int rc, const char *what) {
    if (rc != B200QP_OK) throw std::runtime_error(std::string(what) + ": " + b200qp_last_error());
  }
  int N_, max_batch_;
  b200qp_mpc *h_ = nullptr;
  std::vector<double> in_;
  std::vector<uint8_t> contact_;
  std::vector<b200qp_solver_info> info_;
};

}  // namespace b200qp
