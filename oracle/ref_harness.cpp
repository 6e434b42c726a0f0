// ref_harness.cpp - TEST INFRASTRUCTURE ONLY.  C entry points around the REFERENCE'S OWN sources, compiled where they lie
// under /root/reference by oracle/build_ref.sh into oracle/_ref/libdfki_ref_h<N>.so (one library per compile-time horizon,
// the reference's MPC_PREDICTION_HORIZON, mit_controller_params.hpp:6):
//   ws/src/controllers/src/mit_controller/gait.cpp                       Gait, AdaptiveGait, GaitDatabase
//   ws/src/controllers/src/mit_controller/simple_gait_sequencer.cpp      SimpleGaitSequencer
//   ws/src/controllers/src/mit_controller/adaptive_gait_sequencer.cpp    AdaptiveGaitSequencer
//   ws/src/controllers/src/mit_controller/mpc_trajectory_planner.cpp     MPCTrajectoryPlanner
//   ws/src/controllers/src/mit_controller/raibert_foot_step_planner.cpp  RaibertFootStepPlanner
//   ws/src/controllers/src/mit_controller/mpc.cpp                        MPC (QP data assembly, unpack of the solution)
// What is NOT the reference's: Eigen (ref_shim/mini_eigen.hpp), the acados C interface (implemented below; the QP solve
// itself goes to oracle/mpc_ref.c), the two ROS headers, and the StateInterface / ModelInterface implementations below
// (the reference's own QuadState / QuadModelSymbolic need ROS message types).  Nothing here is linked into the product.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>

#include "acados/ocp_qp/ocp_qp_full_condensing.h"
#include "acados/utils/print.h"
#include "acados_c/ocp_qp_interface.h"
#include "mit_controller/adaptive_gait_sequencer.hpp"
#include "mit_controller/gait.hpp"
#include "mit_controller/mpc.hpp"
#include "mit_controller/simple_gait_sequencer.hpp"

// ---- oracle QP solver (oracle/mpc_ref.c) ----------------------------------------------------------------------------
extern "C" {
typedef struct {
  int N;
  double dt, alpha, w[12], mu, fmin, fmax;
  double rho, sigma, relax, eps_abs, eps_rel, kkt_tol;
  int max_iter, check_every, polish, adaptive_rho;
} ref_params;
typedef struct {
  int status;
  int iters, n_fact, polished;
  double kkt[4];
} ref_info;
int ref_solve_ocp(const ref_params *P, const double *A_cm, const double *const *B_cm, const double *const *q, const double *x0,
                  const double *const *lbu, const double *const *ubu, double *u_out, double *x_out, ref_info *info);
}

#define DFKIREF_API __attribute__((visibility("default")))

namespace {
constexpr int NX = 13, NU = 12, NG = 16;
struct CoutSilencer {  // the reference prints from its constructor and setters
  std::streambuf *old;
  std::ostringstream sink;
  CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~CoutSilencer() { std::cout.rdbuf(old); }
};
}  // namespace

// ---- acados C interface stand-in ------------------------------------------------------------------------------------
extern "C" {

ocp_qp_dims *ocp_qp_dims_create(int N) {
  if (N > DFKIREF_MAX_STAGES) return nullptr;
  ocp_qp_dims *d = new ocp_qp_dims();
  std::memset(d, 0, sizeof(*d));
  d->N = N;
  return d;
}
void ocp_qp_dims_free(void *dims) { delete static_cast<ocp_qp_dims *>(dims); }
ocp_qp_xcond_solver_config *ocp_qp_xcond_solver_config_create(ocp_qp_solver_plan_t plan) {
  auto *c = new ocp_qp_xcond_solver_config();
  c->plan = plan;
  return c;
}
void ocp_qp_xcond_solver_config_free(ocp_qp_xcond_solver_config *config) { delete config; }
void ocp_qp_dims_set(void *, void *dims_, int stage, const char *field, int *value) {
  ocp_qp_dims *d = static_cast<ocp_qp_dims *>(dims_);
  const std::string f(field);
  if (f == "nx") d->nx[stage] = *value;
  else if (f == "nu") d->nu[stage] = *value;
  else if (f == "nbx") d->nbx[stage] = *value;
  else if (f == "nbu") d->nbu[stage] = *value;
  else if (f == "ng") d->ng[stage] = *value;
  else std::fprintf(stderr, "dfkiref acados shim: unknown dimension field %s\n", field);
}

static double *alloc_d(int n) { return n > 0 ? new double[n]() : nullptr; }
ocp_qp_in *ocp_qp_in_create(ocp_qp_dims *d) {
  ocp_qp_in *in = new ocp_qp_in();
  in->own_dim = *d;
  in->dim = &in->own_dim;
  for (int k = 0; k <= d->N; ++k) {
    dfkiref_stage &s = in->stage[k];
    const int nx = d->nx[k], nu = d->nu[k], ng = d->ng[k], nx1 = k < d->N ? d->nx[k + 1] : 0;
    s.A = alloc_d(nx1 * nx);
    s.B = alloc_d(nx1 * nu);
    s.Q = alloc_d(nx * nx);
    s.R = alloc_d(nu * nu);
    s.q = alloc_d(nx);
    s.r = alloc_d(nu);
    s.C = alloc_d(ng * nx);
    s.D = alloc_d(ng * nu);
    s.lg = alloc_d(ng);
    s.ug = alloc_d(ng);
    s.lbx = alloc_d(d->nbx[k]);
    s.ubx = alloc_d(d->nbx[k]);
    s.lbu = alloc_d(d->nbu[k]);
    s.ubu = alloc_d(d->nbu[k]);
    s.Jbx = alloc_d(d->nbx[k] * nx);
    s.Jbu = alloc_d(d->nbu[k] * nu);
  }
  return in;
}
void ocp_qp_in_free(void *in_) {
  ocp_qp_in *in = static_cast<ocp_qp_in *>(in_);
  for (int k = 0; k <= in->dim->N; ++k) {
    dfkiref_stage &s = in->stage[k];
    double *all[] = {s.A, s.B, s.Q, s.R, s.q, s.r, s.C, s.D, s.lg, s.ug, s.lbx, s.ubx, s.lbu, s.ubu, s.Jbx, s.Jbu};
    for (double *p : all) delete[] p;
  }
  delete in;
}
void ocp_qp_in_set(ocp_qp_xcond_solver_config *, ocp_qp_in *in, int k, char *field, void *value) {
  const ocp_qp_dims *d = in->dim;
  dfkiref_stage &s = in->stage[k];
  const int nx = d->nx[k], nu = d->nu[k], ng = d->ng[k], nx1 = k < d->N ? d->nx[k + 1] : 0;
  const std::string f(field);
  double *dst = nullptr;
  int n = 0;
  if (f == "A") dst = s.A, n = nx1 * nx;
  else if (f == "B") dst = s.B, n = nx1 * nu;
  else if (f == "Q") dst = s.Q, n = nx * nx;
  else if (f == "R") dst = s.R, n = nu * nu;
  else if (f == "q") dst = s.q, n = nx;
  else if (f == "r") dst = s.r, n = nu;
  else if (f == "C") dst = s.C, n = ng * nx;
  else if (f == "D") dst = s.D, n = ng * nu;
  else if (f == "lg") dst = s.lg, n = ng;
  else if (f == "ug") dst = s.ug, n = ng;
  else if (f == "lbx") dst = s.lbx, n = d->nbx[k];
  else if (f == "ubx") dst = s.ubx, n = d->nbx[k];
  else if (f == "lbu") dst = s.lbu, n = d->nbu[k];
  else if (f == "ubu") dst = s.ubu, n = d->nbu[k];
  else if (f == "Jbx") dst = s.Jbx, n = d->nbx[k] * nx;
  else if (f == "Jbu") dst = s.Jbu, n = d->nbu[k] * nu;
  else {
    std::fprintf(stderr, "dfkiref acados shim: unknown field %s\n", field);
    return;
  }
  if (dst && n > 0) std::memcpy(dst, value, sizeof(double) * (size_t)n);
}
ocp_qp_xcond_solver_dims *ocp_qp_xcond_solver_dims_create_from_ocp_qp_dims(ocp_qp_xcond_solver_config *, ocp_qp_dims *dims) {
  auto *d = new ocp_qp_xcond_solver_dims();
  d->orig_dims = dims;
  return d;
}
void ocp_qp_xcond_solver_dims_free(ocp_qp_xcond_solver_dims *dims) { delete dims; }
void *ocp_qp_xcond_solver_opts_create(ocp_qp_xcond_solver_config *, ocp_qp_xcond_solver_dims *dims) {
  auto *o = new ocp_qp_xcond_solver_opts();
  o->cond_N = dims->orig_dims->N;
  std::strcpy(o->hpipm_mode, "BALANCE");
  o->warm_start = 0;
  return o;
}
void ocp_qp_xcond_solver_opts_set(ocp_qp_xcond_solver_config *, ocp_qp_xcond_solver_opts *opts, const char *field, void *value) {
  const std::string f(field);
  if (f == "cond_N") opts->cond_N = *static_cast<int *>(value);
  else if (f == "hpipm_mode") std::strncpy(opts->hpipm_mode, static_cast<const char *>(value), sizeof(opts->hpipm_mode) - 1);
  else if (f == "warm_start") opts->warm_start = *static_cast<int *>(value);
}
void ocp_qp_xcond_solver_opts_free(ocp_qp_xcond_solver_opts *opts) { delete opts; }

struct ShimSolverMem {  // first member readable as ocp_qp_xcond_solver_memory / ocp_qp_full_condensing_memory
  void *cond_qp_in;
  ocp_qp_in part_in;
  ocp_qp_dims part_dims;
  dense_qp_in dense_in;
  dense_qp_dims dense_dims;
};
ocp_qp_solver *ocp_qp_create(ocp_qp_xcond_solver_config *config, ocp_qp_xcond_solver_dims *dims, void *opts_) {
  auto *s = new ocp_qp_solver();
  s->config = config;
  s->opts = static_cast<ocp_qp_xcond_solver_opts *>(opts_);
  auto *m = new ShimSolverMem();
  std::memset(m, 0, sizeof(*m));
  const ocp_qp_dims *d = dims->orig_dims;
  if (config->plan.qp_solver >= FULL_CONDENSING_HPIPM) {  // dimensions of the dense QP full condensing produces
    int nv = 0, nb = 0, ng = 0;
    for (int k = 0; k <= d->N; ++k) {
      nv += d->nu[k];
      nb += d->nbu[k];
      ng += d->ng[k];
    }
    m->dense_dims = dense_qp_dims{nv, nb, 0, ng, 0};
    m->dense_in.dim = &m->dense_dims;
    m->cond_qp_in = &m->dense_in;
  } else {  // dimensions of the partially condensed OCP-QP: cond_N blocks of N / cond_N stages (HPIPM block sizes)
    const int N2 = s->opts->cond_N;
    m->part_dims.N = N2;
    int k0 = 0;
    for (int b = 0; b <= N2; ++b) {
      const int len = b < N2 ? d->N / N2 + (b < d->N % N2 ? 1 : 0) : 1;
      m->part_dims.nx[b] = d->nx[k0 < d->N ? k0 : d->N];
      for (int k = k0; k < k0 + len && k <= d->N; ++k) {
        m->part_dims.nu[b] += d->nu[k];
        m->part_dims.nbu[b] += d->nbu[k];
        m->part_dims.ng[b] += d->ng[k];
      }
      k0 += len;
    }
    m->part_dims.nbx[0] = d->nbx[0];
    m->part_in.dim = &m->part_dims;
    m->cond_qp_in = &m->part_in;
  }
  s->mem = m;
  return s;
}
void ocp_qp_solver_destroy(ocp_qp_solver *solver) {
  delete static_cast<ShimSolverMem *>(solver->mem);
  delete solver;
}
ocp_qp_out *ocp_qp_out_create(ocp_qp_dims *d) {
  auto *o = new ocp_qp_out();
  o->own_dim = *d;
  o->dim = &o->own_dim;
  o->ux = new blasfeo_dvec[d->N + 1];
  for (int k = 0; k <= d->N; ++k) {
    o->ux[k].m = d->nu[k] + d->nx[k];
    o->ux[k].pa = alloc_d(o->ux[k].m);
  }
  o->misc = new qp_info();
  std::memset(o->misc, 0, sizeof(qp_info));
  return o;
}
void ocp_qp_out_free(void *out_) {
  ocp_qp_out *o = static_cast<ocp_qp_out *>(out_);
  for (int k = 0; k <= o->dim->N; ++k) delete[] o->ux[k].pa;
  delete[] o->ux;
  delete static_cast<qp_info *>(o->misc);
  delete o;
}
void d_ocp_qp_sol_get_u(int k, ocp_qp_out *o, double *u) { std::memcpy(u, o->ux[k].pa, sizeof(double) * o->dim->nu[k]); }
void d_ocp_qp_sol_get_x(int k, ocp_qp_out *o, double *x) {
  std::memcpy(x, o->ux[k].pa + o->dim->nu[k], sizeof(double) * o->dim->nx[k]);
}
void print_ocp_qp_dims(ocp_qp_dims *) {}  // the reference prints its dimensions from the constructor; kept quiet here
void print_qp_info(qp_info *) {}
void ocp_qp_inf_norm_residuals(ocp_qp_dims *, ocp_qp_in *, ocp_qp_out *out, double *res) {
  for (int i = 0; i < 4; ++i) res[i] = out->kkt[i];
}

// The reference's QP has a fixed structure (mpc.cpp:143-213); it is verified here on the data the reference code set and
// then handed to the oracle solver, which exploits it.  Anything else is reported as ACADOS_QP_FAILURE.
static thread_local std::string g_shim_error;
static int shim_fail(const char *why) {
  g_shim_error = why;
  return ACADOS_QP_FAILURE;
}
static thread_local double g_kkt_tol = 1e-9;
static thread_local int g_max_iter = 20000;
// the acados objects are private members of MPC: the QP the reference assembled is observed at the ocp_qp_solve boundary
static thread_local const ocp_qp_in *g_last_in = nullptr;
static thread_local ocp_qp_out *g_last_out = nullptr;

int ocp_qp_solve(ocp_qp_solver *, ocp_qp_in *in, ocp_qp_out *out) {
  g_last_in = in;
  g_last_out = out;
  const ocp_qp_dims *d = in->dim;
  const int N = d->N;
  if (N > 20) return shim_fail("horizon > 20");
  for (int k = 0; k <= N; ++k)
    if (d->nx[k] != NX || d->nu[k] != (k < N ? NU : 0) || d->ng[k] != (k < N ? NG : 0) || d->nbu[k] != (k < N ? NU : 0) ||
        d->nbx[k] != (k == 0 ? NX : 0))
      return shim_fail("unexpected stage dimensions");
  ref_params P;
  std::memset(&P, 0, sizeof(P));
  P.N = N;
  const dfkiref_stage &s0 = in->stage[0];
  // Q diagonal (last entry 0), identical on all stages; R = alpha I; C = 0; Jbu = I; Jbx = I; lbx = ubx
  for (int k = 0; k <= N; ++k)
    for (int i = 0; i < NX; ++i)
      for (int j = 0; j < NX; ++j) {
        const double v = in->stage[k].Q[j * NX + i];
        if (i != j && v != 0.0) return shim_fail("Q not diagonal");
        if (i == j && v != s0.Q[j * NX + i]) return shim_fail("Q differs between stages");
      }
  if (s0.Q[12 * NX + 12] != 0.0) return shim_fail("Q(12,12) != 0");
  for (int i = 0; i < 12; ++i) P.w[i] = s0.Q[i * NX + i];
  P.alpha = s0.R[0];
  for (int k = 0; k < N; ++k) {
    const dfkiref_stage &s = in->stage[k];
    for (int i = 0; i < NU; ++i)
      for (int j = 0; j < NU; ++j) {
        if (s.R[j * NU + i] != (i == j ? P.alpha : 0.0)) return shim_fail("R != alpha I");
        if (s.Jbu[j * NU + i] != (i == j ? 1.0 : 0.0)) return shim_fail("Jbu != I");
      }
    for (int i = 0; i < NG * NX; ++i)
      if (s.C[i] != 0.0) return shim_fail("C != 0");
    for (int i = 0; i < NU; ++i)
      if (s.r[i] != 0.0) return shim_fail("r != 0");
  }
  for (int i = 0; i < NX; ++i) {
    for (int j = 0; j < NX; ++j)
      if (s0.Jbx[j * NX + i] != (i == j ? 1.0 : 0.0)) return shim_fail("Jbx != I");
    if (s0.lbx[i] != s0.ubx[i]) return shim_fail("lbx != ubx");
  }
  // D = friction pyramid with one mu (mpc.cpp:66-71), lg / ug as mpc.cpp:73-92
  P.mu = s0.D[(0 * 3 + 2) * NG + 1];  // row 1 of foot 0: [1, 0, mu]
  for (int k = 0; k < N; ++k) {
    const dfkiref_stage &s = in->stage[k];
    for (int f = 0; f < 4; ++f)
      for (int r = 0; r < 4; ++r) {
        for (int c = 0; c < NU; ++c) {
          double want = 0.0;
          if (c / 3 == f) {
            const int cc = c % 3;
            if (cc == 0) want = (r < 2) ? 1.0 : 0.0;
            else if (cc == 1) want = (r >= 2) ? 1.0 : 0.0;
            else want = (r % 2 == 0) ? -P.mu : P.mu;
          }
          if (s.D[c * NG + 4 * f + r] != want) return shim_fail("D is not the friction pyramid");
        }
        const double lg = s.lg[4 * f + r], ug = s.ug[4 * f + r];
        if (r % 2 == 0 ? !(lg == -99999999.0 && ug == 0.0) : !(lg == 0.0 && ug == 99999999.0)) return shim_fail("lg / ug pattern");
      }
  }
  // boxes: stance [-fmax, -fmax, fmin] .. [fmax, fmax, fmax], swing 0 (mpc.cpp:94-115)
  bool have = false;
  for (int k = 0; k < N && !have; ++k)
    for (int f = 0; f < 4 && !have; ++f) {
      const double *lb = in->stage[k].lbu + 3 * f, *ub = in->stage[k].ubu + 3 * f;
      if (!(lb[0] == 0 && lb[1] == 0 && lb[2] == 0 && ub[0] == 0 && ub[1] == 0 && ub[2] == 0)) {
        P.fmax = ub[2];
        P.fmin = lb[2];
        have = true;
      }
    }
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      const double *lb = in->stage[k].lbu + 3 * f, *ub = in->stage[k].ubu + 3 * f;
      const bool swing = lb[0] == 0 && lb[1] == 0 && lb[2] == 0 && ub[0] == 0 && ub[1] == 0 && ub[2] == 0;
      if (!swing && !(lb[0] == -P.fmax && lb[1] == -P.fmax && lb[2] == P.fmin && ub[0] == P.fmax && ub[1] == P.fmax && ub[2] == P.fmax))
        return shim_fail("lbu / ubu pattern");
    }
  // one A for all stages (mpc.cpp:355 passes the current yaw for every k)
  for (int k = 1; k < N; ++k)
    if (std::memcmp(in->stage[k].A, s0.A, sizeof(double) * NX * NX) != 0) return shim_fail("A differs between stages");
  P.dt = s0.A[9 * NX + 3];
  // OSQP-style ADMM + polish of oracle/mpc_ref.c, run tight
  P.rho = 3e-4;
  P.sigma = 1e-6;
  P.relax = 1.6;
  P.eps_abs = P.eps_rel = 1e-4;
  P.kkt_tol = g_kkt_tol;
  P.max_iter = g_max_iter;
  P.check_every = 25;
  P.polish = 1;
  P.adaptive_rho = 1;
  const double *Bp[20], *qp[21], *lbp[20], *ubp[20];
  for (int k = 0; k < N; ++k) {
    Bp[k] = in->stage[k].B;
    lbp[k] = in->stage[k].lbu;
    ubp[k] = in->stage[k].ubu;
  }
  for (int k = 0; k <= N; ++k) qp[k] = in->stage[k].q;
  double u[20 * NU], x[21 * NX];
  ref_info info;
  const auto t0 = std::chrono::steady_clock::now();
  const int st = ref_solve_ocp(&P, s0.A, Bp, qp, s0.lbx, lbp, ubp, u, x, &info);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  qp_info *qi = static_cast<qp_info *>(out->misc);
  qi->total_time = qi->solve_QP_time = secs;
  qi->condensing_time = qi->interface_time = 0.0;
  qi->num_iter = info.iters;
  qi->t_computed = 1;
  for (int i = 0; i < 4; ++i) out->kkt[i] = info.kkt[i];
  if (st != 0) return st == 1 ? ACADOS_NAN_DETECTED : ACADOS_MAXITER;
  for (int k = 0; k <= N; ++k) {
    if (k < N) std::memcpy(out->ux[k].pa, u + NU * k, sizeof(double) * NU);
    std::memcpy(out->ux[k].pa + (k < N ? NU : 0), x + NX * k, sizeof(double) * NX);
  }
  return ACADOS_SUCCESS;
}
}  // extern "C"

// ---- StateInterface / ModelInterface implementations ------------------------------------------------------------------
namespace {
class HState : public StateInterface {
 public:
  Eigen::Vector3d position, linear_vel, angular_vel, linear_acc, angular_acc;
  Eigen::Quaterniond orientation;
  TimePoint time;
  std::array<bool, NUM_FEET> contacts{};
  std::array<Eigen::Vector3d, NUM_FEET> contact_forces, feet;  // feet: world-frame foot positions (what the leg FK gives)
  std::array<std::array<double, NUM_JOINT_PER_FOOT>, NUM_FEET> joints{};
  const Eigen::Vector3d &GetPositionInWorld() const override { return position; }
  const Eigen::Quaterniond &GetOrientationInWorld() const override { return orientation; }
  const Eigen::Vector3d &GetLinearVelInWorld() const override { return linear_vel; }
  const Eigen::Vector3d &GetAngularVelInWorld() const override { return angular_vel; }
  const Eigen::Vector3d &GetLinearAccInWorld() const override { return linear_acc; }
  const Eigen::Vector3d &GetAngularAccInWorld() const override { return angular_acc; }
  const TimePoint &GetTime() const override { return time; }
  const std::array<bool, NUM_FEET> &GetFeetContacts() const override { return contacts; }
  const std::array<Eigen::Vector3d, NUM_FEET> &GetContactForces() const override { return contact_forces; }
  const std::array<std::array<double, NUM_JOINT_PER_FOOT>, NUM_FEET> &GetJointPositions() const override { return joints; }
  const std::array<std::array<double, NUM_JOINT_PER_FOOT>, NUM_FEET> &GetJointVelocities() const override { return joints; }
  const std::array<std::array<double, NUM_JOINT_PER_FOOT>, NUM_FEET> &GetJointAccelerations() const override { return joints; }
  const std::array<std::array<double, NUM_JOINT_PER_FOOT>, NUM_FEET> &GetJointTorques() const override { return joints; }
  StateInterface &operator=(const StateInterface &other) override {
    const HState &o = dynamic_cast<const HState &>(other);
    position = o.position;
    linear_vel = o.linear_vel;
    angular_vel = o.angular_vel;
    linear_acc = o.linear_acc;
    angular_acc = o.angular_acc;
    orientation = o.orientation;
    time = o.time;
    contacts = o.contacts;
    contact_forces = o.contact_forces;
    feet = o.feet;
    return *this;
  }
};

class HModel : public ModelInterface {
 public:
  Eigen::Matrix3d inertia;
  double mass = 1.0;
  Eigen::Vector3d com;
  double fk0_z = -0.426;  // z of FK(q = 0) in the body frame (quad_model_symbolic.cpp leg geometry): l1 + l2
  Eigen::Vector3d GetFootPositionInWorld(unsigned int foot_idx, const StateInterface &state) const override {
    return dynamic_cast<const HState &>(state).feet[foot_idx];
  }
  void calcFwdKinLegBody(int, const Eigen::Vector3d &, Eigen::Matrix4d &, Eigen::Vector3d &foot_pos_body) const override {
    foot_pos_body = Eigen::Vector3d(0.0, 0.0, fk0_z);
  }
  Eigen::Matrix3d GetInertia() const override { return inertia; }
  double GetInertia(const int row, const int column) const override { return inertia(row, column); }
  double GetMass() const override { return mass; }
  double GetG() const override { return 9.81; }
  Eigen::Translation3d GetBodyToCOM() const override { return Eigen::Translation3d(com); }
  Eigen::Translation3d GetBodyToIMU() const override { return Eigen::Translation3d::Identity(); }
  Eigen::Translation3d GetBodyToBellyBottom() const override { return Eigen::Translation3d::Identity(); }
  Eigen::Translation3d GetBodyToBellyBottom(unsigned int) const override { return Eigen::Translation3d::Identity(); }
  void SetInertia(const Eigen::Matrix3d &i) override { inertia = i; }
  void SetInertia(const int row, const int column, const double value) override { inertia(row, column) = value; }
  void SetMass(double m) override { mass = m; }
  void SetCOM(const Eigen::Vector3d &c) override { com = c; }
  void SetCOM(const int idx, const double val) override { com(idx) = val; }
  ModelInterface &operator=(const ModelInterface &other) override {
    const HModel &o = dynamic_cast<const HModel &>(other);
    inertia = o.inertia;
    mass = o.mass;
    com = o.com;
    fk0_z = o.fk0_z;
    return *this;
  }
  // not on the path
  [[noreturn]] static void off_path() { throw std::logic_error("dfkiref harness: model function not on the MPC path"); }
  void CalcFootForceVelocityInBodyFrame(int, const Eigen::Ref<const Eigen::Matrix<double, 3, 1>> &,
                                        const Eigen::Ref<const Eigen::Matrix<double, 3, 1>> &,
                                        const Eigen::Ref<const Eigen::Matrix<double, 3, 1>> &,
                                        const Eigen::Ref<const Eigen::Matrix<double, 3, 1>> &, Eigen::Ref<Eigen::Vector3d>,
                                        Eigen::Ref<Eigen::Vector3d>) const override { off_path(); }
  void calcFootForceVelocityBodyFrame(int, const StateInterface &, Eigen::Vector3d &, Eigen::Vector3d &) const override { off_path(); }
  void calcLegInverseKinematicsInBody(int, const Eigen::Vector3d &, const Eigen::Vector3d &, Eigen::Vector3d &) const override { off_path(); }
  void calcLegDiffKinematicsBodyFrame(int, const StateInterface &, Eigen::Vector3d &, Eigen::Vector3d &, Eigen::Vector3d &,
                                      Eigen::Vector3d &) const override { off_path(); }
  void calcJacobianLegBase(int, Eigen::Vector3d, Eigen::Matrix3d &) const override { off_path(); }
  Eigen::Vector3d GetFootPositionInWorld(unsigned int, const Eigen::Vector3d &, const Eigen::Quaterniond &,
                                         const Eigen::Vector3d &) const override { off_path(); }
  Eigen::Vector3d GetFootPositionInBodyFrame(unsigned int, const Eigen::Vector3d &) const override { off_path(); }
  void CalcBaseHeight(const std::array<bool, 4> &, const std::array<const Eigen::Vector3d, 4> &, const Eigen::Quaterniond &,
                      double &) const override { off_path(); }
  bool IsLyingDown(const std::array<const Eigen::Vector3d, 4> &, const std::array<double, 4> &) const override { off_path(); }
  double ComputeKineticEnergy(int, const Eigen::Vector3d &, const Eigen::Vector3d &) const override { off_path(); }
  double ComputeEnergyDerivative(int, const Eigen::Vector3d &, const Eigen::Vector3d &) const override { off_path(); }
};

Eigen::Quaterniond quat_xyzw(const double *q) { return Eigen::Quaterniond(q[3], q[0], q[1], q[2]); }
Eigen::Vector3d vec3(const double *v) { return Eigen::Vector3d(v[0], v[1], v[2]); }

void fill_model(HModel &m, double mass, const double *inertia_cm, const double *com, double max_leg_length) {
  m.mass = mass;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m.inertia(r, c) = inertia_cm[3 * c + r];
  m.com = vec3(com);
  m.fk0_z = -(max_leg_length + 0.02);  // RaibertFootStepPlanner: max_leg_length_ = -FK(0).z - 0.02 (raibert_foot_step_planner.cpp:36-40)
}
// state record: quat xyzw, p, omega, v (the first 13 doubles of the b200qp input record)
void fill_state(HState &s, const double *st) {
  s.orientation = quat_xyzw(st);
  s.position = vec3(st + 4);
  s.angular_vel = vec3(st + 7);
  s.linear_vel = vec3(st + 10);
}

}  // namespace

// ---- C entry points -----------------------------------------------------------------------------------------------------
extern "C" {

DFKIREF_API int dfkiref_horizon(void) { return MPC_PREDICTION_HORIZON; }
DFKIREF_API int dfkiref_gait_sequence_size(void) { return GAIT_SEQUENCE_SIZE; }
DFKIREF_API const char *dfkiref_last_error(void) { return g_shim_error.c_str(); }
DFKIREF_API void dfkiref_set_solver_options(double kkt_tol, int max_iter) {
  g_kkt_tol = kkt_tol;
  g_max_iter = max_iter;
}

// GaitDatabase::getGait (gait.cpp:732-767)
DFKIREF_API void dfkiref_gait_database(int type, double dt, double *period, double *duty, double *offset, char *name32) {
  Gait g = GaitDatabase::getGait(static_cast<GaitDatabase::GaitType>(type), dt);
  *period = g.get_period();
  for (int l = 0; l < N_LEGS; ++l) {
    duty[l] = g.get_duty_factor(l);
    offset[l] = g.get_phase_offset(l);
  }
  if (name32) std::strncpy(name32, g.get_name().c_str(), 31);
}

// Gait::update_sequence (gait.cpp:67-112) for n (gait, phase) pairs; rows 0..steps-1 of the 100-row sequence are returned.
// measured: [n][4] or NULL.  contact [n][steps][4] uint8, swing_time [n][steps][4].
DFKIREF_API void dfkiref_gait_update_sequence(int n, int steps, double dt, const double *period, const double *duty, const double *offset,
                                  const double *phase, const uint8_t *measured, uint8_t *contact, double *swing_time,
                                  double *t_swing, int *keep_mode) {
  static thread_local GaitSequence gs;
  for (int p = 0; p < n; ++p) {
    std::array<double, N_LEGS> df, po;
    for (int l = 0; l < N_LEGS; ++l) {
      df[l] = duty[4 * p + l];
      po[l] = offset[4 * p + l];
    }
    Gait g(period[p], df, po, dt, "");
    gs.sequence_mode = GaitSequence::MOVE;
    if (measured) {
      std::array<bool, N_LEGS> cs;
      for (int l = 0; l < N_LEGS; ++l) cs[l] = measured[4 * p + l] != 0;
      g.update_sequence(gs, phase[p], cs);
    } else {
      g.update_sequence(gs, phase[p]);
    }
    for (int i = 0; i < steps; ++i)
      for (int l = 0; l < N_LEGS; ++l) {
        contact[((size_t)p * steps + i) * 4 + l] = gs.contact_sequence[i][l] ? 1 : 0;
        if (swing_time) swing_time[((size_t)p * steps + i) * 4 + l] = gs.swing_time_sequence[i][l];
      }
    if (t_swing)
      for (int l = 0; l < N_LEGS; ++l) t_swing[4 * p + l] = gs.gait_swing_time[l];
    if (keep_mode) keep_mode[p] = gs.sequence_mode == GaitSequence::KEEP;
  }
}
// Gait::get_contact_sequence / get_swing_time_sequence (gait.cpp:30-65)
DFKIREF_API void dfkiref_gait_contact_sequence(int n, int steps, double dt, const double *period, const double *duty, const double *offset,
                                   const double *phase, uint8_t *contact, double *swing_time) {
  static thread_local std::array<std::array<bool, N_LEGS>, GAIT_SEQUENCE_SIZE> cs;
  static thread_local std::array<std::array<double, N_LEGS>, GAIT_SEQUENCE_SIZE> ss;
  for (int p = 0; p < n; ++p) {
    std::array<double, N_LEGS> df, po;
    for (int l = 0; l < N_LEGS; ++l) {
      df[l] = duty[4 * p + l];
      po[l] = offset[4 * p + l];
    }
    Gait g(period[p], df, po, dt, "");
    g.get_contact_sequence(cs, phase[p]);
    g.get_swing_time_sequence(ss, phase[p]);
    for (int i = 0; i < steps; ++i)
      for (int l = 0; l < N_LEGS; ++l) {
        contact[((size_t)p * steps + i) * 4 + l] = cs[i][l] ? 1 : 0;
        if (swing_time) swing_time[((size_t)p * steps + i) * 4 + l] = ss[i][l];
      }
  }
}

// ---- MPC -------------------------------------------------------------------------------------------------------------
struct dfkiref_mpc {
  std::unique_ptr<MPC> mpc;
  HModel model;
  HState state;
  GaitSequence gs;
  WrenchSequence wrench;
  MPCPrediction pred;
  SolverInformation info;
  // copy of the QP the reference assembled in the last call
  double A[20][NX * NX], B[20][NX * NU], q[21][NX], lbu[20][NU], ubu[20][NU], x0[NX], D[NG * NU], lg[NG], ug[NG], Q[NX * NX], R[NU * NU];
  double kkt[4];
};

DFKIREF_API dfkiref_mpc *dfkiref_mpc_create(double alpha, const double *w12, double mu, double fmin, double fmax, int solver, int cond_N,
                                const char *hpipm_mode, int warm_start, double mass, const double *inertia_cm, const double *com) {
  CoutSilencer quiet;
  auto *h = new dfkiref_mpc();
  fill_model(h->model, mass, inertia_cm, com, 0.406);
  Eigen::Matrix<double, 12, 1> w;
  for (int i = 0; i < 12; ++i) w(i) = w12[i];
  auto st = std::make_unique<HState>();
  auto md = std::make_unique<HModel>(h->model);
  h->mpc = std::make_unique<MPC>(alpha, w, mu, fmin, fmax, std::move(st), std::move(md), static_cast<ocp_qp_solver_t>(solver),
                                 cond_N, std::string(hpipm_mode ? hpipm_mode : "BALANCE"), warm_start);
  return h;
}
DFKIREF_API void dfkiref_mpc_destroy(dfkiref_mpc *h) { delete h; }
DFKIREF_API void dfkiref_mpc_set_state_weights(dfkiref_mpc *h, const double *w12) {
  CoutSilencer quiet;
  Eigen::Matrix<double, 12, 1> w;
  for (int i = 0; i < 12; ++i) w(i) = w12[i];
  h->mpc->SetStateWeights(w);
}
DFKIREF_API void dfkiref_mpc_set_input_weights(dfkiref_mpc *h, double alpha) {
  CoutSilencer quiet;
  h->mpc->SetInputWeights(alpha);
}
DFKIREF_API void dfkiref_mpc_set_fmax(dfkiref_mpc *h, double fmax) {
  CoutSilencer quiet;
  h->mpc->SetFmax(fmax);
}
DFKIREF_API void dfkiref_mpc_set_mu(dfkiref_mpc *h, double mu) {
  CoutSilencer quiet;
  h->mpc->SetMu(mu);
}

// One MPC::UpdateState / UpdateModel / UpdateGaitSequence / GetWrenchSequence cycle (mpc.cpp:284-292, 329-470) on one
// b200qp input record (include/b200qp.h: layout of b200qp_mpc_in_doubles(N)) + contact[N][4].
// forces [N][12], xpred [N+1][13] (MPCPrediction::raw_data), pose [N+1][13] = orientation xyzw, position, angular, linear
// velocity of MPCPrediction (may be NULL).  Outputs are only written on success, as in the reference.
// Returns SolverInformation::return_code; *success = SolverInformation::success.
DFKIREF_API int dfkiref_mpc_solve(dfkiref_mpc *h, const double *rec, const uint8_t *contact, double *forces, double *xpred, double *pose,
                      int *success, int *num_iter) {
  CoutSilencer quiet;
  const int N = MPC_PREDICTION_HORIZON;
  fill_state(h->state, rec);
  fill_model(h->model, rec[13], rec + 14, rec + 23, 0.406);
  GaitSequence &gs = h->gs;
  for (int k = 0; k <= N; ++k) {
    const double *st = rec + 26 + 13 * k;
    gs.desired_reference_trajectory_orientation[k] = quat_xyzw(st);
    gs.desired_reference_trajectory_position[k] = vec3(st + 4);
    gs.reference_trajectory_twist[k] = vec3(st + 7);
    gs.reference_trajectory_velocity[k] = vec3(st + 10);
    gs.reference_trajectory_orientation[k] = gs.desired_reference_trajectory_orientation[k];
    gs.reference_trajectory_position[k] = gs.desired_reference_trajectory_position[k];
  }
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      gs.foot_position_sequence[k][f] = vec3(rec + 26 + 13 * (N + 1) + 12 * k + 3 * f);
      gs.contact_sequence[k][f] = contact[4 * k + f] != 0;
    }
  h->mpc->UpdateState(h->state);
  h->mpc->UpdateModel(h->model);
  h->mpc->UpdateGaitSequence(gs);
  g_last_in = nullptr;
  h->mpc->GetWrenchSequence(h->wrench, h->pred, h->info);
  if (g_last_in) {  // what the reference code handed to ocp_qp_solve
    for (int k = 0; k < N; ++k) {
      std::memcpy(h->A[k], g_last_in->stage[k].A, sizeof(h->A[k]));
      std::memcpy(h->B[k], g_last_in->stage[k].B, sizeof(h->B[k]));
      std::memcpy(h->lbu[k], g_last_in->stage[k].lbu, sizeof(h->lbu[k]));
      std::memcpy(h->ubu[k], g_last_in->stage[k].ubu, sizeof(h->ubu[k]));
    }
    for (int k = 0; k <= N; ++k) std::memcpy(h->q[k], g_last_in->stage[k].q, sizeof(h->q[k]));
    std::memcpy(h->x0, g_last_in->stage[0].lbx, sizeof(h->x0));
    std::memcpy(h->D, g_last_in->stage[0].D, sizeof(h->D));
    std::memcpy(h->lg, g_last_in->stage[0].lg, sizeof(h->lg));
    std::memcpy(h->ug, g_last_in->stage[0].ug, sizeof(h->ug));
    std::memcpy(h->Q, g_last_in->stage[0].Q, sizeof(h->Q));
    std::memcpy(h->R, g_last_in->stage[0].R, sizeof(h->R));
    if (g_last_out) std::memcpy(h->kkt, g_last_out->kkt, sizeof(h->kkt));
  }
  if (success) *success = h->info.success ? 1 : 0;
  if (num_iter) *num_iter = h->info.acados_num_iter;
  if (h->info.success) {
    for (int k = 0; k < N; ++k)
      for (int f = 0; f < 4; ++f)
        for (int c = 0; c < 3; ++c) forces[12 * k + 3 * f + c] = h->wrench.forces[k][f](c);
    for (int k = 0; k <= N; ++k) {
      for (int c = 0; c < 13; ++c) xpred[13 * k + c] = h->pred.raw_data[k](c);
      if (pose) {
        double *o = pose + 13 * k;
        o[0] = h->pred.orientation[k].x();
        o[1] = h->pred.orientation[k].y();
        o[2] = h->pred.orientation[k].z();
        o[3] = h->pred.orientation[k].w();
        for (int c = 0; c < 3; ++c) {
          o[4 + c] = h->pred.position[k](c);
          o[7 + c] = h->pred.angular_velocity[k](c);
          o[10 + c] = h->pred.linear_velocity[k](c);
        }
      }
    }
  }
  return h->info.return_code;
}
// the stage data of the last dfkiref_mpc_solve, column-major as the reference passed it to acados
DFKIREF_API void dfkiref_mpc_last_qp(dfkiref_mpc *h, double *A, double *B, double *q, double *lbu, double *ubu, double *x0, double *D,
                         double *lg, double *ug, double *Q, double *R, double *kkt) {
  const int N = MPC_PREDICTION_HORIZON;
  if (A) std::memcpy(A, h->A, sizeof(double) * N * NX * NX);
  if (B) std::memcpy(B, h->B, sizeof(double) * N * NX * NU);
  if (q) std::memcpy(q, h->q, sizeof(double) * (N + 1) * NX);
  if (lbu) std::memcpy(lbu, h->lbu, sizeof(double) * N * NU);
  if (ubu) std::memcpy(ubu, h->ubu, sizeof(double) * N * NU);
  if (x0) std::memcpy(x0, h->x0, sizeof(h->x0));
  if (D) std::memcpy(D, h->D, sizeof(h->D));
  if (lg) std::memcpy(lg, h->lg, sizeof(h->lg));
  if (ug) std::memcpy(ug, h->ug, sizeof(h->ug));
  if (Q) std::memcpy(Q, h->Q, sizeof(h->Q));
  if (R) std::memcpy(R, h->R, sizeof(h->R));
  if (kkt) std::memcpy(kkt, h->kkt, sizeof(h->kkt));
}

// ---- gait sequencers (stateful across calls, like the controller node's gs_) ---------------------------------------------
struct dfkiref_seq_config {
  int adaptive;            // 0: SimpleGaitSequencer, 1: AdaptiveGaitSequencer
  double period, duty[4], offset[4], dt;  // Gait members (simple) / phase offsets (adaptive)
  double raibert_k;
  double shoulders[12];
  int raibert_filtersize, raibert_z_on_plane;
  int fix_standing_position;
  double fix_position_distance_threshold, fix_position_angular_threshold, fix_position_velocity_threshold;
  int early_contact_detection;
  double mass, inertia_cm[9], com[3], max_leg_length;
  // AdaptiveGait constructor arguments (gait.hpp, mit_controller_node.cpp:250-264)
  double swing_time, zero_velocity_threshold, standing_foot_position_threshold, min_v_cmd_factor, max_correction_cycles,
      correction_period, disturbance_correction, min_v, offset_delay, max_stride_length;
  int filter_size, switch_offsets, correct_all;
};
struct dfkiref_seq {
  std::unique_ptr<GaitSequencerInterface> gs;
  HState state;
  HModel model;
  GaitSequence seq;
  StateInterface::TimePoint t0;
};

DFKIREF_API dfkiref_seq *dfkiref_seq_create(const dfkiref_seq_config *c, const double *state13, const double *feet12, const uint8_t *contacts4) {
  CoutSilencer quiet;
  auto *h = new dfkiref_seq();
  fill_model(h->model, c->mass, c->inertia_cm, c->com, c->max_leg_length);
  fill_state(h->state, state13);
  for (int l = 0; l < 4; ++l) {
    h->state.feet[l] = vec3(feet12 + 3 * l);
    h->state.contacts[l] = contacts4 ? contacts4[l] != 0 : true;
  }
  h->t0 = StateInterface::TimePoint();
  h->state.time = h->t0;
  const std::array<const Eigen::Vector3d, N_LEGS> sh{vec3(c->shoulders), vec3(c->shoulders + 3), vec3(c->shoulders + 6),
                                                     vec3(c->shoulders + 9)};
  std::array<double, N_LEGS> df, po;
  for (int l = 0; l < 4; ++l) {
    df[l] = c->duty[l];
    po[l] = c->offset[l];
  }
  if (!c->adaptive) {
    h->gs = std::make_unique<SimpleGaitSequencer>(Gait(c->period, df, po, c->dt, ""), c->raibert_k, sh, std::make_unique<HState>(h->state),
                                                  std::make_unique<HModel>(h->model), (unsigned)c->raibert_filtersize,
                                                  c->raibert_z_on_plane != 0, c->fix_standing_position != 0,
                                                  c->fix_position_distance_threshold, c->fix_position_angular_threshold,
                                                  c->fix_position_velocity_threshold, c->early_contact_detection != 0);
  } else {
    h->gs = std::make_unique<AdaptiveGaitSequencer>(
        AdaptiveGait(po, c->dt, c->swing_time, (unsigned)c->filter_size, c->zero_velocity_threshold, c->switch_offsets != 0,
                     c->standing_foot_position_threshold, c->min_v_cmd_factor, c->max_correction_cycles, c->correct_all != 0,
                     c->correction_period, c->disturbance_correction, c->min_v, c->offset_delay, c->max_stride_length, ""),
        c->raibert_k, sh, std::make_unique<HState>(h->state), std::make_unique<HModel>(h->model), (unsigned)c->raibert_filtersize,
        c->raibert_z_on_plane != 0, c->fix_standing_position != 0, c->fix_position_distance_threshold,
        c->fix_position_angular_threshold, c->fix_position_velocity_threshold, c->early_contact_detection != 0);
  }
  return h;
}
DFKIREF_API void dfkiref_seq_destroy(dfkiref_seq *h) { delete h; }

// GaitSequencerInterface::UpdateState with a state stamped `time_s` seconds after creation
DFKIREF_API void dfkiref_seq_update_state(dfkiref_seq *h, const double *state13, const double *feet12, const uint8_t *contacts4, double time_s) {
  fill_state(h->state, state13);
  for (int l = 0; l < 4; ++l) {
    h->state.feet[l] = vec3(feet12 + 3 * l);
    if (contacts4) h->state.contacts[l] = contacts4[l] != 0;
  }
  h->state.time = h->t0 + std::chrono::duration_cast<StateInterface::TimePoint::duration>(std::chrono::duration<double>(time_s));
  h->gs->UpdateState(h->state);
}
// GaitSequencerInterface::UpdateTarget.  vals: x y z roll pitch yaw | full_orientation xyzw | x_dot y_dot z_dot |
// hybrid_x_dot hybrid_y_dot hybrid_z_dot | wx wy wz (19 doubles); active: the 16 flags of Target::activateTarget in
// declaration order (target.hpp:28-45).
DFKIREF_API void dfkiref_seq_update_target(dfkiref_seq *h, const double *v, const uint8_t *a) {
  Target t;
  t.x = v[0]; t.y = v[1]; t.z = v[2]; t.roll = v[3]; t.pitch = v[4]; t.yaw = v[5];
  t.full_orientation = quat_xyzw(v + 6);
  t.x_dot = v[10]; t.y_dot = v[11]; t.z_dot = v[12];
  t.hybrid_x_dot = v[13]; t.hybrid_y_dot = v[14]; t.hybrid_z_dot = v[15];
  t.wx = v[16]; t.wy = v[17]; t.wz = v[18];
  t.active.x = a[0]; t.active.y = a[1]; t.active.z = a[2]; t.active.roll = a[3]; t.active.pitch = a[4]; t.active.yaw = a[5];
  t.active.full_orientation = a[6]; t.active.x_dot = a[7]; t.active.y_dot = a[8]; t.active.z_dot = a[9];
  t.active.hybrid_x_dot = a[10]; t.active.hybrid_y_dot = a[11]; t.active.hybrid_z_dot = a[12];
  t.active.wx = a[13]; t.active.wy = a[14]; t.active.wz = a[15];
  h->gs->UpdateTarget(t);
}
// GaitSequencerInterface::GetGaitSequence; the first `steps` rows of every sequence are returned.
//   contact [steps][4], swing_time [steps][4], foot_pos [steps][4][3], ref_pos / des_pos [steps][3], ref_ori / des_ori
//   [steps][4] xyzw, ref_vel / ref_twist [steps][3]; scalars[8] = sequence_mode, target_height, current_height,
//   gait_swing_time[4], (unused); gait_state[18] = period, duty[4], offset[4], phase[4], contact[4], type
DFKIREF_API void dfkiref_seq_get(dfkiref_seq *h, int steps, uint8_t *contact, double *swing_time, double *foot_pos, double *ref_pos,
                     double *ref_ori, double *des_pos, double *des_ori, double *ref_vel, double *ref_twist, double *scalars,
                     double *gait_state) {
  CoutSilencer quiet;
  GaitSequence &s = h->seq;
  h->gs->GetGaitSequence(s);
  for (int i = 0; i < steps; ++i) {
    for (int l = 0; l < 4; ++l) {
      if (contact) contact[4 * i + l] = s.contact_sequence[i][l] ? 1 : 0;
      if (swing_time) swing_time[4 * i + l] = s.swing_time_sequence[i][l];
      if (foot_pos)
        for (int c = 0; c < 3; ++c) foot_pos[12 * i + 3 * l + c] = s.foot_position_sequence[i][l](c);
    }
    for (int c = 0; c < 3; ++c) {
      if (ref_pos) ref_pos[3 * i + c] = s.reference_trajectory_position[i](c);
      if (des_pos) des_pos[3 * i + c] = s.desired_reference_trajectory_position[i](c);
      if (ref_vel) ref_vel[3 * i + c] = s.reference_trajectory_velocity[i](c);
      if (ref_twist) ref_twist[3 * i + c] = s.reference_trajectory_twist[i](c);
    }
    if (ref_ori) {
      const auto &q = s.reference_trajectory_orientation[i];
      ref_ori[4 * i] = q.x(); ref_ori[4 * i + 1] = q.y(); ref_ori[4 * i + 2] = q.z(); ref_ori[4 * i + 3] = q.w();
    }
    if (des_ori) {
      const auto &q = s.desired_reference_trajectory_orientation[i];
      des_ori[4 * i] = q.x(); des_ori[4 * i + 1] = q.y(); des_ori[4 * i + 2] = q.z(); des_ori[4 * i + 3] = q.w();
    }
  }
  if (scalars) {
    scalars[0] = s.sequence_mode == GaitSequence::KEEP ? 0.0 : 1.0;
    scalars[1] = s.target_height;
    scalars[2] = s.current_height;
    for (int l = 0; l < 4; ++l) scalars[3 + l] = s.gait_swing_time[l];
    scalars[7] = 0.0;
  }
  if (gait_state) {
    interfaces::msg::GaitState g;
    h->gs->GetGaitState(g);
    gait_state[0] = g.period;
    for (int l = 0; l < 4; ++l) {
      gait_state[1 + l] = g.duty_factor[l];
      gait_state[5 + l] = g.phase_offset[l];
      gait_state[9 + l] = g.phase[l];
      gait_state[13 + l] = g.contact[l] ? 1.0 : 0.0;
    }
    gait_state[17] = g.gait_sequencer;
  }
}
}  // extern "C"
