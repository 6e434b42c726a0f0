/* Compiled C caller of the C ABI (no Python between the program and libb200qp.so): fills input records by hand, calls
 * b200qp_mpc_solve_batch / b200qp_wbc_solve_batch on the GPU and checks status codes and known answers.
 *   (1) MPC, SURVEY.md Appendix A (i): level stand, four stance feet symmetric about the COM, zero velocities, reference =
 *       current state  =>  sum f_z = m g on the early stages, f_xy ~ 0, equal split over the feet, swing-free;
 *   (2) MPC failure semantics (mpc.cpp:447-455): a NaN record fails, its output slots stay untouched, its neighbours solve;
 *   (3) WBC plugin boundary (wbc_arc_opt.cpp:297): a QP with a closed-form answer.
 * Build + run: tests/test_gpu_c_boundary.py (gcc, links ../../dfki_quad_b200/libb200qp.so).  Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200qp.h"

#define CHECK(cond, ...)                                   \
  do {                                                     \
    if (!(cond)) {                                         \
      fprintf(stderr, "FAILED %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);                        \
      fprintf(stderr, "\n");                               \
      return 1;                                            \
    }                                                      \
  } while (0)

static void fill_standing_record(int N, double *rec, uint8_t *ct) {
  const double mass = 15.019, inertia_cm[9] = {0.202755, 0.00012166, -0.0143742, 0.00012166, 0.490926, -3.12e-05, -0.0143742, -3.12e-05, 0.52697};
  const double feet[4][3] = {{0.18, 0.13, 0.0}, {0.18, -0.13, 0.0}, {-0.18, 0.13, 0.0}, {-0.18, -0.13, 0.0}};
  memset(rec, 0, sizeof(double) * (size_t)b200qp_mpc_in_doubles(N));
  rec[3] = 1.0;               /* quaternion (x, y, z, w) = identity */
  rec[6] = 0.30;              /* height */
  rec[13] = mass;
  memcpy(rec + 14, inertia_cm, sizeof(inertia_cm));
  for (int k = 0; k <= N; ++k) {
    double *st = rec + 26 + 13 * k;
    st[3] = 1.0;
    st[6] = 0.30;
  }
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      memcpy(rec + 26 + 13 * (N + 1) + 12 * k + 3 * f, feet[f], sizeof(double) * 3);
      ct[4 * k + f] = 1;
    }
}

static int test_mpc(int solver, const char *name) {
  enum { N = 10, NB = 3 };
  b200qp_mpc_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.horizon = N;
  cfg.solver = solver;
  cfg.max_batch = NB;
  cfg.device = 0;
  cfg.dt = 0.05;
  cfg.alpha = 1e-5;
  const double w_stand[12] = {30, 30, 30, 30, 30, 30, 1, 1, 1, 1, 1, 1}; /* mpc_state_weights_stand, mit_controller_sim_go2.yaml:27 */
  memcpy(cfg.state_weights, w_stand, sizeof(w_stand));
  cfg.mu = 0.6;
  cfg.fmin = 0.0;
  cfg.fmax = 400.0;
  b200qp_mpc *h = NULL;
  int rc = b200qp_mpc_create(&cfg, &h);
  CHECK(rc == B200QP_OK, "%s: create failed (%d): %s", name, rc, b200qp_last_error());
  const int rd = b200qp_mpc_in_doubles(N);
  double *in = calloc((size_t)NB * rd, sizeof(double)), *forces = malloc(sizeof(double) * NB * N * 12), *xpred = malloc(sizeof(double) * NB * (N + 1) * 13);
  uint8_t *ct = calloc((size_t)NB * N * 4, 1);
  b200qp_solver_info info[NB];
  for (int p = 0; p < NB; ++p) fill_standing_record(N, in + (size_t)p * rd, ct + (size_t)p * N * 4);
  in[(size_t)1 * rd + 5] = NAN; /* problem 1: NaN in the position */
  for (int i = 0; i < NB * N * 12; ++i) forces[i] = 123.0;
  for (int i = 0; i < NB * (N + 1) * 13; ++i) xpred[i] = 321.0;
  rc = b200qp_mpc_solve_batch(h, NB, in, ct, forces, xpred, info);
  CHECK(rc == B200QP_OK, "%s: solve failed (%d): %s", name, rc, b200qp_last_error());
  const double mg = 15.019 * 9.8067;
  for (int p = 0; p < NB; p += 2) {
    CHECK(info[p].status == B200QP_ACADOS_SUCCESS, "%s: problem %d status %d", name, p, info[p].status);
    CHECK(info[p].n_free == 120, "%s: n_free %d", name, info[p].n_free);
    CHECK(info[p].kkt[0] <= 1e-6 && info[p].kkt[2] <= 1e-6 && info[p].kkt[3] <= 1e-6, "%s: KKT %g %g %g", name, info[p].kkt[0], info[p].kkt[2], info[p].kkt[3]);
    const double *F = forces + (size_t)p * N * 12;
    for (int k = 0; k < N; ++k) {
      double fz = 0;
      for (int f = 0; f < 4; ++f) {
        fz += F[12 * k + 3 * f + 2];
        CHECK(fabs(F[12 * k + 3 * f]) < 1e-5 * mg && fabs(F[12 * k + 3 * f + 1]) < 1e-5 * mg, "%s: lateral force at stage %d", name, k);
        CHECK(fabs(F[12 * k + 3 * f + 2] - F[12 * k + 2]) <= 1e-6 * mg, "%s: unequal split at stage %d", name, k);
      }
      if (k < 6) CHECK(fabs(fz - mg) <= 1e-3 * mg, "%s: sum f_z = %g at stage %d, m g = %g", name, fz, k, mg);
    }
    const double *X = xpred + (size_t)p * (N + 1) * 13;
    CHECK(fabs(X[5] - 0.30) < 1e-12 && X[12] == 9.8067, "%s: x0 not reproduced", name);
    CHECK(fabs(X[13 * 3 + 5] - 0.30) < 1e-3, "%s: height drifts: %g", name, X[13 * 3 + 5]);
  }
  CHECK(info[1].status != B200QP_ACADOS_SUCCESS, "%s: NaN problem reported success", name);
  for (int i = 0; i < N * 12; ++i) CHECK(forces[N * 12 + i] == 123.0, "%s: failed problem's forces were touched", name);
  for (int i = 0; i < (N + 1) * 13; ++i) CHECK(xpred[(N + 1) * 13 + i] == 321.0, "%s: failed problem's prediction was touched", name);
  /* setters + second call: a tighter force limit binds */
  CHECK(b200qp_mpc_set_fmax(h, 30.0) == B200QP_OK, "set_fmax");
  in[(size_t)1 * rd + 5] = 0.0;
  rc = b200qp_mpc_solve_batch(h, NB, in, ct, forces, xpred, info);
  CHECK(rc == B200QP_OK, "%s: second solve failed", name);
  for (int p = 0; p < NB; ++p) {
    CHECK(info[p].status == B200QP_ACADOS_SUCCESS, "%s: fmax = 30: status %d", name, info[p].status);
    for (int i = 0; i < N * 4; ++i) CHECK(forces[(size_t)p * N * 12 + 3 * i + 2] <= 30.0 + 1e-7, "%s: fmax violated", name);
    CHECK(fabs(forces[(size_t)p * N * 12 + 2] - 30.0) < 1e-6, "%s: limit should be active (m g / 4 = 36.8 N)", name);
  }
  CHECK(b200qp_mpc_solve_batch(h, NB + 1, in, ct, forces, xpred, info) == B200QP_ERR_ARG, "%s: n > max_batch must be an argument error", name);
  b200qp_mpc_destroy(h);
  free(in); free(forces); free(xpred); free(ct);
  printf("%s: ok\n", name);
  return 0;
}

static int test_wbc(void) {
  enum { NQ = 30, NEQ = 18, NIN = 28, NC = NEQ + NIN, NB = 4 };
  b200qp_wbc_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.nq = NQ; cfg.neq = NEQ; cfg.nin = NIN; cfg.max_batch = NB; cfg.device = 0;
  b200qp_wbc *h = NULL;
  int rc = b200qp_wbc_create(&cfg, &h);
  CHECK(rc == B200QP_OK, "wbc create failed (%d): %s", rc, b200qp_last_error());
  static double H[NB][NQ * NQ], g[NB][NQ], A[NB][NC * NQ], lo[NB][NC], hi[NB][NC], x[NB][NQ], want[NB][NQ];
  b200qp_solver_info info[NB];
  /* min 1/2 |x|^2 - c'x  s.t.  x_i = b_i (i < 18),  -1e20 <= x_{18+j} <= u_j (j < 12), 16 more rows that never bind */
  for (int p = 0; p < NB; ++p) {
    for (int i = 0; i < NQ; ++i) { H[p][i * NQ + i] = 1.0; g[p][i] = -(0.5 * i + p); }
    for (int i = 0; i < NEQ; ++i) { A[p][i * NC + i] = 1.0; lo[p][i] = hi[p][i] = 0.1 * i - p; want[p][i] = lo[p][i]; }
    for (int j = 0; j < NIN; ++j) {
      const int var = NEQ + (j % 12);
      A[p][var * NC + NEQ + j] = 1.0;
      lo[p][NEQ + j] = -1e20;
      hi[p][NEQ + j] = j < 12 ? 10.0 + j : 1e3;
    }
    for (int j = 0; j < 12; ++j) { const double c = 0.5 * (NEQ + j) + p; want[p][NEQ + j] = c < 10.0 + j ? c : 10.0 + j; }
  }
  rc = b200qp_wbc_solve_batch(h, NB, &H[0][0], &g[0][0], &A[0][0], &lo[0][0], &hi[0][0], &x[0][0], info);
  CHECK(rc == B200QP_OK, "wbc solve failed (%d): %s", rc, b200qp_last_error());
  for (int p = 0; p < NB; ++p) {
    CHECK(info[p].status == B200QP_ACADOS_SUCCESS, "wbc problem %d status %d kkt %g %g %g %g", p, info[p].status, info[p].kkt[0], info[p].kkt[1], info[p].kkt[2], info[p].kkt[3]);
    for (int i = 0; i < NQ; ++i) CHECK(fabs(x[p][i] - want[p][i]) <= 1e-6 * (1.0 + fabs(want[p][i])), "wbc problem %d x[%d] = %.9g, expected %.9g", p, i, x[p][i], want[p][i]);
  }
  b200qp_wbc_destroy(h);
  printf("wbc: ok\n");
  return 0;
}

/* The sequencers with their state, the WBC from the robot state and the closed loop on the device, from plain C: a robot standing still on
 * four feet (STAND gait) carries its weight, keeps its height over 20 control periods with the MPC alone and with the WBC in the loop. */
static int test_sequencer_scene_rollout(void) {
  enum { N = 10, NB = 4, PERIODS = 20 };
  b200qp_mpc_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.horizon = N;
  cfg.solver = B200QP_SOLVER_CONDENSED_ADMM;
  cfg.max_batch = NB;
  cfg.dt = 0.05;
  cfg.alpha = 1e-5;
  const double w_stand[12] = {30, 30, 30, 30, 30, 30, 1, 1, 1, 1, 1, 1};
  memcpy(cfg.state_weights, w_stand, sizeof(w_stand));
  cfg.mu = 0.6;
  cfg.fmax = 400.0;
  b200qp_mpc *h = NULL;
  CHECK(b200qp_mpc_create(&cfg, &h) == B200QP_OK, "rollout: create: %s", b200qp_last_error());
  b200qp_seq_model model;
  memset(&model, 0, sizeof(model));
  model.mass = 15.019;
  const double inertia[9] = {0.202755, 0.00012166, -0.0143742, 0.00012166, 0.490926, -3.12e-05, -0.0143742, -3.12e-05, 0.52697};
  memcpy(model.inertia, inertia, sizeof(inertia));
  const double sh[12] = {0.167, 0.1738, 0.0, 0.167, -0.1738, 0.0, -0.197, 0.1738, 0.0, -0.197, -0.1738, 0.0}; /* mit_controller_sim_go2.yaml:33-38 */
  memcpy(model.shoulders, sh, sizeof(sh));
  model.raibert_k = 0.03;
  model.raibert_g = 9.81;
  model.max_leg_length = 0.406;
  static double seq2[NB][B200QP_SEQ2_IN_DOUBLES], hist[PERIODS][NB][B200QP_ROLLOUT_HIST_DOUBLES];
  static double forces[NB][N * 12], xpred[NB][(N + 1) * 13];
  b200qp_solver_info info[NB];
  b200qp_rollout_options ro;
  memset(&ro, 0, sizeof(ro));
  ro.substeps = 10;
  ro.wbc_per_mpc = 5;
  for (int c = 0; c < 3; ++c) { ro.com_pose_Kp[c] = 300.0; ro.com_pose_Kp[3 + c] = 200.0; ro.feet_pose_Kp[c] = 1200.0; }
  for (int c = 0; c < 6; ++c) ro.com_pose_Kd[c] = 100.0;
  ro.feet_pose_Kd[0] = ro.feet_pose_Kd[1] = 500.0; ro.feet_pose_Kd[2] = 400.0;
  CHECK(b200qp_mpc_rollout(h, NULL, NB, 2, &seq2[0][0], &ro, NULL) == B200QP_ERR_ARG, "rollout before configure must be an argument error");
  CHECK(b200qp_mpc_set_model(h, &model) == B200QP_OK, "set_model");
  b200qp_seq_options so;
  memset(&so, 0, sizeof(so));
  so.mode = 0;
  so.raibert_filtersize = 20;
  so.fix_standing_position = 1;
  so.fix_position_distance_threshold = 0.1;
  so.fix_position_angular_threshold = 0.26;
  so.fix_position_velocity_threshold = 0.1;
  CHECK(b200qp_mpc_sequencer_configure(h, &so) == B200QP_OK, "sequencer_configure: %s", b200qp_last_error());
  memset(seq2, 0, sizeof(seq2));
  for (int p = 0; p < NB; ++p) {
    double *S = seq2[p];
    S[3] = 1.0;                      /* identity quaternion */
    S[4] = 0.5 * p; S[6] = 0.30;     /* position */
    for (int f = 0; f < 4; ++f) { S[13 + 3 * f] = S[4] + sh[3 * f]; S[13 + 3 * f + 1] = sh[3 * f + 1]; }
    S[25] = 1.0e6 * p;               /* clock, ns */
    S[26 + 2] = 0.30; S[26 + 9] = 1.0;                 /* Target: z, full_orientation.w */
    S[45] = (double)((1u << 2) | (1u << 3) | (1u << 4) | (1u << 10) | (1u << 11) | (1u << 13) | (1u << 14) | (1u << 15)); /* z roll pitch hybrid_x/y_dot wx wy wz */
    S[46] = 0.5;                                       /* STAND: period, duty 1, offsets 0 (gait.cpp:734-736) */
    for (int f = 0; f < 4; ++f) S[47 + f] = 1.0;
  }
  const double mg = 15.019 * 9.8067;
  int rc = b200qp_mpc_sequencer_step_solve(h, NB, &seq2[0][0], NULL, &forces[0][0], &xpred[0][0], info);
  CHECK(rc == B200QP_OK, "sequencer_step_solve: %s", b200qp_last_error());
  for (int p = 0; p < NB; ++p) {
    CHECK(info[p].status == B200QP_ACADOS_SUCCESS && info[p].n_free == 120, "sequencer_step_solve: robot %d status %d n_free %d", p, info[p].status, info[p].n_free);
    double fz = 0.0;
    for (int f = 0; f < 4; ++f) fz += forces[p][3 * f + 2];
    CHECK(fabs(fz - mg) <= 2e-3 * mg, "sequencer_step_solve: robot %d carries %g N, m g = %g", p, fz, mg);
  }
  CHECK(b200qp_mpc_sequencer_reset(h) == B200QP_OK, "sequencer_reset");
  rc = b200qp_mpc_rollout(h, NULL, NB, PERIODS, &seq2[0][0], &ro, &hist[0][0][0]);
  CHECK(rc == B200QP_OK, "rollout: %s", b200qp_last_error());
  for (int p = 0; p < NB; ++p) {
    CHECK(hist[PERIODS - 1][p][11] == 0.0, "rollout: robot %d MPC status %g", p, hist[PERIODS - 1][p][11]);
    CHECK(fabs(hist[PERIODS - 1][p][2] - 0.30) < 5e-3 && fabs(seq2[p][6] - 0.30) < 5e-3, "rollout: robot %d height %g", p, hist[PERIODS - 1][p][2]);
    CHECK(fabs(hist[PERIODS - 1][p][10] - mg) <= 5e-3 * mg, "rollout: robot %d applied f_z %g", p, hist[PERIODS - 1][p][10]);
    CHECK(fabs(seq2[p][25] - (1.0e6 * p + PERIODS * 5.0e7)) < 0.5, "rollout: robot %d clock %g", p, seq2[p][25]);
    CHECK(hist[0][p][14] == -1.0, "rollout without WBC must report WBC status -1");
  }
  /* the same with the whole-body controller in the loop: scene assembled on the device from the plant state */
  b200qp_wbc_config wc;
  memset(&wc, 0, sizeof(wc));
  wc.nq = 30; wc.neq = 18; wc.nin = 28; wc.max_batch = NB; wc.kkt_tol = 1e-6; wc.max_iter = 40;
  b200qp_wbc *w = NULL;
  CHECK(b200qp_wbc_create(&wc, &w) == B200QP_OK, "wbc create: %s", b200qp_last_error());
  CHECK(b200qp_mpc_rollout(h, w, NB, 2, &seq2[0][0], &ro, NULL) == B200QP_ERR_ARG, "rollout with an unconfigured scene must be an argument error");
  b200qp_wbc_scene_config sc;
  memset(&sc, 0, sizeof(sc));
  sc.mu = 0.45;
  sc.hessian_regularizer = 1e-8;
  for (int c = 0; c < 6; ++c) sc.com_pose_weight[c] = 10.0;
  for (int c = 0; c < 3; ++c) { sc.foot_force_weight[c] = 30.0; sc.foot_pose_weight[c] = 10.0; }
  for (int j = 0; j < 12; ++j) sc.tau_max[j] = (j % 3 == 2) ? 45.43 : 23.7;
  CHECK(b200qp_wbc_scene_configure(w, &sc) == B200QP_OK, "scene_configure: %s", b200qp_last_error());
  CHECK(b200qp_mpc_sequencer_reset(h) == B200QP_OK, "sequencer_reset");
  ro.control_dt = 0.01;
  ro.substeps = 5;
  rc = b200qp_mpc_rollout(h, w, NB, PERIODS, &seq2[0][0], &ro, &hist[0][0][0]);
  CHECK(rc == B200QP_OK, "rollout with WBC: %s", b200qp_last_error());
  for (int p = 0; p < NB; ++p) {
    CHECK(hist[PERIODS - 1][p][11] == 0.0 && hist[PERIODS - 1][p][14] == 0.0, "rollout with WBC: robot %d status %g / %g", p, hist[PERIODS - 1][p][11], hist[PERIODS - 1][p][14]);
    CHECK(fabs(hist[PERIODS - 1][p][2] - 0.30) < 2e-2, "rollout with WBC: robot %d height %g", p, hist[PERIODS - 1][p][2]);
  }
  b200qp_wbc_destroy(w);
  b200qp_mpc_destroy(h);
  printf("sequencer / scene / rollout: ok\n");
  return 0;
}

int main(void) {
  printf("%s, %d CUDA device(s)\n", b200qp_version(), b200qp_device_count());
  if (b200qp_device_count() < 1) {
    fprintf(stderr, "no CUDA device: the library has no CPU fallback\n");
    return 2;
  }
  if (test_mpc(B200QP_SOLVER_CONDENSED_ADMM, "mpc CONDENSED_ADMM")) return 1;
  if (test_mpc(B200QP_SOLVER_SPARSE_RICCATI, "mpc SPARSE_RICCATI")) return 1;
  if (test_mpc(B200QP_SOLVER_PARTIAL_CONDENSING, "mpc PARTIAL_CONDENSING")) return 1;
  if (test_mpc(B200QP_SOLVER_AUTO, "mpc AUTO")) return 1;
  if (test_wbc()) return 1;
  if (test_sequencer_scene_rollout()) return 1;
  printf("all boundary checks passed\n");
  return 0;
}
