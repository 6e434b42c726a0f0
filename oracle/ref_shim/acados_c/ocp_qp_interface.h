/* acados_c/ocp_qp_interface.h - TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * Stand-in for the part of the acados v0.3.6 C interface that the reference's MPC class calls
 * (/root/reference/ws/src/controllers/src/mit_controller/mpc.cpp:175-282, :366-467), so that mpc.cpp compiles
 * UNMODIFIED.  acados (+ HPIPM / BLASFEO / OSQP) is third party, not vendored in /root/reference and not installed here.
 * Names, argument order and the column-major layout of the matrices follow the published acados C API; the structures
 * are this repo's own and only carry what mpc.cpp reads.  ocp_qp_solve() hands the stage data, as the reference code
 * assembled it, to the oracle's QP solver (oracle/mpc_ref.c) - the data assembly is the reference's, the solve is a
 * restatement. */
#ifndef DFKIREF_ACADOS_C_OCP_QP_INTERFACE_H_
#define DFKIREF_ACADOS_C_OCP_QP_INTERFACE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* acados/utils/types.h return_values (names as used at mpc.cpp:413-425; numeric order per upstream) */
enum return_values { ACADOS_SUCCESS, ACADOS_NAN_DETECTED, ACADOS_MAXITER, ACADOS_MINSTEP, ACADOS_QP_FAILURE, ACADOS_READY };

/* acados_c/ocp_qp_interface.h ocp_qp_solver_t: partial-condensing plans first, FULL_CONDENSING_HPIPM is "the first solver
 * after partials" (mpc.cpp:219) */
typedef enum {
  PARTIAL_CONDENSING_HPIPM,
  PARTIAL_CONDENSING_OSQP,
  PARTIAL_CONDENSING_QPDUNES,
  FULL_CONDENSING_HPIPM,
  FULL_CONDENSING_QPOASES,
  FULL_CONDENSING_DAQP,
  INVALID_QP_SOLVER
} ocp_qp_solver_t;

typedef struct { ocp_qp_solver_t qp_solver; } ocp_qp_solver_plan_t;

#define DFKIREF_MAX_STAGES 64
typedef struct ocp_qp_dims {
  int N;
  int nx[DFKIREF_MAX_STAGES + 1], nu[DFKIREF_MAX_STAGES + 1], nbx[DFKIREF_MAX_STAGES + 1], nbu[DFKIREF_MAX_STAGES + 1],
      ng[DFKIREF_MAX_STAGES + 1];
} ocp_qp_dims;

typedef struct ocp_qp_xcond_solver_config { ocp_qp_solver_plan_t plan; } ocp_qp_xcond_solver_config;
typedef struct ocp_qp_xcond_solver_dims { ocp_qp_dims *orig_dims; } ocp_qp_xcond_solver_dims;
typedef struct ocp_qp_xcond_solver_opts {
  int cond_N;
  char hpipm_mode[32];
  int warm_start;
} ocp_qp_xcond_solver_opts;

/* one stage of problem data, column-major like acados' setters expect (mpc.hpp:37-47) */
typedef struct dfkiref_stage {
  double *A, *B, *Q, *R, *q, *r, *C, *D, *lg, *ug, *lbx, *ubx, *lbu, *ubu, *Jbx, *Jbu;
} dfkiref_stage;
typedef struct ocp_qp_in {
  ocp_qp_dims *dim; /* -> own_dim: like acados, in / out carry their own copy (the reference frees the dims first) */
  ocp_qp_dims own_dim;
  dfkiref_stage stage[DFKIREF_MAX_STAGES + 1];
} ocp_qp_in;

struct blasfeo_dvec { double *pa; int m; }; /* ux[k]->pa = [u_k; x_k] as in HPIPM */
typedef struct qp_info {
  double solve_QP_time, condensing_time, interface_time, total_time;
  int num_iter, t_computed;
} qp_info;
typedef struct ocp_qp_out {
  ocp_qp_dims *dim;
  ocp_qp_dims own_dim;
  struct blasfeo_dvec *ux;
  void *misc; /* qp_info */
  double kkt[4];
} ocp_qp_out;

typedef struct dense_qp_dims { int nv, nb, ne, ng, ns; } dense_qp_dims;
typedef struct dense_qp_in { dense_qp_dims *dim; } dense_qp_in;
typedef struct ocp_qp_xcond_solver_memory { void *xcond_qp_in; } ocp_qp_xcond_solver_memory;

typedef struct ocp_qp_solver {
  ocp_qp_xcond_solver_config *config;
  ocp_qp_xcond_solver_opts *opts;
  void *mem; /* ocp_qp_xcond_solver_memory or ocp_qp_full_condensing_memory, as in acados */
} ocp_qp_solver;

ocp_qp_dims *ocp_qp_dims_create(int N);
void ocp_qp_dims_free(void *dims);
ocp_qp_xcond_solver_config *ocp_qp_xcond_solver_config_create(ocp_qp_solver_plan_t plan);
void ocp_qp_xcond_solver_config_free(ocp_qp_xcond_solver_config *config);
void ocp_qp_dims_set(void *config, void *dims, int stage, const char *field, int *value);
ocp_qp_in *ocp_qp_in_create(ocp_qp_dims *dims);
void ocp_qp_in_set(ocp_qp_xcond_solver_config *config, ocp_qp_in *in, int stage, char *field, void *value);
void ocp_qp_in_free(void *in);
ocp_qp_xcond_solver_dims *ocp_qp_xcond_solver_dims_create_from_ocp_qp_dims(ocp_qp_xcond_solver_config *config,
                                                                           ocp_qp_dims *dims);
void ocp_qp_xcond_solver_dims_free(ocp_qp_xcond_solver_dims *dims);
void *ocp_qp_xcond_solver_opts_create(ocp_qp_xcond_solver_config *config, ocp_qp_xcond_solver_dims *dims);
void ocp_qp_xcond_solver_opts_set(ocp_qp_xcond_solver_config *config, ocp_qp_xcond_solver_opts *opts, const char *field,
                                  void *value);
void ocp_qp_xcond_solver_opts_free(ocp_qp_xcond_solver_opts *opts);
ocp_qp_solver *ocp_qp_create(ocp_qp_xcond_solver_config *config, ocp_qp_xcond_solver_dims *dims, void *opts);
void ocp_qp_solver_destroy(ocp_qp_solver *solver);
ocp_qp_out *ocp_qp_out_create(ocp_qp_dims *dims);
void ocp_qp_out_free(void *out);
int ocp_qp_solve(ocp_qp_solver *solver, ocp_qp_in *in, ocp_qp_out *out);
void d_ocp_qp_sol_get_u(int stage, ocp_qp_out *out, double *u);
void d_ocp_qp_sol_get_x(int stage, ocp_qp_out *out, double *x);
void ocp_qp_inf_norm_residuals(ocp_qp_dims *dims, ocp_qp_in *in, ocp_qp_out *out, double *res);

#ifdef __cplusplus
}
#endif
#endif
