/*
 * b200qp.h - C ABI of the B200-native batched MPC / WBC QP engine.
 *
 * Drop-in boundary for ONE hot path of dfki-ric-underactuated-lab/dfki-quad: the MPC solver plugin
 *   class MPCInterface { UpdateState; UpdateModel; UpdateGaitSequence; GetWrenchSequence }
 *   (ws/src/controllers/include/mit_controller/mpc_interface.hpp:22-33), implemented there by class MPC
 *   (ws/src/controllers/include/mit_controller/mpc.hpp:13-144, src/mit_controller/mpc.cpp) on top of acados.
 * The reference has no FFI of its own; these are the entry points a cgo/JNI/ctypes/C++ adapter for that class
 * binds (INTEGRATION.md shows the C++ adapter).  Plain pointers and sizes only; no torch / Eigen / ROS types.
 *
 * Batch semantics: every call processes `n` independent problems (robots / sampled states).  All matrices that
 * cross the ABI are column-major like the reference's acados pointers (mpc.hpp:37-47).
 * Quaternions are (x, y, z, w) = Eigen::Quaterniond::coeffs() order.
 * There is NO CPU fallback: every solve entry point fails with B200QP_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef B200QP_H_
#define B200QP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200QP_NX 13            /* MPC::STATE_SIZE            mpc.hpp:15 */
#define B200QP_NU 12            /* MPC::INPUT_SIZE            mpc.hpp:17 */
#define B200QP_NLEGS 4          /* N_LEGS                     mit_controller_params.hpp:3 */
#define B200QP_MAX_HORIZON 20   /* MPC_PREDICTION_HORIZON of the checkout, mit_controller_params.hpp:6 */
#define B200QP_GAIT_SEQUENCE_SIZE 100 /* GAIT_SEQUENCE_SIZE   mit_controller_params.hpp:5 */

/* function return codes (argument / CUDA errors only; per-problem solver results are in b200qp_solver_info) */
#define B200QP_OK 0
#define B200QP_ERR_ARG (-1)
#define B200QP_ERR_CUDA (-2)
#define B200QP_ERR_UNSUPPORTED (-3)

/* per-problem status = acados `return_values` as named at mpc.cpp:413-425 (numeric order per upstream
 * acados/utils/types.h), plus the reference's own -7777 for NaN-after-success (mpc.cpp:452). */
#define B200QP_ACADOS_SUCCESS 0
#define B200QP_ACADOS_NAN_DETECTED 1
#define B200QP_ACADOS_MAXITER 2
#define B200QP_ACADOS_MINSTEP 3
#define B200QP_ACADOS_QP_FAILURE 4
#define B200QP_ACADOS_READY 5
#define B200QP_NAN_AFTER_SUCCESS (-7777)

/* solver selection, named after the reference's acados plans (mit_controller_node.cpp:502-519) */
#define B200QP_SOLVER_CONDENSED_ADMM 0 /* replaces PARTIAL_CONDENSING_OSQP with small cond_N / FULL_CONDENSING_* */
#define B200QP_SOLVER_SPARSE_RICCATI 1 /* replaces PARTIAL_CONDENSING_HPIPM near the sparse end: Riccati interior-point     */
                                       /* method, one warp per problem; cond_N = N is the sparse formulation proper, the   */
                                       /* default (cond_N = 0) condenses pairs of stages; blocks are capped at 12 inputs   */
#define B200QP_SOLVER_PARTIAL_CONDENSING 3 /* replaces PARTIAL_CONDENSING_HPIPM for ANY cond_N in 1..N (mpc.cpp:219-225):  */
                                       /* block-condensed Riccati interior-point kernel, b200qp_mpc_config.cond_N blocks;  */
                                       /* cond_N = N is the sparse formulation, cond_N = 1 the fully condensed one          */
#define B200QP_SOLVER_AUTO 2           /* per problem: condensed kernel when the reduced size fits (<= 156 free     */
                                       /* variables), sparse Riccati kernel otherwise; same optimum either way;     */
                                       /* problems the condensed kernel gives up on are retried on the Riccati one  */

/* Replaces the MPC constructor arguments (mpc.hpp:117-127) + compile-time constants (mit_controller_params.hpp). */
typedef struct b200qp_mpc_config {
  int32_t horizon;            /* N, 1..B200QP_MAX_HORIZON (reference: compile-time MPC_PREDICTION_HORIZON)        */
  int32_t solver;             /* B200QP_SOLVER_*                                                                 */
  int32_t max_batch;          /* device buffers are sized for this many problems at create                      */
  int32_t device;             /* CUDA device ordinal                                                            */
  double dt;                  /* MPC_DT                                                                         */
  double alpha;               /* input weight, R = alpha I        (mpc.cpp:166)                                  */
  double state_weights[12];   /* Q = diag(w, 0)                   (mpc.cpp:161-165)                              */
  double mu;                  /* friction coefficient             (mpc.cpp:66-71)                                */
  double fmin;                /* (mpc.cpp:94-115)                                                                */
  double fmax;
  /* solver options (0 => library default) */
  double admm_rho;            /* initial ADMM penalty            default 5 alpha in [1e-6, 5e-4] (5e-5 at 1e-5) */
  double admm_sigma;          /* ADMM proximal weight                                     default 1e-6          */
  double admm_alpha;          /* ADMM over-relaxation                                     default 1.6           */
  double admm_eps;            /* relative tolerance at which ADMM hands over to polish    default 1e-3          */
  double kkt_tol;             /* accept when all four KKT inf-norms <= kkt_tol            default 1e-7          */
  int32_t admm_max_iter;      /* total ADMM iteration cap per problem                     default 2000          */
  int32_t admm_check_every;   /* residual check / rho adaptation interval                 default 15            */
  int32_t polish_max_rounds;  /* active-set repair sweeps per polish                      default 8             */
  int32_t ipm_max_iter;       /* Riccati IPM iteration cap                                default 40            */
  int32_t warm_start;         /* MPC::MPC warm_start (mpc.hpp:127, `mpc_warm_start`, mit_controller_node.cpp:65): 0 cold start;   */
                              /* 1 warm start - the ADMM iterate of problem slot p starts from the previous solution of slot p    */
                              /* (primal, duals) of this handle; 2 hot start - the previous active set is tried first (one        */
                              /* polish round, no factorisation of K) and ADMM only runs, warm-started, if that is not optimal.   */
                              /* Condensed kernel only; the Riccati IPM always starts cold.                                       */
  int32_t cond_N;             /* B200QP_SOLVER_PARTIAL_CONDENSING: number of blocks the horizon is condensed into, the reference's */
                              /* `condensing_N` constructor argument (mpc.hpp:125, mpc.cpp:219-225; HPIPM block sizes: N / cond_N,  */
                              /* the first N mod cond_N blocks one longer).  0 => library default (blocks of 4 stages).  Every      */
                              /* value gives the same optimum; blocks with more than 128 stacked stance inputs are split.           */
  int32_t precision;          /* condensed kernel only.  0: FP64 throughout (the product; the reference computes in double).        */
                              /* 1, 2: FP32 ERROR STUDY (north star: "with the FP32 error stated separately") - every stored          */
                              /* intermediate of (1) K^-1 and the ADMM iterations, with assembly and polish still in FP64 ("FP32       */
                              /* factor + FP64 refinement"), or (2) of everything incl. assembly and polish, is rounded to binary32.  */
                              /* With 2 and kkt_tol == 0 the acceptance tolerance is 5e-2 (an FP32 polish cannot reach 1e-7).         */
} b200qp_mpc_config;

/* Input record of one problem = the subset of {StateInterface, ModelInterface, GaitSequence} that
 * MPC::GetWrenchSequence reads (mpc.cpp:332-396), as a flat FP64 array of b200qp_mpc_in_doubles(N) values:
 *   [0:4)    body orientation quaternion (x,y,z,w)          StateInterface::GetOrientationInWorld
 *   [4:7)    body position                                  GetPositionInWorld
 *   [7:10)   angular velocity, world frame                  GetAngularVelInWorld
 *   [10:13)  linear velocity, world frame                   GetLinearVelInWorld
 *   [13]     mass                                           ModelInterface::GetMass
 *   [14:23)  body inertia 3x3, column-major                 GetInertia
 *   [23:26)  body -> COM translation                        GetBodyToCOM().translation()
 *   then for k = 0..N (13 doubles each):
 *            desired_reference_trajectory_orientation[k] (4), desired_reference_trajectory_position[k] (3),
 *            reference_trajectory_twist[k] (3), reference_trajectory_velocity[k] (3)      gait_sequence.hpp:22-29
 *   then for k = 0..N-1 (12 doubles each): foot_position_sequence[k][leg 0..3] (3 each)   gait_sequence.hpp:22
 * plus a separate uint8 array contact[n][N][4] = contact_sequence rows 0..N-1             gait_sequence.hpp:21 */
static inline int32_t b200qp_mpc_in_doubles(int32_t horizon) { return 26 + 13 * (horizon + 1) + 12 * horizon; }

/* Per-problem result record; replaces SolverInformation (mpc_interface.hpp:10-20). */
typedef struct b200qp_solver_info {
  int32_t status;     /* B200QP_ACADOS_* / B200QP_NAN_AFTER_SUCCESS; success <=> status == 0 (mpc.cpp:445)      */
  int32_t iterations; /* ADMM or IPM iterations (acados_num_iter, mpc.cpp:442)                                  */
  int32_t polish_rounds; /* active-set solves done by the polish (0 for the IPM)                                */
  int32_t n_free;     /* number of free (stance) force variables of the reduced QP                              */
  double kkt[4];      /* inf-norms: stationarity, dynamics, inequality, complementarity (the 4-tuple of the
                         commented-out ocp_qp_inf_norm_residuals call, mpc.cpp:430-433)                         */
} b200qp_solver_info;

typedef struct b200qp_mpc b200qp_mpc; /* opaque handle: one per host thread / per GPU; not thread-safe           */

/* MPC::MPC(...)  mpc.cpp:117-272 */
int32_t b200qp_mpc_create(const b200qp_mpc_config *cfg, b200qp_mpc **out);
/* MPC::~MPC()    mpc.cpp:274-282 */
void b200qp_mpc_destroy(b200qp_mpc *h);
/* MPC::SetStateWeights / SetInputWeights / SetFmax / SetMu   mpc.cpp:294-326, mpc.hpp:138-141 */
int32_t b200qp_mpc_set_state_weights(b200qp_mpc *h, const double weights[12]);
int32_t b200qp_mpc_set_input_weights(b200qp_mpc *h, double alpha);
int32_t b200qp_mpc_set_fmax(b200qp_mpc *h, double fmax);
int32_t b200qp_mpc_set_mu(b200qp_mpc *h, double mu);
/* Warm-start mode (see b200qp_mpc_config.warm_start); any call also forgets the stored solutions of all slots. */
int32_t b200qp_mpc_set_warm_start(b200qp_mpc *h, int32_t mode);

/* MPC::GetWrenchSequence for `n` problems, HOST buffers (mpc.cpp:329-470).
 *   in       [n][b200qp_mpc_in_doubles(N)]   input records (see above)
 *   contact  [n][N][4]                        uint8 0/1
 *   forces   [n][N][4][3]                     WrenchSequence::forces (wrench_sequence.hpp:9); world-frame GRFs
 *   xpred    [n][N+1][13]                     MPCPrediction::raw_data (mpc_prediction.hpp:12)
 *   info     [n]
 * Problems whose status != 0 leave their forces/xpred slots untouched (mpc.cpp:447-455).
 * Copies H2D, runs the kernels, copies D2H, synchronises. */
int32_t b200qp_mpc_solve_batch(b200qp_mpc *h, int32_t n, const double *in, const uint8_t *contact, double *forces,
                               double *xpred, b200qp_solver_info *info);
/* Same with DEVICE pointers (inputs already resident in HBM); asynchronous on the handle's stream unless sync != 0. */
int32_t b200qp_mpc_solve_batch_device(b200qp_mpc *h, int32_t n, const double *d_in, const uint8_t *d_contact,
                                      double *d_forces, double *d_xpred, b200qp_solver_info *d_info, int32_t sync);
/* Optional diagnostics of the last device solve: duals in the reference's row order, per problem
 *   lam_box [n][N][12]  multipliers of lbu/ubu rows (>0 at upper, <0 at lower),  lam_g [n][N][16] of lg/ug rows.
 * Host buffers. Valid for problems with status == 0. */
int32_t b200qp_mpc_get_duals(b200qp_mpc *h, int32_t n, double *lam_box, double *lam_g);
/* Debug/parity hook: condensed reduced Hessian and gradient of problem `index` of the last solve
 * (SURVEY.md section 7 step 5).  H is [n_free x n_free] column-major, g is [n_free]; free variables are ordered by
 * (stage, leg, xyz) over stance feet.  Returns n_free or a negative error. Requires the handle to have been
 * created with max_batch problems of scratch (always true). */
int32_t b200qp_mpc_get_condensed(b200qp_mpc *h, int32_t index, double *H, int32_t ldh, double *g);
/* Device time (ms, CUDA events on the handle's stream) of the kernels of the last solve call, and launch count. */
int32_t b200qp_mpc_last_timing(b200qp_mpc *h, float *kernel_ms, int32_t *kernel_launches);
/* CUDA stream of the handle as a void* (cudaStream_t), for callers that time with their own events. */
void *b200qp_mpc_stream(b200qp_mpc *h);

/* ---- one batch, several GPUs ----------------------------------------------------------------
 * The problems of ONE host array are cut into contiguous slices of ceil(n / G) problems, slice g goes to devices[g]; every
 * GPU has its own handle, stream, pinned staging buffers and host thread, results land in disjoint ranges of the caller's
 * output arrays; there is no collective and no inter-GPU traffic on the solve path.  cfg->max_batch is the size of the
 * whole batch, cfg->device is ignored.  Page-locked caller buffers (cudaHostAlloc / cudaHostRegister with the portable
 * flag, torch pin_memory) are used in place by all GPUs. */
typedef struct b200qp_mpc_sharded b200qp_mpc_sharded;
int32_t b200qp_mpc_create_sharded(const b200qp_mpc_config *cfg, const int32_t *devices, int32_t n_devices, b200qp_mpc_sharded **out);
void b200qp_mpc_destroy_sharded(b200qp_mpc_sharded *h);
/* arguments as b200qp_mpc_solve_batch */
int32_t b200qp_mpc_solve_batch_sharded(b200qp_mpc_sharded *h, int32_t n, const double *in, const uint8_t *contact, double *forces,
                                       double *xpred, b200qp_solver_info *info);
/* the per-GPU handle of slice g (for the setters: b200qp_mpc_set_state_weights etc. are applied per handle), or NULL */
b200qp_mpc *b200qp_mpc_sharded_handle(b200qp_mpc_sharded *h, int32_t g);
int32_t b200qp_mpc_sharded_devices(const b200qp_mpc_sharded *h);
/* the slicing rule itself (pure arithmetic, no CUDA): problems [*lo, *hi) of n belong to slice `rank` of `world` */
int32_t b200qp_shard_range(int32_t n, int32_t rank, int32_t world, int32_t *lo, int32_t *hi);

/* ---- on-device gait sequencer -----------------------------------------------------
 * Replaces, for n robots at once, the caller side of MPC::UpdateGaitSequence: one SimpleGaitSequencer pass
 *   SimpleGaitSequencer::GetGaitSequence            simple_gait_sequencer.cpp:43-60
 *   Gait::update_sequence                           gait.cpp:67-112
 *   MPCTrajectoryPlanner::plan_trajectory           mpc_trajectory_planner.cpp:28-140 (+ plan_desired_trajectory,
 *       :142-264, first call; Target activation of the controller node, mit_controller_node.cpp:145-165)
 *   RaibertFootStepPlanner::get_foot_position_sequence  raibert_foot_step_planner.cpp:43-169 (z_on_plane = false)
 * so the host->device boundary is (state, feet, Target, gait, phase) = 400 B per robot instead of the 2.3 KB record.
 * Sequencer input per robot, B200QP_SEQ_IN_DOUBLES doubles:
 *   [0:4) quat xyzw  [4:7) position  [7:10) angular velocity (world)  [10:13) linear velocity   StateInterface
 *   [13:25) current foot positions, world, leg-major                  StateInterface::GetFeetPositions
 *   [25:28) filtered velocity (RaibertFootStepPlanner's moving average, kept by the caller)
 *   [28:32) last stance height per foot                               RaibertFootStepPlanner::last_foot_heights_
 *   [32:40) Target: hybrid_x_dot, hybrid_y_dot, wz, wx, wy, roll, pitch, z   interfaces::msg::QuadControlTarget
 *   [40] gait phase  [41] period  [42:46) duty factor per leg  [46:50) phase offset per leg      Gait members */
#define B200QP_SEQ_IN_DOUBLES 50
typedef struct b200qp_seq_model {
  double mass;            /* QuadModel mass                                   quad_model_symbolic.cpp        */
  double inertia[9];      /* body inertia, row-major                                                          */
  double com[3];          /* body-frame COM offset                                                            */
  double shoulders[12];   /* shoulder offsets in the body frame, leg-major    raibert_foot_step_planner.cpp:28-34 */
  double raibert_k;       /* raibert.k                                        mit_controller_sim_go2.yaml    */
  double raibert_g;       /* RaibertFootStepPlanner g                         raibert_foot_step_planner.hpp  */
  double max_leg_length;  /* FK(0) leg length - 0.02                          raibert_foot_step_planner.cpp:36-40 */
} b200qp_seq_model;
/* Model constants used by the sequencer entry points (MPC::UpdateModel + planner constructor arguments). */
int32_t b200qp_mpc_set_model(b200qp_mpc *h, const b200qp_seq_model *model);
/* Sequencer only: HOST seq_in[n][50], measured_contact[n][4] or NULL -> HOST records[n][b200qp_mpc_in_doubles(N)] and
 * contact[n][N][4] in exactly the layout b200qp_mpc_solve_batch consumes (parity hook for the sequencer kernel). */
int32_t b200qp_mpc_sequence_batch(b200qp_mpc *h, int32_t n, const double *seq_in, const uint8_t *measured_contact,
                                  double *records, uint8_t *contact);
/* Sequencer + QP assembly + solve in one call; outputs and failure semantics as b200qp_mpc_solve_batch. */
int32_t b200qp_mpc_sequence_solve_batch(b200qp_mpc *h, int32_t n, const double *seq_in,
                                        const uint8_t *measured_contact, double *forces, double *xpred,
                                        b200qp_solver_info *info);
/* Same with DEVICE pointers on the handle's stream (d_measured_contact may be NULL). */
int32_t b200qp_mpc_sequence_solve_batch_device(b200qp_mpc *h, int32_t n, const double *d_seq_in,
                                               const uint8_t *d_measured_contact, double *d_forces, double *d_xpred,
                                               b200qp_solver_info *d_info, int32_t sync);

/* ---- the gait sequencers with their state across calls (kernel K5b) ------------------------------------------------------
 * What b200qp_mpc_sequence_* above does for ONE call of a freshly constructed SimpleGaitSequencer, these entry points do for
 * the sequencers as the controller node runs them, call after call: the state the reference keeps in its objects lives in
 * a per-robot record on the device (slot p of the handle <-> robot p) -
 *   SimpleGaitSequencer::phase_ / time_ (phase = fmod((t - t0) / T, 1), 0 for a standing gait)   simple_gait_sequencer.cpp:62-80
 *   MPCTrajectoryPlanner::sequence_initialized_, foot_heights_ and the previous desired row 1 (stand-still latch)
 *                                                                                      mpc_trajectory_planner.cpp:28-264
 *   RaibertFootStepPlanner::velocity_filter_, last_foot_heights_                          raibert_foot_step_planner.cpp:25-129
 *   AdaptiveGait: leg phases, period, duty factors, phase offsets, velocity filter       gait.cpp:363-700 (mode 1)
 * and every field / activation flag of `Target` (target.hpp) is honoured.
 * Input per robot and call, B200QP_SEQ2_IN_DOUBLES doubles:
 *   [0:13) quat xyzw, position, angular velocity (world), linear velocity     [13:25) current foot positions, world
 *   [25]   StateInterface::GetTime() in NANOSECONDS (an integer held in a double)
 *   [26:45) Target values: x y z roll pitch yaw | full_orientation xyzw | x_dot y_dot z_dot | hybrid_x_dot hybrid_y_dot
 *           hybrid_z_dot | wx wy wz
 *   [45]   Target::active as a bit mask (bit i = i-th flag of Target::activateTarget in declaration order, target.hpp:28-45)
 *   [46] period  [47:51) duty factors  [51:55) phase offsets of the Gait (mode 0; ignored by the adaptive gait)
 * measured_contact [n][4] = StateInterface::GetFeetContacts() (NULL: all feet in contact). */
#define B200QP_SEQ2_IN_DOUBLES 64
typedef struct b200qp_seq_options {
  int32_t mode;                     /* 0: SimpleGaitSequencer, 1: AdaptiveGaitSequencer                                     */
  int32_t raibert_filtersize;       /* raibert.filtersize (mit_controller_sim_go2.yaml: 20), <= 32                          */
  int32_t fix_standing_position;    /* fix_standing_position                                                                */
  int32_t early_contact_detection;  /* early_contact_detection                                                              */
  double fix_position_distance_threshold, fix_position_angular_threshold, fix_position_velocity_threshold;
  /* AdaptiveGait constructor arguments (gait.hpp:123-139; defaults mit_controller_node.cpp:250-264) */
  double swing_time, zero_velocity_threshold, standing_foot_position_threshold, min_v_cmd_factor, max_correction_cycles,
      correction_period, disturbance_correction, min_v, offset_delay, max_stride_length, phase_offset[4];
  double control_dt;                /* MPC_CONTROL_DT, the dt AdaptiveGaitSequencer passes to update_sequence                */
  int32_t filter_size, switch_offsets, correct_all;
} b200qp_seq_options;
/* Sets the options (b200qp_mpc_set_model must have been called) and forgets the state of every slot. */
int32_t b200qp_mpc_sequencer_configure(b200qp_mpc *h, const b200qp_seq_options *opt);
/* Forgets the state of every slot (as if the sequencers were constructed anew). */
int32_t b200qp_mpc_sequencer_reset(b200qp_mpc *h);
/* One UpdateState + UpdateTarget + GetGaitSequence per robot -> records / contact in the layout b200qp_mpc_solve_batch
 * consumes (HOST buffers; parity hook). */
int32_t b200qp_mpc_sequencer_step(b200qp_mpc *h, int32_t n, const double *seq2_in, const uint8_t *measured_contact, double *records,
                                  uint8_t *contact);
/* The same followed by QP assembly + solve; outputs and failure semantics as b200qp_mpc_solve_batch. */
int32_t b200qp_mpc_sequencer_step_solve(b200qp_mpc *h, int32_t n, const double *seq2_in, const uint8_t *measured_contact, double *forces,
                                        double *xpred, b200qp_solver_info *info);

/* ---- gait schedule -> contact mask -> bound indexing (kernel K4, bit-exact) -------------------------------------
 * Gait::update_sequence (gait.cpp:67-112) for n robots with per-robot gait parameters and phase, followed by
 * MPC::GetUBounds (mpc.cpp:94-115).
 *   period[n], duty[n][4], offset[n][4], phase[n]         Gait members / SimpleGaitSequencer::phase_
 *   measured_contact[n][4] or NULL                        StateInterface::GetFeetContacts (early-contact rule)
 *   steps                                                 rows to produce (<= B200QP_GAIT_SEQUENCE_SIZE)
 *   contact[n][steps][4] uint8, swing_time[n][steps][4] (may be NULL), lbu/ubu[n][steps][12] (may be NULL)
 * HOST buffers. */
int32_t b200qp_gait_schedule(int32_t device, int32_t n, int32_t steps, double dt, const double *period,
                             const double *duty, const double *offset, const double *phase,
                             const uint8_t *measured_contact, double fmin, double fmax, uint8_t *contact,
                             double *swing_time, double *lbu, double *ubu);
/* Same with DEVICE pointers on `stream` (cudaStream_t as void*, may be NULL). */
int32_t b200qp_gait_schedule_device(int32_t n, int32_t steps, double dt, const double *d_period, const double *d_duty,
                                    const double *d_offset, const double *d_phase, const uint8_t *d_measured_contact,
                                    double fmin, double fmax, uint8_t *d_contact, double *d_swing_time, double *d_lbu,
                                    double *d_ubu, void *stream);

/* ---- whole-body-control task-space QP (kernel K3) ------------------------------------------------------------------
 * Replaces the `wbc::QPSolver` plugin that ARC-OPT's scene calls in `wbc_scene_->solve(qp)`
 * (ws/src/controllers/src/mit_controller/wbc_arc_opt.cpp:51-74,297) for n QPs at once:
 *     min 1/2 x'Hx + g'x   s.t.   A[0:neq] x = lower_y[0:neq],   lower_y[neq:] <= A[neq:] x <= upper_y[neq:]
 * Layouts follow ARC-OPT's wbc::QuadraticProgram (Eigen, column-major): per problem H [nq x nq], g [nq],
 * A [(neq+nin) x nq] column-major, lower_y / upper_y [neq+nin]; |bound| >= 1e19 means "no bound". */
typedef struct b200qp_wbc_config {
  int32_t nq;         /* decision variables (reduced TSID on Go2: 18 accelerations + 12 contact forces = 30), <= 48 */
  int32_t neq;        /* equality rows (6 floating-base dynamics + 3 per foot = 18), <= 32                          */
  int32_t nin;        /* two-sided inequality rows (12 torque limits + 16 friction pyramid = 28), <= 64             */
  int32_t max_batch;
  int32_t device;
  int32_t max_iter;   /* IPM iteration cap, 0 => 40                                                                */
  double kkt_tol;     /* acceptance tolerance on the KKT inf-norms, 0 => 1e-6 (internal target 1e-4 of it)          */
} b200qp_wbc_config;
typedef struct b200qp_wbc b200qp_wbc;
int32_t b200qp_wbc_create(const b200qp_wbc_config *cfg, b200qp_wbc **out);
void b200qp_wbc_destroy(b200qp_wbc *h);
/* HOST buffers; x [n][nq] is written only for problems with info.status == 0 (on a solver failure the reference
 * falls back to zero torque, wbc_arc_opt.cpp:298-309 - the caller's job).  info.kkt = stationarity, equality,
 * inequality, complementarity inf-norms.  info.iterations counts both attempts where the kernel's aggressive first
 * attempt was given up and the QP solved again conservatively (a few QPs per 10 000).  Batches >= 8192 are cut into
 * chunks that alternate between two internal streams; the call returns after everything has landed, and work queued on
 * b200qp_wbc_stream() before / after the call stays ordered before / after it. */
int32_t b200qp_wbc_solve_batch(b200qp_wbc *h, int32_t n, const double *H, const double *g, const double *A,
                               const double *lower_y, const double *upper_y, double *x, b200qp_solver_info *info);
/* DEVICE pointers, asynchronous on the handle's stream unless sync != 0 */
int32_t b200qp_wbc_solve_batch_device(b200qp_wbc *h, int32_t n, const double *d_H, const double *d_g, const double *d_A,
                                      const double *d_lower_y, const double *d_upper_y, double *d_x,
                                      b200qp_solver_info *d_info, int32_t sync);
void *b200qp_wbc_stream(b200qp_wbc *h);

/* ---- the WBC scene assembled on the device (kernel K6) -------------------------------------------------------------------
 * What `scene->update()` computes on the host in the reference before `scene->solve()` (wbc_arc_opt.cpp:292-297) - Pinocchio's
 * M(q), h(q, nu), foot Jacobians and Jdot nu for the Go2 (wbc_arc_opt.cpp:30-47,236-241; URDF go2_description.urdf) and the
 * reduced-TSID QP (task set / weights wbc_arc_opt.cpp:78-123, mit_controller_sim_go2.yaml:62-72) - runs on the GPU, so a solve
 * takes the robot's state (576 B) instead of the assembled QP (19.5 KB).  Handle must be created with nq = 30, neq = 18, nin = 28.
 * scene_in per robot, B200QP_WBC_SCENE_IN_DOUBLES doubles:
 *   [0:4) base quaternion xyzw  [4:7) base position  [7:10) base linear velocity (world)  [10:13) base angular velocity (world)
 *   [13:25) joint positions (fl, fr, bl, br x abad, hip, knee)  [25:37) joint velocities
 *   [37:43) desired base acceleration (linear, angular; CartesianPosPDController output, wbc_arc_opt.cpp:256-289)
 *   [43:55) desired foot accelerations  [55:67) wrench references = MPC GRFs of stage 0 (mit_controller_node.cpp:1006)
 *   [67] feet in contact, bit f = foot f (UpdateFootContact) */
#define B200QP_WBC_SCENE_IN_DOUBLES 72
typedef struct b200qp_wbc_scene_config {
  double mu;                      /* wbc.friction_coefficient (0.45)                */
  double com_pose_weight[6];      /* wbc.com_pose_weight: linear xyz, angular xyz   */
  double foot_force_weight[3];    /* wbc.feet_force_weight                          */
  double foot_pose_weight[3];     /* wbc.feet_pose_weight                           */
  double hessian_regularizer;     /* ARC-OPT scene regulariser (1e-8)               */
  double tau_max[12];             /* joint effort limits of the URDF                */
} b200qp_wbc_scene_config;
int32_t b200qp_wbc_scene_configure(b200qp_wbc *h, const b200qp_wbc_scene_config *cfg);
/* HOST scene_in[n][72] -> x[n][30] = [nu_dot(18), f(12)], tau[n][12] = joint torques of the solution (may be NULL), info[n].
 * A QP whose solve fails leaves its x / tau rows untouched. */
int32_t b200qp_wbc_scene_solve_batch(b200qp_wbc *h, int32_t n, const double *scene_in, double *x, double *tau, b200qp_solver_info *info);
/* Assembly only (parity hook): HOST scene_in -> HOST H[n][900], g[n][30], A[n][1380], lower_y / upper_y[n][46] in the layout of
 * b200qp_wbc_solve_batch, feet[n][24] = foot positions and velocities (world; may be NULL). */
int32_t b200qp_wbc_scene_assemble(b200qp_wbc *h, int32_t n, const double *scene_in, double *H, double *g, double *A, double *lower_y,
                                  double *upper_y, double *feet);
/* Device-resident variant: d_scene_in[n][72] -> d_x, d_tau (may be NULL), d_info; work is queued on the handle's stream. */
int32_t b200qp_wbc_scene_solve_batch_device(b200qp_wbc *h, int32_t n, const double *d_scene_in, double *d_x, double *d_tau,
                                            b200qp_solver_info *d_info, int32_t sync);

/* ---- the closed loop on the device (kernels K7) -----------------------------------------------------------
 * n single-rigid-body robots (the plant of potato_sim.cpp:92-140: stage-0 GRFs applied at the foot positions + gravity) stepped for
 * `periods` control periods (length control_dt, default the MPC stage length) with the controller in the loop the way mit_controller_node.cpp
 * drives one robot: UpdateState -> GetGaitSequence (stateful sequencer, K5b) -> GetWrenchSequence (QP solve) -> and, when a WBC
 * handle is given, `wbc_per_mpc` times per period: stage-1 prediction / stage-0 GRFs / stage-0 contacts -> WBC target
 * (mit_controller_node.cpp:1000-1006) -> scene (K6) -> WBC QP (K3) -> the WBC's contact forces drive the plant for control_dt / wbc_per_mpc.
 * Nothing is copied to or from the host inside the loop.  b200qp_mpc_set_model and b200qp_mpc_sequencer_configure must have been
 * called (and b200qp_wbc_scene_configure on the WBC handle, same device, max_batch >= n).
 * seq2_inout [n][B200QP_SEQ2_IN_DOUBLES]: initial state, feet, clock, Target and gait of every robot; on return the final ones
 * (column 55 is used by the loop: contact mask of the last period + 16).
 * history (may be NULL) [periods][n][B200QP_ROLLOUT_HIST_DOUBLES]: pos 0:3, vel 3:6, quat 6:10, sum of applied f_z 10 (MPC stage-0
 * forces), MPC status 11, iterations 12, polish rounds 13, status of the period's last WBC solve 14 (-1 without WBC). */
#define B200QP_ROLLOUT_HIST_DOUBLES 16
typedef struct b200qp_rollout_options {
  int32_t substeps;       /* plant integration steps per control period (10)                                            */
  int32_t plant_inertia;  /* 0: Iw = R Ib R' (physical), 1: the reference MPC model's own world inertia (mpc.cpp:60-64)  */
  int32_t wbc_per_mpc;    /* WBC solves per MPC period when a WBC handle is given (5 = 500 Hz under a 100 Hz MPC)        */
  int32_t reserved;
  double com_pose_Kp[6], com_pose_Kd[6];    /* wbc.com_pose_Kp / Kd (linear xyz, angular xyz)  mit_controller_sim_go2.yaml:66-67 */
  double feet_pose_Kp[3], feet_pose_Kd[3];  /* wbc.feet_pose_Kp / Kd                                                    */
  double control_dt;      /* length of a control period = time between MPC solves (MPC_CONTROL_DT = 0.01 in the reference,
                             mit_controller_params.hpp:8); 0: the MPC stage length dt                                   */
} b200qp_rollout_options;
int32_t b200qp_mpc_rollout(b200qp_mpc *h, b200qp_wbc *wbc, int32_t n, int32_t periods, double *seq2_inout,
                           const b200qp_rollout_options *opt, double *history);

/* library / device info */
const char *b200qp_version(void);
int32_t b200qp_device_count(void);
const char *b200qp_last_error(void);
/* Measured dense FP64 FMA throughput of the device (TFLOP/s, DFMA microbenchmark): the roofline denominator of the
 * on-chip solve kernels (MEASURED_PEAKS.json carries HBM and bf16 only). */
int32_t b200qp_fp64_peak_tflops(int32_t device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* B200QP_H_ */
