// C ABI of libb200qp.so (include/b200qp.h): handle management, staging buffers, kernel launches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200qp {
cudaError_t launch_gait_schedule(int n, int steps, double dt, const double *period, const double *duty,
                                 const double *offset, const double *phase, const uint8_t *measured, double fmin,
                                 double fmax, uint8_t *contact, double *swing_time, double *lbu, double *ubu,
                                 cudaStream_t st);
cudaError_t launch_mpc_bin(int n, int N, const uint8_t *contact, int *bin_count, int *bin_lists,
                           b200qp_solver_info *info, int overflow_bin, cudaStream_t st);
cudaError_t launch_mpc_bin_list(const int *src, const int *src_count, int n, int N, const uint8_t *contact, int *bin_count, int *bin_lists,
                                cudaStream_t st);
cudaError_t launch_mpc_admm(int bin, const MpcParams &P, const MpcBuffers &B, int grid, cudaStream_t st);
int admm_block_size(int bin);
cudaError_t fp64_peak(int num_sms, double *tflops);
int admm_occupancy(int bin);
cudaError_t launch_mpc_riccati(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces,
                               double *xpred, b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch,
                               int *next, int grid, cudaStream_t st, long long *dbg_cycles, const int *list,
                               const int *list_count);
int riccati_grid(int num_sms, int horizon);
cudaError_t launch_mpc_pcond(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                             b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, int *next, int grid,
                             cudaStream_t st, const int *list, const int *list_count, int condN, int nub_limit, long long *dbg);
int pcond_grid(int num_sms, int N, int condN, int nub_limit);
cudaError_t launch_mpc_pcw(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                           b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, int *next, int grid, cudaStream_t st,
                           const int *list, const int *list_count, int *rej, int *rej_count, int condN, int fs, long long *dbg);
int pcw_grid(int num_sms, int N, int fs, int condN);
cudaError_t launch_smem_polluter(int num_sms, cudaStream_t st);
size_t pcw_scratch_doubles(int N, int fs);
void pcw_hot_region(int N, int fs, int grid, double *scratch, double **ptr, size_t *bytes);
size_t pcond_scratch_doubles(int N);
cudaError_t launch_sequencer(int n, const SeqModel &M, const double *seq_in, const uint8_t *measured, double *rec_out,
                             uint8_t *contact_out, int in_stride, cudaStream_t st);
cudaError_t launch_sequencer_full(int n, const SeqModel &M, const SeqOptions &O, const double *seq_in, const uint8_t *measured,
                                  double *state, double *rec_out, uint8_t *contact_out, int in_stride, cudaStream_t st);
struct WbcParams {
  int nq, neq, nin;
  double kkt_tol;
  int max_iter;
  double tau, sig0;  // first attempt: fraction of the step to the boundary that is taken, end-game centring floor
};
struct WbcSceneConfig {
  double mu, com_pose_weight[6], foot_force_weight[3], foot_pose_weight[3], hessian_regularizer, tau_max[12];
};
cudaError_t launch_wbc_scene(int n, const WbcSceneConfig &C, const double *scene_in, double *H, double *g, double *A, double *lo, double *hi,
                             double *feet, cudaStream_t st);
cudaError_t launch_wbc_torque(int n, const double *A, const double *lo, const double *hi, const double *x, double *tau, cudaStream_t st);
struct RolloutOptions {
  int substeps, plant_inertia, wbc_per_mpc, reserved;
  double com_pose_Kp[6], com_pose_Kd[6], feet_pose_Kp[3], feet_pose_Kd[3];
};
cudaError_t launch_rollout_feet(int n, int N, int in_stride, const double *rec, const uint8_t *contact, const double *forces,
                                const b200qp_solver_info *info, double *seq2, uint8_t *measured, double *f0, cudaStream_t st);
cudaError_t launch_rollout_plant(int n, const SeqModel &M, int plant_inertia, double h_total, int substeps, const double *f_app, int f_stride,
                                 int f_off, const b200qp_solver_info *f_info, const double *f_fallback, double *seq2, cudaStream_t st);
cudaError_t launch_rollout_finish(int n, double dt_ns, const double *f0, const b200qp_solver_info *info, const b200qp_solver_info *winfo,
                                  double *seq2, double *hist, cudaStream_t st);
cudaError_t launch_rollout_wbc_target(int n, int N, const RolloutOptions &O, const double *seq2, const double *xpred, const double *f0,
                                      double *scene, cudaStream_t st);
size_t wbc_smem_bytes(const WbcParams &P);
int wbc_grid(const WbcParams &P, int num_sms);
cudaError_t launch_wbc_qp(const WbcParams &P, int n, const double *H, const double *g, const double *A, const double *lo,
                          const double *hi, double *x, b200qp_solver_info *info, int *next, int grid, cudaStream_t st);
}  // namespace b200qp

using namespace b200qp;

static thread_local std::string g_err;
static void set_err(const char *where, cudaError_t e) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
}
#define CK(call)                     \
  do {                               \
    cudaError_t e__ = (call);        \
    if (e__ != cudaSuccess) {        \
      set_err(#call, e__);           \
      return B200QP_ERR_CUDA;        \
    }                                \
  } while (0)

struct b200qp_mpc {
  b200qp_mpc_config cfg;
  MpcParams P;
  cudaStream_t stream = nullptr;
  cudaStream_t bin_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // bins 1..3 and the Riccati overflow run beside bin 0
  cudaStream_t copy_stream = nullptr;  // chunked H2D of the records, overlapped with the solve
  cudaEvent_t ev_copy = nullptr;
  int *d_ready = nullptr, *h_one = nullptr;  // per-chunk "records have landed" flags, pinned constant 1
  int stream_chunk = 0;                      // problems per chunk while a streamed solve is being issued, else 0
  // warm / hot start: previous solution per problem slot
  double *d_warm_x = nullptr, *d_warm_y = nullptr;
  double ipm_warm_mu = 0.1;  // B200QP_IPM_WARM_MU overrides (experiments)
  uint8_t *d_warm_valid = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0;
  int async_engines = 0;  // copy engines that can run beside a kernel (0: streamed inputs are not used)
  // device buffers sized for max_batch
  double *d_in = nullptr, *d_forces = nullptr, *d_xpred = nullptr, *d_lam_box = nullptr, *d_lam_g = nullptr;
  uint8_t *d_contact = nullptr;
  b200qp_solver_info *d_info = nullptr;
  int *d_bin_count = nullptr;  // [5] counts (4 condensed bins + overflow), [5..8] condensed work counters, [9] Riccati counter
  int *d_bin_lists = nullptr;  // [5][max_batch]
  int *d_retry = nullptr;      // [max_batch] solver AUTO: problems handed from the condensed to the Riccati kernel
  double *d_hscratch[4] = {nullptr, nullptr, nullptr, nullptr};
  int grid[4] = {0, 0, 0, 0};
  double *d_ric_scratch = nullptr;
  int ric_grid = 0;
  double *d_pc_scratch = nullptr;  // partial-condensing kernel: block constants (+ factors that do not fit shared memory)
  int pc_grid = 0, pc_cond = 0, pc_nub = 0;
  // warp-per-problem Riccati kernel (solver SPARSE_RICCATI): one launch for problems with <= 32 stance foot-stages, one for <= 64,
  // the block kernel for the rest
  double *d_pcw_scratch[2] = {nullptr, nullptr};
  int pcw_grid[2] = {0, 0}, pcw_cond = 0;
  int *d_pcw_rej = nullptr;  // [2][max_batch] hand-over lists
  bool poison_smem = false;  // B200QP_POISON_SMEM=1 (development)
  bool ipm_refine = true;    // SPARSE_RICCATI: second opinion by the active-set kernel for stall-accepted problems (B200QP_IPM_REFINE=0: off)
  int *d_refine = nullptr;   // [max_batch] list + [16] counters (list length, 4 bin counts, 4 work counters) at the end
  double *d_refine_x = nullptr, *d_refine_y = nullptr;  // hot start of the second opinion: forces / duals of the interior-point answer
  uint8_t *d_refine_valid = nullptr;
  size_t l2_window_max = 0;
  size_t l2_persist = 0;     // bytes of L2 set aside for persisting accesses (hot scratch of the warp Riccati kernel), 0: not available
  int in_chunks = 4;         // record chunks of the streamed H2D copy (B200QP_IN_CHUNKS, 1..16)
  bool ric_legacy = false;   // B200QP_RICCATI_LEGACY=1: block kernel only (A/B runs)
  // pinned staging
  double *h_in = nullptr, *h_forces = nullptr, *h_xpred = nullptr;
  uint8_t *h_contact = nullptr;
  b200qp_solver_info *h_info = nullptr;
  // last call
  const double *last_in = nullptr;
  const uint8_t *last_contact = nullptr;
  int last_n = 0;
  int last_launches = 0;
  bool want_duals = false;
  bool have_timing = false;
  long long *d_dbg_cycles = nullptr;  // optional per-phase cycle accounting (B200QP_PHASE_CYCLES=1)
  // on-device gait sequencer (b200qp_mpc_sequence_*)
  SeqModel model;
  bool have_model = false;
  bool no_zero_copy = false;  // B200QP_NO_ZERO_COPY=1: always stage outputs through device buffers + DMA
  double *d_seq = nullptr, *h_seq = nullptr;
  uint8_t *d_measured = nullptr;
  // stateful sequencers (b200qp_mpc_sequencer_*)
  SeqOptions seq_opt;
  bool have_seq_opt = false;
  double *d_seq_state = nullptr, *d_seq2 = nullptr, *h_seq2 = nullptr;
};

static bool is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

static void fill_params(b200qp_mpc *h) {
  const b200qp_mpc_config &c = h->cfg;
  MpcParams &P = h->P;
  P.N = c.horizon;
  P.in_stride = b200qp_mpc_in_doubles(c.horizon);
  P.dt = c.dt;
  P.alpha = c.alpha;
  for (int i = 0; i < 12; ++i) P.w[i] = c.state_weights[i];
  P.mu = c.mu;
  P.fmin = c.fmin;
  P.fmax = c.fmax;
  // the best initial penalty scales with the input weight (measured: alpha 1e-6 / 1e-5 / 1e-4 / 1e-3 -> rho0 5e-7 / 5e-5 / 5e-4 / 5e-4)
  P.rho0 = c.admm_rho > 0 ? c.admm_rho : fmin(fmax(5.0 * c.alpha, 1e-6), 5e-4);
  P.sigma = c.admm_sigma > 0 ? c.admm_sigma : 1e-6;
  P.relax = c.admm_alpha > 0 ? c.admm_alpha : 1.6;
  P.eps = c.admm_eps > 0 ? c.admm_eps : 1e-3;
  P.kkt_tol = c.kkt_tol > 0 ? c.kkt_tol : 1e-7;
  P.max_iter = c.admm_max_iter > 0 ? c.admm_max_iter : 2000;
  P.check_every = c.admm_check_every > 0 ? c.admm_check_every : 15;
  P.polish_rounds = c.polish_max_rounds > 0 ? c.polish_max_rounds : 8;
  P.ipm_max_iter = c.ipm_max_iter > 0 ? c.ipm_max_iter : 40;
  P.precision = c.precision;
  P.ipm_warm_x = nullptr;
  P.ipm_warm_valid = nullptr;
  P.ipm_warm_mu = 1.0;
  P.ipm_tau = (getenv("B200QP_IPM_TAU") && atof(getenv("B200QP_IPM_TAU")) > 0) ? atof(getenv("B200QP_IPM_TAU")) : 0.9999;
  P.ipm_cold_mu = (getenv("B200QP_IPM_COLD_MU") && atof(getenv("B200QP_IPM_COLD_MU")) > 0) ? atof(getenv("B200QP_IPM_COLD_MU")) : 0.2;
  P.ipm_refine_list = nullptr;
  P.ipm_refine_count = nullptr;
  P.ipm_refine_thr = 1e-10;
  P.ipm_refine_x = nullptr;
  P.ipm_refine_y = nullptr;
  P.ipm_refine_valid = nullptr;
  P.ipm_flags = getenv("B200QP_IPM_FLAGS") ? atoi(getenv("B200QP_IPM_FLAGS")) : 3;  // separate primal / dual step lengths + two-cycle guard
  if (c.precision == 2 && !(c.kkt_tol > 0)) P.kkt_tol = 5e-2;  // FP32 polish: stationarity of O(1e-3) is all there is
}

extern "C" {

static int32_t ensure_warm_buffers(b200qp_mpc *h);
void b200qp_mpc_destroy(b200qp_mpc *h);
void b200qp_wbc_destroy(b200qp_wbc *h);
}
// a create() that fails half-way must not leak what it already allocated
namespace {
struct MpcGuard {
  b200qp_mpc *h;
  ~MpcGuard() { if (h) b200qp_mpc_destroy(h); }
};
struct WbcGuard {
  b200qp_wbc *h;
  ~WbcGuard() { if (h) b200qp_wbc_destroy(h); }
};
// device temporaries of the diagnostic entry points: released on every exit path
struct DevTmp {
  void *p[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int n = 0;
  template <class T>
  cudaError_t alloc(T **out, size_t bytes) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(out), bytes);
    if (e == cudaSuccess && n < 6) p[n++] = *out;
    return e;
  }
  ~DevTmp() {
    for (int i = 0; i < n; ++i) cudaFree(p[i]);
  }
};
}  // namespace
extern "C" {

const char *b200qp_version(void) { return "b200qp 0.1 (sm_100a)"; }
const char *b200qp_last_error(void) { return g_err.c_str(); }
int32_t b200qp_fp64_peak_tflops(int32_t device, double *tflops) {
  if (!tflops) return B200QP_ERR_ARG;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return B200QP_ERR_CUDA;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  CK(fp64_peak(prop.multiProcessorCount, tflops));
  return B200QP_OK;
}
int32_t b200qp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int32_t b200qp_mpc_create(const b200qp_mpc_config *cfg, b200qp_mpc **out) {
  if (!cfg || !out) return B200QP_ERR_ARG;
  *out = nullptr;
  if (cfg->horizon < 1 || cfg->horizon > B200QP_MAX_HORIZON || cfg->max_batch < 1 || !(cfg->dt > 0) ||
      !(cfg->alpha > 0) || !(cfg->mu >= 0) || !(cfg->fmax >= cfg->fmin)) {
    g_err = "b200qp_mpc_create: invalid configuration";
    return B200QP_ERR_ARG;
  }
  if (cfg->precision < 0 || cfg->precision > 2 || (cfg->precision != 0 && cfg->solver != B200QP_SOLVER_CONDENSED_ADMM)) {
    g_err = "b200qp_mpc_create: precision must be 0 (FP64); 1 / 2 (FP32 error study) need solver CONDENSED_ADMM";
    return B200QP_ERR_ARG;
  }
  if (cfg->solver != B200QP_SOLVER_CONDENSED_ADMM && cfg->solver != B200QP_SOLVER_SPARSE_RICCATI &&
      cfg->solver != B200QP_SOLVER_AUTO && cfg->solver != B200QP_SOLVER_PARTIAL_CONDENSING) {
    g_err = "b200qp_mpc_create: unknown solver";
    return B200QP_ERR_ARG;
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_err = "b200qp_mpc_create: no such CUDA device (there is no CPU fallback)";
    return B200QP_ERR_CUDA;
  }
  CK(cudaSetDevice(cfg->device));
  b200qp_mpc *h = new (std::nothrow) b200qp_mpc();
  if (!h) return B200QP_ERR_ARG;
  MpcGuard guard{h};
  h->cfg = *cfg;
  fill_params(h);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  h->num_sms = prop.multiProcessorCount;
  h->async_engines = prop.asyncEngineCount;
  const size_t nb = (size_t)cfg->max_batch;
  const int N = cfg->horizon;
  const size_t in_d = (size_t)b200qp_mpc_in_doubles(N);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&h->d_ready, 16 * sizeof(int)));
  CK(cudaMallocHost(&h->h_one, sizeof(int)));
  *h->h_one = 1;
  for (int b = 0; b < 4; ++b) {
    CK(cudaStreamCreateWithFlags(&h->bin_stream[b], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_join[b], cudaEventDisableTiming));
  }
  CK(cudaMalloc(&h->d_in, nb * in_d * sizeof(double)));
  CK(cudaMalloc(&h->d_contact, nb * N * 4));
  CK(cudaMalloc(&h->d_forces, nb * N * 12 * sizeof(double)));
  CK(cudaMalloc(&h->d_xpred, nb * (N + 1) * 13 * sizeof(double)));
  CK(cudaMalloc(&h->d_lam_box, nb * N * 12 * sizeof(double)));
  CK(cudaMalloc(&h->d_lam_g, nb * N * 16 * sizeof(double)));
  CK(cudaMalloc(&h->d_info, nb * sizeof(b200qp_solver_info)));
  CK(cudaMalloc(&h->d_bin_count, 16 * sizeof(int)));
  CK(cudaMalloc(&h->d_bin_lists, 5 * nb * sizeof(int)));
  CK(cudaMalloc(&h->d_retry, nb * sizeof(int)));
  CK(cudaMemset(h->d_forces, 0, nb * N * 12 * sizeof(double)));
  CK(cudaMemset(h->d_xpred, 0, nb * (N + 1) * 13 * sizeof(double)));
  if (cfg->solver == B200QP_SOLVER_PARTIAL_CONDENSING) {
    if (cfg->cond_N < 0 || cfg->cond_N > cfg->horizon) {
      g_err = "b200qp_mpc_create: cond_N must be 0 (default) or 1..horizon";
      return B200QP_ERR_ARG;
    }
    h->pc_cond = cfg->cond_N > 0 ? cfg->cond_N : (cfg->horizon + 3) / 4;  // default: blocks of 4 stages
    if (const char *e = getenv("B200QP_PCOND_NUB")) h->pc_nub = atoi(e);  // development: cap on the stacked inputs of a block
    int g = pcond_grid(h->num_sms, cfg->horizon, h->pc_cond, h->pc_nub);
    if (const char *e = getenv("B200QP_BLOCKS_PER_SM")) g = h->num_sms * (atoi(e) > 0 ? atoi(e) : 1);
    if ((size_t)g > nb) g = (int)nb;
    h->pc_grid = g;
    CK(cudaMalloc(&h->d_pc_scratch, (size_t)g * pcond_scratch_doubles(cfg->horizon) * sizeof(double)));
  }
  if (cfg->solver == B200QP_SOLVER_SPARSE_RICCATI) {
    h->ipm_refine = !(getenv("B200QP_IPM_REFINE") && atoi(getenv("B200QP_IPM_REFINE")) == 0) && getenv("B200QP_RICCATI_LEGACY") == nullptr;
    if (!getenv("B200QP_NO_L2_PERSIST") && prop.persistingL2CacheMaxSize > 0) {
      size_t want = (size_t)32 << 20;
      if (want > (size_t)prop.persistingL2CacheMaxSize) want = (size_t)prop.persistingL2CacheMaxSize;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) h->l2_persist = want;
      else cudaGetLastError();
      h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    }
    if (h->ipm_refine) {
      CK(cudaMalloc(&h->d_refine, (nb + 16) * sizeof(int)));
      if (!(getenv("B200QP_IPM_REFINE_COLD"))) {
        CK(cudaMalloc(&h->d_refine_x, nb * N * 12 * sizeof(double)));
        CK(cudaMalloc(&h->d_refine_y, nb * N * 28 * sizeof(double)));
        CK(cudaMalloc(&h->d_refine_valid, nb));
        CK(cudaMemset(h->d_refine_valid, 0, nb));
      }
    }
  }
  if (cfg->solver == B200QP_SOLVER_CONDENSED_ADMM || cfg->solver == B200QP_SOLVER_AUTO ||
      (cfg->solver == B200QP_SOLVER_SPARSE_RICCATI && h->ipm_refine)) {
    for (int b = 0; b < 4; ++b) {
      const int bs = admm_block_size(b);
      int g = h->num_sms * admm_occupancy(b);
      if (const char *e = getenv("B200QP_BLOCKS_PER_SM")) g = h->num_sms * (atoi(e) > 0 ? atoi(e) : 1);  // development
      if ((size_t)g > nb) g = (int)nb;
      h->grid[b] = g;
      CK(cudaMalloc(&h->d_hscratch[b], (size_t)g * ((size_t)bs * bs + B200QP_MAX_HORIZON * 36) * sizeof(double)));
    }
  }
  if (cfg->solver == B200QP_SOLVER_SPARSE_RICCATI) {
    h->ric_legacy = getenv("B200QP_RICCATI_LEGACY") != nullptr;
    h->poison_smem = getenv("B200QP_POISON_SMEM") != nullptr;
    if (const char *ic = getenv("B200QP_IN_CHUNKS")) {
      const int v = atoi(ic);
      if (v >= 1 && v <= 16) h->in_chunks = v;
    }
    if (const char *wm = getenv("B200QP_IPM_WARM_MU")) h->ipm_warm_mu = atof(wm) > 0 ? atof(wm) : h->ipm_warm_mu;
    // blocks of the warp-per-problem kernel: cond_N as given (cond_N = horizon: one stage per block, the sparse formulation proper);
    // default: two stages per block - same Newton direction, fewer and better filled block factorisations
    h->pcw_cond = (cfg->cond_N > 0 && cfg->cond_N <= cfg->horizon) ? cfg->cond_N : (cfg->horizon + 1) / 2;
    for (int v = 0; v < 1 && !h->ric_legacy; ++v) {  // the two-foot-stages-per-lane instantiation is slower than the block kernel: not used
      int g = pcw_grid(h->num_sms, cfg->horizon, v + 1, h->pcw_cond);
      if (const char *e = getenv("B200QP_BLOCKS_PER_SM")) g = h->num_sms * (atoi(e) > 0 ? atoi(e) : 1);
      if ((size_t)g > nb) g = (int)nb;
      h->pcw_grid[v] = g;
      CK(cudaMalloc(&h->d_pcw_scratch[v], (size_t)g * pcw_scratch_doubles(cfg->horizon, v + 1) * sizeof(double)));
      if (getenv("B200QP_POISON_SCRATCH"))  // development: NaN-fill the whole scratch (read-before-write hunts)
        CK(cudaMemset(h->d_pcw_scratch[v], 0xFF, (size_t)g * pcw_scratch_doubles(cfg->horizon, v + 1) * sizeof(double)));
    }
    CK(cudaMalloc(&h->d_pcw_rej, 2 * nb * sizeof(int)));
  }
  if (cfg->solver == B200QP_SOLVER_SPARSE_RICCATI || cfg->solver == B200QP_SOLVER_AUTO) {
    int g = riccati_grid(h->num_sms, cfg->horizon);
    if ((size_t)g > nb) g = (int)nb;
    h->ric_grid = g;
    CK(cudaMalloc(&h->d_ric_scratch, (size_t)g * B200QP_MAX_HORIZON * 288 * sizeof(double)));
  }
  h->no_zero_copy = getenv("B200QP_NO_ZERO_COPY") != nullptr;
  if (getenv("B200QP_PHASE_CYCLES")) {
    CK(cudaMalloc(&h->d_dbg_cycles, nb * 10 * sizeof(long long)));
    CK(cudaMemset(h->d_dbg_cycles, 0, nb * 10 * sizeof(long long)));
  }
  CK(cudaMallocHost(&h->h_in, nb * in_d * sizeof(double)));
  CK(cudaMallocHost(&h->h_contact, nb * N * 4));
  CK(cudaMallocHost(&h->h_forces, nb * N * 12 * sizeof(double)));
  CK(cudaMallocHost(&h->h_xpred, nb * (N + 1) * 13 * sizeof(double)));
  CK(cudaMallocHost(&h->h_info, nb * sizeof(b200qp_solver_info)));
  if (cfg->warm_start < 0 || cfg->warm_start > 2) {
    g_err = "b200qp_mpc_create: warm_start must be 0, 1 or 2";
    return B200QP_ERR_ARG;
  }
  if (cfg->warm_start > 0) {
    const int32_t wrc = ensure_warm_buffers(h);
    if (wrc != B200QP_OK) return wrc;
  }
  CK(cudaMalloc(&h->d_seq, nb * B200QP_SEQ_IN_DOUBLES * sizeof(double)));
  CK(cudaMalloc(&h->d_measured, nb * 4));
  CK(cudaMallocHost(&h->h_seq, nb * B200QP_SEQ_IN_DOUBLES * sizeof(double) + nb * 4));
  guard.h = nullptr;
  *out = h;
  return B200QP_OK;
}

void b200qp_mpc_destroy(b200qp_mpc *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->d_in);
  cudaFree(h->d_contact);
  cudaFree(h->d_forces);
  cudaFree(h->d_xpred);
  cudaFree(h->d_lam_box);
  cudaFree(h->d_lam_g);
  cudaFree(h->d_info);
  cudaFree(h->d_bin_count);
  cudaFree(h->d_bin_lists);
  cudaFree(h->d_retry);
  for (int b = 0; b < 4; ++b) cudaFree(h->d_hscratch[b]);
  cudaFree(h->d_ric_scratch);
  cudaFree(h->d_pc_scratch);
  cudaFree(h->d_pcw_scratch[0]);
  cudaFree(h->d_pcw_scratch[1]);
  cudaFree(h->d_refine);
  cudaFree(h->d_refine_x);
  cudaFree(h->d_refine_y);
  cudaFree(h->d_refine_valid);
  cudaFree(h->d_pcw_rej);
  cudaFreeHost(h->h_in);
  cudaFreeHost(h->h_contact);
  cudaFreeHost(h->h_forces);
  cudaFreeHost(h->h_xpred);
  cudaFreeHost(h->h_info);
  cudaFree(h->d_seq);
  cudaFree(h->d_seq_state);
  cudaFree(h->d_seq2);
  cudaFreeHost(h->h_seq2);
  cudaFree(h->d_measured);
  cudaFree(h->d_warm_x);
  cudaFree(h->d_warm_y);
  cudaFree(h->d_warm_valid);
  cudaFreeHost(h->h_seq);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_copy) cudaEventDestroy(h->ev_copy);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  cudaFree(h->d_ready);
  cudaFreeHost(h->h_one);
  for (int b = 0; b < 4; ++b) {
    if (h->ev_join[b]) cudaEventDestroy(h->ev_join[b]);
    if (h->bin_stream[b]) cudaStreamDestroy(h->bin_stream[b]);
  }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static int32_t ensure_warm_buffers(b200qp_mpc *h) {
  const size_t nb = (size_t)h->cfg.max_batch, N = (size_t)h->cfg.horizon;
  if (h->d_warm_x == nullptr) {
    CK(cudaMalloc(&h->d_warm_x, nb * N * 12 * sizeof(double)));
    CK(cudaMalloc(&h->d_warm_y, nb * N * 28 * sizeof(double)));
    CK(cudaMalloc(&h->d_warm_valid, nb));
  }
  CK(cudaMemsetAsync(h->d_warm_valid, 0, nb, h->stream));  // forget every stored solution
  return B200QP_OK;
}

int32_t b200qp_mpc_set_warm_start(b200qp_mpc *h, int32_t mode) {
  if (!h || mode < 0 || mode > 2) return B200QP_ERR_ARG;
  CK(cudaSetDevice(h->cfg.device));
  h->cfg.warm_start = mode;
  if (mode > 0 || h->d_warm_x != nullptr) return ensure_warm_buffers(h);
  return B200QP_OK;
}

int32_t b200qp_mpc_set_state_weights(b200qp_mpc *h, const double weights[12]) {
  if (!h || !weights) return B200QP_ERR_ARG;
  for (int i = 0; i < 12; ++i) h->cfg.state_weights[i] = weights[i];
  fill_params(h);
  return B200QP_OK;
}
int32_t b200qp_mpc_set_input_weights(b200qp_mpc *h, double alpha) {
  if (!h || !(alpha > 0)) return B200QP_ERR_ARG;
  h->cfg.alpha = alpha;
  fill_params(h);
  return B200QP_OK;
}
int32_t b200qp_mpc_set_fmax(b200qp_mpc *h, double fmax) {
  if (!h || !(fmax >= h->cfg.fmin)) return B200QP_ERR_ARG;
  h->cfg.fmax = fmax;
  fill_params(h);
  return B200QP_OK;
}
int32_t b200qp_mpc_set_mu(b200qp_mpc *h, double mu) {
  if (!h || !(mu >= 0)) return B200QP_ERR_ARG;
  h->cfg.mu = mu;
  fill_params(h);
  return B200QP_OK;
}

static int32_t solve_device_impl(b200qp_mpc *h, int32_t n, const double *d_in, const uint8_t *d_contact,
                                 double *d_forces, double *d_xpred, b200qp_solver_info *d_info, int dbg_index,
                                 double *dbg_H, double *dbg_g) {
  const int N = h->cfg.horizon;
  int launches = 0;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaMemsetAsync(h->d_bin_count, 0, 16 * sizeof(int), h->stream));
  const bool use_admm = h->cfg.solver == B200QP_SOLVER_CONDENSED_ADMM || h->cfg.solver == B200QP_SOLVER_AUTO;
  const bool autosel = h->cfg.solver == B200QP_SOLVER_AUTO;
  if (h->cfg.solver == B200QP_SOLVER_PARTIAL_CONDENSING) {  // block-condensed Riccati IPM over the whole batch
    CK(launch_mpc_pcond(h->P, n, d_in, d_contact, d_forces, d_xpred, d_info, h->want_duals ? h->d_lam_box : nullptr,
                        h->want_duals ? h->d_lam_g : nullptr, h->d_pc_scratch, h->d_bin_count + 9, h->pc_grid, h->stream, nullptr,
                        nullptr, h->pc_cond, h->pc_nub, h->d_dbg_cycles));
    ++launches;
  }
  // The size bins (and the Riccati overflow list) are independent launches: bin 0 stays on the handle's stream, the others
  // fork onto their own streams after the binning kernel and join before the end event, so their tails overlap instead of
  // adding up (their persistent grids are sized for a whole GPU each; the hardware interleaves the resident blocks).
  bool forked[4] = {false, false, false, false};
  if (use_admm) {
    CK(launch_mpc_bin(n, N, d_contact, h->d_bin_count, h->d_bin_lists, d_info, autosel ? 1 : 0, h->stream));
    ++launches;
    CK(cudaEventRecord(h->ev_fork, h->stream));
    for (int b = 0; b < 4; ++b) {
      MpcBuffers B;
      B.in = d_in;
      B.contact = d_contact;
      B.forces = d_forces;
      B.xpred = d_xpred;
      B.info = d_info;
      B.lam_box = h->want_duals ? h->d_lam_box : nullptr;
      B.lam_g = h->want_duals ? h->d_lam_g : nullptr;
      B.bin_list = h->d_bin_lists + (size_t)b * n;
      B.bin_count = h->d_bin_count + b;
      B.bin_next = h->d_bin_count + 5 + b;
      B.hscratch = h->d_hscratch[b];
      B.retry_list = autosel ? h->d_retry : nullptr;
      B.retry_count = autosel ? h->d_bin_count + 10 : nullptr;
      B.refine = 0;
      B.warm_mode = (h->cfg.warm_start > 0 && h->d_warm_x != nullptr && dbg_index < 0 && !h->want_duals) ? h->cfg.warm_start : 0;
      B.warm_x = h->d_warm_x;
      B.warm_y = h->d_warm_y;
      B.warm_valid = h->d_warm_valid;
      B.ready = h->stream_chunk > 0 ? h->d_ready : nullptr;
      B.ready_chunk = h->stream_chunk;
      B.dbg_index = dbg_index;
      B.dbg_H = dbg_H;
      B.dbg_g = dbg_g;
      B.dbg_cycles = h->d_dbg_cycles;
      int g = h->grid[b];
      if (g > n) g = n;
      cudaStream_t st = h->stream;
      if (b > 0) {
        st = h->bin_stream[b - 1];
        CK(cudaStreamWaitEvent(st, h->ev_fork, 0));
        forked[b - 1] = true;
      }
      MpcParams Pb = h->P;
      // AUTO: a straggler that needs more than a few hundred ADMM iterations (tight force limits with a tiny input weight:
      // up to 1800) is cheaper on the interior-point kernel, and it would otherwise stretch the whole launch
      if (autosel && Pb.max_iter > 300) Pb.max_iter = 300;
      CK(launch_mpc_admm(b, Pb, B, g, st));
      ++launches;
    }
  }
  if (h->cfg.solver == B200QP_SOLVER_SPARSE_RICCATI && !h->ric_legacy) {
    // warp-per-problem kernel: <= 32 stance foot-stages first, what does not fit is handed on to the two-per-lane instantiation and
    // from there to the block kernel (d_bin_count[12], [13] = lengths of the hand-over lists, [14], [15], [9] = work counters)
    double *lb = h->want_duals ? h->d_lam_box : nullptr, *lg = h->want_duals ? h->d_lam_g : nullptr;
    int *rej0 = h->d_pcw_rej;
    if (h->poison_smem) CK(launch_smem_polluter(148, h->stream));
    MpcParams Pw = h->P;
    if (h->cfg.warm_start > 0 && h->d_warm_x != nullptr && dbg_index < 0) {  // IPM warm start: previous forces of the slot
      Pw.ipm_warm_x = h->d_warm_x;
      Pw.ipm_warm_valid = h->d_warm_valid;
      Pw.ipm_warm_mu = h->ipm_warm_mu;
    }
    const bool refine = h->ipm_refine && h->d_refine != nullptr && dbg_index < 0;
    int *ref_cnt = h->d_refine + h->cfg.max_batch;  // [0] list length, [1..4] bin counts, [5..8] work counters
    if (refine) {
      CK(cudaMemsetAsync(ref_cnt, 0, 16 * sizeof(int), h->stream));
      Pw.ipm_refine_list = h->d_refine;
      Pw.ipm_refine_count = ref_cnt;
      Pw.ipm_refine_x = h->d_refine_x;
      Pw.ipm_refine_y = h->d_refine_y;
      Pw.ipm_refine_valid = h->d_refine_valid;
    }
    bool window = false;
    if (h->l2_persist > 0) {  // keep the per-iteration scratch of the resident warps in L2 for the duration of the launch
      double *hp = nullptr;
      size_t hb = 0;
      pcw_hot_region(N, 1, h->pcw_grid[0] < n ? h->pcw_grid[0] : n, h->d_pcw_scratch[0], &hp, &hb);
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.base_ptr = hp;
      av.accessPolicyWindow.num_bytes = hb < h->l2_window_max ? hb : h->l2_window_max;
      av.accessPolicyWindow.hitRatio = hb <= h->l2_persist ? 1.0f : (float)((double)h->l2_persist / (double)hb);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      window = cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
      if (!window) cudaGetLastError();
    }
    CK(launch_mpc_pcw(Pw, n, d_in, d_contact, d_forces, d_xpred, d_info, lb, lg, h->d_pcw_scratch[0], h->d_bin_count + 14, h->pcw_grid[0],
                      h->stream, nullptr, nullptr, rej0, h->d_bin_count + 12, h->pcw_cond, 1, h->d_dbg_cycles));
    if (window) {  // later launches on this stream run without the window
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.num_bytes = 0;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);
    }
    CK(launch_mpc_riccati(Pw, n, d_in, d_contact, d_forces, d_xpred, d_info, lb, lg, h->d_ric_scratch, h->d_bin_count + 9, h->ric_grid,
                          h->stream, nullptr, rej0, h->d_bin_count + 12));
    launches += 2;
    if (refine) {
      // second opinion: what the interior-point kernel only solved by its stall rule goes through the condensed ADMM + polish kernel
      // (where it fits, n <= 156); that kernel overwrites forces / prediction / record only when it succeeds
      CK(launch_mpc_bin_list(h->d_refine, ref_cnt, n, N, d_contact, ref_cnt + 1, h->d_bin_lists, h->stream));
      for (int b = 0; b < 4; ++b) {
        MpcBuffers B;
        B.in = d_in;
        B.contact = d_contact;
        B.forces = d_forces;
        B.xpred = d_xpred;
        B.info = d_info;
        B.lam_box = lb;
        B.lam_g = lg;
        B.bin_list = h->d_bin_lists + (size_t)b * n;
        B.bin_count = ref_cnt + 1 + b;
        B.bin_next = ref_cnt + 5 + b;
        B.hscratch = h->d_hscratch[b];
        B.retry_list = nullptr;
        B.retry_count = nullptr;
        B.refine = 1;
        B.warm_mode = h->d_refine_x != nullptr ? 2 : 0;  // hot: the interior-point answer's active set goes straight to the polish
        B.warm_x = h->d_refine_x;
        B.warm_y = h->d_refine_y;
        B.warm_valid = h->d_refine_valid;
        B.ready = nullptr;
        B.ready_chunk = 0;
        B.dbg_index = -1;
        B.dbg_H = nullptr;
        B.dbg_g = nullptr;
        B.dbg_cycles = nullptr;
        int g = h->grid[b] < 2 * h->num_sms ? h->grid[b] : 2 * h->num_sms;  // a few hundred problems at most
        if (g > n) g = n;
        CK(launch_mpc_admm(b, h->P, B, g, h->stream));
      }
      launches += 5;
    }
  } else if (h->cfg.solver == B200QP_SOLVER_SPARSE_RICCATI || autosel) {  // whole batch, or (AUTO) the problems too large for the condensed kernel
    cudaStream_t st = h->stream;
    if (autosel) {
      st = h->bin_stream[3];
      CK(cudaStreamWaitEvent(st, h->ev_fork, 0));
      forked[3] = true;
    }
    CK(launch_mpc_riccati(h->P, n, d_in, d_contact, d_forces, d_xpred, d_info,
                          h->want_duals ? h->d_lam_box : nullptr, h->want_duals ? h->d_lam_g : nullptr,
                          h->d_ric_scratch, h->d_bin_count + 9, h->ric_grid, st, h->d_dbg_cycles,
                          autosel ? h->d_bin_lists + (size_t)4 * n : nullptr, autosel ? h->d_bin_count + 4 : nullptr));
    ++launches;
  }
  for (int b = 0; b < 4; ++b)
    if (forked[b]) {
      CK(cudaEventRecord(h->ev_join[b], h->bin_stream[b]));
      CK(cudaStreamWaitEvent(h->stream, h->ev_join[b], 0));
    }
  if (autosel) {  // second opinion: what the condensed kernel gave up on goes through the interior-point kernel (usually nothing)
    int g = h->ric_grid < 64 ? h->ric_grid : 64;
    CK(launch_mpc_riccati(h->P, n, d_in, d_contact, d_forces, d_xpred, d_info,
                          h->want_duals ? h->d_lam_box : nullptr, h->want_duals ? h->d_lam_g : nullptr,
                          h->d_ric_scratch, h->d_bin_count + 11, g, h->stream, nullptr, h->d_retry, h->d_bin_count + 10));
    ++launches;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->have_timing = true;
  h->last_launches = launches;
  h->last_in = d_in;
  h->last_contact = d_contact;
  h->last_n = n;
  return B200QP_OK;
}

int32_t b200qp_mpc_solve_batch_device(b200qp_mpc *h, int32_t n, const double *d_in, const uint8_t *d_contact,
                                      double *d_forces, double *d_xpred, b200qp_solver_info *d_info, int32_t sync) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!d_in || !d_contact || !d_forces || !d_xpred || !d_info))) {
    g_err = "b200qp_mpc_solve_batch_device: bad arguments";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  int32_t rc = solve_device_impl(h, n, d_in, d_contact, d_forces, d_xpred, d_info, -1, nullptr, nullptr);
  if (rc != B200QP_OK) return rc;
  if (sync) CK(cudaStreamSynchronize(h->stream));
  return B200QP_OK;
}

// Device results -> caller buffers through DMA copies (pageable or unmappable caller buffers).
static int32_t fetch_outputs(b200qp_mpc *h, int32_t n, double *forces, double *xpred, b200qp_solver_info *info) {
  const int N = h->cfg.horizon;
  const size_t f_b = (size_t)n * N * 12 * sizeof(double), x_b = (size_t)n * (N + 1) * 13 * sizeof(double);
  // The per-problem status comes back first: failed problems must leave the caller's output slots untouched
  // (mpc.cpp:447-455), so the outputs can only be written in place when every problem succeeded.
  CK(cudaMemcpyAsync(h->h_info, h->d_info, (size_t)n * sizeof(b200qp_solver_info), cudaMemcpyDeviceToHost, h->stream));
  const bool out_pinned = is_pinned(forces) && is_pinned(xpred);
  if (!out_pinned) {
    CK(cudaMemcpyAsync(h->h_forces, h->d_forces, f_b, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_xpred, h->d_xpred, x_b, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  const size_t f1 = (size_t)N * 12, x1 = (size_t)(N + 1) * 13;
  bool all_ok = true;
  for (int p = 0; p < n; ++p) all_ok &= (h->h_info[p].status == B200QP_ACADOS_SUCCESS);
  if (out_pinned && all_ok) {
    CK(cudaMemcpyAsync(forces, h->d_forces, f_b, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(xpred, h->d_xpred, x_b, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  } else {
    if (out_pinned) {
      CK(cudaMemcpyAsync(h->h_forces, h->d_forces, f_b, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpyAsync(h->h_xpred, h->d_xpred, x_b, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
    }
    if (all_ok) {
      memcpy(forces, h->h_forces, f_b);
      memcpy(xpred, h->h_xpred, x_b);
    } else {
      for (int p = 0; p < n; ++p)
        if (h->h_info[p].status == B200QP_ACADOS_SUCCESS) {
          memcpy(forces + p * f1, h->h_forces + p * f1, f1 * sizeof(double));
          memcpy(xpred + p * x1, h->h_xpred + p * x1, x1 * sizeof(double));
        }
    }
  }
  memcpy(info, h->h_info, (size_t)n * sizeof(b200qp_solver_info));
  return B200QP_OK;
}

// Solve the records resident in h->d_in / h->d_contact and deliver the results to HOST buffers.  Page-locked caller
// buffers are mapped into the device address space and written by the kernels themselves as each problem finishes
// (posted PCIe writes overlap the remaining solves; a failed problem's slots are never touched by the kernels, which
// is the reference's failure semantics, mpc.cpp:447-455), so the call ends with one stream synchronisation and no
// device->host copy.  Pageable buffers go through the handle's pinned staging copies.
static int32_t solve_and_fetch(b200qp_mpc *h, int32_t n, double *forces, double *xpred, b200qp_solver_info *info) {
  void *mf = nullptr, *mx = nullptr, *mi = nullptr;
  const bool direct = !h->no_zero_copy && is_pinned(forces) && is_pinned(xpred) &&
                      cudaHostGetDevicePointer(&mf, forces, 0) == cudaSuccess &&
                      cudaHostGetDevicePointer(&mx, xpred, 0) == cudaSuccess &&
                      cudaHostGetDevicePointer(&mi, h->h_info, 0) == cudaSuccess;
  if (!direct) {
    cudaGetLastError();
    int32_t rc = solve_device_impl(h, n, h->d_in, h->d_contact, h->d_forces, h->d_xpred, h->d_info, -1, nullptr, nullptr);
    if (rc != B200QP_OK) return rc;
    return fetch_outputs(h, n, forces, xpred, info);
  }
  int32_t rc = solve_device_impl(h, n, h->d_in, h->d_contact, static_cast<double *>(mf), static_cast<double *>(mx),
                                 static_cast<b200qp_solver_info *>(mi), -1, nullptr, nullptr);
  if (rc != B200QP_OK) return rc;
  CK(cudaStreamSynchronize(h->stream));
  memcpy(info, h->h_info, (size_t)n * sizeof(b200qp_solver_info));
  return B200QP_OK;
}

int32_t b200qp_mpc_solve_batch(b200qp_mpc *h, int32_t n, const double *in, const uint8_t *contact, double *forces,
                               double *xpred, b200qp_solver_info *info) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!in || !contact || !forces || !xpred || !info))) {
    g_err = "b200qp_mpc_solve_batch: bad arguments";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const int N = h->cfg.horizon;
  const size_t in_b = (size_t)n * b200qp_mpc_in_doubles(N) * sizeof(double);
  // Inputs: DMA straight from the caller's buffers when they are page-locked (cudaHostAlloc / cudaHostRegister /
  // torch pin_memory), else stage through the handle's pinned buffers.
  const bool in_pinned = is_pinned(in) && is_pinned(contact);
  const double *src_in = in;
  const uint8_t *src_ct = contact;
  if (!in_pinned) {
    memcpy(h->h_in, in, in_b);
    memcpy(h->h_contact, contact, (size_t)n * N * 4);
    src_in = h->h_in;
    src_ct = h->h_contact;
  }
  CK(cudaMemcpyAsync(h->d_contact, src_ct, (size_t)n * N * 4, cudaMemcpyHostToDevice, h->stream));
  // The records (2.3 KB per problem, the bulk of the input) go up in chunks on the copy stream while the condensed kernel is
  // already solving the chunks that have landed: a per-chunk flag, written by a 4-byte copy queued behind the chunk's data,
  // is what the kernel polls before it stages a record.  The Riccati kernel has no such wait, so only CONDENSED_ADMM streams.
  const int chunks = (h->cfg.solver == B200QP_SOLVER_CONDENSED_ADMM && !h->no_zero_copy && h->async_engines >= 1 && n >= 2048) ? h->in_chunks : 1;
  if (chunks == 1) {
    CK(cudaMemcpyAsync(h->d_in, src_in, in_b, cudaMemcpyHostToDevice, h->stream));
    return solve_and_fetch(h, n, forces, xpred, info);
  }
  const int per = (n + chunks - 1) / chunks;
  const size_t rec_d = (size_t)b200qp_mpc_in_doubles(N);
  CK(cudaMemsetAsync(h->d_ready, 0, 16 * sizeof(int), h->stream));
  CK(cudaEventRecord(h->ev_copy, h->stream));
  CK(cudaStreamWaitEvent(h->copy_stream, h->ev_copy, 0));
  for (int c = 0; c < chunks; ++c) {
    const int lo = c * per, m = (lo + per <= n ? per : n - lo);
    if (m <= 0) break;
    CK(cudaMemcpyAsync(h->d_in + lo * rec_d, src_in + lo * rec_d, m * rec_d * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_ready + c, h->h_one, sizeof(int), cudaMemcpyHostToDevice, h->copy_stream));
  }
  h->stream_chunk = per;
  const int32_t rc = solve_and_fetch(h, n, forces, xpred, info);
  h->stream_chunk = 0;
  CK(cudaStreamSynchronize(h->copy_stream));
  return rc;
}

// ---- one batch over several GPUs: independent slices, one host thread per GPU ------------------------------------------
int32_t b200qp_shard_range(int32_t n, int32_t rank, int32_t world, int32_t *lo, int32_t *hi) {
  if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return B200QP_ERR_ARG;
  const int64_t per = ((int64_t)n + world - 1) / world;  // contiguous slices of ceil(n / G); the last ones may be short or empty
  int64_t a = per * rank, b = a + per;
  if (a > n) a = n;
  if (b > n) b = n;
  *lo = (int32_t)a;
  *hi = (int32_t)b;
  return B200QP_OK;
}
}  // extern "C"
struct b200qp_mpc_sharded {
  std::vector<b200qp_mpc *> h;
  int max_batch = 0;
};
extern "C" {
int32_t b200qp_mpc_create_sharded(const b200qp_mpc_config *cfg, const int32_t *devices, int32_t n_devices, b200qp_mpc_sharded **out) {
  if (!cfg || !devices || !out || n_devices < 1 || cfg->max_batch < 1) return B200QP_ERR_ARG;
  *out = nullptr;
  b200qp_mpc_sharded *s = new (std::nothrow) b200qp_mpc_sharded();
  if (!s) return B200QP_ERR_ARG;
  s->max_batch = cfg->max_batch;
  int32_t lo = 0, hi = 0;
  b200qp_shard_range(cfg->max_batch, 0, n_devices, &lo, &hi);
  for (int g = 0; g < n_devices; ++g) {
    b200qp_mpc_config c = *cfg;
    c.device = devices[g];
    c.max_batch = hi - lo > 0 ? hi - lo : 1;  // slice 0 is the largest
    b200qp_mpc *hg = nullptr;
    const int32_t rc = b200qp_mpc_create(&c, &hg);
    if (rc != B200QP_OK) {
      for (b200qp_mpc *p : s->h) b200qp_mpc_destroy(p);
      delete s;
      return rc;
    }
    s->h.push_back(hg);
  }
  *out = s;
  return B200QP_OK;
}
void b200qp_mpc_destroy_sharded(b200qp_mpc_sharded *s) {
  if (!s) return;
  for (b200qp_mpc *p : s->h) b200qp_mpc_destroy(p);
  delete s;
}
b200qp_mpc *b200qp_mpc_sharded_handle(b200qp_mpc_sharded *s, int32_t g) {
  return (s && g >= 0 && g < (int32_t)s->h.size()) ? s->h[g] : nullptr;
}
int32_t b200qp_mpc_sharded_devices(const b200qp_mpc_sharded *s) { return s ? (int32_t)s->h.size() : 0; }

int32_t b200qp_mpc_solve_batch_sharded(b200qp_mpc_sharded *s, int32_t n, const double *in, const uint8_t *contact, double *forces,
                                       double *xpred, b200qp_solver_info *info) {
  if (!s || n < 0 || n > s->max_batch || (n > 0 && (!in || !contact || !forces || !xpred || !info))) {
    g_err = "b200qp_mpc_solve_batch_sharded: bad arguments";
    return B200QP_ERR_ARG;
  }
  const int G = (int)s->h.size();
  std::vector<int32_t> rc(G, B200QP_OK);
  std::vector<std::string> err(G);
  std::vector<std::thread> th;
  th.reserve(G);
  for (int g = 0; g < G; ++g) {
    int32_t lo = 0, hi = 0;
    b200qp_shard_range(n, g, G, &lo, &hi);
    if (hi <= lo) continue;
    b200qp_mpc *hg = s->h[g];
    const int N = hg->cfg.horizon;
    const size_t rec_d = (size_t)b200qp_mpc_in_doubles(N);
    th.emplace_back([=, &rc, &err]() {  // one host thread per GPU: its own device context, stream and staging buffers
      rc[g] = b200qp_mpc_solve_batch(hg, hi - lo, in + (size_t)lo * rec_d, contact + (size_t)lo * N * 4, forces + (size_t)lo * N * 12,
                                     xpred + (size_t)lo * (N + 1) * 13, info + lo);
      if (rc[g] != B200QP_OK) err[g] = b200qp_last_error();  // thread-local message of the worker
    });
  }
  for (std::thread &t : th) t.join();
  for (int g = 0; g < G; ++g)
    if (rc[g] != B200QP_OK) {
      g_err = "slice " + std::to_string(g) + ": " + err[g];
      return rc[g];
    }
  return B200QP_OK;
}

int32_t b200qp_mpc_set_model(b200qp_mpc *h, const b200qp_seq_model *m) {
  if (!h || !m || !(m->mass > 0) || !(m->raibert_g > 0) || !(m->max_leg_length > 0)) return B200QP_ERR_ARG;
  SeqModel &M = h->model;
  M.mass = m->mass;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M.inertia[c * 3 + r] = m->inertia[r * 3 + c];  // record layout is column-major
  for (int c = 0; c < 3; ++c) M.com[c] = m->com[c];
  for (int c = 0; c < 12; ++c) M.shoulders[c] = m->shoulders[c];
  M.raibert_k = m->raibert_k;
  M.raibert_g = m->raibert_g;
  M.max_leg_length = m->max_leg_length;
  h->have_model = true;
  return B200QP_OK;
}

// seq_in (+ optional measured contacts) host -> device, then the sequencer kernel into the handle's record buffers.
static int32_t sequence_impl(b200qp_mpc *h, int32_t n, const double *seq_in, const uint8_t *measured, bool on_device) {
  if (!h->have_model) {
    g_err = "b200qp_mpc_sequence: call b200qp_mpc_set_model first";
    return B200QP_ERR_ARG;
  }
  h->model.dt = h->cfg.dt;
  h->model.N = h->cfg.horizon;
  const double *d_seq = seq_in;
  const uint8_t *d_meas = measured;
  if (!on_device) {
    const size_t sb = (size_t)n * B200QP_SEQ_IN_DOUBLES * sizeof(double);
    const double *src = seq_in;
    if (!is_pinned(seq_in)) {
      memcpy(h->h_seq, seq_in, sb);
      src = h->h_seq;
    }
    CK(cudaMemcpyAsync(h->d_seq, src, sb, cudaMemcpyHostToDevice, h->stream));
    d_seq = h->d_seq;
    if (measured) {
      uint8_t *stage = reinterpret_cast<uint8_t *>(h->h_seq + (size_t)h->cfg.max_batch * B200QP_SEQ_IN_DOUBLES);
      memcpy(stage, measured, (size_t)n * 4);
      CK(cudaMemcpyAsync(h->d_measured, stage, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
      d_meas = h->d_measured;
    }
  }
  CK(launch_sequencer(n, h->model, d_seq, d_meas, h->d_in, h->d_contact, h->P.in_stride, h->stream));
  return B200QP_OK;
}

// ---- stateful sequencers --------------------------------------------------------------------------------------------------
int32_t b200qp_mpc_sequencer_configure(b200qp_mpc *h, const b200qp_seq_options *o) {
  if (!h || !o || (o->mode != 0 && o->mode != 1) || o->raibert_filtersize < 1 || o->raibert_filtersize > 32) return B200QP_ERR_ARG;
  if (o->mode == 1 && (o->filter_size < 1 || o->filter_size > 24 || !(o->swing_time > 0) || !(o->correction_period > 0) || !(o->control_dt > 0)))
    return B200QP_ERR_ARG;
  if (!h->have_model) {
    g_err = "b200qp_mpc_sequencer_configure: call b200qp_mpc_set_model first";
    return B200QP_ERR_ARG;
  }
  CK(cudaSetDevice(h->cfg.device));
  SeqOptions &O = h->seq_opt;
  O.mode = o->mode;
  O.raibert_filtersize = o->raibert_filtersize;
  O.fix_standing_position = o->fix_standing_position;
  O.early_contact_detection = o->early_contact_detection;
  O.dist_thr = o->fix_position_distance_threshold;
  O.ang_thr = o->fix_position_angular_threshold;
  O.vel_thr = o->fix_position_velocity_threshold;
  O.swing_time = o->swing_time;
  O.zero_velocity_threshold = o->zero_velocity_threshold;
  O.standing_foot_position_threshold = o->standing_foot_position_threshold;
  O.min_v_cmd_factor = o->min_v_cmd_factor;
  O.max_correction_cycles = o->max_correction_cycles;
  O.correction_period = o->correction_period;
  O.disturbance_correction = o->disturbance_correction;
  O.min_v = o->min_v;
  O.offset_delay = o->offset_delay;
  O.max_stride_length = o->max_stride_length;
  for (int l = 0; l < 4; ++l) O.phase_offset[l] = o->phase_offset[l];
  O.control_dt = o->control_dt;
  O.filter_size = o->filter_size;
  O.switch_offsets = o->switch_offsets;
  O.correct_all = o->correct_all;
  const size_t nb = (size_t)h->cfg.max_batch;
  if (!h->d_seq_state) {
    CK(cudaMalloc(&h->d_seq_state, nb * 192 * sizeof(double)));
    CK(cudaMalloc(&h->d_seq2, nb * B200QP_SEQ2_IN_DOUBLES * sizeof(double)));
    CK(cudaMallocHost(&h->h_seq2, nb * B200QP_SEQ2_IN_DOUBLES * sizeof(double) + nb * 4));
  }
  h->have_seq_opt = true;
  return b200qp_mpc_sequencer_reset(h);
}
int32_t b200qp_mpc_sequencer_reset(b200qp_mpc *h) {
  if (!h || !h->d_seq_state) return B200QP_ERR_ARG;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->d_seq_state, 0, (size_t)h->cfg.max_batch * 192 * sizeof(double), h->stream));
  return B200QP_OK;
}
static int32_t sequencer_step_impl(b200qp_mpc *h, int32_t n, const double *seq2_in, const uint8_t *measured) {
  if (!h->have_seq_opt) {
    g_err = "b200qp_mpc_sequencer_step: call b200qp_mpc_sequencer_configure first";
    return B200QP_ERR_ARG;
  }
  h->model.dt = h->cfg.dt;
  h->model.N = h->cfg.horizon;
  const size_t sb = (size_t)n * B200QP_SEQ2_IN_DOUBLES * sizeof(double);
  const double *src = seq2_in;
  if (!is_pinned(seq2_in)) {
    memcpy(h->h_seq2, seq2_in, sb);
    src = h->h_seq2;
  }
  CK(cudaMemcpyAsync(h->d_seq2, src, sb, cudaMemcpyHostToDevice, h->stream));
  const uint8_t *d_meas = nullptr;
  if (measured) {
    uint8_t *stage = reinterpret_cast<uint8_t *>(h->h_seq2 + (size_t)h->cfg.max_batch * B200QP_SEQ2_IN_DOUBLES);
    memcpy(stage, measured, (size_t)n * 4);
    CK(cudaMemcpyAsync(h->d_measured, stage, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
    d_meas = h->d_measured;
  }
  CK(launch_sequencer_full(n, h->model, h->seq_opt, h->d_seq2, d_meas, h->d_seq_state, h->d_in, h->d_contact, h->P.in_stride, h->stream));
  return B200QP_OK;
}
int32_t b200qp_mpc_sequencer_step(b200qp_mpc *h, int32_t n, const double *seq2_in, const uint8_t *measured_contact, double *records,
                                  uint8_t *contact) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!seq2_in || !records || !contact))) return B200QP_ERR_ARG;
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const int32_t rc = sequencer_step_impl(h, n, seq2_in, measured_contact);
  if (rc != B200QP_OK) return rc;
  const int N = h->cfg.horizon;
  CK(cudaMemcpyAsync(h->h_in, h->d_in, (size_t)n * h->P.in_stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_contact, h->d_contact, (size_t)n * N * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(records, h->h_in, (size_t)n * h->P.in_stride * sizeof(double));
  memcpy(contact, h->h_contact, (size_t)n * N * 4);
  return B200QP_OK;
}
int32_t b200qp_mpc_sequencer_step_solve(b200qp_mpc *h, int32_t n, const double *seq2_in, const uint8_t *measured_contact, double *forces,
                                        double *xpred, b200qp_solver_info *info) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!seq2_in || !forces || !xpred || !info))) return B200QP_ERR_ARG;
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  int32_t rc = sequencer_step_impl(h, n, seq2_in, measured_contact);
  if (rc != B200QP_OK) return rc;
  rc = solve_and_fetch(h, n, forces, xpred, info);
  h->last_launches += 1;
  return rc;
}

int32_t b200qp_mpc_sequence_batch(b200qp_mpc *h, int32_t n, const double *seq_in, const uint8_t *measured_contact,
                                  double *records, uint8_t *contact) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!seq_in || !records || !contact))) {
    g_err = "b200qp_mpc_sequence_batch: bad arguments";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  int32_t rc = sequence_impl(h, n, seq_in, measured_contact, false);
  if (rc != B200QP_OK) return rc;
  const int N = h->cfg.horizon;
  CK(cudaMemcpyAsync(h->h_in, h->d_in, (size_t)n * h->P.in_stride * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_contact, h->d_contact, (size_t)n * N * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(records, h->h_in, (size_t)n * h->P.in_stride * sizeof(double));
  memcpy(contact, h->h_contact, (size_t)n * N * 4);
  return B200QP_OK;
}

int32_t b200qp_mpc_sequence_solve_batch(b200qp_mpc *h, int32_t n, const double *seq_in, const uint8_t *measured_contact,
                                        double *forces, double *xpred, b200qp_solver_info *info) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!seq_in || !forces || !xpred || !info))) {
    g_err = "b200qp_mpc_sequence_solve_batch: bad arguments";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  int32_t rc = sequence_impl(h, n, seq_in, measured_contact, false);
  if (rc != B200QP_OK) return rc;
  rc = solve_and_fetch(h, n, forces, xpred, info);
  h->last_launches += 1;
  return rc;
}

int32_t b200qp_mpc_sequence_solve_batch_device(b200qp_mpc *h, int32_t n, const double *d_seq_in,
                                               const uint8_t *d_measured_contact, double *d_forces, double *d_xpred,
                                               b200qp_solver_info *d_info, int32_t sync) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!d_seq_in || !d_forces || !d_xpred || !d_info))) {
    g_err = "b200qp_mpc_sequence_solve_batch_device: bad arguments";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  int32_t rc = sequence_impl(h, n, d_seq_in, d_measured_contact, true);
  if (rc != B200QP_OK) return rc;
  rc = solve_device_impl(h, n, h->d_in, h->d_contact, d_forces, d_xpred, d_info, -1, nullptr, nullptr);
  if (rc != B200QP_OK) return rc;
  h->last_launches += 1;
  if (sync) CK(cudaStreamSynchronize(h->stream));
  return B200QP_OK;
}

int32_t b200qp_mpc_get_duals(b200qp_mpc *h, int32_t n, double *lam_box, double *lam_g) {
  if (!h || n < 0 || n > h->last_n || !lam_box || !lam_g || !h->last_in) return B200QP_ERR_ARG;
  CK(cudaSetDevice(h->cfg.device));
  // Re-runs the last batch with dual output enabled (diagnostic path, not timed).  The inputs of the last solve must
  // still be alive: after b200qp_mpc_solve_batch they are the handle's own buffers; after ..._solve_batch_device they are
  // the CALLER's device pointers, which must not have been freed or overwritten in between.
  const int N = h->cfg.horizon;
  DevTmp tmp;
  double *tf = nullptr, *tx = nullptr;
  b200qp_solver_info *ti = nullptr;
  CK(tmp.alloc(&tf, (size_t)n * N * 12 * sizeof(double)));
  CK(tmp.alloc(&tx, (size_t)n * (N + 1) * 13 * sizeof(double)));
  CK(tmp.alloc(&ti, (size_t)n * sizeof(b200qp_solver_info)));
  CK(cudaMemsetAsync(h->d_lam_box, 0, (size_t)n * N * 12 * sizeof(double), h->stream));
  CK(cudaMemsetAsync(h->d_lam_g, 0, (size_t)n * N * 16 * sizeof(double), h->stream));
  h->want_duals = true;
  const int32_t rc = solve_device_impl(h, n, h->last_in, h->last_contact, tf, tx, ti, -1, nullptr, nullptr);
  h->want_duals = false;
  if (rc == B200QP_OK) {
    CK(cudaMemcpyAsync(lam_box, h->d_lam_box, (size_t)n * N * 12 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(lam_g, h->d_lam_g, (size_t)n * N * 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return rc;
}

int32_t b200qp_mpc_get_condensed(b200qp_mpc *h, int32_t index, double *H, int32_t ldh, double *g) {
  if (!h || !H || !g || index < 0 || index >= h->last_n || !h->last_in) return B200QP_ERR_ARG;
  if (h->cfg.solver == B200QP_SOLVER_SPARSE_RICCATI || h->cfg.solver == B200QP_SOLVER_PARTIAL_CONDENSING) return B200QP_ERR_UNSUPPORTED;
  CK(cudaSetDevice(h->cfg.device));
  const int N = h->cfg.horizon;
  const size_t in_d = (size_t)b200qp_mpc_in_doubles(N);
  DevTmp tmp;
  double *dH = nullptr, *dg = nullptr, *tf = nullptr, *tx = nullptr;
  b200qp_solver_info *ti = nullptr;
  const int nmax = 12 * N;
  CK(tmp.alloc(&dH, (size_t)nmax * nmax * sizeof(double)));
  CK(tmp.alloc(&dg, (size_t)nmax * sizeof(double)));
  CK(tmp.alloc(&tf, (size_t)N * 12 * sizeof(double)));
  CK(tmp.alloc(&tx, (size_t)(N + 1) * 13 * sizeof(double)));
  CK(tmp.alloc(&ti, sizeof(b200qp_solver_info)));
  CK(cudaMemsetAsync(dH, 0, (size_t)nmax * nmax * sizeof(double), h->stream));
  CK(cudaMemsetAsync(dg, 0, (size_t)nmax * sizeof(double), h->stream));
  const double *sv_in = h->last_in;
  const uint8_t *sv_ct = h->last_contact;
  const int sv_n = h->last_n;
  const int32_t rc = solve_device_impl(h, 1, sv_in + (size_t)index * in_d, sv_ct + (size_t)index * N * 4, tf, tx, ti, 0, dH, dg);
  h->last_in = sv_in;
  h->last_contact = sv_ct;
  h->last_n = sv_n;
  if (rc != B200QP_OK) return rc;
  b200qp_solver_info inf;
  memset(&inf, 0, sizeof(inf));
  CK(cudaMemcpyAsync(&inf, ti, sizeof(inf), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const int nfree = inf.n_free;
  // only the condensed kernel forms H and g: a problem that went to the Riccati kernel (solver AUTO, more than 156 free
  // variables) has none to report
  if (nfree > 156) return B200QP_ERR_UNSUPPORTED;
  if (nfree <= 0) return nfree;
  if (nfree > nmax || ldh < nfree) return B200QP_ERR_ARG;
  CK(cudaMemcpy2D(H, (size_t)ldh * sizeof(double), dH, (size_t)nfree * sizeof(double), (size_t)nfree * sizeof(double), nfree,
                  cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(g, dg, (size_t)nfree * sizeof(double), cudaMemcpyDeviceToHost));
  return nfree;
}

int32_t b200qp_mpc_last_timing(b200qp_mpc *h, float *kernel_ms, int32_t *kernel_launches) {
  if (!h) return B200QP_ERR_ARG;
  if (!h->have_timing) {  // nothing has been launched on this handle yet
    if (kernel_ms) *kernel_ms = 0.f;
    if (kernel_launches) *kernel_launches = 0;
    return B200QP_OK;
  }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (kernel_ms) *kernel_ms = ms;
  if (kernel_launches) *kernel_launches = h->last_launches;
  return B200QP_OK;
}

void *b200qp_mpc_stream(b200qp_mpc *h) { return h ? (void *)h->stream : nullptr; }

/* development hook (not part of the public header): per-phase cycle counters of the last solve, [n][10] */
int32_t b200qp_dev_phase_cycles(b200qp_mpc *h, int32_t n, long long *out) {
  if (!h || !h->d_dbg_cycles || !out) return B200QP_ERR_ARG;
  CK(cudaMemcpy(out, h->d_dbg_cycles, (size_t)n * 10 * sizeof(long long), cudaMemcpyDeviceToHost));
  return B200QP_OK;
}

int32_t b200qp_gait_schedule_device(int32_t n, int32_t steps, double dt, const double *d_period, const double *d_duty,
                                    const double *d_offset, const double *d_phase, const uint8_t *d_measured_contact,
                                    double fmin, double fmax, uint8_t *d_contact, double *d_swing_time, double *d_lbu,
                                    double *d_ubu, void *stream) {
  if (n < 0 || steps < 1 || steps > B200QP_GAIT_SEQUENCE_SIZE || !d_period || !d_duty || !d_offset || !d_phase ||
      !d_contact || ((d_lbu == nullptr) != (d_ubu == nullptr)))
    return B200QP_ERR_ARG;
  CK(launch_gait_schedule(n, steps, dt, d_period, d_duty, d_offset, d_phase, d_measured_contact, fmin, fmax, d_contact,
                          d_swing_time, d_lbu, d_ubu, (cudaStream_t)stream));
  return B200QP_OK;
}

int32_t b200qp_gait_schedule(int32_t device, int32_t n, int32_t steps, double dt, const double *period,
                             const double *duty, const double *offset, const double *phase,
                             const uint8_t *measured_contact, double fmin, double fmax, uint8_t *contact,
                             double *swing_time, double *lbu, double *ubu) {
  if (n < 0 || steps < 1 || steps > B200QP_GAIT_SEQUENCE_SIZE || !period || !duty || !offset || !phase || !contact ||
      ((lbu == nullptr) != (ubu == nullptr)))
    return B200QP_ERR_ARG;
  if (n == 0) return B200QP_OK;
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) {
    g_err = "b200qp_gait_schedule: no such CUDA device (there is no CPU fallback)";
    return B200QP_ERR_CUDA;
  }
  CK(cudaSetDevice(device));
  const size_t nn = (size_t)n;
  double *d_per = nullptr, *d_duty = nullptr, *d_off = nullptr, *d_ph = nullptr, *d_sw = nullptr, *d_lb = nullptr,
         *d_ub = nullptr;
  uint8_t *d_meas = nullptr, *d_ct = nullptr;
  int32_t rc = B200QP_OK;
  cudaError_t e = cudaSuccess;
#define TRY(x)                                   \
  if (e == cudaSuccess) {                        \
    e = (x);                                     \
    if (e != cudaSuccess) set_err(#x, e);        \
  }
  TRY(cudaMalloc(&d_per, nn * 8));
  TRY(cudaMalloc(&d_duty, nn * 32));
  TRY(cudaMalloc(&d_off, nn * 32));
  TRY(cudaMalloc(&d_ph, nn * 8));
  TRY(cudaMalloc(&d_ct, nn * steps * 4));
  if (measured_contact) TRY(cudaMalloc(&d_meas, nn * 4));
  if (swing_time) TRY(cudaMalloc(&d_sw, nn * steps * 32));
  if (lbu) {
    TRY(cudaMalloc(&d_lb, nn * steps * 96));
    TRY(cudaMalloc(&d_ub, nn * steps * 96));
  }
  TRY(cudaMemcpy(d_per, period, nn * 8, cudaMemcpyHostToDevice));
  TRY(cudaMemcpy(d_duty, duty, nn * 32, cudaMemcpyHostToDevice));
  TRY(cudaMemcpy(d_off, offset, nn * 32, cudaMemcpyHostToDevice));
  TRY(cudaMemcpy(d_ph, phase, nn * 8, cudaMemcpyHostToDevice));
  if (measured_contact) TRY(cudaMemcpy(d_meas, measured_contact, nn * 4, cudaMemcpyHostToDevice));
  TRY(launch_gait_schedule(n, steps, dt, d_per, d_duty, d_off, d_ph, d_meas, fmin, fmax, d_ct, d_sw, d_lb, d_ub, 0));
  TRY(cudaMemcpy(contact, d_ct, nn * steps * 4, cudaMemcpyDeviceToHost));
  if (swing_time) TRY(cudaMemcpy(swing_time, d_sw, nn * steps * 32, cudaMemcpyDeviceToHost));
  if (lbu) {
    TRY(cudaMemcpy(lbu, d_lb, nn * steps * 96, cudaMemcpyDeviceToHost));
    TRY(cudaMemcpy(ubu, d_ub, nn * steps * 96, cudaMemcpyDeviceToHost));
  }
#undef TRY
  if (e != cudaSuccess) rc = B200QP_ERR_CUDA;
  cudaFree(d_per);
  cudaFree(d_duty);
  cudaFree(d_off);
  cudaFree(d_ph);
  cudaFree(d_ct);
  cudaFree(d_meas);
  cudaFree(d_sw);
  cudaFree(d_lb);
  cudaFree(d_ub);
  return rc;
}


// ---- WBC QP (kernel K3) ---------------------------------------------------------------------------------------------
}  // extern "C"

struct b200qp_wbc {
  b200qp_wbc_config cfg;
  WbcParams P;
  cudaStream_t stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;  // aux: odd chunks of the host-buffer call
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int grid = 0;
  double *d_H = nullptr, *d_g = nullptr, *d_A = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_x = nullptr;
  b200qp_solver_info *d_info = nullptr, *h_info = nullptr;
  int *d_next = nullptr;
  // on-device scene assembly (b200qp_wbc_scene_*)
  WbcSceneConfig scene;
  bool have_scene = false;
  double *d_scene = nullptr, *d_tau = nullptr, *d_feet = nullptr, *h_stage = nullptr;
};

extern "C" {

int32_t b200qp_wbc_create(const b200qp_wbc_config *cfg, b200qp_wbc **out) {
  if (!cfg || !out) return B200QP_ERR_ARG;
  *out = nullptr;
  if (cfg->nq < 1 || cfg->nq > 48 || cfg->neq < 0 || cfg->neq > 32 || cfg->nin < 0 || cfg->nin > 64 || cfg->max_batch < 1) {
    g_err = "b200qp_wbc_create: invalid configuration";
    return B200QP_ERR_ARG;
  }
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_err = "b200qp_wbc_create: no such CUDA device (there is no CPU fallback)";
    return B200QP_ERR_CUDA;
  }
  CK(cudaSetDevice(cfg->device));
  b200qp_wbc *h = new (std::nothrow) b200qp_wbc();
  if (!h) return B200QP_ERR_ARG;
  WbcGuard guard{h};
  h->cfg = *cfg;
  h->P.nq = cfg->nq;
  h->P.neq = cfg->neq;
  h->P.nin = cfg->nin;
  h->P.kkt_tol = cfg->kkt_tol > 0 ? cfg->kkt_tol : 1e-6;
  h->P.max_iter = cfg->max_iter > 0 ? cfg->max_iter : 40;
  h->P.tau = (getenv("B200QP_WBC_TAU") && atof(getenv("B200QP_WBC_TAU")) > 0) ? atof(getenv("B200QP_WBC_TAU")) : 0.9999;
  h->P.sig0 = getenv("B200QP_WBC_SIG0") ? atof(getenv("B200QP_WBC_SIG0")) : 0.01;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  h->grid = wbc_grid(h->P, prop.multiProcessorCount);
  if (h->grid <= 0) {
    g_err = "b200qp_wbc_create: problem dimensions do not fit shared memory";
    return B200QP_ERR_UNSUPPORTED;  // the guard releases h
  }
  const size_t nb = cfg->max_batch, nq = cfg->nq, nc = cfg->neq + cfg->nin;
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
  for (int c = 0; c < 8; ++c) CK(cudaEventCreateWithFlags(&h->ev[c], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaMallocHost(&h->h_info, (size_t)cfg->max_batch * sizeof(b200qp_solver_info)));
  CK(cudaMalloc(&h->d_H, nb * nq * nq * sizeof(double)));
  CK(cudaMalloc(&h->d_g, nb * nq * sizeof(double)));
  CK(cudaMalloc(&h->d_A, nb * nc * nq * sizeof(double)));
  CK(cudaMalloc(&h->d_lo, nb * nc * sizeof(double)));
  CK(cudaMalloc(&h->d_hi, nb * nc * sizeof(double)));
  CK(cudaMalloc(&h->d_x, nb * nq * sizeof(double)));
  CK(cudaMalloc(&h->d_info, nb * sizeof(b200qp_solver_info)));
  CK(cudaMalloc(&h->d_next, 8 * sizeof(int)));  // one work counter per chunk in flight
  guard.h = nullptr;
  *out = h;
  return B200QP_OK;
}

void b200qp_wbc_destroy(b200qp_wbc *h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->d_H);
  cudaFree(h->d_g);
  cudaFree(h->d_A);
  cudaFree(h->d_lo);
  cudaFree(h->d_hi);
  cudaFree(h->d_x);
  cudaFree(h->d_info);
  cudaFree(h->d_next);
  cudaFree(h->d_scene);
  cudaFree(h->d_tau);
  cudaFree(h->d_feet);
  cudaFreeHost(h->h_stage);
  cudaFreeHost(h->h_info);
  for (int c = 0; c < 8; ++c)
    if (h->ev[c]) cudaEventDestroy(h->ev[c]);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int32_t b200qp_wbc_solve_batch_device(b200qp_wbc *h, int32_t n, const double *d_H, const double *d_g, const double *d_A,
                                      const double *d_lower_y, const double *d_upper_y, double *d_x,
                                      b200qp_solver_info *d_info, int32_t sync) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!d_H || !d_g || !d_A || !d_lower_y || !d_upper_y || !d_x || !d_info)))
    return B200QP_ERR_ARG;
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->d_next, 0, sizeof(int), h->stream));
  CK(launch_wbc_qp(h->P, n, d_H, d_g, d_A, d_lower_y, d_upper_y, d_x, d_info, h->d_next, h->grid, h->stream));
  if (sync) CK(cudaStreamSynchronize(h->stream));
  return B200QP_OK;
}

int32_t b200qp_wbc_solve_batch(b200qp_wbc *h, int32_t n, const double *H, const double *g, const double *A,
                               const double *lower_y, const double *upper_y, double *x, b200qp_solver_info *info) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!H || !g || !A || !lower_y || !upper_y || !x || !info)))
    return B200QP_ERR_ARG;
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const size_t nn = n, nq = h->cfg.nq, nc = h->cfg.neq + h->cfg.nin;
  // Large batches are cut into chunks: the H2D copy of chunk c+1 (19.5 KB per QP - this call is PCIe-bound) runs on the
  // copy stream while chunk c is being solved. Even and odd chunks are solved on two streams, so the blocks of chunk c+1
  // fill the SMs the tail of chunk c leaves idle (4096 QPs are only 4.6 per resident block); each chunk has its own work counter.
  static const int wbc_chunks = [] {
    const char *e = getenv("B200QP_WBC_CHUNKS");
    const int v = e ? atoi(e) : 8;
    return (v >= 1 && v <= 8) ? v : 8;
  }();
  const int chunks = (n >= 8192) ? wbc_chunks : (n >= 2048 ? (wbc_chunks < 2 ? wbc_chunks : 2) : 1);
  for (int c = 0; c < chunks; ++c) {  // all copies are queued first: a D2H into pageable memory would block the host
    const size_t lo = nn * c / chunks, hi = nn * (c + 1) / chunks, m = hi - lo;
    if (m == 0) continue;
    CK(cudaMemcpyAsync(h->d_H + lo * nq * nq, H + lo * nq * nq, m * nq * nq * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_g + lo * nq, g + lo * nq, m * nq * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_A + lo * nc * nq, A + lo * nc * nq, m * nc * nq * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_lo + lo * nc, lower_y + lo * nc, m * nc * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_hi + lo * nc, upper_y + lo * nc, m * nc * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(h->d_x + lo * nq, x + lo * nq, m * nq * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));  // keep caller's x on failure
    CK(cudaEventRecord(h->ev[c], h->copy_stream));
  }
  CK(cudaMemsetAsync(h->d_next, 0, 8 * sizeof(int), h->stream));
  if (chunks > 1) {
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
  }
  for (int c = 0; c < chunks; ++c) {
    const size_t lo = nn * c / chunks, hi = nn * (c + 1) / chunks, m = hi - lo;
    if (m == 0) continue;
    cudaStream_t ks = (c & 1) ? h->aux_stream : h->stream;
    CK(cudaStreamWaitEvent(ks, h->ev[c], 0));
    CK(launch_wbc_qp(h->P, (int)m, h->d_H + lo * nq * nq, h->d_g + lo * nq, h->d_A + lo * nc * nq, h->d_lo + lo * nc, h->d_hi + lo * nc,
                     h->d_x + lo * nq, h->d_info + lo, h->d_next + c, h->grid, ks));
    CK(cudaMemcpyAsync(h->h_info + lo, h->d_info + lo, m * sizeof(b200qp_solver_info), cudaMemcpyDeviceToHost, ks));
  }
  if (chunks > 1) {
    CK(cudaEventRecord(h->ev_join, h->aux_stream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  }
  CK(cudaMemcpyAsync(x, h->d_x, nn * nq * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  memcpy(info, h->h_info, nn * sizeof(b200qp_solver_info));
  return B200QP_OK;
}

void *b200qp_wbc_stream(b200qp_wbc *h) { return h ? (void *)h->stream : nullptr; }

// ---- scene assembled on the device ----------------------------------------------------------------------------------------
int32_t b200qp_wbc_scene_configure(b200qp_wbc *h, const b200qp_wbc_scene_config *c) {
  if (!h || !c) return B200QP_ERR_ARG;
  if (h->cfg.nq != 30 || h->cfg.neq != 18 || h->cfg.nin != 28) {
    g_err = "b200qp_wbc_scene_configure: the Go2 reduced-TSID scene needs a handle with nq = 30, neq = 18, nin = 28";
    return B200QP_ERR_ARG;
  }
  if (!(c->mu > 0) || !(c->hessian_regularizer >= 0)) return B200QP_ERR_ARG;
  CK(cudaSetDevice(h->cfg.device));
  h->scene.mu = c->mu;
  for (int i = 0; i < 6; ++i) h->scene.com_pose_weight[i] = c->com_pose_weight[i];
  for (int i = 0; i < 3; ++i) {
    h->scene.foot_force_weight[i] = c->foot_force_weight[i];
    h->scene.foot_pose_weight[i] = c->foot_pose_weight[i];
  }
  h->scene.hessian_regularizer = c->hessian_regularizer;
  for (int i = 0; i < 12; ++i) h->scene.tau_max[i] = c->tau_max[i];
  const size_t nb = (size_t)h->cfg.max_batch;
  if (!h->d_scene) {
    CK(cudaMalloc(&h->d_scene, nb * B200QP_WBC_SCENE_IN_DOUBLES * sizeof(double)));
    CK(cudaMalloc(&h->d_tau, nb * 12 * sizeof(double)));
    CK(cudaMalloc(&h->d_feet, nb * 24 * sizeof(double)));
    CK(cudaMallocHost(&h->h_stage, nb * (B200QP_WBC_SCENE_IN_DOUBLES + 30 + 12) * sizeof(double)));
  }
  h->have_scene = true;
  return B200QP_OK;
}

// Rows [lo, lo + n) of the handle's QP buffers are assembled, solved and turned into torques on stream st with work counter `slot`.
static int32_t wbc_scene_queue(b200qp_wbc *h, int32_t n, const double *d_scene, double *d_x, double *d_tau, b200qp_solver_info *d_info,
                               size_t lo = 0, cudaStream_t st = nullptr, int slot = 0) {
  if (!st) st = h->stream;
  const size_t nq = 30, nc = 46;
  double *H = h->d_H + lo * nq * nq, *g = h->d_g + lo * nq, *A = h->d_A + lo * nc * nq, *l = h->d_lo + lo * nc, *u = h->d_hi + lo * nc;
  CK(launch_wbc_scene(n, h->scene, d_scene, H, g, A, l, u, h->d_feet + lo * 24, st));
  CK(cudaMemsetAsync(h->d_next + slot, 0, sizeof(int), st));
  CK(launch_wbc_qp(h->P, n, H, g, A, l, u, d_x, d_info, h->d_next + slot, h->grid, st));
  if (d_tau) CK(launch_wbc_torque(n, A, l, u, d_x, d_tau, st));
  return B200QP_OK;
}

int32_t b200qp_wbc_scene_solve_batch_device(b200qp_wbc *h, int32_t n, const double *d_scene_in, double *d_x, double *d_tau,
                                            b200qp_solver_info *d_info, int32_t sync) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!d_scene_in || !d_x || !d_info))) return B200QP_ERR_ARG;
  if (!h->have_scene) {
    g_err = "b200qp_wbc_scene_solve_batch_device: call b200qp_wbc_scene_configure first";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const int32_t rc = wbc_scene_queue(h, n, d_scene_in, d_x, d_tau, d_info);
  if (rc != B200QP_OK) return rc;
  if (sync) CK(cudaStreamSynchronize(h->stream));
  return B200QP_OK;
}

int32_t b200qp_wbc_scene_solve_batch(b200qp_wbc *h, int32_t n, const double *scene_in, double *x, double *tau, b200qp_solver_info *info) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!scene_in || !x || !info))) return B200QP_ERR_ARG;
  if (!h->have_scene) {
    g_err = "b200qp_wbc_scene_solve_batch: call b200qp_wbc_scene_configure first";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const size_t nn = n, rin = B200QP_WBC_SCENE_IN_DOUBLES, nq = 30, nc = 46;
  double *st_in = h->h_stage, *st_x = h->h_stage + (size_t)h->cfg.max_batch * rin;
  // The robot states go up in two halves (staged into pinned memory; K6 assembles half 0 while the host stages half 1 - K6 is a
  // thread-per-robot kernel whose duration barely depends on the rows it is given, so it is not cut finer). The QPs are then solved
  // in 8 chunks that alternate between two streams: the blocks of chunk c+1 fill the SMs the tail of chunk c leaves idle, and the
  // host copies the results of chunk c-1 out meanwhile. The torques of a chunk come back into its (by then consumed) staging rows.
  const int in_chunks = (n >= 8192) ? 2 : 1, chunks = (n >= 8192) ? 8 : (n >= 2048 ? 2 : 1);
  for (int c = 0; c < in_chunks; ++c) {
    const size_t lo = nn * c / in_chunks, hi = nn * (c + 1) / in_chunks, m = hi - lo;
    memcpy(st_in + lo * rin, scene_in + lo * rin, m * rin * sizeof(double));
    CK(cudaMemcpyAsync(h->d_scene + lo * rin, st_in + lo * rin, m * rin * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CK(launch_wbc_scene((int)m, h->scene, h->d_scene + lo * rin, h->d_H + lo * nq * nq, h->d_g + lo * nq, h->d_A + lo * nc * nq,
                        h->d_lo + lo * nc, h->d_hi + lo * nc, h->d_feet + lo * 24, h->stream));
  }
  CK(cudaMemsetAsync(h->d_next, 0, 8 * sizeof(int), h->stream));
  CK(cudaEventRecord(h->ev_fork, h->stream));
  CK(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
  for (int c = 0; c < chunks; ++c) {
    const size_t lo = nn * c / chunks, hi = nn * (c + 1) / chunks, m = hi - lo;
    cudaStream_t ks = (c & 1) ? h->aux_stream : h->stream;
    CK(launch_wbc_qp(h->P, (int)m, h->d_H + lo * nq * nq, h->d_g + lo * nq, h->d_A + lo * nc * nq, h->d_lo + lo * nc, h->d_hi + lo * nc,
                     h->d_x + lo * nq, h->d_info + lo, h->d_next + c, h->grid, ks));
    if (tau) CK(launch_wbc_torque((int)m, h->d_A + lo * nc * nq, h->d_lo + lo * nc, h->d_hi + lo * nc, h->d_x + lo * nq, h->d_tau + lo * 12, ks));
    CK(cudaMemcpyAsync(st_x + lo * nq, h->d_x + lo * nq, m * nq * sizeof(double), cudaMemcpyDeviceToHost, ks));
    CK(cudaMemcpyAsync(h->h_info + lo, h->d_info + lo, m * sizeof(b200qp_solver_info), cudaMemcpyDeviceToHost, ks));
    if (tau) CK(cudaMemcpyAsync(st_in + lo * rin, h->d_tau + lo * 12, m * 12 * sizeof(double), cudaMemcpyDeviceToHost, ks));
    CK(cudaEventRecord(h->ev[c], ks));
  }
  CK(cudaEventRecord(h->ev_join, h->aux_stream));  // later work on the handle's stream stays behind the odd chunks
  CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
  for (int c = 0; c < chunks; ++c) {
    const size_t lo = nn * c / chunks, hi = nn * (c + 1) / chunks, m = hi - lo;
    CK(cudaEventSynchronize(h->ev[c]));
    memcpy(info + lo, h->h_info + lo, m * sizeof(b200qp_solver_info));
    size_t solved = 0;
    for (size_t p = 0; p < m; ++p) solved += h->h_info[lo + p].status == B200QP_ACADOS_SUCCESS;
    const double *tn = st_in + lo * rin;
    if (solved == m) {
      memcpy(x + lo * nq, st_x + lo * nq, m * nq * sizeof(double));
      if (tau) memcpy(tau + 12 * lo, tn, m * 12 * sizeof(double));
    } else {  // a failed QP keeps the caller's rows (the kernel leaves its row of d_x as it was, which is not the caller's here)
      for (size_t p = 0; p < m; ++p) {
        if (h->h_info[lo + p].status != B200QP_ACADOS_SUCCESS) continue;
        memcpy(x + nq * (lo + p), st_x + nq * (lo + p), nq * sizeof(double));
        if (tau) memcpy(tau + 12 * (lo + p), tn + 12 * p, 12 * sizeof(double));
      }
    }
  }
  return B200QP_OK;
}

int32_t b200qp_wbc_scene_assemble(b200qp_wbc *h, int32_t n, const double *scene_in, double *H, double *g, double *A, double *lower_y,
                                  double *upper_y, double *feet) {
  if (!h || n < 0 || n > h->cfg.max_batch || (n > 0 && (!scene_in || !H || !g || !A || !lower_y || !upper_y))) return B200QP_ERR_ARG;
  if (!h->have_scene) {
    g_err = "b200qp_wbc_scene_assemble: call b200qp_wbc_scene_configure first";
    return B200QP_ERR_ARG;
  }
  if (n == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  const size_t nn = n;
  CK(cudaMemcpyAsync(h->d_scene, scene_in, nn * B200QP_WBC_SCENE_IN_DOUBLES * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(launch_wbc_scene(n, h->scene, h->d_scene, h->d_H, h->d_g, h->d_A, h->d_lo, h->d_hi, h->d_feet, h->stream));
  CK(cudaMemcpyAsync(H, h->d_H, nn * 900 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(g, h->d_g, nn * 30 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(A, h->d_A, nn * 1380 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(lower_y, h->d_lo, nn * 46 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(upper_y, h->d_hi, nn * 46 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (feet) CK(cudaMemcpyAsync(feet, h->d_feet, nn * 24 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return B200QP_OK;
}

// ---- closed loop on the device ---------------------------------------------------------------------------------------------
int32_t b200qp_mpc_rollout(b200qp_mpc *h, b200qp_wbc *w, int32_t n, int32_t periods, double *seq2_inout, const b200qp_rollout_options *opt,
                           double *history) {
  if (!h || !opt || n < 0 || periods < 0 || n > h->cfg.max_batch || (n > 0 && !seq2_inout) || opt->substeps < 1 ||
      (opt->plant_inertia != 0 && opt->plant_inertia != 1))
    return B200QP_ERR_ARG;
  if (!h->have_seq_opt) {
    g_err = "b200qp_mpc_rollout: call b200qp_mpc_set_model and b200qp_mpc_sequencer_configure first";
    return B200QP_ERR_ARG;
  }
  if (w && (!w->have_scene || w->cfg.device != h->cfg.device || w->cfg.max_batch < n || opt->wbc_per_mpc < 1 || opt->substeps % opt->wbc_per_mpc)) {
    g_err = "b200qp_mpc_rollout: the WBC handle needs a configured scene on the same device, max_batch >= n, and wbc_per_mpc must divide substeps";
    return B200QP_ERR_ARG;
  }
  if (n == 0 || periods == 0) return B200QP_OK;
  CK(cudaSetDevice(h->cfg.device));
  RolloutOptions O;
  O.substeps = opt->substeps;
  O.plant_inertia = opt->plant_inertia;
  O.wbc_per_mpc = opt->wbc_per_mpc;
  O.reserved = 0;
  for (int c = 0; c < 6; ++c) {
    O.com_pose_Kp[c] = opt->com_pose_Kp[c];
    O.com_pose_Kd[c] = opt->com_pose_Kd[c];
  }
  for (int c = 0; c < 3; ++c) {
    O.feet_pose_Kp[c] = opt->feet_pose_Kp[c];
    O.feet_pose_Kd[c] = opt->feet_pose_Kd[c];
  }
  const int N = h->cfg.horizon;
  h->model.dt = h->cfg.dt;
  h->model.N = N;
  const size_t nn = n, sb = nn * B200QP_SEQ2_IN_DOUBLES * sizeof(double);
  DevTmp tmp;
  double *d_f0 = nullptr, *d_hist = nullptr;
  CK(tmp.alloc(&d_f0, nn * 12 * sizeof(double)));
  CK(tmp.alloc(&d_hist, history ? (size_t)periods * nn * B200QP_ROLLOUT_HIST_DOUBLES * sizeof(double) : 8));
  cudaStream_t st = h->stream;
  CK(cudaMemcpyAsync(h->d_seq2, seq2_inout, sb, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(h->d_measured, 1, nn * 4, st));  // first period: every foot measured in contact
  const double dt = opt->control_dt > 0 ? opt->control_dt : h->cfg.dt, dt_ns = (double)llround(dt * 1e9);
  int launches = 0;
  for (int k = 0; k < periods; ++k) {
    CK(launch_sequencer_full(n, h->model, h->seq_opt, h->d_seq2, h->d_measured, h->d_seq_state, h->d_in, h->d_contact, h->P.in_stride, st));
    const int32_t rc = solve_device_impl(h, n, h->d_in, h->d_contact, h->d_forces, h->d_xpred, h->d_info, -1, nullptr, nullptr);
    if (rc != B200QP_OK) return rc;
    launches += 1 + h->last_launches;
    CK(launch_rollout_feet(n, N, h->P.in_stride, h->d_in, h->d_contact, h->d_forces, h->d_info, h->d_seq2, h->d_measured, d_f0, st));
    if (!w) {
      CK(launch_rollout_plant(n, h->model, O.plant_inertia, dt, O.substeps, d_f0, 12, 0, nullptr, d_f0, h->d_seq2, st));
      launches += 3;
    } else {
      for (int s = 0; s < O.wbc_per_mpc; ++s) {
        CK(launch_rollout_wbc_target(n, N, O, h->d_seq2, h->d_xpred, d_f0, w->d_scene, st));
        CK(launch_wbc_scene(n, w->scene, w->d_scene, w->d_H, w->d_g, w->d_A, w->d_lo, w->d_hi, w->d_feet, st));
        CK(cudaMemsetAsync(w->d_next, 0, sizeof(int), st));
        CK(launch_wbc_qp(w->P, n, w->d_H, w->d_g, w->d_A, w->d_lo, w->d_hi, w->d_x, w->d_info, w->d_next, w->grid, st));
        // the WBC's contact forces (x[18:30]) drive the plant; a robot whose WBC QP failed keeps the MPC's forces for this slice
        CK(launch_rollout_plant(n, h->model, O.plant_inertia, dt / O.wbc_per_mpc, O.substeps / O.wbc_per_mpc, w->d_x, 30, 18, w->d_info, d_f0,
                                h->d_seq2, st));
        launches += 4;
      }
      launches += 2;
    }
    CK(launch_rollout_finish(n, dt_ns, d_f0, h->d_info, w ? w->d_info : nullptr, h->d_seq2,
                             history ? d_hist + (size_t)k * nn * B200QP_ROLLOUT_HIST_DOUBLES : nullptr, st));
  }
  CK(cudaMemcpyAsync(seq2_inout, h->d_seq2, sb, cudaMemcpyDeviceToHost, st));
  if (history) CK(cudaMemcpyAsync(history, d_hist, (size_t)periods * nn * B200QP_ROLLOUT_HIST_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  h->last_launches = launches;
  return B200QP_OK;
}

}  // extern "C"
