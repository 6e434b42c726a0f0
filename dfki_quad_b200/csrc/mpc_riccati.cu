// Kernel K2b: batched Riccati-based primal-dual interior-point solve of the SPARSE OCP-QP (HPIPM-style), one thread
// block per problem, FP64.  Replaces, per problem, MPC::GetWrenchSequence with the PARTIAL_CONDENSING_HPIPM plan at
// cond_N == N (ws/src/controllers/src/mit_controller/mpc.cpp:329-470, solver call :410).
//
// Formulation: min sum_k 1/2 x'Qx - (Q xref_k)'x + 1/2 alpha u'u  s.t. x_{k+1} = A x_k + B_k u_k,
// per stance foot the 10 one-sided rows G f <= h  (boxes mpc.cpp:103-109, pyramid mpc.cpp:66-92), swing feet f = 0.
// Mehrotra predictor-corrector; every Newton system is an LQR problem in (dx, du) with R~ = alpha I + G'WG
// (block diagonal per foot), solved by a Riccati recursion that exploits A = I + M (M sparse):
//     backward  Huu = R~ + B'PB,  Hux = B'PA,  K = -Huu^-1 Hux,  P <- Q + A'PA + Hux'K
//     vectors   hu = r~ + B'p,  k = -Huu^-1 hu,  p <- A'p + K'hu            (twice per iteration: predictor, corrector)
//     forward   du = K dx + k,  dx <- A dx + B du
// The iterate always satisfies the dynamics and the x-stationarity exactly (x is re-rolled from u, the costate is the
// exact adjoint), so only u-stationarity, primal slack and complementarity residuals drive the iteration.
// Thread t <-> (stage k = t / 4, foot f = t % 4) owns that foot's 10 (s, z) pairs in registers; the 12x12 stage
// matrices live in shared memory and are processed by all threads; K_k and Huu_k^-1 of every stage go to a per-block
// global scratch (L2 resident) between the backward and forward sweeps.
#include "common.cuh"

namespace b200qp {

namespace {
constexpr double BIGR = 1.0e300;

template <int V, class Op>
__device__ __forceinline__ void block_reduce(double (&v)[V], double *red, Op op) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < V; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] = op(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < V; ++k) red[k * 8 + warp] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < V; ++k) {
    double m = red[k * 8];
    for (int w = 1; w < nw; ++w) m = op(m, red[k * 8 + w]);
    v[k] = m;
  }
}
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };
struct OpMin { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct OpSum { __device__ double operator()(double a, double b) const { return a + b; } };

// G f for one foot; rows: +x -x +y -y +z -z  x-mu z  -x-mu z  y-mu z  -y-mu z
__device__ __forceinline__ void G_apply(double fx, double fy, double fz, double mu, double (&g)[10]) {
  g[0] = fx; g[1] = -fx; g[2] = fy; g[3] = -fy; g[4] = fz; g[5] = -fz;
  g[6] = fx - mu * fz; g[7] = -fx - mu * fz; g[8] = fy - mu * fz; g[9] = -fy - mu * fz;
}
__device__ __forceinline__ void GT_apply(const double (&v)[10], double mu, double (&o)[3]) {
  o[0] = (v[0] - v[1]) + (v[6] - v[7]);
  o[1] = (v[2] - v[3]) + (v[8] - v[9]);
  o[2] = (v[4] - v[5]) - mu * ((v[6] + v[7]) + (v[8] + v[9]));
}
}  // namespace

// NH = horizon capacity of the instantiation (shared-memory stage arrays are sized by it), MINB = resident blocks per SM
// the register allocation is tuned for: the kernel is a chain of dependent FP64 instructions, so resident warps are what
// hides its latency.
template <int BS, int NH, int MINB, bool TM = false>
__global__ void __launch_bounds__(BS, MINB) mpc_riccati_kernel(const MpcParams P, int nprob, const double *__restrict__ in,
                                                         const uint8_t *__restrict__ contact, double *__restrict__ forces,
                                                         double *__restrict__ xpred, b200qp_solver_info *__restrict__ info,
                                                         double *__restrict__ lam_box, double *__restrict__ lam_g,
                                                         double *__restrict__ scratch, int *next, long long *dbg_cycles,
                                                         const int *__restrict__ list, const int *__restrict__ list_count) {
  constexpr int NS = 12;  // controllable state dimension (the 13th state is the constant g, mpc.cpp:396)
  __shared__ double sP[144], sPn[144], sT[144], sPB[144], sHuu[144], sHux[144];
  __shared__ double sKbuf[288], sLi[144];
  __shared__ double sBw[NH * 36];
  __shared__ double sU[NH * 12], sDU[NH * 12], sRt12[NH * 12], sKK[NH * 12], sUb[NH * 12];
  __shared__ double sX[(NH + 1) * 12], sXref[(NH + 1) * 12], sQe[(NH + 1) * 12];
  __shared__ double sLw[(NH + 2) * 3], sLv[(NH + 2) * 3];
  __shared__ double sRb[NH * 4 * 6];  // R~ blocks per foot: xx yy zz xz yz (xy = 0) + pad
  __shared__ double sBwsum[NH * 6];
  __shared__ double sYaw[NH + 1], sX0[13], sRot[9], sVec[64], red[4 * 8];
  __shared__ unsigned char sOn[NH * 4], sAct[NH * 12];
  __shared__ int sNu[NH];
  __shared__ int sSlot;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.N;
  const double dt = P.dt, mu_f = P.mu;
  double *Ks = scratch + (size_t)blockIdx.x * MAXH * 288;
  static_assert(4 * NH <= BS, "one thread per (stage, foot)");

  for (;;) {
    __syncthreads();
    if (tid == 0) sSlot = atomicAdd(next, 1);
    __syncthreads();
    // work list (solver AUTO: the problems too large for the condensed kernel) or the whole batch
    if (sSlot >= (list != nullptr ? *list_count : nprob)) break;
    const int prob = list != nullptr ? list[sSlot] : sSlot;
    const double *rec = in + (size_t)prob * P.in_stride;
    const uint8_t *ct = contact + (size_t)prob * N * 4;
    const double mass = rec[13], dm = dt / mass;
    constexpr bool timing = TM;  // phase accounting is its own instantiation: the product kernel carries no counters
    long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = timing ? clock64() : 0;
#define RPH(idx) do { if (timing && tid == 0) { const long long t__ = clock64(); cyc[idx] += t__ - tlast; tlast = t__; } } while (0)

    // ---- setup (same data as mpc.cpp:332-398) ---------------------------------------------------------------------
    if (tid <= N) {
      double R[3][3];
      quat_to_rot(rec + 26 + 13 * tid, R);
      sYaw[tid] = atan2(R[1][0], R[0][0]);
    }
    if (tid == 32) {
      double R[3][3], e[3];
      quat_to_rot(rec, R);
      rot_to_euler(R, e);
      for (int c = 0; c < 3; ++c) {
        sX0[c] = e[c];
        sX0[3 + c] = rec[4 + c] + rec[23 + c];
        sX0[6 + c] = rec[7 + c];
        sX0[9 + c] = rec[10 + c];
      }
      sX0[12] = GRAVITY_CONSTANT;
      const double psi0 = atan2(R[1][0], R[0][0]);
      const double cs = cos(psi0), sn = sin(psi0);
      sRot[0] = cs; sRot[1] = sn; sRot[2] = 0.0;
      sRot[3] = -sn; sRot[4] = cs; sRot[5] = 0.0;
      sRot[6] = 0.0; sRot[7] = 0.0; sRot[8] = 1.0;
      sVec[63] = psi0;
    }
    if (tid < 4 * N) sOn[tid] = ct[tid] != 0;
    __syncthreads();
    if (tid < N) {  // compact list of the stance inputs of stage tid
      int m = 0;
      for (int j = 0; j < 12; ++j)
        if (sOn[tid * 4 + j / 3]) sAct[12 * tid + m++] = (unsigned char)j;
      sNu[tid] = m;
    }
    if (tid == 0) {  // yaw unwrap, mpc.cpp:334-344
      double before = sVec[63];
      for (int k = 0; k <= N; ++k) {
        double y = sYaw[k];
        const double d = y - before;
        if (d < -M_PI) y += 2 * M_PI;
        else if (d > M_PI) y -= 2 * M_PI;
        before = y;
        sYaw[k] = y;
      }
    }
    __syncthreads();
    if (tid < N) {  // Bw_kf = dt Iw^-1 skew(r) (mpc.cpp:33-43,60-64)
      const int k = tid;
      const double psi = sYaw[k], cs = cos(psi), sn = sin(psi);
      const double Rt[3][3] = {{cs, sn, 0.0}, {-sn, cs, 0.0}, {0.0, 0.0, 1.0}};
      double Ib[3][3], T[3][3], Iw[3][3], Ii[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ib[r][c] = rec[14 + 3 * c + r];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[r][c] = Rt[r][0] * Ib[0][c] + Rt[r][1] * Ib[1][c] + Rt[r][2] * Ib[2][c];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Iw[r][c] = T[r][0] * Rt[c][0] + T[r][1] * Rt[c][1] + T[r][2] * Rt[c][2];
      const double c00 = Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1], c01 = Iw[1][2] * Iw[2][0] - Iw[1][0] * Iw[2][2];
      const double c02 = Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0];
      const double id = 1.0 / (Iw[0][0] * c00 + Iw[0][1] * c01 + Iw[0][2] * c02);
      Ii[0][0] = c00 * id; Ii[0][1] = (Iw[0][2] * Iw[2][1] - Iw[0][1] * Iw[2][2]) * id; Ii[0][2] = (Iw[0][1] * Iw[1][2] - Iw[0][2] * Iw[1][1]) * id;
      Ii[1][0] = c01 * id; Ii[1][1] = (Iw[0][0] * Iw[2][2] - Iw[0][2] * Iw[2][0]) * id; Ii[1][2] = (Iw[0][2] * Iw[1][0] - Iw[0][0] * Iw[1][2]) * id;
      Ii[2][0] = c02 * id; Ii[2][1] = (Iw[0][1] * Iw[2][0] - Iw[0][0] * Iw[2][1]) * id; Ii[2][2] = (Iw[0][0] * Iw[1][1] - Iw[0][1] * Iw[1][0]) * id;
      const double *dp = rec + 26 + 13 * k + 4, *fp = rec + 26 + 13 * (N + 1) + 12 * k;
      for (int f = 0; f < 4; ++f) {
        double *Bw = sBw + (k * 4 + f) * 9;
        if (sOn[k * 4 + f]) {
          const double rx = fp[3 * f] - (dp[0] + rec[23]), ry = fp[3 * f + 1] - (dp[1] + rec[24]), rz = fp[3 * f + 2] - (dp[2] + rec[25]);
          const double S[3][3] = {{0.0, -rz, ry}, {rz, 0.0, -rx}, {-ry, rx, 0.0}};
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Bw[3 * r + c] = (Ii[r][0] * S[0][c] + Ii[r][1] * S[1][c] + Ii[r][2] * S[2][c]) * dt;
        } else {
          for (int t = 0; t < 9; ++t) Bw[t] = 0.0;
        }
      }
    }
    if (tid >= 32 && tid - 32 <= N) {  // reference states (mpc.cpp:372-379)
      const int k = tid - 32;
      double R[3][3], e[3];
      const double *st = rec + 26 + 13 * k;
      quat_to_rot(st, R);
      rot_to_euler(R, e);
      e[2] = sYaw[k];
      for (int c = 0; c < 3; ++c) {
        sXref[12 * k + c] = e[c];
        sXref[12 * k + 3 + c] = st[4 + c] + rec[23 + c];
        sXref[12 * k + 6 + c] = st[7 + c];
        sXref[12 * k + 9 + c] = st[10 + c];
      }
    }
    // starting point: every stance foot carries its share of the weight (strictly inside the pyramid and the box),
    // slacks s = h - G u, duals on the central path z = 1 / s: ~30 % fewer iterations than u = 0, s = max(h, 1), z = 1
    for (int e = tid; e < N * 12; e += BS) {
      const int k = e / 12, j = e - 12 * k;
      double v = 0.0;
      if ((j % 3) == 2 && sOn[k * 4 + j / 3]) {
        const int ns = sOn[k * 4] + sOn[k * 4 + 1] + sOn[k * 4 + 2] + sOn[k * 4 + 3];
        const double span = P.fmax - P.fmin;
        v = fmin(fmax(mass * GRAVITY_CONSTANT / (double)ns, P.fmin + 0.05 * span), P.fmax - 0.05 * span);
      }
      sU[e] = v;
    }
    __syncthreads();

    // ---- IPM state of this thread's foot ---------------------------------------------------------------------------
    const bool on = tid < 4 * N && sOn[tid];
    const int kf = tid, ks = tid >> 2;
    double hh[10] = {P.fmax, P.fmax, P.fmax, P.fmax, P.fmax, -P.fmin, 0.0, 0.0, 0.0, 0.0};
    double s[10], z[10];
    double inv_s[10], inv_z[10];  // 1 / s, 1 / z of the current iterate
#pragma unroll
    for (int r = 0; r < 10; ++r) inv_s[r] = inv_z[r] = 1.0;
    {
      double g0[10];
      G_apply(on ? sU[3 * kf] : 0.0, on ? sU[3 * kf + 1] : 0.0, on ? sU[3 * kf + 2] : 0.0, mu_f, g0);
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        s[r] = fmax(hh[r] - g0[r], 1e-2);
        z[r] = 1.0 / s[r];
      }
    }
    double cnt[1] = {on ? 10.0 : 0.0};
    block_reduce<1>(cnt, red, OpSum());
    const double mtot = fmax(cnt[0], 1.0);
    int status = B200QP_ACADOS_MAXITER, iter = 0;
    double k_stat = 0.0, k_ineq = 0.0, k_comp = 0.0;
    const double target = 1e-4 * P.kkt_tol;  // the cost is flat (alpha = 1e-5) along some directions: GRF parity needs ~1e-11
    double best = BIGR, b_stat = 0.0, b_ineq = 0.0, b_comp = 0.0;
    double sb[10], zb[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) { sb[r] = s[r]; zb[r] = z[r]; }
    int stall = 0;
    bool final_pass = false;

    RPH(0);
    for (iter = 0; iter <= P.ipm_max_iter + 1; ++iter) {
      // (A) roll-out x(u) and exact adjoint -----------------------------------------------------------------------
      if (tid < N) {
        double dw[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
        for (int f = 0; f < 4; ++f) {
          if (!sOn[tid * 4 + f]) continue;
          const double *Bw = sBw + (tid * 4 + f) * 9;
          const double u0 = sU[tid * 12 + 3 * f], u1 = sU[tid * 12 + 3 * f + 1], u2 = sU[tid * 12 + 3 * f + 2];
          for (int r = 0; r < 3; ++r) dw[r] += Bw[3 * r] * u0 + Bw[3 * r + 1] * u1 + Bw[3 * r + 2] * u2;
          dv[0] += u0; dv[1] += u1; dv[2] += u2;
        }
        for (int r = 0; r < 3; ++r) {
          sBwsum[6 * tid + r] = dw[r];
          sBwsum[6 * tid + 3 + r] = dv[r] * dm;
        }
      }
      __syncthreads();
      if (warp == 0) {
        if (lane < 6) {  // omega (0..2) and v (3..5) prefix sums
          double a = sX0[6 + lane];
          sX[6 + lane] = a;
          for (int k = 0; k < N; ++k) {
            a += sBwsum[6 * k + lane];
            if (lane == 5) a -= dt * sX0[12];
            sX[12 * (k + 1) + 6 + lane] = a;
          }
        }
        __syncwarp();
        if (lane < 6) {  // theta (0..2) and p (3..5)
          double a = sX0[lane];
          sX[lane] = a;
          for (int k = 0; k < N; ++k) {
            if (lane < 3) a += dt * (sRot[3 * lane] * sX[12 * k + 6] + sRot[3 * lane + 1] * sX[12 * k + 7] + sRot[3 * lane + 2] * sX[12 * k + 8]);
            else a += dt * sX[12 * k + 6 + lane];
            sX[12 * (k + 1) + lane] = a;
          }
        }
      }
      __syncthreads();
      for (int e = tid; e < N * 12; e += BS) {
        const int k = e / 12 + 1, c = e - 12 * (k - 1);
        sQe[12 * k + c] = P.w[c] * (sX[12 * k + c] - sXref[12 * k + c]);
      }
      __syncthreads();
      if (tid < 6) {
        const int c = tid % 3;
        const bool lin = tid >= 3;
        double sA = 0.0, ssA = 0.0, sB = 0.0;
        for (int k = N; k >= 1; --k) {
          ssA += sA;
          sA += sQe[12 * k + (lin ? 3 : 0) + c];
          sB += sQe[12 * k + (lin ? 9 : 6) + c];
          if (lin) sLv[3 * k + c] = sB + dt * ssA;
          else {
            sLw[3 * k + c] = sB;
            sQe[12 * k + c] = ssA;
          }
        }
      }
      __syncthreads();
      if (tid < N) {
        const int k = tid + 1;
        const double a0 = sQe[12 * k], a1 = sQe[12 * k + 1], a2 = sQe[12 * k + 2];
        for (int c = 0; c < 3; ++c) sLw[3 * k + c] += dt * (sRot[c] * a0 + sRot[3 + c] * a1 + sRot[6 + c] * a2);
      }
      __syncthreads();

      RPH(1);
      // (B) residuals ------------------------------------------------------------------------------------------------
      double rd[3] = {0, 0, 0}, rp[10], Gu[10];
      double rs[3] = {0.0, 0.0, 0.0};  // max |rd|,|rp| ; max s*z ; max violation
      double gap[1] = {0.0};
      bool bad = false;
      if (on) {
        const double fx = sU[kf * 3], fy = sU[kf * 3 + 1], fz = sU[kf * 3 + 2];
        G_apply(fx, fy, fz, mu_f, Gu);
        double gz[3];
        GT_apply(z, mu_f, gz);
        const double *Bw = sBw + kf * 9;
        const double *lw = sLw + 3 * (ks + 1), *lv = sLv + 3 * (ks + 1);
        const double uu[3] = {fx, fy, fz};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          rd[c] = P.alpha * uu[c] + Bw[c] * lw[0] + Bw[3 + c] * lw[1] + Bw[6 + c] * lw[2] + dm * lv[c] + gz[c];
          rs[0] = fmax(rs[0], fabs(rd[c]));
          bad |= !(fabs(rd[c]) < BIGR);
        }
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          rp[r] = Gu[r] + s[r] - hh[r];
          rs[0] = fmax(rs[0], fabs(rp[r]));
          rs[1] = fmax(rs[1], s[r] * z[r]);
          rs[2] = fmax(rs[2], Gu[r] - hh[r]);
          gap[0] += s[r] * z[r];
        }
      }
      const int anybad = __syncthreads_or(bad ? 1 : 0);
      block_reduce<3>(rs, red, OpMax());
      block_reduce<1>(gap, red, OpSum());
      const double rnow = anybad ? BIGR : fmax(rs[0], rs[1]);
      if (final_pass) {  // restored best iterate: report it
        status = (!anybad && rnow <= P.kkt_tol) ? B200QP_ACADOS_SUCCESS : B200QP_ACADOS_MINSTEP;
        k_stat = rs[0]; k_comp = rs[1]; k_ineq = fmax(rs[2], 0.0);
        break;
      }
      bool give_up = anybad || iter == P.ipm_max_iter;
      if (!anybad) {
        k_stat = rs[0]; k_comp = rs[1]; k_ineq = fmax(rs[2], 0.0);
        if ((rs[0] <= target && rs[1] <= target) || cnt[0] == 0.0) {
          status = B200QP_ACADOS_SUCCESS;
          break;
        }
        if (rnow < 0.9 * best) {  // remember the best iterate (the endgame can break down numerically)
          best = rnow; stall = 0;
          b_stat = k_stat; b_ineq = k_ineq; b_comp = k_comp;
#pragma unroll
          for (int r = 0; r < 10; ++r) { sb[r] = s[r]; zb[r] = z[r]; }
          for (int e = tid; e < N * 12; e += BS) sUb[e] = sU[e];
        } else if (++stall >= 3 && best <= P.kkt_tol) {
          give_up = true;  // numerical floor reached
        }
      }
      if (give_up) {
        if (best <= P.kkt_tol) {  // fall back to the best iterate and re-evaluate it
          __syncthreads();
#pragma unroll
          for (int r = 0; r < 10; ++r) { s[r] = sb[r]; z[r] = zb[r]; }
          for (int e = tid; e < N * 12; e += BS) sU[e] = sUb[e];
          final_pass = true;
          __syncthreads();
          continue;
        }
        status = anybad ? B200QP_ACADOS_NAN_DETECTED : B200QP_ACADOS_MAXITER;
        break;
      }
      const double mu_c = gap[0] / mtot;

      RPH(2);
      // (C) R~ blocks ---------------------------------------------------------------------------------------------------
      if (tid < 4 * N) {
        double *Rb = sRb + 6 * tid;
        if (on) {
          double w[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) {  // the only FP64 divisions of the iteration: 1/s and 1/z, reused by both passes below
            inv_s[r] = 1.0 / s[r];
            inv_z[r] = 1.0 / z[r];
            w[r] = z[r] * inv_s[r];
          }
          Rb[0] = P.alpha + (w[0] + w[1]) + (w[6] + w[7]);
          Rb[1] = P.alpha + (w[2] + w[3]) + (w[8] + w[9]);
          Rb[2] = P.alpha + (w[4] + w[5]) + mu_f * mu_f * ((w[6] + w[7]) + (w[8] + w[9]));
          Rb[3] = -mu_f * (w[6] - w[7]);
          Rb[4] = -mu_f * (w[8] - w[9]);
        } else {
          Rb[0] = Rb[1] = Rb[2] = 1.0;
          Rb[3] = Rb[4] = 0.0;
        }
      }
      for (int e = tid; e < 144; e += BS) sP[e] = ((e / 12) == (e % 12)) ? P.w[e / 12] : 0.0;
      __syncthreads();

      // (D) Riccati backward sweep (matrices), compacted to the nu_k stance inputs of every stage ---------------------------
      double *Pc = sP, *Pnx = sPn;
      for (int k = N - 1; k >= 0; --k) {
        const double *Bwk = sBw + k * 36;
        const int nu = sNu[k];
        const unsigned char *act = sAct + 12 * k;
        for (int e = tid; e < 144; e += BS) {
          const int i = e / 12, j = e - 12 * i;
          if (j < nu) {  // PB[i][m], m = compact input index
            const int jj = act[j], f = jj / 3, c = jj - 3 * f;
            const double *Bw = Bwk + 9 * f;
            sPB[e] = Pc[12 * i + 6] * Bw[c] + Pc[12 * i + 7] * Bw[3 + c] + Pc[12 * i + 8] * Bw[6 + c] + dm * Pc[12 * i + 9 + c];
          }
          double t = Pc[e];  // T = P A
          if (j >= 6 && j < 9) t += dt * (Pc[12 * i] * sRot[j - 6] + Pc[12 * i + 1] * sRot[3 + j - 6] + Pc[12 * i + 2] * sRot[6 + j - 6]);
          else if (j >= 9) t += dt * Pc[12 * i + j - 6];
          sT[e] = t;
        }
        __syncthreads();
        for (int e = tid; e < 144; e += BS) {
          const int a = e / 12, j = e - 12 * a;
          if (a >= nu) continue;
          const int aa = act[a], fa = aa / 3, ca = aa - 3 * fa;
          const double *Bw = Bwk + 9 * fa;
          if (j < nu) {  // Huu[a][j] on the compact index set
            const int jj = act[j], fj = jj / 3, cj = jj - 3 * fj;
            double h = Bw[ca] * sPB[12 * 6 + j] + Bw[3 + ca] * sPB[12 * 7 + j] + Bw[6 + ca] * sPB[12 * 8 + j] + dm * sPB[12 * (9 + ca) + j];
            if (fa == fj) {
              const double *Rb = sRb + 6 * (k * 4 + fa);
              if (ca == cj) h += Rb[ca];
              else if (ca + cj == 2) h += Rb[3];  // (x,z)
              else if (ca + cj == 3) h += Rb[4];  // (y,z)
            }
            sHuu[e] = h;
          }
          const int col = j;  // Hux[a][col] = sum_i PB[i][a] A[i][col]
          double x = sPB[12 * col + a];
          if (col >= 6 && col < 9) x += dt * (sPB[a] * sRot[col - 6] + sPB[12 + a] * sRot[3 + col - 6] + sPB[24 + a] * sRot[6 + col - 6]);
          else if (col >= 9) x += dt * sPB[12 * (col - 6) + a];
          sHux[e] = x;
        }
        __syncthreads();
        RPH(3);
        if (warp == 0) {  // Cholesky Huu = L L' and L^-1 inside one warp (explicit Huu^-1 loses the IPM endgame accuracy)
          const double od = (lane < nu) ? sHuu[13 * lane] : 1.0;  // original diagonal, for the round-off pivot floor
          // Everything in this branch is one serial dependency chain (the other warp waits at the barrier), so the
          // instruction count per pivot is what matters: the (row, column) of the trailing-update entries a lane owns
          // is computed once per stage instead of with an integer division per pivot, and there is one rsqrt per pivot.
          int ei[5], ej[5];
          const float rnu = 1.0f / (float)(nu > 0 ? nu : 1);
#pragma unroll
          for (int m = 0; m < 5; ++m) {
            const int e = lane + 32 * m;
            int qi = (int)((float)e * rnu);
            if (qi * nu > e) --qi;
            if ((qi + 1) * nu <= e) ++qi;
            ei[m] = (e < nu * nu) ? qi : -1;
            ej[m] = e - qi * nu;
          }
          const int nm = (nu * nu + 31) >> 5;
          for (int p = 0; p < nu; ++p) {
            const double flo = 1e-13 * __shfl_sync(0xffffffffu, od, p);
            const double piv0 = sHuu[13 * p];
            const double piv = (piv0 > flo) ? piv0 : flo;
            const double rd = fast_rsqrt(piv);
            __syncwarp();
            if (lane < nu) {
              if (lane == p) {
                sHuu[13 * p] = piv * rd;
                sVec[48 + p] = rd;  // 1 / L[p][p]
              } else if (lane > p) {
                sHuu[12 * lane + p] *= rd;
              }
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 5; ++m)
              if (m < nm && ei[m] > p && ej[m] > p && ej[m] <= ei[m])
                sHuu[12 * ei[m] + ej[m]] -= sHuu[12 * ei[m] + p] * sHuu[12 * ej[m] + p];
            __syncwarp();
          }
          RPH(8);
          if (lane < nu) {  // column `lane` of X = L^-1 by forward substitution
            const int j = lane;
            for (int i = 0; i < j; ++i) sLi[12 * i + j] = 0.0;
            for (int i = j; i < nu; ++i) {
              double acc = (i == j) ? 1.0 : 0.0;
              for (int kk2 = j; kk2 < i; ++kk2) acc -= sHuu[12 * i + kk2] * sLi[12 * kk2 + j];
              sLi[12 * i + j] = acc * sVec[48 + i];
            }
          }
          RPH(9);
        } else {  // Pn = Q + A' T
          for (int e = tid - 32; e < 144; e += BS - 32) {
            const int i = e / 12, j = e - 12 * i;
            double v = sT[e];
            if (i >= 6 && i < 9) v += dt * (sRot[i - 6] * sT[j] + sRot[3 + i - 6] * sT[12 + j] + sRot[6 + i - 6] * sT[24 + j]);
            else if (i >= 9) v += dt * sT[12 * (i - 6) + j];
            if (i == j) v += P.w[i];
            Pnx[e] = v;
          }
        }
        __syncthreads();
        RPH(4);
        for (int e = tid; e < 12 * nu; e += BS) {  // Y = L^-1 Hux   (nu x 12)
          const int j = e / 12, col = e - 12 * j;
          double acc = 0.0;
          for (int a = 0; a <= j; ++a) acc = fma(sLi[12 * j + a], sHux[12 * a + col], acc);
          sT[e] = acc;
        }
        __syncthreads();
        for (int e = tid; e < 144; e += BS) {  // K = -L^-T Y ;  P <- Pn - Y'Y (symmetric by construction) ; spill K, L^-1
          const int j = e / 12, col = e - 12 * j;
          double pp = Pnx[e];
          for (int a = 0; a < nu; ++a) pp = fma(-sT[12 * a + j], sT[12 * a + col], pp);
          Pnx[e] = pp;
          if (j < nu) {
            double acc = 0.0;
            for (int a = j; a < nu; ++a) acc = fma(sLi[12 * a + j], sT[12 * a + col], acc);
            Ks[(size_t)k * 288 + e] = -acc;
            Ks[(size_t)k * 288 + 144 + e] = (col < nu) ? sLi[e] : 0.0;
          }
        }
        __syncthreads();
        double *tmp = Pc;
        Pc = Pnx;
        Pnx = tmp;
        RPH(5);
      }

      // (E)/(F) predictor and corrector ---------------------------------------------------------------------------------
      double ds[10], dz[10], sigma_mu = 0.0;
#pragma unroll
      for (int r = 0; r < 10; ++r) ds[r] = dz[r] = 0.0;
      for (int pass = 0; pass < 2; ++pass) {
        if (tid < 4 * N) {
          double rt[3] = {0.0, 0.0, 0.0};
          if (on) {
            double t10[10];
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              const double rc = (pass == 0) ? s[r] * z[r] : (s[r] * z[r] + ds[r] * dz[r] - sigma_mu);
              t10[r] = (z[r] * rp[r] - rc) * inv_s[r];
            }
            GT_apply(t10, mu_f, rt);
            rt[0] += rd[0]; rt[1] += rd[1]; rt[2] += rd[2];
          }
          sRt12[3 * tid] = rt[0]; sRt12[3 * tid + 1] = rt[1]; sRt12[3 * tid + 2] = rt[2];
        }
        __syncthreads();
        RPH(6);
        if (warp == 0) {
          // vector backward sweep: lanes 0..11 hold p; compact input quantities live in lanes < nu
          double pv = 0.0;  // p[lane]
          double pre[9];
          for (int m = 0; m < 9; ++m) pre[m] = Ks[(size_t)(N - 1) * 288 + lane + 32 * m];
          for (int k = N - 1; k >= 0; --k) {
            const int nu = sNu[k];
            for (int m = 0; m < 9; ++m) sKbuf[lane + 32 * m] = pre[m];
            if (k > 0)
              for (int m = 0; m < 9; ++m) pre[m] = Ks[(size_t)(k - 1) * 288 + lane + 32 * m];
            if (lane < 12) sVec[lane] = pv;
            __syncwarp();
            if (lane < nu) {
              const int jj = sAct[12 * k + lane], f = jj / 3, c = jj - 3 * f;
              const double *Bw = sBw + (k * 4 + f) * 9;
              sVec[16 + lane] = sRt12[12 * k + jj] + Bw[c] * sVec[6] + Bw[3 + c] * sVec[7] + Bw[6 + c] * sVec[8] + dm * sVec[9 + c];
            }
            __syncwarp();
            if (lane < nu) {
              double yv = 0.0;
              for (int a = 0; a <= lane; ++a) yv = fma(sKbuf[144 + 12 * lane + a], sVec[16 + a], yv);
              sVec[32 + lane] = yv;  // y = L^-1 hu
            }
            __syncwarp();
            if (lane < 12) {
              double pn = sVec[lane];
              for (int a = 0; a < nu; ++a) pn = fma(sKbuf[12 * a + lane], sVec[16 + a], pn);
              if (lane >= 6 && lane < 9) pn += dt * (sRot[lane - 6] * sVec[0] + sRot[3 + lane - 6] * sVec[1] + sRot[6 + lane - 6] * sVec[2]);
              else if (lane >= 9) pn += dt * sVec[lane - 6];
              pv = pn;
              if (lane < nu) {
                double kk = 0.0;
                for (int a = lane; a < nu; ++a) kk = fma(sKbuf[144 + 12 * a + lane], sVec[32 + a], kk);
                sKK[12 * k + lane] = -kk;
              }
            }
            __syncwarp();
          }
          // forward sweep
          double dx = 0.0;
          for (int m = 0; m < 5; ++m)
            if (lane + 32 * m < 144) pre[m] = Ks[lane + 32 * m];
          for (int k = 0; k < N; ++k) {
            const int nu = sNu[k];
            for (int m = 0; m < 5; ++m)
              if (lane + 32 * m < 144) sKbuf[lane + 32 * m] = pre[m];
            if (k + 1 < N)
              for (int m = 0; m < 5; ++m)
                if (lane + 32 * m < 144) pre[m] = Ks[(size_t)(k + 1) * 288 + lane + 32 * m];
            if (lane < 12) {
              sVec[lane] = dx;
              sVec[16 + lane] = 0.0;
              sDU[12 * k + lane] = 0.0;
            }
            __syncwarp();
            if (lane < nu) {
              double du = sKK[12 * k + lane];
#pragma unroll
              for (int i = 0; i < 12; ++i) du = fma(sKbuf[12 * lane + i], sVec[i], du);
              const int jj = sAct[12 * k + lane];
              sDU[12 * k + jj] = du;
              sVec[16 + jj] = du;
            }
            __syncwarp();
            if (lane < 12) {
              double nx = sVec[lane];
              if (lane < 3) nx += dt * (sRot[3 * lane] * sVec[6] + sRot[3 * lane + 1] * sVec[7] + sRot[3 * lane + 2] * sVec[8]);
              else if (lane < 6) nx += dt * sVec[lane + 6];
              else if (lane < 9) {
                const int r = lane - 6;
                for (int f = 0; f < 4; ++f) {
                  const double *Bw = sBw + (k * 4 + f) * 9;
                  nx += Bw[3 * r] * sVec[16 + 3 * f] + Bw[3 * r + 1] * sVec[16 + 3 * f + 1] + Bw[3 * r + 2] * sVec[16 + 3 * f + 2];
                }
              } else {
                const int c = lane - 9;
                nx += dm * (sVec[16 + c] + sVec[19 + c] + sVec[22 + c] + sVec[25 + c]);
              }
              dx = nx;
            }
            __syncwarp();
          }
        }
        __syncthreads();
        RPH(7);
        // step in (s, z)
        double st[1] = {1.0};
        double dsn[10], dzn[10];
        if (on) {
          double Gd[10], tmax = 0.0;
          G_apply(sDU[3 * kf], sDU[3 * kf + 1], sDU[3 * kf + 2], mu_f, Gd);
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            const double rc = (pass == 0) ? s[r] * z[r] : (s[r] * z[r] + ds[r] * dz[r] - sigma_mu);
            dsn[r] = -rp[r] - Gd[r];
            dzn[r] = (-rc - z[r] * dsn[r]) * inv_s[r];
            tmax = fmax(tmax, fmax(-dsn[r] * inv_s[r], -dzn[r] * inv_z[r]));  // largest relative decrease
          }
          if (tmax > 1.0) st[0] = 1.0 / tmax;  // step to the boundary = 1 / max(-ds/s, -dz/z), capped at 1
        }
        block_reduce<1>(st, red, OpMin());
        if (pass == 0) {
          const double a = st[0];
          double g2[1] = {0.0};
          if (on) {
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              g2[0] += (s[r] + a * dsn[r]) * (z[r] + a * dzn[r]);
              ds[r] = dsn[r];
              dz[r] = dzn[r];
            }
          }
          block_reduce<1>(g2, red, OpSum());
          const double mu_aff = g2[0] / mtot;
          const double ratio = mu_aff / fmax(mu_c, 1e-300);
          sigma_mu = ratio * ratio * ratio * mu_c;
        } else {
          double a = fmin(1.0, 0.995 * st[0]);
          // Monotone-gap safeguard.  Once feasibility is ahead of complementarity, a step whose second-order term makes the
          // gap GROW traps Mehrotra's method in a two-cycle (seen at alpha ~ 1e-6 on single-foot stance patterns: step
          // lengths 0.05 / 0.6 alternating, gap 0.3 / 0.65 alternating, 200 iterations without progress).  The gap along the
          // step is the quadratic gap + a c1 + a^2 c2: cut the step where it stops falling by at least a / 10.
          if (rs[0] < rs[1]) {
            double c12[2] = {0.0, 0.0};
            if (on) {
#pragma unroll
              for (int r = 0; r < 10; ++r) {
                c12[0] += s[r] * dzn[r] + z[r] * dsn[r];
                c12[1] += dsn[r] * dzn[r];
              }
            }
            block_reduce<2>(c12, red, OpSum());
            const double lin = c12[0] + 0.1 * gap[0];
            if (c12[1] > 0.0 && lin < 0.0) a = fmin(a, -lin / c12[1]);
          }
          if (on) {
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              s[r] += a * dsn[r];
              z[r] += a * dzn[r];
            }
            sU[3 * kf] += a * sDU[3 * kf];
            sU[3 * kf + 1] += a * sDU[3 * kf + 1];
            sU[3 * kf + 2] += a * sDU[3 * kf + 2];
          }
        }
        __syncthreads();
      }
    }

    RPH(8);
    if (timing && tid == 0)
      for (int q = 0; q < 10; ++q) dbg_cycles[(size_t)prob * 10 + q] = cyc[q];
    // ---- outputs ------------------------------------------------------------------------------------------------------
    __syncthreads();
    if (status == B200QP_ACADOS_SUCCESS) {
      bool nanout = false;
      for (int e = tid; e < (N + 1) * 12; e += BS) nanout |= !(fabs(sX[e]) < BIGR);
      if (__syncthreads_or(nanout ? 1 : 0)) status = B200QP_NAN_AFTER_SUCCESS;
    }
    if (status == B200QP_ACADOS_SUCCESS) {
      double *fo = forces + (size_t)prob * N * 12;
      for (int e = tid; e < N * 12; e += BS) fo[e] = sOn[e / 3] ? sU[e] : 0.0;
      double *xo = xpred + (size_t)prob * (N + 1) * 13;
      for (int e = tid; e < (N + 1) * 13; e += BS) {
        const int k = e / 13, c = e - 13 * k;
        xo[e] = (c < 12) ? sX[12 * k + c] : sX0[12];
      }
      if (lam_box != nullptr && tid < 4 * N) {
        double *lb = lam_box + (size_t)prob * N * 12 + 3 * kf, *lg = lam_g + (size_t)prob * N * 16 + 4 * kf;
        if (on) {
          lb[0] = z[0] - z[1]; lb[1] = z[2] - z[3]; lb[2] = z[4] - z[5];
          lg[0] = z[6]; lg[1] = -z[7]; lg[2] = z[8]; lg[3] = -z[9];
        } else {
          lb[0] = lb[1] = lb[2] = 0.0;
          lg[0] = lg[1] = lg[2] = lg[3] = 0.0;
        }
      }
    }
    if (tid == 0) {
      b200qp_solver_info inf;
      inf.status = status;
      inf.iterations = iter;
      inf.polish_rounds = 0;
      inf.n_free = (int)(cnt[0] * 0.3 + 0.5);
      inf.kkt[0] = k_stat;
      inf.kkt[1] = 0.0;
      inf.kkt[2] = k_ineq;
      inf.kkt[3] = k_comp;
      info[prob] = inf;
      if (P.ipm_refine_valid != nullptr) P.ipm_refine_valid[prob] = 0;  // this kernel hands no starting point over: cold second opinion
      if (P.ipm_refine_list != nullptr && ((status == B200QP_ACADOS_SUCCESS && fmax(k_stat, k_comp) > P.ipm_refine_thr) ||
                                           status == B200QP_ACADOS_MAXITER || status == B200QP_ACADOS_MINSTEP))
        P.ipm_refine_list[atomicAdd(P.ipm_refine_count, 1)] = prob;  // second opinion by the active-set kernel (see mpc_pcw.cu)
    }
  }
  (void)NS;
}

template <int BS, int NH, int MINB>
static int ric_occ() {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mpc_riccati_kernel<BS, NH, MINB>, BS, 0) != cudaSuccess) {
    cudaGetLastError();
    occ = 1;
  }
  return occ < 1 ? 1 : occ;
}

int riccati_grid(int num_sms, int horizon) {
  const int occ = horizon <= 10 ? ric_occ<64, 10, 8>() : horizon <= 16 ? ric_occ<64, 16, 6>() : ric_occ<96, 20, 3>();
  return num_sms * occ;
}

cudaError_t launch_mpc_riccati(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces,
                               double *xpred, b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch,
                               int *next, int grid, cudaStream_t st, long long *dbg_cycles, const int *list, const int *list_count) {
  if (grid > n) grid = n;
  if (P.N <= 10) {
    if (dbg_cycles != nullptr) mpc_riccati_kernel<64, 10, 8, true><<<grid, 64, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
    else mpc_riccati_kernel<64, 10, 8><<<grid, 64, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
  } else if (P.N <= 16) {
    if (dbg_cycles != nullptr) mpc_riccati_kernel<64, 16, 6, true><<<grid, 64, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
    else mpc_riccati_kernel<64, 16, 6><<<grid, 64, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
  } else {
    if (dbg_cycles != nullptr) mpc_riccati_kernel<96, 20, 3, true><<<grid, 96, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
    else mpc_riccati_kernel<96, 20, 3><<<grid, 96, 0, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, dbg_cycles, list, list_count);
  }
  return cudaGetLastError();
}

}  // namespace b200qp
