// Kernel K2c: PARTIALLY CONDENSED Riccati primal-dual interior-point solve, one thread block per problem, FP64.
// Replaces, per problem, MPC::GetWrenchSequence with the PARTIAL_CONDENSING_HPIPM plan for any cond_N in 1..N
// (ws/src/controllers/src/mit_controller/mpc.cpp:219-225 `cond_N`, :226-234 `hpipm_mode`, solver call :410; experiment axis
// ws/src/solver_experiments/solver_experiments/solver_experiments.py:619-640).
//
// The horizon is cut into cond_N consecutive blocks of M = N / cond_N stages (the first N mod cond_N blocks get one more, as
// HPIPM's d_part_cond_qp does).  Each block is condensed onto its first state, x_{s+M} = A^M x_s + Bbar ubar, so the Newton
// system of the interior-point method is an LQR problem with cond_N stages whose inputs are the stacked stance forces of a
// block.  cond_N = N is the sparse formulation of mpc_riccati.cu, cond_N = 1 the fully condensed one; all of them have the
// same Newton direction, hence the same iterates up to round-off.
//     Huu = Rc_b + W_b + Bbar' P Bbar      Hux = Sc_b + Bbar' P Abar      Hxx = Qc_b + Abar' P Abar
//     L L' = Huu,  X = L^-1,  Y = X Hux,  P <- Hxx - Y'Y                          (matrix sweep, once per iteration)
//     hu = r~ + Bbar' p,  y = X hu,  p <- Abar' p - Y' y                            (vector sweeps, predictor + corrector)
//     t = Y dx + y,  du = -X' t,  dx <- Abar dx + Bbar du
// Rc / Sc / Qc (cost of the states inside a block) do not change between iterations: they are formed once per problem in
// closed form - A = I + M with M^2 = 0 on the 12 controllable states, so A^t = I + t M - and parked in an L2-resident
// scratch; W (barrier weights, 3x3 per foot) and the P terms are added each iteration.  The dense block algebra (blocked
// Cholesky with one thread per row, explicit L^-1 so that every sweep is a mat-vec) keeps all threads of the block busy -
// the sparse kernel's stage factorisations are a one-warp dependency chain - and the factors X, Y stay in shared memory
// between the matrix sweep and the vector sweeps.
// The interior-point shell (Mehrotra predictor-corrector, exact dynamics / adjoint so that only u-stationarity, slack and
// complementarity residuals drive the iteration, monotone-gap safeguard, best-iterate fallback) is the one of
// mpc_riccati.cu; thread t <-> (stage t / 4, foot t % 4) owns that foot's 10 (s, z) pairs in registers.
#include "pcond_common.cuh"

namespace b200qp {

using namespace pc;

// per free input a of the problem (ordered by stage, foot, component): geometry needed by the closed forms
struct PcIn {
  double bw[3];  // column c of Bw_if = dt Iw^-1 skew(r): effect of this force component on omega
  double rb[3];  // Rt(psi0) bw: effect on the Euler-angle rates (x dt per stage travelled)
};

// BS threads per block (>= 4 N and >= nubcap + 13), NH = horizon capacity.  Dynamic shared memory:
// [K: (nubcap + 13) * ld][PB: 12 * nubcap][lst: 4 * BS][factors: fcap]; nubcap multiple of 4, ld odd >= nubcap + 13.
template <int BS, int NH>
__global__ void __launch_bounds__(BS) mpc_pcond_kernel(const MpcParams P, int nprob, const double *__restrict__ in,
                                                       const uint8_t *__restrict__ contact, double *__restrict__ forces,
                                                       double *__restrict__ xpred, b200qp_solver_info *__restrict__ info,
                                                       double *__restrict__ lam_box, double *__restrict__ lam_g,
                                                       double *__restrict__ scratch, size_t scratch_stride, int *next,
                                                       const int *__restrict__ list, const int *__restrict__ list_count, int condN,
                                                       int nubcap, int ld, int fcap, long long *dbg_cycles) {
  extern __shared__ __align__(16) double dsm[];
  const int nacap = nubcap + 13;          // rows of the augmented block matrix [inputs | 12 states | rhs]
  double *sK = dsm;                       // [nacap * ld] column-major lower triangle, ld odd >= nacap
  double *sPB = sK + ((nacap * ld + 1) & ~1);  // [12 * nubcap] row-major: P Bbar (16-byte aligned: double2 staging below)
  double *sLst = sPB + 12 * nubcap;       // [4 * BS]
  double *sFac = sLst + 4 * BS;           // [fcap] packed factors (L, Y', y') of all blocks

  __shared__ double sP[144], sPA[144], sDst[20], sYst[8];
  __shared__ double sBw[NH * 36];
  __shared__ PcIn sIn[NH * 12];
  __shared__ double sU[NH * 12], sDU[NH * 12], sRt12[NH * 12], sUb[NH * 12], sYv[NH * 12], sDul[NH * 12];
  __shared__ double sX[(NH + 1) * 12], sXref[(NH + 1) * 12], sQe[(NH + 1) * 12];
  __shared__ double sLw[(NH + 2) * 3], sLv[(NH + 2) * 3];
  __shared__ double sRb[NH * 4 * 6];
  __shared__ double sBwsum[NH * 6];
  __shared__ double sYaw[NH + 1], sX0[13], sRot[9], sVec[64], red[4 * 8];
  __shared__ unsigned char sOn[NH * 4];
  __shared__ short sFree[NH * 12];   // free input a -> global input index 12 k + 3 f + c
  __shared__ short sFidx[NH * 12];   // global input index -> free index or -1
  __shared__ short sBlkS[NH + 2];    // first stage of block b (b = 0..nb), sBlkS[nb] = N
  __shared__ short sBlkA[NH + 2];    // first free input of block b
  __shared__ int sFoff[NH + 2];      // offset of block b's factors in the factor storage
  __shared__ int sCoff[NH + 2];      // offset of block b's constants in the scratch
  __shared__ int sSlot, sNfree, sFit;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = P.N;
  const double dt = P.dt, mu_f = P.mu;
  double *Cs = scratch + (size_t)blockIdx.x * scratch_stride;
  static_assert(4 * NH <= BS, "one thread per (stage, foot)");
  (void)lane;
  (void)warp;

  for (;;) {
    __syncthreads();
    if (tid == 0) sSlot = atomicAdd(next, 1);
    __syncthreads();
    if (sSlot >= (list != nullptr ? *list_count : nprob)) break;
    const int prob = list != nullptr ? list[sSlot] : sSlot;
    const double *rec = in + (size_t)prob * P.in_stride;
    const uint8_t *ct = contact + (size_t)prob * N * 4;
    const double mass = rec[13], dm = dt / mass;
    const bool timing = dbg_cycles != nullptr;
    long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = timing ? clock64() : 0;
#define CPH(idx) do { if (timing && tid == 0) { const long long t__ = clock64(); cyc[idx] += t__ - tlast; tlast = t__; } } while (0)

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
      sRot[0] = cs; sRot[1] = sn; sRot[2] = 0.0;   // MPC::GetR (mpc.cpp:46-58), row-major
      sRot[3] = -sn; sRot[4] = cs; sRot[5] = 0.0;
      sRot[6] = 0.0; sRot[7] = 0.0; sRot[8] = 1.0;
      sVec[63] = psi0;
    }
    if (tid < 4 * N) sOn[tid] = ct[tid] != 0;
    __syncthreads();
    if (tid == 0) {
      // yaw unwrap, mpc.cpp:334-344
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
    if (tid == 32) {
      // free-input list and block partition.  HPIPM's partial condensing: cond_N blocks, the first N mod cond_N one longer; a
      // block whose stance inputs do not fit the factorisation area is split further (same Newton direction, see header)
      int nfree = 0;
      for (int e = 0; e < 12 * N; ++e) {
        const bool on = sOn[e / 3] != 0;
        sFidx[e] = on ? (short)nfree : (short)-1;
        if (on) sFree[nfree++] = (short)e;
      }
      sNfree = nfree;
      const int cN = condN < 1 ? 1 : (condN > N ? N : condN);
      int nb = 0, k0 = 0;
      for (int b = 0; b < cN; ++b) {
        int len = N / cN + (b < N % cN ? 1 : 0);
        while (len > 0) {  // split until the block's inputs fit
          int take = len, cnt;
          for (;;) {
            cnt = 0;
            for (int e = 4 * k0; e < 4 * (k0 + take); ++e) cnt += 3 * sOn[e];
            if (cnt <= nubcap || take == 1) break;
            take = (take + 1) / 2;
          }
          sBlkS[nb++] = (short)k0;
          k0 += take;
          len -= take;
        }
      }
      sBlkS[nb] = (short)N;
      sVec[62] = (double)nb;
      int foff = 0, coff = 0;
      {  // first free input of every block = number of free inputs before its first stage
        int pre = 0, bn = 0;
        for (int k = 0; k <= N; ++k) {
          while (bn <= nb && sBlkS[bn] == k) sBlkA[bn++] = (short)pre;
          if (k < N) pre += 3 * (sOn[4 * k] + sOn[4 * k + 1] + sOn[4 * k + 2] + sOn[4 * k + 3]);
        }
      }
      for (int b = 0; b < nb; ++b) {
        const int nu = sBlkA[b + 1] - sBlkA[b];
        sFoff[b] = foff;
        sCoff[b] = coff;
        const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
        foff += nu4 * na - (nu4 * (nu4 - 1)) / 2;  // packed columns 0 .. nu4-1, rows a .. na-1
        coff += nu * nu + 12 * nu + 144;
      }
      sFoff[nb] = foff;
      sFit = foff <= fcap;
    }
    __syncthreads();
    const int nb = (int)sVec[62];
    const int nfree = sNfree;
    // factors live in shared memory when the whole problem fits, else behind the constants in the global scratch
    double *Fac = sFit ? sFac : (Cs + (scratch_stride >> 1));

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
    // starting point: every stance foot carries its share of the weight (strictly inside the pyramid and the box)
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
    for (int a = tid; a < nfree; a += BS) {  // geometry of the free inputs
      const int e = sFree[a], kf = e / 3, c = e - 3 * kf;
      const double *Bw = sBw + 9 * kf;
      const double b0 = Bw[c], b1 = Bw[3 + c], b2 = Bw[6 + c];
      sIn[a].bw[0] = b0; sIn[a].bw[1] = b1; sIn[a].bw[2] = b2;
      for (int r = 0; r < 3; ++r) sIn[a].rb[r] = sRot[3 * r] * b0 + sRot[3 * r + 1] * b1 + sRot[3 * r + 2] * b2;
    }
    __syncthreads();
    // ---- block constants: cost of the states inside each block, in closed form (A^t = I + t M) -------------------------
    for (int b = 0; b < nb; ++b) {
      const int s = sBlkS[b], M = sBlkS[b + 1] - s, a0 = sBlkA[b], nu = sBlkA[b + 1] - a0;
      double *Rc = Cs + sCoff[b], *Sc = Rc + nu * nu, *Qc = Sc + 12 * nu;
      for (int e = tid; e < nu * nu; e += BS) {  // Rc[a][a2], column-major like Huu (only the lower triangle is used)
        const int a2 = e / nu, a = e - a2 * nu;
        if (a < a2) continue;
        const int ea = sFree[a0 + a], eb = sFree[a0 + a2];
        const int ia = ea / 12 - s, ib = eb / 12 - s, ca = ea % 3, cb = eb % 3;
        const int m = ia > ib ? ia : ib;
        // inner states x_{s+t}, t = m+1 .. M-1, reached from the two inputs after (t-1-ia) and (t-1-ib) further stages
        const int cnt = M - 1 - m;
        double S2 = 0.0;
        for (int t = m + 1; t <= M - 1; ++t) S2 += (double)((t - 1 - ia) * (t - 1 - ib));
        const PcIn &A = sIn[a0 + a], &B2 = sIn[a0 + a2];
        double v = 0.0;
        if (cnt > 0) {
          v = (double)cnt * (A.bw[0] * P.w[6] * B2.bw[0] + A.bw[1] * P.w[7] * B2.bw[1] + A.bw[2] * P.w[8] * B2.bw[2]) +
              S2 * dt * dt * (A.rb[0] * P.w[0] * B2.rb[0] + A.rb[1] * P.w[1] * B2.rb[1] + A.rb[2] * P.w[2] * B2.rb[2]);
          if (ca == cb) v += dm * dm * ((double)cnt * P.w[9 + ca] + S2 * dt * dt * P.w[3 + ca]);
        }
        if (a == a2) v += P.alpha;
        Rc[a2 * nu + a] = v;
      }
      for (int e = tid; e < 12 * nu; e += BS) {  // Sc[a][col]
        const int a = e / 12, col = e - 12 * a;
        const int ea = sFree[a0 + a], ia = ea / 12 - s, ca = ea % 3;
        const PcIn &A = sIn[a0 + a];
        const int n1 = M - 1 - ia;  // inner states after this input: t = ia+1 .. M-1, kd = t-1-ia = 0 .. n1-1
        double skd = 0.0, stkd = 0.0;
        for (int kd = 0; kd < n1; ++kd) {
          skd += (double)kd;
          stkd += (double)(kd * (kd + 1 + ia));
        }
        double v = 0.0;
        if (n1 > 0) {
          if (col < 3) v = skd * dt * A.rb[col] * P.w[col];
          else if (col < 6) v = (col - 3 == ca) ? skd * dt * dm * P.w[col] : 0.0;
          else if (col < 9) {
            const int j = col - 6;
            v = stkd * dt * dt * (A.rb[0] * P.w[0] * sRot[j] + A.rb[1] * P.w[1] * sRot[3 + j] + A.rb[2] * P.w[2] * sRot[6 + j]) +
                (double)n1 * A.bw[j] * P.w[6 + j];
          } else {
            const int j = col - 9;
            v = (j == ca) ? dm * (stkd * dt * dt * P.w[3 + j] + (double)n1 * P.w[9 + j]) : 0.0;
          }
        }
        Sc[e] = v;
      }
      for (int e = tid; e < 144; e += BS) {  // Qc = sum_{t=0}^{M-1} (A^t)' Q A^t
        const int r = e / 12, c = e - 12 * r;
        const double s0 = (double)M, s1 = 0.5 * (double)(M * (M - 1)), s2 = (double)((M - 1) * M * (2 * M - 1)) / 6.0;
        double v = 0.0;
        if (r == c && r < 6) v = s0 * P.w[r];
        else if (r < 3 && c >= 6 && c < 9) v = s1 * dt * P.w[r] * sRot[3 * r + (c - 6)];
        else if (c < 3 && r >= 6 && r < 9) v = s1 * dt * P.w[c] * sRot[3 * c + (r - 6)];
        else if (r >= 6 && r < 9 && c >= 6 && c < 9) {
          const int i = r - 6, j = c - 6;
          v = s2 * dt * dt * (sRot[i] * P.w[0] * sRot[j] + sRot[3 + i] * P.w[1] * sRot[3 + j] + sRot[6 + i] * P.w[2] * sRot[6 + j]);
          if (i == j) v += s0 * P.w[6 + i];
        } else if (r >= 3 && r < 6 && c == r + 6) v = s1 * dt * P.w[r];
        else if (c >= 3 && c < 6 && r == c + 6) v = s1 * dt * P.w[c];
        else if (r == c && r >= 9) v = s0 * P.w[r] + s2 * dt * dt * P.w[r - 6];
        Qc[e] = v;
      }
    }
    __syncthreads();

    // ---- IPM state of this thread's foot ---------------------------------------------------------------------------
    const bool on = tid < 4 * N && sOn[tid];
    const int kf = tid, ks = tid >> 2;
    double hh[10] = {P.fmax, P.fmax, P.fmax, P.fmax, P.fmax, -P.fmin, 0.0, 0.0, 0.0, 0.0};
    double s[10], z[10];
    double inv_s[10], inv_z[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) inv_s[r] = inv_z[r] = 1.0;
    {
      double g0[10];
      Gf(on ? sU[3 * kf] : 0.0, on ? sU[3 * kf + 1] : 0.0, on ? sU[3 * kf + 2] : 0.0, mu_f, g0);
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        s[r] = fmax(hh[r] - g0[r], 1e-2);
        z[r] = 1.0 / s[r];
      }
    }
    double cnt[1] = {on ? 10.0 : 0.0};
    blk_reduce<1>(cnt, red, PSum());
    const double mtot = fmax(cnt[0], 1.0);
    int status = B200QP_ACADOS_MAXITER, iter = 0;
    double k_stat = 0.0, k_ineq = 0.0, k_comp = 0.0;
    const double target = 1e-4 * P.kkt_tol;
    double best = BIGP;
    double sb[10], zb[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) { sb[r] = s[r]; zb[r] = z[r]; }
    int stall = 0;
    bool final_pass = false;

    CPH(0);
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
      if (tid < 32) {
        if (tid < 6) {  // omega (0..2) and v (3..5) prefix sums
          double a = sX0[6 + tid];
          sX[6 + tid] = a;
          for (int k = 0; k < N; ++k) {
            a += sBwsum[6 * k + tid];
            if (tid == 5) a -= dt * sX0[12];
            sX[12 * (k + 1) + 6 + tid] = a;
          }
        }
        __syncwarp();
        if (tid < 6) {  // theta (0..2) and p (3..5)
          double a = sX0[tid];
          sX[tid] = a;
          for (int k = 0; k < N; ++k) {
            if (tid < 3) a += dt * (sRot[3 * tid] * sX[12 * k + 6] + sRot[3 * tid + 1] * sX[12 * k + 7] + sRot[3 * tid + 2] * sX[12 * k + 8]);
            else a += dt * sX[12 * k + 6 + tid];
            sX[12 * (k + 1) + tid] = a;
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

      CPH(1);
      // (B) residuals ------------------------------------------------------------------------------------------------
      double rd[3] = {0, 0, 0}, rp[10], Gu[10];
      double rs[3] = {0.0, 0.0, 0.0};
      double gap[1] = {0.0};
      bool bad = false;
      if (on) {
        const double fx = sU[kf * 3], fy = sU[kf * 3 + 1], fz = sU[kf * 3 + 2];
        Gf(fx, fy, fz, mu_f, Gu);
        double gz[3];
        GTv(z, mu_f, gz);
        const double *Bw = sBw + kf * 9;
        const double *lw = sLw + 3 * (ks + 1), *lv = sLv + 3 * (ks + 1);
        const double uu[3] = {fx, fy, fz};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          rd[c] = P.alpha * uu[c] + Bw[c] * lw[0] + Bw[3 + c] * lw[1] + Bw[6 + c] * lw[2] + dm * lv[c] + gz[c];
          rs[0] = fmax(rs[0], fabs(rd[c]));
          bad |= !(fabs(rd[c]) < BIGP);
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
      blk_reduce<3>(rs, red, PMax());
      blk_reduce<1>(gap, red, PSum());
      const double rnow = anybad ? BIGP : fmax(rs[0], rs[1]);
      if (final_pass) {
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
        if (rnow < 0.9 * best) {
          best = rnow; stall = 0;
#pragma unroll
          for (int r = 0; r < 10; ++r) { sb[r] = s[r]; zb[r] = z[r]; }
          for (int e = tid; e < N * 12; e += BS) sUb[e] = sU[e];
        } else if (++stall >= 3 && best <= P.kkt_tol) {
          give_up = true;
        }
      }
      if (give_up) {
        if (best <= P.kkt_tol) {
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

      CPH(2);
      // (C) barrier weights: R~ blocks per foot ----------------------------------------------------------------------
      if (tid < 4 * N) {
        double *Rb = sRb + 6 * tid;
        if (on) {
          double w[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            inv_s[r] = 1.0 / s[r];
            inv_z[r] = 1.0 / z[r];
            w[r] = z[r] * inv_s[r];
          }
          Rb[0] = (w[0] + w[1]) + (w[6] + w[7]);
          Rb[1] = (w[2] + w[3]) + (w[8] + w[9]);
          Rb[2] = (w[4] + w[5]) + mu_f * mu_f * ((w[6] + w[7]) + (w[8] + w[9]));
          Rb[3] = -mu_f * (w[6] - w[7]);
          Rb[4] = -mu_f * (w[8] - w[9]);
        } else {
          Rb[0] = Rb[1] = Rb[2] = 1.0;
          Rb[3] = Rb[4] = 0.0;
        }
      }
      // terminal cost-to-go P_N = Q, p_N = 0 (the linear state cost is carried by the exact adjoint)
      for (int e = tid; e < 144; e += BS) sP[e] = ((e / 12) == (e % 12)) ? P.w[e / 12] : 0.0;
      if (tid < 12) sVec[tid] = 0.0;
      // predictor right-hand side r~ = rd + G' ((z rp - s z) / s)
      double ds[10], dz[10], sigma_mu = 0.0;
#pragma unroll
      for (int r = 0; r < 10; ++r) ds[r] = dz[r] = 0.0;
      if (tid < 4 * N) {
        double rt[3] = {0.0, 0.0, 0.0};
        if (on) {
          double t10[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) t10[r] = (z[r] * rp[r] - s[r] * z[r]) * inv_s[r];
          GTv(t10, mu_f, rt);
          rt[0] += rd[0]; rt[1] += rd[1]; rt[2] += rd[2];
        }
        sRt12[3 * tid] = rt[0]; sRt12[3 * tid + 1] = rt[1]; sRt12[3 * tid + 2] = rt[2];
      }
      __syncthreads();

      // (D) block Riccati backward sweep: matrices and the predictor's vector sweep in one augmented factorisation per block
      for (int b = nb - 1; b >= 0; --b) {
        const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, a0 = sBlkA[b], nu = sBlkA[b + 1] - a0;
        const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
        const double Mdt = (double)M * dt;
        const double *Rc = Cs + sCoff[b], *Sc = Rc + nu * nu, *Qc = Sc + 12 * nu;
        // PA = P Abar (all threads) and PB = P Bbar (thread a: column a).  Bbar[:, a] = [kd rb ; kd dm e_c ; bw ; dm e_c] with
        // kd = dt x (stages left in the block after the input's own)
        for (int e = tid; e < 144; e += BS) {
          const int r = e / 12, c = e - 12 * r;
          double v = sP[e];
          if (c >= 6 && c < 9) v += Mdt * (sP[12 * r] * sRot[c - 6] + sP[12 * r + 1] * sRot[3 + c - 6] + sP[12 * r + 2] * sRot[6 + c - 6]);
          else if (c >= 9) v += Mdt * sP[12 * r + c - 6];
          sPA[e] = v;
        }
        int ea = 0, cc = 0;
        double kd = 0.0;
        if (tid < nu) {
          ea = sFree[a0 + tid];
          cc = ea % 3;
          kd = (double)(M - 1 - (ea / 12 - s0)) * dt;
          const PcIn &A = sIn[a0 + tid];
#pragma unroll
          for (int r = 0; r < 12; ++r) {
            const double *Pr = sP + 12 * r;
            sPB[r * nubcap + tid] = kd * (Pr[0] * A.rb[0] + Pr[1] * A.rb[1] + Pr[2] * A.rb[2] + dm * Pr[3 + cc]) +
                                    (Pr[6] * A.bw[0] + Pr[7] * A.bw[1] + Pr[8] * A.bw[2]) + dm * Pr[9 + cc];
          }
        }
        __syncthreads();
        // augmented matrix, lower triangle: [Huu | Hux' | hu] rows = inputs (nu, padded to nu4), 12 states, rhs
        if (tid < nu) {
          const PcIn &A = sIn[a0 + tid];
          double *Krow = sK + tid;
          const double *Rb = sRb + 6 * (ea / 3);
          for (int a2 = 0; a2 <= tid; ++a2) {  // Huu[a][a2] = Rc + W + Bbar_a' (P Bbar)_a2
            double v = __ldcg(Rc + a2 * nu + tid);
            v += kd * (A.rb[0] * sPB[a2] + A.rb[1] * sPB[nubcap + a2] + A.rb[2] * sPB[2 * nubcap + a2] + dm * sPB[(3 + cc) * nubcap + a2]) +
                 (A.bw[0] * sPB[6 * nubcap + a2] + A.bw[1] * sPB[7 * nubcap + a2] + A.bw[2] * sPB[8 * nubcap + a2]) +
                 dm * sPB[(9 + cc) * nubcap + a2];
            const int eb = sFree[a0 + a2];
            if (ea / 3 == eb / 3) {  // same foot: barrier block
              const int cb = eb % 3;
              if (cc == cb) v += Rb[cc];
              else if (cc + cb == 2) v += Rb[3];
              else if (cc + cb == 3) v += Rb[4];
            }
            Krow[a2 * ld] = v;
          }
          double *Kcol = sK + tid * ld + nu4;  // rows below: Hux[a][0..11], hu_a
#pragma unroll
          for (int col = 0; col < 12; ++col)
            Kcol[col] = __ldcg(Sc + 12 * tid + col) +
                        kd * (A.rb[0] * sPA[col] + A.rb[1] * sPA[12 + col] + A.rb[2] * sPA[24 + col] + dm * sPA[12 * (3 + cc) + col]) +
                        (A.bw[0] * sPA[72 + col] + A.bw[1] * sPA[84 + col] + A.bw[2] * sPA[96 + col]) + dm * sPA[12 * (9 + cc) + col];
          // hu = r~ + Bbar' p
          const double q0 = sRot[0] * sVec[0] + sRot[3] * sVec[1] + sRot[6] * sVec[2];  // Rt' p_theta
          const double q1 = sRot[1] * sVec[0] + sRot[4] * sVec[1] + sRot[7] * sVec[2];
          const double q2 = sRot[2] * sVec[0] + sRot[5] * sVec[1] + sRot[8] * sVec[2];
          Kcol[12] = sRt12[ea] + A.bw[0] * (kd * q0 + sVec[6]) + A.bw[1] * (kd * q1 + sVec[7]) + A.bw[2] * (kd * q2 + sVec[8]) +
                     dm * (kd * sVec[3 + cc] + sVec[9 + cc]);
        } else if (tid < nu4) {  // identity padding
          for (int a2 = 0; a2 < tid; ++a2) sK[a2 * ld + tid] = 0.0;
          sK[tid * ld + tid] = 1.0;
          for (int r = nu4; r < na; ++r) sK[tid * ld + r] = 0.0;
        } else if (tid < nu4 + 12) {  // Hxx = Qc + Abar' PA, lower triangle of row r
          const int r = tid - nu4;
          for (int c = 0; c <= r; ++c) {
            double v = __ldcg(Qc + 12 * r + c) + sPA[12 * r + c];
            if (r >= 6 && r < 9) v += Mdt * (sRot[r - 6] * sPA[c] + sRot[3 + r - 6] * sPA[12 + c] + sRot[6 + r - 6] * sPA[24 + c]);
            else if (r >= 9) v += Mdt * sPA[12 * (r - 6) + c];
            sK[(nu4 + c) * ld + tid] = v;
          }
        } else if (tid == nu4 + 12) {  // hx = Abar' p
          for (int c = 0; c < 12; ++c) {
            double v = sVec[c];
            if (c >= 6 && c < 9) v += Mdt * (sRot[c - 6] * sVec[0] + sRot[3 + c - 6] * sVec[1] + sRot[6 + c - 6] * sVec[2]);
            else if (c >= 9) v += Mdt * sVec[c - 6];
            sK[(nu4 + c) * ld + tid] = v;
          }
          sK[tid * ld + tid] = 1.0;
        }
        __syncthreads();
        CPH(3);
        chol_aug(sK, ld, nu4, na, sLst, sDst);
        CPH(4);
        // park the factor columns (thread i: its row), pick up the cost-to-go of the block's first state
        double *F = Fac + sFoff[b];
        if (tid < na) {
          const int lim = tid < nu4 ? tid : nu4 - 1;
          for (int a = 0; a <= lim; ++a) F[fcol(a, na) + tid - a] = sK[a * ld + tid];
        }
        for (int e = tid; e < 144; e += BS) {
          const int r = e / 12, c = e - 12 * r;
          sP[e] = (r >= c) ? sK[(nu4 + c) * ld + nu4 + r] : sK[(nu4 + r) * ld + nu4 + c];
        }
        if (tid < 12) sVec[tid] = sK[(nu4 + tid) * ld + nu4 + 12];
        __syncthreads();
        CPH(5);
      }

      // (E)/(F) predictor forward sweep, then the corrector's backward + forward sweeps ------------------------------------
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
          if (tid < 4 * N) {
            double rt[3] = {0.0, 0.0, 0.0};
            if (on) {
              double t10[10];
#pragma unroll
              for (int r = 0; r < 10; ++r) t10[r] = (z[r] * rp[r] - (s[r] * z[r] + ds[r] * dz[r] - sigma_mu)) * inv_s[r];
              GTv(t10, mu_f, rt);
              rt[0] += rd[0]; rt[1] += rd[1]; rt[2] += rd[2];
            }
            sRt12[3 * tid] = rt[0]; sRt12[3 * tid + 1] = rt[1]; sRt12[3 * tid + 2] = rt[2];
          }
          if (tid < 12) sVec[tid] = 0.0;
          __syncthreads();
          for (int b = nb - 1; b >= 0; --b) {  // y = L^-1 (r~ + Bbar' p),  p <- Abar' p - Y' y  (rows below the factor)
            const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, a0 = sBlkA[b], nu = sBlkA[b + 1] - a0;
            const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
            const double Mdt = (double)M * dt;
            const double *F = Fac + sFoff[b];
            double acc = 0.0;
            if (tid < nu) {
              const int ea = sFree[a0 + tid], cc = ea % 3;
              const double kd = (double)(M - 1 - (ea / 12 - s0)) * dt;
              const PcIn &A = sIn[a0 + tid];
              const double q0 = sRot[0] * sVec[0] + sRot[3] * sVec[1] + sRot[6] * sVec[2];
              const double q1 = sRot[1] * sVec[0] + sRot[4] * sVec[1] + sRot[7] * sVec[2];
              const double q2 = sRot[2] * sVec[0] + sRot[5] * sVec[1] + sRot[8] * sVec[2];
              acc = sRt12[ea] + A.bw[0] * (kd * q0 + sVec[6]) + A.bw[1] * (kd * q1 + sVec[7]) + A.bw[2] * (kd * q2 + sVec[8]) +
                    dm * (kd * sVec[3 + cc] + sVec[9 + cc]);
            } else if (tid >= nu4 && tid < nu4 + 12) {
              const int c = tid - nu4;
              acc = sVec[c];
              if (c >= 6 && c < 9) acc += Mdt * (sRot[c - 6] * sVec[0] + sRot[3 + c - 6] * sVec[1] + sRot[6 + c - 6] * sVec[2]);
              else if (c >= 9) acc += Mdt * sVec[c - 6];
            }
            const double yv = tri_fwd(F, nu4, na, nu4 + 12, acc, sYst);  // starts with a barrier after the reads of sVec above
            if (tid < nu) sYv[a0 + tid] = yv;
            if (tid >= nu4 && tid < nu4 + 12) sVec[tid - nu4] = yv;
            __syncthreads();
          }
        }
        // forward sweep: du = -L^-T (Y dx + y),  dx <- Abar dx + Bbar du
        if (tid < 12) sVec[tid] = 0.0;
        for (int e = tid; e < 12 * N; e += BS) sDU[e] = 0.0;
        __syncthreads();
        for (int b = 0; b < nb; ++b) {
          const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, a0 = sBlkA[b], nu = sBlkA[b + 1] - a0;
          const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
          const double Mdt = (double)M * dt;
          const double *F = Fac + sFoff[b];
          double acc = 0.0;
          if (tid < nu) {
            const double *Ya = F + fcol(tid, na) + nu4 - tid;  // Y[a][0..11], then the predictor's y_a
            acc = (pass == 0) ? Ya[12] : sYv[a0 + tid];
#pragma unroll
            for (int c = 0; c < 12; ++c) acc = fma(Ya[c], sVec[c], acc);
          }
          const double du = -tri_bwd(F, nu4, na, acc, sYst);
          if (tid < nu) {
            sDU[sFree[a0 + tid]] = du;
            sDul[tid] = du;
          }
          __syncthreads();
          if (tid < 12) {  // thread r: row r of Abar dx + Bbar du
            double nx = sVec[tid];
            if (tid < 3) nx += Mdt * (sRot[3 * tid] * sVec[6] + sRot[3 * tid + 1] * sVec[7] + sRot[3 * tid + 2] * sVec[8]);
            else if (tid < 6) nx += Mdt * sVec[tid + 6];
            double a0_ = 0.0, a1_ = 0.0;
            for (int a = 0; a < nu; ++a) {
              const int ea = sFree[a0 + a], cc = ea % 3;
              const double kd = (double)(M - 1 - (ea / 12 - s0)) * dt;
              const PcIn &A = sIn[a0 + a];
              double bv;
              if (tid < 3) bv = kd * A.rb[tid];
              else if (tid < 6) bv = (tid - 3 == cc) ? kd * dm : 0.0;
              else if (tid < 9) bv = A.bw[tid - 6];
              else bv = (tid - 9 == cc) ? dm : 0.0;
              if (a & 1) a1_ = fma(bv, sDul[a], a1_);
              else a0_ = fma(bv, sDul[a], a0_);
            }
            sVec[32 + tid] = nx + (a0_ + a1_);
          }
          __syncthreads();
          if (tid < 12) sVec[tid] = sVec[32 + tid];
          __syncthreads();
        }
        CPH(7);
        // step in (s, z)
        double st[1] = {1.0};
        double dsn[10], dzn[10];
        if (on) {
          double Gd[10], tmax = 0.0;
          Gf(sDU[3 * kf], sDU[3 * kf + 1], sDU[3 * kf + 2], mu_f, Gd);
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            const double rc = (pass == 0) ? s[r] * z[r] : (s[r] * z[r] + ds[r] * dz[r] - sigma_mu);
            dsn[r] = -rp[r] - Gd[r];
            dzn[r] = (-rc - z[r] * dsn[r]) * inv_s[r];
            tmax = fmax(tmax, fmax(-dsn[r] * inv_s[r], -dzn[r] * inv_z[r]));
          }
          if (tmax > 1.0) st[0] = 1.0 / tmax;
        }
        blk_reduce<1>(st, red, PMin());
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
          blk_reduce<1>(g2, red, PSum());
          const double mu_aff = g2[0] / mtot;
          const double ratio = mu_aff / fmax(mu_c, 1e-300);
          sigma_mu = ratio * ratio * ratio * mu_c;
        } else {
          double a = fmin(1.0, 0.995 * st[0]);
          if (rs[0] < rs[1]) {  // monotone-gap safeguard (see mpc_riccati.cu)
            double c12[2] = {0.0, 0.0};
            if (on) {
#pragma unroll
              for (int r = 0; r < 10; ++r) {
                c12[0] += s[r] * dzn[r] + z[r] * dsn[r];
                c12[1] += dsn[r] * dzn[r];
              }
            }
            blk_reduce<2>(c12, red, PSum());
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
      CPH(8);
    }

    if (timing && tid == 0)
      for (int q = 0; q < 10; ++q) dbg_cycles[(size_t)prob * 10 + q] = cyc[q];
    // ---- outputs ------------------------------------------------------------------------------------------------------
    __syncthreads();
    if (status == B200QP_ACADOS_SUCCESS) {
      bool nanout = false;
      for (int e = tid; e < (N + 1) * 12; e += BS) nanout |= !(fabs(sX[e]) < BIGP);
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
      inf.n_free = nfree;
      inf.kkt[0] = k_stat;
      inf.kkt[1] = 0.0;
      inf.kkt[2] = k_ineq;
      inf.kkt[3] = k_comp;
      info[prob] = inf;
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
// One launch shape per (horizon, cond_N): nubcap = largest block (stacked stance inputs, multiple of 4) the factorisation
// area takes - larger blocks are split by the kernel; threads >= 4 N and >= nubcap + 13 (one thread per row of the augmented
// block matrix).
struct PcondShape {
  int bs, nubcap, ld, fcap;
  size_t smem;
};
static PcondShape pcond_shape(int N, int condN, int nub_limit) {
  const int cN = condN < 1 ? 1 : (condN > N ? N : condN);
  const int M = (N + cN - 1) / cN;
  int nub = 12 * M;
  if (nub_limit > 0 && nub > nub_limit) nub = nub_limit;
  nub = (nub + 3) & ~3;
  if (nub < 12) nub = 12;
  if (nub > 112) nub = 112;
  PcondShape s;
  if (N <= 16) s.bs = nub <= 48 ? 64 : 128;
  else s.bs = nub <= 80 ? 96 : 128;
  s.nubcap = nub;
  const int nacap = nub + 13;
  s.ld = nacap | 1;
  // packed factors of every block in shared memory when all feet are in stance, capped at 24 KB (else: global scratch)
  const int nblk = (12 * N + nub - 1) / nub;
  int fc = nblk * (nub * nacap - (nub * (nub - 1)) / 2);
  if (fc > 3072) fc = 3072;
  s.fcap = fc;
  s.smem = sizeof(double) * ((((size_t)nacap * s.ld + 1) & ~(size_t)1) + 12 * (size_t)nub + 4 * (size_t)s.bs + (size_t)s.fcap);
  return s;
}
size_t pcond_scratch_doubles(int N) {
  // first half: block constants, sum_b (nu_b^2 + 12 nu_b + 144) <= 112 * 12 N + 144 N + 144 N;
  // second half: packed factors that do not fit shared memory, sum_b (nu4_b (nu4_b + 13) - ...) <= (112 + 13 + 3) * (12 N + 3 N)
  const size_t half = (size_t)128 * 15 * (size_t)N + 288 * (size_t)N + 1024;
  return 2 * half;
}

template <int BS, int NH>
static cudaError_t pcond_launch_t(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                                  b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, size_t stride, int *next,
                                  const int *list, const int *list_count, int condN, const PcondShape &S, int grid, cudaStream_t st,
                                  long long *dbg) {
  cudaError_t e = cudaFuncSetAttribute(mpc_pcond_kernel<BS, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.smem);
  if (e != cudaSuccess) return e;
  mpc_pcond_kernel<BS, NH><<<grid, BS, S.smem, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, stride, next, list,
                                                     list_count, condN, S.nubcap, S.ld, S.fcap, dbg);
  return cudaGetLastError();
}
template <int BS, int NH>
static int pcond_occ_t(const PcondShape &S) {
  int occ = 0;
  if (cudaFuncSetAttribute(mpc_pcond_kernel<BS, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mpc_pcond_kernel<BS, NH>, BS, S.smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return occ;
}

int pcond_grid(int num_sms, int N, int condN, int nub_limit) {
  const PcondShape S = pcond_shape(N, condN, nub_limit);
  int occ = 0;
  if (S.bs == 64) occ = pcond_occ_t<64, 16>(S);
  else if (S.bs == 96) occ = pcond_occ_t<96, 20>(S);
  else occ = pcond_occ_t<128, 20>(S);
  return num_sms * (occ < 1 ? 1 : occ);
}

cudaError_t launch_mpc_pcond(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                             b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, int *next, int grid,
                             cudaStream_t st, const int *list, const int *list_count, int condN, int nub_limit, long long *dbg) {
  const PcondShape S = pcond_shape(P.N, condN, nub_limit);
  const size_t stride = pcond_scratch_doubles(P.N);
  if (grid > n) grid = n;
  if (S.bs == 64)
    return pcond_launch_t<64, 16>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, stride, next, list, list_count, condN, S, grid, st, dbg);
  if (S.bs == 96)
    return pcond_launch_t<96, 20>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, stride, next, list, list_count, condN, S, grid, st, dbg);
  return pcond_launch_t<128, 20>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, stride, next, list, list_count, condN, S, grid, st, dbg);
}

}  // namespace b200qp
