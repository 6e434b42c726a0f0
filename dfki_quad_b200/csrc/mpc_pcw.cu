// Kernel K2d: block-condensed Riccati primal-dual interior-point solve, ONE WARP PER PROBLEM (32-thread blocks), FP64.
// Same formulation, same interior-point shell and same augmented block factorisation as mpc_pcond.cu (see there):
//   [Huu Hux' hu; . Hxx hx] --(n4 pivots)--> L, Y', y', (P, p)      one factorisation per block = Riccati step + predictor sweep
// but mapped for the latency that bounds these solves: a problem is a chain of dependent FP64 operations, so what counts is
// (a) how short that chain is and (b) how many problems an SM keeps in flight.
//   * one warp owns one problem and never meets another warp: no block barrier, no parked warp (mpc_riccati.cu leaves its
//     second warp at a barrier for half of the time); 32-thread blocks, as many resident problems per SM as shared memory allows;
//   * lane j owns STANCE foot-stage j (the 10 slack / multiplier pairs of a foot in stance at one stage, in registers): the
//     compact list has <= 32 FS entries (FS = 1: trot / pace / bound at N <= 16), swing feet cost nothing;
//   * lane a owns row a of the augmented block matrix (<= 12 stacked inputs + 12 states + 1 rhs = 25 rows), so the blocked
//     elimination and the triangular sweeps of pcond_common.cuh run unchanged with one warp;
//   * the factors of all blocks stay in shared memory between the matrix sweep and the vector sweeps (nothing of the
//     iteration goes to HBM); blocks of M = 1 stage need no block constants at all.
// cond_N (b200qp_mpc_config.cond_N) picks the block length as in mpc_pcond.cu; blocks with more than 12 stacked stance inputs are
// split (e.g. all-stance stages always form blocks of one stage).
#include "pcond_common.cuh"

namespace b200qp {

using namespace pc;

namespace {
template <class Op>
__device__ __forceinline__ double wred(double v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// (row, column) of the e-th entry of a lower triangle enumerated row by row, e = r (r + 1) / 2 + c, for up to 16 rows
__constant__ unsigned char TRI_R[136] = {
    0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8,
    9, 9, 9, 9, 9, 9, 9, 9, 9, 9, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13, 13,
    14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 14, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15};
__constant__ unsigned char TRI_C[136] = {
    0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6, 0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3, 4, 5, 6, 7, 8,
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11,
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13,
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
constexpr int NUBW = 12;        // stacked inputs of a block: 4 stance foot-stages (two trot stages, or one all-stance stage)
constexpr int NAW = NUBW + 13;  // rows of the augmented matrix
constexpr int LDW = 25;         // odd leading dimension
}  // namespace

// NH: horizon capacity, FS: stance foot-stages per lane.  Dynamic shared memory per block (= per problem), doubles:
//   K [NAW * LDW] | PB [12 * NUBW] (aliases the 4 x 32 staging of the elimination) | factors [fcap]
template <int NH, int FS, bool TM>
__global__ void __launch_bounds__(32) mpc_pcw_kernel(const MpcParams P, int nprob, const double *__restrict__ in,
                                                     const uint8_t *__restrict__ contact, double *__restrict__ forces,
                                                     double *__restrict__ xpred, b200qp_solver_info *__restrict__ info,
                                                     double *__restrict__ lam_box, double *__restrict__ lam_g,
                                                     double *__restrict__ scratch, size_t scratch_stride, int *next,
                                                     const int *__restrict__ list, const int *__restrict__ list_count,
                                                     int *__restrict__ reject_list, int *__restrict__ reject_count, int condN, int fcap,
                                                     long long *dbg_cycles) {
  constexpr int MAXFS = 32 * FS;  // stance foot-stages a problem may have
  constexpr int MAXA = 3 * MAXFS;  // free inputs
  extern __shared__ __align__(16) double dsm[];
  double *sK = dsm;                    // [NAW * LDW] column-major lower triangle
  double *sPB = sK + ((NAW * LDW + 1) & ~1);  // [12 * NUBW] P Bbar (row-major) / 4 x 32 staging of the elimination
  double *sLst = sPB;
  double *sFac = sPB + 12 * NUBW;      // [fcap]
  static_assert(12 * NUBW >= 4 * 32, "staging area");
  static_assert(NAW <= 32 && (LDW & 1) && LDW >= NAW, "one lane per row of the augmented block matrix");

  __shared__ double sP[144], sPA[144], sDst[20], sYst[8], sVec[64];
  __shared__ double sIn[NUBW * 8];  // per input of the current block: bw[3], rb[3], kd
  __shared__ unsigned char sTriR[80], sTriC[80], sCc[NUBW];  // lower-triangle enumeration (12 rows + 2 spare), component of an input
  __shared__ double sU[MAXA], sDul[NUBW];
  double *sQe = sK;  // [(NH + 1) * 12] weighted state error: only alive in the roll-out / adjoint phase, before any block is assembled
  static_assert((NH + 1) * 12 <= NAW * LDW, "sQe alias");
  __shared__ double sYaw[NH + 1], sX0[13], sRot[9];
  __shared__ unsigned char sOn[NH * 4], sFs[MAXFS];
  __shared__ short sSt[NH + 2];      // first stance foot-stage of stage k (prefix counts), sSt[N] = nfs
  __shared__ short sBlkS[NH + 2];    // first stage of block b
  __shared__ int sFoff[NH + 2], sCoff[NH + 2];
  __shared__ int sSlot, sNb, sFit, sNfs;

  const int lane = threadIdx.x;
  const int N = P.N;
  const double dt = P.dt, mu_f = P.mu;
  // scratch = [block constants of every resident block | hot tables of every resident block]: the second part (written every
  // iteration, 18 KB per block) is one contiguous range that the host marks L2-persisting for the launch, so that it is not written
  // back to DRAM each time the streaming inputs / outputs push it out
  constexpr size_t HOT = 23 * MAXFS + 21 * 12 + 14 * MAXFS + 12 * 22 + 64 + 12 * 21 + 9 * MAXFS;
  const size_t front_stride = scratch_stride - HOT;
  double *Cs = scratch + (size_t)blockIdx.x * front_stride;
  double *hotb = scratch + (size_t)gridDim.x * front_stride + (size_t)blockIdx.x * HOT;
  double *Bk = hotb + HOT - (20 * MAXFS + MAXA);  // best-iterate backup: (s, z) of every lane's foot-stages, forces
  double *gXref = Bk - (NH + 1) * 12;                      // reference states (read once per iteration)
  // Per-problem tables that are written once (set-up) or once per iteration and read through L1: keeping them out of shared memory
  // is what lets a seventh block reside on the SM (36.2 KB -> 31 KB per block); the latency-bound kernel scales with resident warps.
  double *const sBwc = gXref - MAXFS * 9;   // Bw of the stance foot-stages, row-major 3x3
  double *const sRb = sBwc - MAXFS * 5;
  double *const sBwsum = sRb - NH * 6;
  double *const sLw = sBwsum - (NH + 2) * 3, *const sLv = sLw - (NH + 2) * 3;
  double *const sX = sLv - (NH + 1) * 12, *const sDU = sX - MAXA, *const sRt = sDU - MAXA, *const sYv = sRt - MAXA;

  for (int e = lane; e < 80; e += 32) {  // constant memory serialises lane-dependent indices: keep the tables in shared memory
    sTriR[e] = TRI_R[e];
    sTriC[e] = TRI_C[e];
  }
  for (;;) {
    __syncwarp();
    if (lane == 0) sSlot = atomicAdd(next, 1);
    __syncwarp();
    if (sSlot >= (list != nullptr ? *list_count : nprob)) break;
    const int prob = list != nullptr ? list[sSlot] : sSlot;
    const double *rec = in + (size_t)prob * P.in_stride;
    const uint8_t *ct = contact + (size_t)prob * N * 4;
    const double mass = rec[13], dm = dt / mass;
    constexpr bool timing = TM;  // per-phase clock64 accounting: a separate instantiation, so the product kernel carries no counters
    long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = timing ? clock64() : 0;
#define WPH(idx) do { if (timing && lane == 0) { const long long t__ = clock64(); cyc[idx] += t__ - tlast; tlast = t__; } } while (0)

    // ---- setup (same data as mpc.cpp:332-398) ---------------------------------------------------------------------
    for (int k = lane; k <= N; k += 32) {
      double R[3][3];
      quat_to_rot(rec + 26 + 13 * k, R);
      sYaw[k] = atan2(R[1][0], R[0][0]);
    }
    if (lane == 31) {
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
    {  // compact list of the stance foot-stages, ordered by (stage, foot)
      int base = 0;
      for (int e0 = 0; e0 < 4 * N; e0 += 32) {
        const int e = e0 + lane;
        const bool o = (e < 4 * N) && (ct[e] != 0);
        const unsigned bal = __ballot_sync(0xffffffffu, o);
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (e < 4 * N) sOn[e] = o;
        if (o && pos < MAXFS) sFs[pos] = (unsigned char)e;
        base += __popc(bal);
      }
      if (lane == 0) sNfs = base;
    }
    __syncwarp();
    const int nfs = sNfs;
    if (nfs > MAXFS) {  // more stance foot-stages than this instantiation carries: hand the problem to the next kernel in line
      if (lane == 0) {
        if (reject_list != nullptr) reject_list[atomicAdd(reject_count, 1)] = prob;
        else {
          b200qp_solver_info inf = {B200QP_ACADOS_QP_FAILURE, 0, 0, 3 * nfs, {0, 0, 0, 0}};
          info[prob] = inf;
        }
      }
      continue;
    }
    for (int k = lane; k <= N; k += 32) {  // prefix counts
      int c = 0;
      for (int e = 0; e < 4 * k; ++e) c += sOn[e];
      sSt[k] = (short)c;
    }
    if (lane == 0) {
      double before = sVec[63];  // yaw unwrap, mpc.cpp:334-344
      for (int k = 0; k <= N; ++k) {
        double y = sYaw[k];
        const double d = y - before;
        if (d < -M_PI) y += 2 * M_PI;
        else if (d > M_PI) y -= 2 * M_PI;
        before = y;
        sYaw[k] = y;
      }
    }
    __syncwarp();
    if (lane == 1) {
      // block partition.  HPIPM's partial condensing: cond_N blocks, the first N mod cond_N one longer; a block whose stacked
      // stance inputs exceed 15 is split further (same Newton direction)
      const int cN = condN < 1 ? 1 : (condN > N ? N : condN);
      int nb = 0, k0 = 0;
      for (int b = 0; b < cN; ++b) {
        int len = N / cN + (b < N % cN ? 1 : 0);
        while (len > 0) {
          int take = len;
          while (take > 1 && 3 * (sSt[k0 + take] - sSt[k0]) > NUBW) take = (take + 1) / 2;
          sBlkS[nb++] = (short)k0;
          k0 += take;
          len -= take;
        }
      }
      sBlkS[nb] = (short)N;
      sNb = nb;
      int foff = 0, coff = 0;
      for (int b = 0; b < nb; ++b) {
        const int nu = 3 * (sSt[sBlkS[b + 1]] - sSt[sBlkS[b]]);
        const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
        sFoff[b] = foff;
        sCoff[b] = coff;
        foff += nu4 * na - (nu4 * (nu4 - 1)) / 2;
        if (sBlkS[b + 1] - sBlkS[b] > 1) coff += nu * nu + 12 * nu + 144;  // single-stage blocks have no inner states
      }
      sFit = foff <= fcap;
    }
    for (int k = lane; k < N; k += 32) {  // Bw_kf = dt Iw^-1 skew(r) (mpc.cpp:33-43,60-64) for the stance feet of stage k
      if (sSt[k + 1] == sSt[k]) continue;
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
      int j = sSt[k];
      for (int f = 0; f < 4; ++f) {
        if (!sOn[k * 4 + f]) continue;
        double *Bw = sBwc + 9 * j++;
        const double rx = fp[3 * f] - (dp[0] + rec[23]), ry = fp[3 * f + 1] - (dp[1] + rec[24]), rz = fp[3 * f + 2] - (dp[2] + rec[25]);
        const double S[3][3] = {{0.0, -rz, ry}, {rz, 0.0, -rx}, {-ry, rx, 0.0}};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) Bw[3 * r + c] = (Ii[r][0] * S[0][c] + Ii[r][1] * S[1][c] + Ii[r][2] * S[2][c]) * dt;
      }
    }
    for (int k = lane; k <= N; k += 32) {  // reference states (mpc.cpp:372-379)
      double R[3][3], e[3];
      const double *st = rec + 26 + 13 * k;
      quat_to_rot(st, R);
      rot_to_euler(R, e);
      e[2] = sYaw[k];
      for (int c = 0; c < 3; ++c) {
        gXref[12 * k + c] = e[c];
        gXref[12 * k + 3 + c] = st[4 + c] + rec[23 + c];
        gXref[12 * k + 6 + c] = st[7 + c];
        gXref[12 * k + 9 + c] = st[10 + c];
      }
    }
    // starting point: every stance foot carries its share of the weight (strictly inside the pyramid and the box).  Warm start
    // (mpc.cpp:236-237): 0.9 x the slot's previous solution + 0.1 x that point - a convex combination of a feasible and a strictly
    // interior point is strictly interior; a foot-stage that was in swing before (zero force) starts cold.
    const bool warm = P.ipm_warm_x != nullptr && P.ipm_warm_valid[prob] != 0;
    for (int j = lane; j < nfs; j += 32) {
      const int k = sFs[j] >> 2;
      const int ns = sSt[k + 1] - sSt[k];
      const double span = P.fmax - P.fmin;
      const double fz0 = fmin(fmax(mass * GRAVITY_CONSTANT / (double)ns, P.fmin + 0.05 * span), P.fmax - 0.05 * span);
      double u0 = 0.0, u1 = 0.0, u2 = fz0;
      if (warm) {
        const double *wp = P.ipm_warm_x + (size_t)prob * N * 12 + 3 * sFs[j];
        const double w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const bool usable = w2 > 0.0 && w2 <= P.fmax && fabs(w0) <= mu_f * w2 * (1.0 + 1e-9) && fabs(w1) <= mu_f * w2 * (1.0 + 1e-9) &&
                            fabs(w0) <= P.fmax && fabs(w1) <= P.fmax && w2 >= P.fmin;
        if (usable) {
          u0 = 0.9 * w0;
          u1 = 0.9 * w1;
          u2 = 0.9 * w2 + 0.1 * fz0;
        }
      }
      sU[3 * j] = u0;
      sU[3 * j + 1] = u1;
      sU[3 * j + 2] = u2;
    }
    __syncwarp();
    const int nb = sNb;
    if (!sFit) {  // the packed factors of this problem do not fit the shared-memory area: hand it to the next kernel in line
      if (lane == 0) {
        if (reject_list != nullptr) reject_list[atomicAdd(reject_count, 1)] = prob;
        else {
          b200qp_solver_info inf = {B200QP_ACADOS_QP_FAILURE, 0, 0, 3 * nfs, {0, 0, 0, 0}};
          info[prob] = inf;
        }
      }
      continue;
    }
    double *Fac = sFac;

    // ---- block constants of the multi-stage blocks: cost of the states inside a block, closed form (A^t = I + t M) ----------
    for (int b = 0; b < nb; ++b) {
      const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, j0 = sSt[s0], nu = 3 * (sSt[s0 + M] - j0);
      if (M == 1) continue;
      double *Rc = Cs + sCoff[b], *Sc = Rc + nu * nu, *Qc = Sc + 12 * nu;
      for (int e = lane; e < nu * nu; e += 32) {
        const int a2 = e / nu, a = e - a2 * nu;
        if (a < a2) continue;
        const int ja = j0 + a / 3, jb = j0 + a2 / 3, ca = a % 3, cb = a2 % 3;
        const int ia = (sFs[ja] >> 2) - s0, ib = (sFs[jb] >> 2) - s0;
        const int m = ia > ib ? ia : ib;
        const int cnt = M - 1 - m;
        double S2 = 0.0;
        for (int t = m + 1; t <= M - 1; ++t) S2 += (double)((t - 1 - ia) * (t - 1 - ib));
        const double *Ba = sBwc + 9 * ja, *Bb = sBwc + 9 * jb;
        const double bwa[3] = {Ba[ca], Ba[3 + ca], Ba[6 + ca]}, bwb[3] = {Bb[cb], Bb[3 + cb], Bb[6 + cb]};
        double rba[3], rbb[3];
        for (int r = 0; r < 3; ++r) {
          rba[r] = sRot[3 * r] * bwa[0] + sRot[3 * r + 1] * bwa[1] + sRot[3 * r + 2] * bwa[2];
          rbb[r] = sRot[3 * r] * bwb[0] + sRot[3 * r + 1] * bwb[1] + sRot[3 * r + 2] * bwb[2];
        }
        double v = 0.0;
        if (cnt > 0) {
          v = (double)cnt * (bwa[0] * P.w[6] * bwb[0] + bwa[1] * P.w[7] * bwb[1] + bwa[2] * P.w[8] * bwb[2]) +
              S2 * dt * dt * (rba[0] * P.w[0] * rbb[0] + rba[1] * P.w[1] * rbb[1] + rba[2] * P.w[2] * rbb[2]);
          if (ca == cb) v += dm * dm * ((double)cnt * P.w[9 + ca] + S2 * dt * dt * P.w[3 + ca]);
        }
        if (a == a2) v += P.alpha;
        Rc[a2 * nu + a] = v;
      }
      for (int e = lane; e < 12 * nu; e += 32) {
        const int a = e / 12, col = e - 12 * a;
        const int ja = j0 + a / 3, ca = a % 3, ia = (sFs[ja] >> 2) - s0;
        const double *Ba = sBwc + 9 * ja;
        const double bwa[3] = {Ba[ca], Ba[3 + ca], Ba[6 + ca]};
        double rba[3];
        for (int r = 0; r < 3; ++r) rba[r] = sRot[3 * r] * bwa[0] + sRot[3 * r + 1] * bwa[1] + sRot[3 * r + 2] * bwa[2];
        const int n1 = M - 1 - ia;
        double skd = 0.0, stkd = 0.0;
        for (int kd = 0; kd < n1; ++kd) {
          skd += (double)kd;
          stkd += (double)(kd * (kd + 1 + ia));
        }
        double v = 0.0;
        if (n1 > 0) {
          if (col < 3) v = skd * dt * rba[col] * P.w[col];
          else if (col < 6) v = (col - 3 == ca) ? skd * dt * dm * P.w[col] : 0.0;
          else if (col < 9) {
            const int jj = col - 6;
            v = stkd * dt * dt * (rba[0] * P.w[0] * sRot[jj] + rba[1] * P.w[1] * sRot[3 + jj] + rba[2] * P.w[2] * sRot[6 + jj]) +
                (double)n1 * bwa[jj] * P.w[6 + jj];
          } else {
            const int jj = col - 9;
            v = (jj == ca) ? dm * (stkd * dt * dt * P.w[3 + jj] + (double)n1 * P.w[9 + jj]) : 0.0;
          }
        }
        Sc[e] = v;
      }
      for (int e = lane; e < 144; e += 32) {
        const int r = e / 12, c = e - 12 * r;
        const double q0 = (double)M, q1 = 0.5 * (double)(M * (M - 1)), q2 = (double)((M - 1) * M * (2 * M - 1)) / 6.0;
        double v = 0.0;
        if (r == c && r < 6) v = q0 * P.w[r];
        else if (r < 3 && c >= 6 && c < 9) v = q1 * dt * P.w[r] * sRot[3 * r + (c - 6)];
        else if (c < 3 && r >= 6 && r < 9) v = q1 * dt * P.w[c] * sRot[3 * c + (r - 6)];
        else if (r >= 6 && r < 9 && c >= 6 && c < 9) {
          const int i = r - 6, jj = c - 6;
          v = q2 * dt * dt * (sRot[i] * P.w[0] * sRot[jj] + sRot[3 + i] * P.w[1] * sRot[3 + jj] + sRot[6 + i] * P.w[2] * sRot[6 + jj]);
          if (i == jj) v += q0 * P.w[6 + i];
        } else if (r >= 3 && r < 6 && c == r + 6) v = q1 * dt * P.w[r];
        else if (c >= 3 && c < 6 && r == c + 6) v = q1 * dt * P.w[c];
        else if (r == c && r >= 9) v = q0 * P.w[r] + q2 * dt * dt * P.w[r - 6];
        Qc[e] = v;
      }
    }
    __syncwarp();

    // ---- IPM state of this lane's stance foot-stages -----------------------------------------------------------------
    const double hh[10] = {P.fmax, P.fmax, P.fmax, P.fmax, P.fmax, -P.fmin, 0.0, 0.0, 0.0, 0.0};
    double s[FS][10], z[FS][10], inv_s[FS][10];
#pragma unroll
    for (int q = 0; q < FS; ++q) {
      const int j = lane + 32 * q;
      const bool on = j < nfs;
      double g0[10];
      Gf(on ? sU[3 * j] : 0.0, on ? sU[3 * j + 1] : 0.0, on ? sU[3 * j + 2] : 0.0, mu_f, g0);
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        s[q][r] = fmax(hh[r] - g0[r], 1e-2);
        z[q][r] = (warm ? P.ipm_warm_mu : P.ipm_cold_mu) / s[q][r];
        inv_s[q][r] = 1.0;
      }
    }
    const double mtot = fmax(10.0 * (double)nfs, 1.0);
    int status = B200QP_ACADOS_MAXITER, iter = 0;
    double k_stat = 0.0, k_ineq = 0.0, k_comp = 0.0;
    const double target = 1e-4 * P.kkt_tol;
    double best = BIGP;
    int stall = 0;
    bool final_pass = false;
    double mu_prev = BIGP;
    int damp = 0;

    WPH(0);
    for (iter = 0; iter <= P.ipm_max_iter + 1; ++iter) {
      // (A) roll-out x(u) and exact adjoint -----------------------------------------------------------------------
      for (int k = lane; k < N; k += 32) {
        double dw[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
        for (int j = sSt[k]; j < sSt[k + 1]; ++j) {
          const double *Bw = sBwc + 9 * j;
          const double u0 = sU[3 * j], u1 = sU[3 * j + 1], u2 = sU[3 * j + 2];
          for (int r = 0; r < 3; ++r) dw[r] += Bw[3 * r] * u0 + Bw[3 * r + 1] * u1 + Bw[3 * r + 2] * u2;
          dv[0] += u0; dv[1] += u1; dv[2] += u2;
        }
        for (int r = 0; r < 3; ++r) {
          sBwsum[6 * k + r] = dw[r];
          sBwsum[6 * k + 3 + r] = dv[r] * dm;
        }
      }
      __syncwarp();
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
      __syncwarp();
      for (int e = lane; e < N * 12; e += 32) {
        const int k = e / 12 + 1, c = e - 12 * (k - 1);
        sQe[12 * k + c] = P.w[c] * (sX[12 * k + c] - __ldcg(gXref + 12 * k + c));
      }
      __syncwarp();
      if (lane < 6) {
        const int c = lane % 3;
        const bool lin = lane >= 3;
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
      __syncwarp();
      for (int k = lane + 1; k <= N; k += 32) {
        const double a0 = sQe[12 * k], a1 = sQe[12 * k + 1], a2 = sQe[12 * k + 2];
        for (int c = 0; c < 3; ++c) sLw[3 * k + c] += dt * (sRot[c] * a0 + sRot[3 + c] * a1 + sRot[6 + c] * a2);
      }
      __syncwarp();

      WPH(1);
      // (B) residuals ------------------------------------------------------------------------------------------------
      double rd[FS][3], rp[FS][10];
      double r0 = 0.0, r1 = 0.0, r2 = 0.0, gap = 0.0;
      bool bad = false;
#pragma unroll
      for (int q = 0; q < FS; ++q) {
        const int j = lane + 32 * q;
        if (j < nfs) {
          const int ks = sFs[j] >> 2;
          const double fx = sU[3 * j], fy = sU[3 * j + 1], fz = sU[3 * j + 2];
          double Gu[10], gz[3];
          Gf(fx, fy, fz, mu_f, Gu);
          GTv(z[q], mu_f, gz);
          const double *Bw = sBwc + 9 * j;
          const double *lw = sLw + 3 * (ks + 1), *lv = sLv + 3 * (ks + 1);
          const double uu[3] = {fx, fy, fz};
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            rd[q][c] = P.alpha * uu[c] + Bw[c] * lw[0] + Bw[3 + c] * lw[1] + Bw[6 + c] * lw[2] + dm * lv[c] + gz[c];
            r0 = fmax(r0, fabs(rd[q][c]));
            bad |= !(fabs(rd[q][c]) < BIGP);
          }
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            rp[q][r] = Gu[r] + s[q][r] - hh[r];
            r0 = fmax(r0, fabs(rp[q][r]));
            r1 = fmax(r1, s[q][r] * z[q][r]);
            r2 = fmax(r2, Gu[r] - hh[r]);
            gap += s[q][r] * z[q][r];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) rd[q][c] = 0.0;
#pragma unroll
          for (int r = 0; r < 10; ++r) rp[q][r] = 0.0;
        }
      }
      const bool anybad = __any_sync(0xffffffffu, bad);
      r0 = wred(r0, PMax());
      r1 = wred(r1, PMax());
      r2 = wred(r2, PMax());
      gap = wred(gap, PSum());
      const double rnow = anybad ? BIGP : fmax(r0, r1);
      if (final_pass) {
        status = (!anybad && rnow <= P.kkt_tol) ? B200QP_ACADOS_SUCCESS : B200QP_ACADOS_MINSTEP;
        k_stat = r0; k_comp = r1; k_ineq = fmax(r2, 0.0);
        break;
      }
      bool give_up = anybad || iter == P.ipm_max_iter;
      if (!anybad) {
        k_stat = r0; k_comp = r1; k_ineq = fmax(r2, 0.0);
        if ((r0 <= target && r1 <= target) || nfs == 0) {
          status = B200QP_ACADOS_SUCCESS;
          break;
        }
        if (rnow < 0.9 * best) {  // remember the best iterate (the end game can break down numerically): parked in the scratch
          best = rnow; stall = 0;
#pragma unroll
          for (int q = 0; q < FS; ++q)
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              Bk[(20 * q + r) * 32 + lane] = s[q][r];
              Bk[(20 * q + 10 + r) * 32 + lane] = z[q][r];
            }
          for (int e = lane; e < 3 * nfs; e += 32) Bk[20 * MAXFS + e] = sU[e];
        } else if (++stall >= 3 && best <= P.kkt_tol) {
          give_up = true;
        }
      }
      if (give_up) {
        if (best <= P.kkt_tol) {
          __syncwarp();
#pragma unroll
          for (int q = 0; q < FS; ++q)
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              s[q][r] = Bk[(20 * q + r) * 32 + lane];
              z[q][r] = Bk[(20 * q + 10 + r) * 32 + lane];
            }
          for (int e = lane; e < 3 * nfs; e += 32) sU[e] = Bk[20 * MAXFS + e];
          final_pass = true;
          __syncwarp();
          continue;
        }
        status = anybad ? B200QP_ACADOS_NAN_DETECTED : B200QP_ACADOS_MAXITER;
        break;
      }
      const double mu_c = gap / mtot;
      if ((P.ipm_flags & 2) && mu_c > 0.9 * mu_prev) damp = 2;  // the gap did not fall by 10 %: centre the next two steps (two-cycle guard, cf. wbc_qp.cu)
      mu_prev = mu_c;

      WPH(2);
      // (C) barrier weights per foot-stage, predictor right-hand side -------------------------------------------------------
      double dsdz[FS][10], sigma_mu = 0.0;
#pragma unroll
      for (int q = 0; q < FS; ++q) {
        const int j = lane + 32 * q;
#pragma unroll
        for (int r = 0; r < 10; ++r) dsdz[q][r] = 0.0;
        if (j < nfs) {
          double w[10], t10[10], rt[3];
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            inv_s[q][r] = fast_rcp(s[q][r]);
            w[r] = z[q][r] * inv_s[q][r];
            t10[r] = (z[q][r] * rp[q][r] - s[q][r] * z[q][r]) * inv_s[q][r];
          }
          double *Rb = sRb + 5 * j;
          Rb[0] = (w[0] + w[1]) + (w[6] + w[7]);
          Rb[1] = (w[2] + w[3]) + (w[8] + w[9]);
          Rb[2] = (w[4] + w[5]) + mu_f * mu_f * ((w[6] + w[7]) + (w[8] + w[9]));
          Rb[3] = -mu_f * (w[6] - w[7]);
          Rb[4] = -mu_f * (w[8] - w[9]);
          GTv(t10, mu_f, rt);
          sRt[3 * j] = rt[0] + rd[q][0];
          sRt[3 * j + 1] = rt[1] + rd[q][1];
          sRt[3 * j + 2] = rt[2] + rd[q][2];
        }
      }
      for (int e = lane; e < 144; e += 32) sP[e] = ((e / 12) == (e % 12)) ? P.w[e / 12] : 0.0;  // P_N = Q, p_N = 0
      if (lane < 12) sVec[lane] = 0.0;
      __syncwarp();
      WPH(9);

      // (D) block Riccati backward sweep: matrices + the predictor's vector sweep in one augmented factorisation per block ------
      for (int b = nb - 1; b >= 0; --b) {
        const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, j0 = sSt[s0], nu = 3 * (sSt[s0 + M] - j0);
        const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
        const double Mdt = (double)M * dt;
        const double *Rc = Cs + sCoff[b], *Sc = Rc + nu * nu, *Qc = Sc + 12 * nu;
        // geometry of the block's inputs: Bbar[:, a] = [kd rb ; kd dm e_c ; bw ; dm e_c],  kd = dt x (stages left after the input's own)
        if (lane < nu) {
          const int ja = j0 + lane / 3, cc = lane % 3;
          const double *Ba = sBwc + 9 * ja;
          const double b0 = Ba[cc], b1 = Ba[3 + cc], b2 = Ba[6 + cc];
          double *I = sIn + 8 * lane;
          I[0] = b0; I[1] = b1; I[2] = b2;
#pragma unroll
          for (int r = 0; r < 3; ++r) I[3 + r] = sRot[3 * r] * b0 + sRot[3 * r + 1] * b1 + sRot[3 * r + 2] * b2;
          I[6] = (double)(M - 1 - ((sFs[ja] >> 2) - s0)) * dt;
          sCc[lane] = (unsigned char)cc;
        }
        for (int e = lane; e < 144; e += 32) {  // PA = P Abar
          const int r = e / 12, c = e - 12 * r;
          double v = sP[e];
          if (c >= 6 && c < 9) v += Mdt * (sP[12 * r] * sRot[c - 6] + sP[12 * r + 1] * sRot[3 + c - 6] + sP[12 * r + 2] * sRot[6 + c - 6]);
          else if (c >= 9) v += Mdt * sP[12 * r + c - 6];
          sPA[e] = v;
        }
        __syncwarp();
        for (int e = lane; e < 12 * nu; e += 32) {  // PB = P Bbar, element-parallel
          const int r = e / nu, a = e - r * nu;
          const double *I = sIn + 8 * a, *Pr = sP + 12 * r;
          const int cc = sCc[a];
          sPB[r * NUBW + a] = I[6] * (Pr[0] * I[3] + Pr[1] * I[4] + Pr[2] * I[5] + dm * Pr[3 + cc]) +
                              (Pr[6] * I[0] + Pr[7] * I[1] + Pr[8] * I[2]) + dm * Pr[9 + cc];
        }
        __syncwarp();
        WPH(6);
        // augmented matrix [Huu | Hux' | hu ; . | Hxx | hx], lower triangle, every lane takes entries
        for (int e = lane; e < (nu * (nu + 1)) / 2; e += 32) {  // Huu[a][a2] = Rc + W + Bbar_a' (P Bbar)_a2
          const int a = sTriR[e], a2 = sTriC[e];
          const double *I = sIn + 8 * a;
          const int cc = sCc[a];
          double v = (M > 1) ? __ldcg(Rc + a2 * nu + a) : (a2 == a ? P.alpha : 0.0);
          v += I[6] * (I[3] * sPB[a2] + I[4] * sPB[NUBW + a2] + I[5] * sPB[2 * NUBW + a2] + dm * sPB[(3 + cc) * NUBW + a2]) +
               (I[0] * sPB[6 * NUBW + a2] + I[1] * sPB[7 * NUBW + a2] + I[2] * sPB[8 * NUBW + a2]) + dm * sPB[(9 + cc) * NUBW + a2];
          if (a2 / 3 == a / 3) {  // same foot-stage: barrier block
            const double *Rb = sRb + 5 * (j0 + a / 3);
            const int cb = a2 % 3;
            if (cc == cb) v += Rb[cc];
            else if (cc + cb == 2) v += Rb[3];
            else if (cc + cb == 3) v += Rb[4];
          }
          sK[a2 * LDW + a] = v;
        }
        {
          const double q0 = sRot[0] * sVec[0] + sRot[3] * sVec[1] + sRot[6] * sVec[2];  // Rt' p_theta
          const double q1 = sRot[1] * sVec[0] + sRot[4] * sVec[1] + sRot[7] * sVec[2];
          const double q2 = sRot[2] * sVec[0] + sRot[5] * sVec[1] + sRot[8] * sVec[2];
          for (int e = lane; e < 13 * nu; e += 32) {  // Hux[a][col] = Sc + Bbar_a' (P Abar)[:, col];  col 12: hu = r~ + Bbar' p
            const int a = e / 13, col = e - 13 * a;
            const double *I = sIn + 8 * a;
            const int cc = sCc[a];
            double v;
            if (col < 12)
              v = ((M > 1) ? __ldcg(Sc + 12 * a + col) : 0.0) +
                  I[6] * (I[3] * sPA[col] + I[4] * sPA[12 + col] + I[5] * sPA[24 + col] + dm * sPA[12 * (3 + cc) + col]) +
                  (I[0] * sPA[72 + col] + I[1] * sPA[84 + col] + I[2] * sPA[96 + col]) + dm * sPA[12 * (9 + cc) + col];
            else
              v = sRt[3 * j0 + a] + I[0] * (I[6] * q0 + sVec[6]) + I[1] * (I[6] * q1 + sVec[7]) + I[2] * (I[6] * q2 + sVec[8]) +
                  dm * (I[6] * sVec[3 + cc] + sVec[9 + cc]);
            sK[a * LDW + nu4 + col] = v;
          }
        }
        for (int e = lane; e < 78 + 12; e += 32) {  // Hxx = Qc + Abar' PA (lower triangle), hx = Abar' p
          if (e < 78) {
            const int r = sTriR[e], c = sTriC[e];
            double v = ((M > 1) ? __ldcg(Qc + 12 * r + c) : (r == c ? P.w[r] : 0.0)) + sPA[12 * r + c];
            if (r >= 6 && r < 9) v += Mdt * (sRot[r - 6] * sPA[c] + sRot[3 + r - 6] * sPA[12 + c] + sRot[6 + r - 6] * sPA[24 + c]);
            else if (r >= 9) v += Mdt * sPA[12 * (r - 6) + c];
            sK[(nu4 + c) * LDW + nu4 + r] = v;
          } else {
            const int c = e - 78;
            double v = sVec[c];
            if (c >= 6 && c < 9) v += Mdt * (sRot[c - 6] * sVec[0] + sRot[3 + c - 6] * sVec[1] + sRot[6 + c - 6] * sVec[2]);
            else if (c >= 9) v += Mdt * sVec[c - 6];
            sK[(nu4 + c) * LDW + nu4 + 12] = v;
          }
        }
        if (lane >= nu && lane < nu4) {  // identity padding
          for (int a2 = 0; a2 < lane; ++a2) sK[a2 * LDW + lane] = 0.0;
          sK[lane * LDW + lane] = 1.0;
          for (int r = nu4; r < na; ++r) sK[lane * LDW + r] = 0.0;
        }
        if (lane == 31) sK[(nu4 + 12) * LDW + nu4 + 12] = 1.0;
        __syncwarp();
        WPH(3);
        chol_aug(sK, LDW, nu4, na, sLst, sDst);
        WPH(4);
        double *F = Fac + sFoff[b];
        if (lane < na) {
          const int lim = lane < nu4 ? lane : nu4 - 1;
          for (int a = 0; a <= lim; ++a) F[fcol(a, na) + lane - a] = sK[a * LDW + lane];
        }
        for (int e = lane; e < 144; e += 32) {
          const int r = e / 12, c = e - 12 * r;
          sP[e] = (r >= c) ? sK[(nu4 + c) * LDW + nu4 + r] : sK[(nu4 + r) * LDW + nu4 + c];
        }
        if (lane < 12) sVec[lane] = sK[(nu4 + lane) * LDW + nu4 + 12];
        __syncwarp();
        WPH(5);
      }

      // (E)/(F) predictor forward sweep, then the corrector's backward + forward sweeps ------------------------------------
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
#pragma unroll
          for (int q = 0; q < FS; ++q) {
            const int j = lane + 32 * q;
            if (j < nfs) {
              double t10[10], rt[3];
#pragma unroll
              for (int r = 0; r < 10; ++r) t10[r] = (z[q][r] * rp[q][r] - (s[q][r] * z[q][r] + dsdz[q][r] - sigma_mu)) * inv_s[q][r];
              GTv(t10, mu_f, rt);
              sRt[3 * j] = rt[0] + rd[q][0];
              sRt[3 * j + 1] = rt[1] + rd[q][1];
              sRt[3 * j + 2] = rt[2] + rd[q][2];
            }
          }
          if (lane < 12) sVec[lane] = 0.0;
          __syncwarp();
          for (int b = nb - 1; b >= 0; --b) {  // y = L^-1 (r~ + Bbar' p),  p <- Abar' p - Y' y
            const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, j0 = sSt[s0], nu = 3 * (sSt[s0 + M] - j0);
            const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
            const double Mdt = (double)M * dt;
            const double *F = Fac + sFoff[b];
            double acc = 0.0;
            if (lane < nu) {
              const int ja = j0 + lane / 3, cc = lane % 3;
              const double kd = (double)(M - 1 - ((sFs[ja] >> 2) - s0)) * dt;
              const double *Ba = sBwc + 9 * ja;
              const double q0 = sRot[0] * sVec[0] + sRot[3] * sVec[1] + sRot[6] * sVec[2];
              const double q1 = sRot[1] * sVec[0] + sRot[4] * sVec[1] + sRot[7] * sVec[2];
              const double q2 = sRot[2] * sVec[0] + sRot[5] * sVec[1] + sRot[8] * sVec[2];
              acc = sRt[3 * j0 + lane] + Ba[cc] * (kd * q0 + sVec[6]) + Ba[3 + cc] * (kd * q1 + sVec[7]) + Ba[6 + cc] * (kd * q2 + sVec[8]) +
                    dm * (kd * sVec[3 + cc] + sVec[9 + cc]);
            } else if (lane >= nu4 && lane < nu4 + 12) {
              const int c = lane - nu4;
              acc = sVec[c];
              if (c >= 6 && c < 9) acc += Mdt * (sRot[c - 6] * sVec[0] + sRot[3 + c - 6] * sVec[1] + sRot[6 + c - 6] * sVec[2]);
              else if (c >= 9) acc += Mdt * sVec[c - 6];
            }
            const double yv = tri_fwd(F, nu4, na, nu4 + 12, acc, sYst);
            if (lane < nu) sYv[3 * j0 + lane] = yv;
            if (lane >= nu4 && lane < nu4 + 12) sVec[lane - nu4] = yv;
            __syncwarp();
          }
        }
        // forward sweep: du = -L^-T (Y dx + y),  dx <- Abar dx + Bbar du
        if (lane < 12) sVec[lane] = 0.0;
        __syncwarp();
        for (int b = 0; b < nb; ++b) {
          const int s0 = sBlkS[b], M = sBlkS[b + 1] - s0, j0 = sSt[s0], nu = 3 * (sSt[s0 + M] - j0);
          const int nu4 = (nu + 3) & ~3, na = nu4 + 13;
          const double Mdt = (double)M * dt;
          const double *F = Fac + sFoff[b];
          double acc = 0.0;
          if (lane < nu) {
            const double *Ya = F + fcol(lane, na) + nu4 - lane;  // Y[a][0..11], then the predictor's y_a
            acc = (pass == 0) ? Ya[12] : sYv[3 * j0 + lane];
#pragma unroll
            for (int c = 0; c < 12; ++c) acc = fma(Ya[c], sVec[c], acc);
          }
          const double du = -tri_bwd(F, nu4, na, acc, sYst);
          if (lane < nu) {  // publish du and this input's column of Bbar du (8 non-zeros) for the 12 state lanes
            const int ja = j0 + lane / 3, cc = lane % 3;
            const double kd = (double)(M - 1 - ((sFs[ja] >> 2) - s0)) * dt;
            const double *Ba = sBwc + 9 * ja;
            const double b0 = Ba[cc], b1 = Ba[3 + cc], b2 = Ba[6 + cc];
            sDU[3 * j0 + lane] = du;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              sPB[r * NUBW + lane] = kd * du * (sRot[3 * r] * b0 + sRot[3 * r + 1] * b1 + sRot[3 * r + 2] * b2);
              sPB[(3 + r) * NUBW + lane] = (r == cc) ? kd * dm * du : 0.0;
              sPB[(9 + r) * NUBW + lane] = (r == cc) ? dm * du : 0.0;
            }
            sPB[6 * NUBW + lane] = b0 * du;
            sPB[7 * NUBW + lane] = b1 * du;
            sPB[8 * NUBW + lane] = b2 * du;
          }
          __syncwarp();
          double nx = 0.0;
          if (lane < 12) {
            nx = sVec[lane];
            if (lane < 3) nx += Mdt * (sRot[3 * lane] * sVec[6] + sRot[3 * lane + 1] * sVec[7] + sRot[3 * lane + 2] * sVec[8]);
            else if (lane < 6) nx += Mdt * sVec[lane + 6];
            double a0_ = 0.0, a1_ = 0.0;
            int a = 0;
            for (; a + 1 < nu; a += 2) {
              a0_ += sPB[lane * NUBW + a];
              a1_ += sPB[lane * NUBW + a + 1];
            }
            if (a < nu) a0_ += sPB[lane * NUBW + a];
            nx += a0_ + a1_;
          }
          __syncwarp();
          if (lane < 12) sVec[lane] = nx;
          __syncwarp();
        }
        WPH(7);
        // step in (s, z)
        double stp = 1.0, stq = 1.0;  // step to the boundary in s (primal) and in z (dual)
        double dsn[FS][10], dzn[FS][10];
#pragma unroll
        for (int q = 0; q < FS; ++q) {
          const int j = lane + 32 * q;
          if (j < nfs) {
            double Gd[10], tmax = 0.0, tmaz = 0.0;
            Gf(sDU[3 * j], sDU[3 * j + 1], sDU[3 * j + 2], mu_f, Gd);
#pragma unroll
            for (int r = 0; r < 10; ++r) {
              const double rc = (pass == 0) ? s[q][r] * z[q][r] : (s[q][r] * z[q][r] + dsdz[q][r] - sigma_mu);
              dsn[q][r] = -rp[q][r] - Gd[r];
              dzn[q][r] = (-rc - z[q][r] * dsn[q][r]) * inv_s[q][r];
              tmax = fmax(tmax, -dsn[q][r] * inv_s[q][r]);
              tmaz = fmax(tmaz, -dzn[q][r] * fast_rcp(z[q][r]));
            }
            if (tmax > 1.0) stp = fmin(stp, 1.0 / tmax);
            if (tmaz > 1.0) stq = fmin(stq, 1.0 / tmaz);
          } else {
#pragma unroll
            for (int r = 0; r < 10; ++r) dsn[q][r] = dzn[q][r] = 0.0;
          }
        }
        stp = wred(stp, PMin());
        stq = wred(stq, PMin());
        if (!(P.ipm_flags & 1)) stp = stq = fmin(stp, stq);  // one step length for the whole iterate
        if (pass == 0) {
          double g2 = 0.0;
#pragma unroll
          for (int q = 0; q < FS; ++q)
            if (lane + 32 * q < nfs) {
#pragma unroll
              for (int r = 0; r < 10; ++r) {
                g2 += (s[q][r] + stp * dsn[q][r]) * (z[q][r] + stq * dzn[q][r]);
                dsdz[q][r] = dsn[q][r] * dzn[q][r];
              }
            }
          g2 = wred(g2, PSum());
          const double mu_aff = g2 / mtot;
          const double ratio = mu_aff / fmax(mu_c, 1e-300);
          sigma_mu = ratio * ratio * ratio * mu_c;
          if (damp > 0) sigma_mu = fmax(sigma_mu, 0.2 * mu_c);
        } else {
          double a = fmin(1.0, P.ipm_tau * stp), az = fmin(1.0, P.ipm_tau * stq);
          if (damp > 0) --damp;
          if (r0 < r1) {  // monotone-gap safeguard (see mpc_riccati.cu)
            double c1 = 0.0, c2 = 0.0;
#pragma unroll
            for (int q = 0; q < FS; ++q)
              if (lane + 32 * q < nfs) {
#pragma unroll
                for (int r = 0; r < 10; ++r) {
                  c1 += s[q][r] * dzn[q][r] + z[q][r] * dsn[q][r];
                  c2 += dsn[q][r] * dzn[q][r];
                }
              }
            c1 = wred(c1, PSum());
            c2 = wred(c2, PSum());
            const double lin = c1 + 0.1 * gap;
            if (c2 > 0.0 && lin < 0.0) {
              a = fmin(a, -lin / c2);
              az = fmin(az, -lin / c2);
            }
          }
#pragma unroll
          for (int q = 0; q < FS; ++q) {
            const int j = lane + 32 * q;
            if (j < nfs) {
#pragma unroll
              for (int r = 0; r < 10; ++r) {
                s[q][r] += a * dsn[q][r];
                z[q][r] += az * dzn[q][r];
              }
              sU[3 * j] += a * sDU[3 * j];
              sU[3 * j + 1] += a * sDU[3 * j + 1];
              sU[3 * j + 2] += a * sDU[3 * j + 2];
            }
          }
        }
        __syncwarp();
      }
      WPH(8);
    }

    if (timing && lane == 0)
      for (int q = 0; q < 10; ++q) dbg_cycles[(size_t)prob * 10 + q] = cyc[q];
    // ---- outputs ------------------------------------------------------------------------------------------------------
    __syncwarp();
    if (status == B200QP_ACADOS_SUCCESS) {
      bool nanout = false;
      for (int e = lane; e < (N + 1) * 12; e += 32) nanout |= !(fabs(sX[e]) < BIGP);
      if (__any_sync(0xffffffffu, nanout)) status = B200QP_NAN_AFTER_SUCCESS;
    }
    if (status == B200QP_ACADOS_SUCCESS) {
      double *fo = forces + (size_t)prob * N * 12;
      for (int e = lane; e < N * 12; e += 32) fo[e] = 0.0;
      __syncwarp();
      for (int j = lane; j < nfs; j += 32) {
        double *f3 = fo + 3 * sFs[j];
        f3[0] = sU[3 * j]; f3[1] = sU[3 * j + 1]; f3[2] = sU[3 * j + 2];
      }
      if (P.ipm_warm_x != nullptr) {  // remember this solution for the next call on the same slot
        double *wo = P.ipm_warm_x + (size_t)prob * N * 12;
        for (int e = lane; e < N * 12; e += 32) wo[e] = 0.0;
        __syncwarp();
        for (int j = lane; j < nfs; j += 32) {
          double *f3 = wo + 3 * sFs[j];
          f3[0] = sU[3 * j]; f3[1] = sU[3 * j + 1]; f3[2] = sU[3 * j + 2];
        }
      }
      double *xo = xpred + (size_t)prob * (N + 1) * 13;
      for (int e = lane; e < (N + 1) * 13; e += 32) {
        const int k = e / 13, c = e - 13 * k;
        xo[e] = (c < 12) ? sX[12 * k + c] : sX0[12];
      }
      if (P.ipm_refine_x != nullptr && fmax(k_stat, k_comp) > P.ipm_refine_thr) {  // hand the answer to the second opinion (hot start)
        double *rx = P.ipm_refine_x + (size_t)prob * N * 12, *ry = P.ipm_refine_y + (size_t)prob * N * 28;
        for (int e = lane; e < N * 12; e += 32) rx[e] = 0.0;
        for (int e = lane; e < N * 28; e += 32) ry[e] = 0.0;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < FS; ++q) {
          const int j = lane + 32 * q;
          if (j < nfs) {
            double *f3 = rx + 3 * sFs[j], *y7 = ry + 7 * sFs[j];
            f3[0] = sU[3 * j]; f3[1] = sU[3 * j + 1]; f3[2] = sU[3 * j + 2];
            y7[0] = z[q][0] - z[q][1]; y7[1] = z[q][2] - z[q][3]; y7[2] = z[q][4] - z[q][5];
            y7[3] = z[q][6]; y7[4] = -z[q][7]; y7[5] = z[q][8]; y7[6] = -z[q][9];
          }
        }
        if (lane == 0) P.ipm_refine_valid[prob] = 1;
      }
      if (lam_box != nullptr) {
        double *lb = lam_box + (size_t)prob * N * 12, *lg = lam_g + (size_t)prob * N * 16;
        for (int e = lane; e < N * 12; e += 32) lb[e] = 0.0;
        for (int e = lane; e < N * 16; e += 32) lg[e] = 0.0;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < FS; ++q) {
          const int j = lane + 32 * q;
          if (j < nfs) {
            double *l3 = lb + 3 * sFs[j], *l4 = lg + 4 * sFs[j];
            l3[0] = z[q][0] - z[q][1]; l3[1] = z[q][2] - z[q][3]; l3[2] = z[q][4] - z[q][5];
            l4[0] = z[q][6]; l4[1] = -z[q][7]; l4[2] = z[q][8]; l4[3] = -z[q][9];
          }
        }
      }
    }
    if (lane == 0) {
      if (P.ipm_warm_x != nullptr) P.ipm_warm_valid[prob] = (status == B200QP_ACADOS_SUCCESS) ? 1 : 0;
      // solved, but only by the stall rule (residual between the target and the tolerance): these carry the GRF error tail of the
      // interior-point path (degenerate vertices, cost flat along directions only alpha penalises) - ask the active-set kernel
      // ... and so do the problems it gave up on (iteration cap, step collapse): a failure then needs both kernels to fail
      if (P.ipm_refine_valid != nullptr && status != B200QP_ACADOS_SUCCESS) P.ipm_refine_valid[prob] = 0;  // gave up: the second kernel starts cold
      if (P.ipm_refine_list != nullptr && ((status == B200QP_ACADOS_SUCCESS && fmax(k_stat, k_comp) > P.ipm_refine_thr) ||
                                           status == B200QP_ACADOS_MAXITER || status == B200QP_ACADOS_MINSTEP))
        P.ipm_refine_list[atomicAdd(P.ipm_refine_count, 1)] = prob;
      b200qp_solver_info inf;
      inf.status = status;
      inf.iterations = iter;
      inf.polish_rounds = 0;
      inf.n_free = 3 * nfs;
      inf.kkt[0] = k_stat;
      inf.kkt[1] = 0.0;
      inf.kkt[2] = k_ineq;
      inf.kkt[3] = k_comp;
      info[prob] = inf;
    }
  }
}

// development aid (B200QP_POISON_SMEM=1): leaves NaN patterns in the shared memory of every SM, so that a read of shared memory the
// kernel has not written shows up as a changed result instead of depending on which kernel ran before
__global__ void smem_polluter_kernel(int nd) {
  extern __shared__ __align__(16) double pdsm[];
  for (int e = threadIdx.x; e < nd; e += blockDim.x) pdsm[e] = __longlong_as_double(0xFFF8000000000000ULL | (unsigned long long)e);
  __syncthreads();
  if (pdsm[(threadIdx.x * 7) % nd] == 1.0) pdsm[0] = 0.0;  // keep the stores
}
cudaError_t launch_smem_polluter(int num_sms, cudaStream_t st) {
  const int bytes = 200 * 1024;
  cudaError_t e = cudaFuncSetAttribute(smem_polluter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  smem_polluter_kernel<<<num_sms * 2, 256, bytes, st>>>(bytes / 8);
  return cudaGetLastError();
}

// ---- host side --------------------------------------------------------------------------------------------------------
size_t pcw_scratch_doubles(int N, int fs);
static size_t pcw_smem(int fcap) { return sizeof(double) * (size_t)(((NAW * LDW + 1) & ~1) + 12 * NUBW + fcap); }
// packed factors of a problem: sum over blocks of nu4 (nu4 + 13) - nu4 (nu4 - 1) / 2.  Shared-memory budget: fs stance foot-stages in
// single-stage blocks of two feet (nu = 6 -> nu4 = 8: 140 doubles) or two-stage blocks (nu = 12: 234 doubles per 4 foot-stages)
static int pcw_fcap(int fs, int N, int condN) {
  // shared-memory budget for the packed factors: trot-like schedules (two stance feet per stage), blocks of M = N / cond_N
  // stages capped at 15 stacked inputs; problems that need more keep their factors in the L2-resident scratch
  const int cN = condN < 1 ? 1 : (condN > N ? N : condN);
  int M = (N + cN - 1) / cN;
  if (M > 2) M = 2;
  const int nu4 = M == 1 ? 8 : 12, nblk = (N + M - 1) / M;
  int c = nblk * (nu4 * (nu4 + 13) - (nu4 * (nu4 - 1)) / 2) + 64;
  if (fs > 1) c = c * 3 / 2;
  return c > 3200 ? 3200 : c;
}
// the L2-persisting part of the scratch for a launch with `grid` blocks: [*ptr, *ptr + *bytes)
void pcw_hot_region(int N, int fs, int grid, double *scratch, double **ptr, size_t *bytes) {
  const size_t hot = 23 * (size_t)(32 * fs) + 21 * 12 + 14 * (size_t)(32 * fs) + 12 * 22 + 64 + 12 * 21 + 9 * (size_t)(32 * fs);
  const size_t front = pcw_scratch_doubles(N, fs) - hot;
  *ptr = scratch + (size_t)grid * front;
  *bytes = (size_t)grid * hot * sizeof(double);
}
size_t pcw_scratch_doubles(int N, int fs) {
  const size_t half = (size_t)16 * 15 * (size_t)N + 288 * (size_t)N + 1024;  // block constants / factors that do not fit shared memory
  return 2 * half + 23 * (size_t)(32 * fs) + 21 * 12 + 14 * (size_t)(32 * fs) + 12 * 22 + 64 + 12 * 21 + 9 * (size_t)(32 * fs);
}

template <int NH, int FS>
static cudaError_t pcw_launch_t(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                                b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, int *next, const int *list,
                                const int *list_count, int *rej, int *rej_count, int condN, int grid, cudaStream_t st, long long *dbg) {
  const int fcap = pcw_fcap(FS, P.N, condN);
  const size_t sm = pcw_smem(fcap);
  if (grid > n) grid = n;
  if (dbg != nullptr) {  // phase accounting (scripts/phase_cycles_pcw.py): the instrumented instantiation
    cudaError_t e = cudaFuncSetAttribute(mpc_pcw_kernel<NH, FS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
    mpc_pcw_kernel<NH, FS, true><<<grid, 32, sm, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, pcw_scratch_doubles(P.N, FS), next,
                                               list, list_count, rej, rej_count, condN, fcap, dbg);
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(mpc_pcw_kernel<NH, FS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  mpc_pcw_kernel<NH, FS, false><<<grid, 32, sm, st>>>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, pcw_scratch_doubles(P.N, FS), next,
                                               list, list_count, rej, rej_count, condN, fcap, dbg);
  return cudaGetLastError();
}
template <int NH, int FS>
static int pcw_occ_t(int N, int condN) {
  const size_t sm = pcw_smem(pcw_fcap(FS, N, condN));
  int occ = 0;
  if (cudaFuncSetAttribute(mpc_pcw_kernel<NH, FS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mpc_pcw_kernel<NH, FS, false>, 32, sm) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return occ;
}
// fs = 1: problems with <= 32 stance foot-stages (trot / pace / bound up to N = 16), fs = 2: <= 64 (every gait up to N = 16,
// trot-like gaits at N = 20)
int pcw_grid(int num_sms, int N, int fs, int condN) {
  int occ;
  if (N <= 16) occ = fs == 1 ? pcw_occ_t<16, 1>(N, condN) : pcw_occ_t<16, 2>(N, condN);
  else occ = fs == 1 ? pcw_occ_t<20, 1>(N, condN) : pcw_occ_t<20, 2>(N, condN);
  return num_sms * (occ < 1 ? 1 : occ);
}
cudaError_t launch_mpc_pcw(const MpcParams &P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                           b200qp_solver_info *info, double *lam_box, double *lam_g, double *scratch, int *next, int grid, cudaStream_t st,
                           const int *list, const int *list_count, int *rej, int *rej_count, int condN, int fs, long long *dbg) {
  if (P.N <= 16) {
    if (fs == 1) return pcw_launch_t<16, 1>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, list, list_count, rej, rej_count, condN, grid, st, dbg);
    return pcw_launch_t<16, 2>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, list, list_count, rej, rej_count, condN, grid, st, dbg);
  }
  if (fs == 1) return pcw_launch_t<20, 1>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, list, list_count, rej, rej_count, condN, grid, st, dbg);
  return pcw_launch_t<20, 2>(P, n, in, contact, forces, xpred, info, lam_box, lam_g, scratch, next, list, list_count, rej, rej_count, condN, grid, st, dbg);
}

}  // namespace b200qp
