// Kernels K1 + K2a, fused: SRBD discretisation, condensed (stance-only) Hessian/gradient assembly, OSQP-style ADMM
// with an explicit K^-1 in shared memory, null-space active-set polish, roll-out of the predicted states.
// One thread block per problem, one thread per free force variable (FP64 throughout).
//
// Replaces, per problem: MPC::GetWrenchSequence (ws/src/controllers/src/mit_controller/mpc.cpp:329-470) =
//   GetA/GetB/GetIWorld/GetUBounds (mpc.cpp:11-115) + acados condensing + OSQP behind ocp_qp_solve (mpc.cpp:410)
//   + unpack of u_k / x_k (mpc.cpp:457-467).
//
// Math.  With M = A - I nilpotent the condensed Hessian has the closed form
//   H[(j,f),(l,f')] = Bw_jf^T (cnt Qw + S2 dt^2 Rt^T Qth Rt) Bw_lf' + (dt/m)^2 diag(cnt Qv + S2 dt^2 Qp) + alpha I
//   cnt = N - max(j,l),  S2 = sum_{t=max(j,l)}^{N-1} (t-j)(t-l),  Bw_kf = dt Iw_k^-1 skew(r_kf)
// and g = B^T lambda with lambda the adjoint of the free response; swing-leg forces (pinned to 0 by
// lbu = ubu = 0, mpc.cpp:110-113) are eliminated, which leaves the optimum unchanged.
#include "common.cuh"

#ifndef B200QP_TC_GJ
#define B200QP_TC_GJ 0
#endif
#ifndef B200QP_CHOL_POLISH
#define B200QP_CHOL_POLISH 0
#endif
namespace b200qp {

// optional per-phase cycle accounting (B.dbg_cycles != nullptr): thread 0 accumulates clock64() deltas per phase
#define PH(idx)                                  \
  do {                                           \
    if (timing && tid == 0) {                    \
      const long long t__ = clock64();           \
      cyc[idx] += t__ - tlast;                   \
      tlast = t__;                               \
    }                                            \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// per-foot constraint rows (reduced problem), in the order
//   0: fx in [-fmax, fmax]   1: fy in [-fmax, fmax]   2: fz in [fmin, fmax]            (mpc.cpp:103-109)
//   3: fx - mu fz <= 0       4: fx + mu fz >= 0       5: fy - mu fz <= 0   6: fy + mu fz >= 0   (mpc.cpp:66-92)
// rows 3..6 are the reference's D rows 0..3 of that foot.
constexpr double BIG = 1.0e300;

// FP32 error study (b200qp_mpc_config.precision): every stored intermediate is rounded to binary32, so the data flow has the
// rounding behaviour of an FP32 build of the same algorithm (storage and every operation result in FP32; the four-term FMA
// chains accumulate in FP64 before their single rounding, i.e. marginally better than a pure FP32 chain).
template <bool F32>
__device__ __forceinline__ double rnd(double v) {
  return F32 ? (double)(float)v : v;
}

__device__ __forceinline__ void foot_rows_eval(double fx, double fy, double fz, double mu, double (&ax)[7]) {
  ax[0] = fx;
  ax[1] = fy;
  ax[2] = fz;
  ax[3] = fx - mu * fz;
  ax[4] = fx + mu * fz;
  ax[5] = fy - mu * fz;
  ax[6] = fy + mu * fz;
}
__device__ __forceinline__ void foot_row_vec(int r, double mu, double (&a)[3]) {
  a[0] = (r == 0 || r == 3 || r == 4) ? 1.0 : 0.0;
  a[1] = (r == 1 || r == 5 || r == 6) ? 1.0 : 0.0;
  a[2] = (r == 2) ? 1.0 : ((r == 3 || r == 5) ? -mu : ((r == 4 || r == 6) ? mu : 0.0));
}
__device__ __forceinline__ void foot_bounds(double fmin, double fmax, double (&lo)[7], double (&hi)[7]) {
  lo[0] = -fmax; hi[0] = fmax;
  lo[1] = -fmax; hi[1] = fmax;
  lo[2] = fmin;  hi[2] = fmax;
  lo[3] = -BIG;  hi[3] = 0.0;
  lo[4] = 0.0;   hi[4] = BIG;
  lo[5] = -BIG;  hi[5] = 0.0;
  lo[6] = 0.0;   hi[6] = BIG;
}
// (A^T w)_c for the variable with component c of a foot
__device__ __forceinline__ double foot_at(int c, const double (&w)[7], double mu) {
  return (c == 0) ? (w[0] + w[3] + w[4]) : (c == 1) ? (w[1] + w[5] + w[6]) : (w[2] + mu * ((w[4] - w[3]) + (w[6] - w[5])));
}

// Null space of the active rows of one foot in closed form (register-only; a generic 7x4 rref through local memory
// cost 20 % of the kernel): returns nz = 3 - rank, particular solution cf and basis Z (column q at Z[3q..3q+2]).
// Rows: 0/1/2 boxes on x/y/z, 3: x - mu z = 0, 4: x + mu z = 0, 5: y - mu z = 0, 6: y + mu z = 0.  Inconsistent
// combinations (only produced by a bad active-set guess) are resolved by priority: z box, cone apex, x box + cone.
__device__ __forceinline__ int foot_nullspace(unsigned actl, unsigned actu, double mu, const double (&lo)[7],
                                              const double (&hi)[7], double *cf, double *Z) {  // cf, Z: shared memory
  const unsigned act = actl | actu;
  const bool bx = act & 1u, by = act & 2u, bz = act & 4u;
  const double vx = (actl & 1u) ? lo[0] : hi[0], vy = (actl & 2u) ? lo[1] : hi[1], vz = (actl & 4u) ? lo[2] : hi[2];
  const bool x3 = act & 8u, x4 = act & 16u, y5 = act & 32u, y6 = act & 64u;
  const double imu = (mu > 0.0) ? 1.0 / mu : 0.0;
  // slope of the tie x = sx mu z (0 when not tied to z by exactly one cone row)
  const double sx = (x3 && !x4) ? 1.0 : ((x4 && !x3) ? -1.0 : 0.0);
  const double sy = (y5 && !y6) ? 1.0 : ((y6 && !y5) ? -1.0 : 0.0);
  bool zfix = bz;
  double zval = vz;
  if (!zfix && ((x3 && x4) || (y5 && y6))) { zfix = true; zval = 0.0; }          // apex: x (or y) = 0 and z = 0
  if (!zfix && bx && sx != 0.0 && mu > 0.0) { zfix = true; zval = sx * vx * imu; }  // x box and one x-cone row pin z
  if (!zfix && by && sy != 0.0 && mu > 0.0) { zfix = true; zval = sy * vy * imu; }
  int nz = 0;
  // x
  double xc = 0.0, xt = 0.0;  // x = xc + xt * z_free
  bool xfree = false;
  if (x3 && x4) xc = 0.0;
  else if (bx) xc = vx;
  else if (sx != 0.0) { if (zfix) xc = sx * mu * zval; else xt = sx * mu; }
  else xfree = true;
  double yc = 0.0, yt = 0.0;
  bool yfree = false;
  if (y5 && y6) yc = 0.0;
  else if (by) yc = vy;
  else if (sy != 0.0) { if (zfix) yc = sy * mu * zval; else yt = sy * mu; }
  else yfree = true;
  cf[0] = xc; cf[1] = yc; cf[2] = zfix ? zval : 0.0;
  if (xfree) { Z[3 * nz] = 1.0; Z[3 * nz + 1] = 0.0; Z[3 * nz + 2] = 0.0; ++nz; }
  if (yfree) { Z[3 * nz] = 0.0; Z[3 * nz + 1] = 1.0; Z[3 * nz + 2] = 0.0; ++nz; }
  if (!zfix) { Z[3 * nz] = xt; Z[3 * nz + 1] = yt; Z[3 * nz + 2] = 1.0; ++nz; }
  return nz;
}

// Least squares on <= 3 columns (normal equations, Gaussian elimination with partial pivoting).
__device__ __noinline__ bool ls_small(const double (*M)[3], const int *idx, int k, const double (&b)[3], double *s) {
  double G[3][4];
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j < k; ++j)
      G[i][j] = M[idx[i]][0] * M[idx[j]][0] + M[idx[i]][1] * M[idx[j]][1] + M[idx[i]][2] * M[idx[j]][2];
    G[i][3] = M[idx[i]][0] * b[0] + M[idx[i]][1] * b[1] + M[idx[i]][2] * b[2];
  }
  for (int col = 0; col < k; ++col) {
    int p = col;
    for (int q = col + 1; q < k; ++q)
      if (fabs(G[q][col]) > fabs(G[p][col])) p = q;
    if (fabs(G[p][col]) < 1e-14) return false;
    if (p != col)
      for (int t = 0; t < 4; ++t) {
        const double tmp = G[col][t];
        G[col][t] = G[p][t];
        G[p][t] = tmp;
      }
    const double inv = 1.0 / G[col][col];
    for (int q = 0; q < k; ++q)
      if (q != col) {
        const double f = G[q][col] * inv;
        for (int t = col; t < 4; ++t) G[q][t] -= f * G[col][t];
      }
  }
  for (int i = 0; i < k; ++i) s[i] = G[i][3] / G[i][i];
  return true;
}

// Lawson-Hanson NNLS for min || M c - b ||, c >= 0, M = [3 x m] given as m columns.  Returns residual b - M c.
__device__ __noinline__ void nnls3(const double (*M)[3], int m, const double (&b)[3], double *coef, double (&res)[3]) {
  bool P[14], excl[14];
  for (int j = 0; j < m; ++j) {
    coef[j] = 0.0;
    P[j] = false;
    excl[j] = false;
  }
  const double bn = fmax(fmax(fabs(b[0]), fabs(b[1])), fabs(b[2]));
  for (int outer = 0; outer < 3 * m + 6; ++outer) {
    res[0] = b[0]; res[1] = b[1]; res[2] = b[2];
    for (int j = 0; j < m; ++j)
      if (coef[j] != 0.0) {
        res[0] -= coef[j] * M[j][0];
        res[1] -= coef[j] * M[j][1];
        res[2] -= coef[j] * M[j][2];
      }
    int js = -1;
    double wbest = 1e-13 * (bn + 1e-300);
    int np = 0;
    for (int j = 0; j < m; ++j) {
      if (P[j]) {
        ++np;
        continue;
      }
      if (excl[j]) continue;
      const double w = M[j][0] * res[0] + M[j][1] * res[1] + M[j][2] * res[2];
      if (w > wbest) {
        wbest = w;
        js = j;
      }
    }
    if (js < 0 || np >= 3) break;
    P[js] = true;
    for (int inner = 0; inner < 8; ++inner) {
      int idx[3];
      int k = 0;
      for (int j = 0; j < m && k < 3; ++j)
        if (P[j]) idx[k++] = j;
      double s[3];
      if (k == 0) break;
      if (!ls_small(M, idx, k, b, s)) {  // dependent column: drop the newest one and stop
        P[js] = false;
        excl[js] = true;
        break;
      }
      bool allpos = true;
      for (int i = 0; i < k; ++i) allpos &= (s[i] > 0.0);
      if (allpos) {
        for (int i = 0; i < k; ++i) coef[idx[i]] = s[i];
        break;
      }
      double al = 1.0;
      for (int i = 0; i < k; ++i)
        if (s[i] <= 0.0) {
          const double d = coef[idx[i]] - s[i];
          if (d > 0.0) al = fmin(al, coef[idx[i]] / d);
          else al = 0.0;
        }
      for (int i = 0; i < k; ++i) {
        coef[idx[i]] += al * (s[i] - coef[idx[i]]);
        if (coef[idx[i]] <= 1e-300 || (s[i] <= 0.0 && coef[idx[i]] <= 1e-14 * (bn + 1.0))) {
          coef[idx[i]] = 0.0;
          P[idx[i]] = false;
        }
      }
    }
  }
  res[0] = b[0]; res[1] = b[1]; res[2] = b[2];
  for (int j = 0; j < m; ++j)
    if (coef[j] != 0.0) {
      res[0] -= coef[j] * M[j][0];
      res[1] -= coef[j] * M[j][1];
      res[2] -= coef[j] * M[j][2];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// dense helpers on the shared-memory matrix K (column-major, odd leading dimension LD => row and column
// accesses across a warp are both bank-conflict free for 8-byte words)

// In-place Gauss-Jordan inverse of an SPD matrix (no pivoting), thread i owns row i.
// Blocked: 4 pivots per step.  With D = A[P,P], the block step is
//     A[P,P] <- D^-1,  A[P,R] <- D^-1 A[P,R],  A[R,P] <- -A[R,P] D^-1,  A[R,R] <- A[R,R] - A[R,P] D^-1 A[P,R]
// so every row element is read and written once per FOUR pivots, one reciprocal pair and two barriers per four pivots
// (the scalar sweep was bound by exactly that per-pivot latency).  The old pivot rows are staged in prow4[j][4]; the new
// pivot rows are produced column-wise (thread j writes A[P, j]) so that no thread runs a long divergent branch.
// The last n mod 4 pivots use the scalar step.
#ifndef B200QP_GJ_SYM
#define B200QP_GJ_SYM 0
#endif
// Symmetric variant of the sweep below (-DB200QP_GJ_SYM=1; correct - the whole GPU suite passes with it - but measured SLOWER, 3.04 vs
// 3.19 M solves/s: halving the multiply-adds and the shared-memory traffic does not shorten a block step, whose length is set by the
// longest line (line n-1 still has n entries) and by the serial prelude): between block steps the matrix satisfies M(a,b) = M(b,a) when a and b are both
// already pivoted or both not, and M(a,b) = -M(b,a) when exactly one of them is (A[R,P] carries the minus sign of the sweep).  So only
// the lower triangle (b <= a) is kept and updated - thread i owns M(i, 0..i), contiguous at K[i * LD] - and an element of the upper
// triangle is taken from its mirror with that sign: half the multiply-adds and half the shared-memory traffic of the sweep.  The full
// symmetric inverse is written out at the end.
template <int LD, bool F32 = false>
__device__ __noinline__ void gj_invert_sym(double *K, int n, double *prow4) {
  static_assert((LD % 2) == 0, "128-bit accesses to a thread's run");
  const int i = threadIdx.x;
  double *Ki = K + (size_t)i * LD;
  int p0 = 0;
  for (; p0 + 3 < n; p0 += 4) {
    double st[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < n) {  // stage M(p0+q, i): stored in the pivot line when i <= p0+q, else the mirror in this thread's own line (both unpivoted: +)
#pragma unroll
      for (int q = 0; q < 4; ++q) st[q] = (i <= p0 + q) ? K[(p0 + q) * LD + i] : Ki[p0 + q];
      double2 a, b;
      a.x = st[0]; a.y = st[1]; b.x = st[2]; b.y = st[3];
      *reinterpret_cast<double2 *>(prow4 + 4 * i) = a;
      *reinterpret_cast<double2 *>(prow4 + 4 * i + 2) = b;
    }
    __syncthreads();
    if (i < n) {
      double D[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 a = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c));
        const double2 b = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c) + 2);
        D[0][c] = a.x; D[1][c] = a.y; D[2][c] = b.x; D[3][c] = b.y;
      }
      const double r1 = fast_rcp(D[0][0] * D[1][1] - D[0][1] * D[1][0]);
      const double i00 = D[1][1] * r1, i01 = -D[0][1] * r1, i10 = -D[1][0] * r1, i11 = D[0][0] * r1;
      const double t00 = D[2][0] * i00 + D[2][1] * i10, t01 = D[2][0] * i01 + D[2][1] * i11;
      const double t10 = D[3][0] * i00 + D[3][1] * i10, t11 = D[3][0] * i01 + D[3][1] * i11;
      const double s00 = D[2][2] - (t00 * D[0][2] + t01 * D[1][2]), s01 = D[2][3] - (t00 * D[0][3] + t01 * D[1][3]);
      const double s10 = D[3][2] - (t10 * D[0][2] + t11 * D[1][2]), s11 = D[3][3] - (t10 * D[0][3] + t11 * D[1][3]);
      const double r2 = fast_rcp(s00 * s11 - s01 * s10);
      const double j00 = s11 * r2, j01 = -s01 * r2, j10 = -s10 * r2, j11 = s00 * r2;
      const double u00 = i00 * D[0][2] + i01 * D[1][2], u01 = i00 * D[0][3] + i01 * D[1][3];
      const double u10 = i10 * D[0][2] + i11 * D[1][2], u11 = i10 * D[0][3] + i11 * D[1][3];
      double E[4][4];
      E[2][2] = j00; E[2][3] = j01; E[3][2] = j10; E[3][3] = j11;
      E[0][2] = -(u00 * j00 + u01 * j10); E[0][3] = -(u00 * j01 + u01 * j11);
      E[1][2] = -(u10 * j00 + u11 * j10); E[1][3] = -(u10 * j01 + u11 * j11);
      E[2][0] = -(j00 * t00 + j01 * t10); E[2][1] = -(j00 * t01 + j01 * t11);
      E[3][0] = -(j10 * t00 + j11 * t10); E[3][1] = -(j10 * t01 + j11 * t11);
      E[0][0] = i00 - (E[0][2] * t00 + E[0][3] * t10); E[0][1] = i01 - (E[0][2] * t01 + E[0][3] * t11);
      E[1][0] = i10 - (E[1][2] * t00 + E[1][3] * t10); E[1][1] = i11 - (E[1][2] * t01 + E[1][3] * t11);
      if (F32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) E[q][q2] = rnd<F32>(E[q][q2]);
      }
      const bool inP = (i >= p0) && (i < p0 + 4);
      // new entries of the pivot lines at column i, where they are stored (i <= p0 + q)
      if (i < p0 + 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (i <= p0 + q) {
            const double v = inP ? E[q][i - p0] : (E[q][0] * st[0] + E[q][1] * st[1] + E[q][2] * st[2] + E[q][3] * st[3]);
            K[(p0 + q) * LD + i] = rnd<F32>(v);
          }
        }
      }
      if (!inP) {
        const bool below = i > p0 + 3;  // this line is not pivoted yet: its entries at the pivot columns are stored in the line itself
        double c0, c1, c2, c3;
        if (below) {
          const double2 c01 = *reinterpret_cast<const double2 *>(Ki + p0), c23 = *reinterpret_cast<const double2 *>(Ki + p0 + 2);
          c0 = c01.x; c1 = c01.y; c2 = c23.x; c3 = c23.y;
        } else {  // already pivoted line: M(i, p0+q) = -M(p0+q, i)
          c0 = -st[0]; c1 = -st[1]; c2 = -st[2]; c3 = -st[3];
        }
        const double w0 = rnd<F32>(c0 * E[0][0] + c1 * E[1][0] + c2 * E[2][0] + c3 * E[3][0]);
        const double w1 = rnd<F32>(c0 * E[0][1] + c1 * E[1][1] + c2 * E[2][1] + c3 * E[3][1]);
        const double w2 = rnd<F32>(c0 * E[0][2] + c1 * E[1][2] + c2 * E[2][2] + c3 * E[3][2]);
        const double w3 = rnd<F32>(c0 * E[0][3] + c1 * E[1][3] + c2 * E[2][3] + c3 * E[3][3]);
        int j = 0;
        for (; j + 3 <= i; j += 4) {  // columns j..j+3 <= i
          double2 pa[4], pb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pa[u] = *reinterpret_cast<const double2 *>(prow4 + 4 * (j + u));
            pb[u] = *reinterpret_cast<const double2 *>(prow4 + 4 * (j + u) + 2);
          }
          double2 k01 = *reinterpret_cast<const double2 *>(Ki + j), k23 = *reinterpret_cast<const double2 *>(Ki + j + 2);
          k01.x = rnd<F32>(fma(-w3, pb[0].y, fma(-w2, pb[0].x, fma(-w1, pa[0].y, fma(-w0, pa[0].x, k01.x)))));
          k01.y = rnd<F32>(fma(-w3, pb[1].y, fma(-w2, pb[1].x, fma(-w1, pa[1].y, fma(-w0, pa[1].x, k01.y)))));
          k23.x = rnd<F32>(fma(-w3, pb[2].y, fma(-w2, pb[2].x, fma(-w1, pa[2].y, fma(-w0, pa[2].x, k23.x)))));
          k23.y = rnd<F32>(fma(-w3, pb[3].y, fma(-w2, pb[3].x, fma(-w1, pa[3].y, fma(-w0, pa[3].x, k23.y)))));
          *reinterpret_cast<double2 *>(Ki + j) = k01;
          *reinterpret_cast<double2 *>(Ki + j + 2) = k23;
        }
        for (; j <= i; ++j) {
          const double2 a0 = *reinterpret_cast<const double2 *>(prow4 + 4 * j);
          const double2 b0 = *reinterpret_cast<const double2 *>(prow4 + 4 * j + 2);
          Ki[j] = rnd<F32>(fma(-w3, b0.y, fma(-w2, b0.x, fma(-w1, a0.y, fma(-w0, a0.x, Ki[j])))));
        }
        if (below) {
          double2 m01, m23;
          m01.x = -w0; m01.y = -w1; m23.x = -w2; m23.y = -w3;
          *reinterpret_cast<double2 *>(Ki + p0) = m01;
          *reinterpret_cast<double2 *>(Ki + p0 + 2) = m23;
        }
      }
    }
    __syncthreads();
  }
  double *prow = prow4;
  for (int p = p0; p < n; ++p) {  // scalar tail, same rules with a 1 x 1 pivot block
    double sp = 0.0;
    if (i < n) {
      sp = (i <= p) ? K[p * LD + i] : Ki[p];
      prow[i] = sp;
    }
    __syncthreads();
    if (i < n) {
      const double d = rnd<F32>(1.0 / prow[p]);
      if (i <= p) K[p * LD + i] = (i == p) ? d : rnd<F32>(sp * d);
      if (i != p) {
        const double c = (i > p) ? Ki[p] : -sp;
        const double f = rnd<F32>(c * d);
        for (int j = 0; j <= i; ++j)
          if (j != p) Ki[j] = rnd<F32>(fma(-f, prow[j], Ki[j]));
        if (i > p) Ki[p] = -f;
      }
    }
    __syncthreads();
  }
  // the inverse is symmetric: fill the upper triangle (the mat-vecs read whole runs)
  if (i < n)
    for (int j = i + 1; j < n; ++j) Ki[j] = K[j * LD + i];
  __syncthreads();
}

template <int LD, bool F32 = false>
__device__ __noinline__ void gj_invert(double *K, int n, double *prow4) {
  // Access pattern (round 2): the matrix is symmetric before and after, and M(i,j) = +-M(j,i) in between, so the sweep may as well be
  // run on the transpose: thread i owns the CONTIGUOUS run K[i * LD + 0..n) and reads / writes it with 128-bit accesses (LD is even
  // and = 14 or 2 mod 16: conflict-free per quarter warp); the pivot rows are gathered with stride LD (lanes adjacent: conflict-free).
  static_assert((LD % 2) == 0, "128-bit accesses to a thread's run");
  const int i = threadIdx.x;
  double *Ki = K + (size_t)i * LD;
  int p0 = 0;
  for (; p0 + 3 < n; p0 += 4) {
    if (i < n) {  // stage entry i of the 4 pivot lines: prow4[i][q] = M[p0+q][i]
      double2 a, b;
      a.x = K[p0 * LD + i];
      a.y = K[(p0 + 1) * LD + i];
      b.x = K[(p0 + 2) * LD + i];
      b.y = K[(p0 + 3) * LD + i];
      *reinterpret_cast<double2 *>(prow4 + 4 * i) = a;
      *reinterpret_cast<double2 *>(prow4 + 4 * i + 2) = b;
    }
    __syncthreads();
    if (i < n) {
      // D[q][q'] = M[p0+q][p0+q'] = prow4[p0+q'][q]; 4x4 inverse through 2x2 blocks (two reciprocals)
      double D[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 a = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c));
        const double2 b = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c) + 2);
        D[0][c] = a.x; D[1][c] = a.y; D[2][c] = b.x; D[3][c] = b.y;
      }
      const double r1 = fast_rcp(D[0][0] * D[1][1] - D[0][1] * D[1][0]);
      const double i00 = D[1][1] * r1, i01 = -D[0][1] * r1, i10 = -D[1][0] * r1, i11 = D[0][0] * r1;  // A11^-1
      // T = A21 A11^-1
      const double t00 = D[2][0] * i00 + D[2][1] * i10, t01 = D[2][0] * i01 + D[2][1] * i11;
      const double t10 = D[3][0] * i00 + D[3][1] * i10, t11 = D[3][0] * i01 + D[3][1] * i11;
      // S = A22 - T A12
      const double s00 = D[2][2] - (t00 * D[0][2] + t01 * D[1][2]), s01 = D[2][3] - (t00 * D[0][3] + t01 * D[1][3]);
      const double s10 = D[3][2] - (t10 * D[0][2] + t11 * D[1][2]), s11 = D[3][3] - (t10 * D[0][3] + t11 * D[1][3]);
      const double r2 = fast_rcp(s00 * s11 - s01 * s10);
      const double j00 = s11 * r2, j01 = -s01 * r2, j10 = -s10 * r2, j11 = s00 * r2;  // S^-1
      // U = A11^-1 A12 ; E12 = -U S^-1 ; E21 = -S^-1 T ; E11 = A11^-1 + U S^-1 T
      const double u00 = i00 * D[0][2] + i01 * D[1][2], u01 = i00 * D[0][3] + i01 * D[1][3];
      const double u10 = i10 * D[0][2] + i11 * D[1][2], u11 = i10 * D[0][3] + i11 * D[1][3];
      double E[4][4];
      E[2][2] = j00; E[2][3] = j01; E[3][2] = j10; E[3][3] = j11;
      E[0][2] = -(u00 * j00 + u01 * j10); E[0][3] = -(u00 * j01 + u01 * j11);
      E[1][2] = -(u10 * j00 + u11 * j10); E[1][3] = -(u10 * j01 + u11 * j11);
      E[2][0] = -(j00 * t00 + j01 * t10); E[2][1] = -(j00 * t01 + j01 * t11);
      E[3][0] = -(j10 * t00 + j11 * t10); E[3][1] = -(j10 * t01 + j11 * t11);
      E[0][0] = i00 - (E[0][2] * t00 + E[0][3] * t10); E[0][1] = i01 - (E[0][2] * t01 + E[0][3] * t11);
      E[1][0] = i10 - (E[1][2] * t00 + E[1][3] * t10); E[1][1] = i11 - (E[1][2] * t01 + E[1][3] * t11);
      if (F32) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) E[q][q2] = rnd<F32>(E[q][q2]);
      }
      const bool inP = (i >= p0) && (i < p0 + 4);
      // gather duty: new entries i of the four pivot lines
      {
        const double2 a = *reinterpret_cast<const double2 *>(prow4 + 4 * i);
        const double2 b = *reinterpret_cast<const double2 *>(prow4 + 4 * i + 2);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double v = inP ? E[q][i - p0] : (E[q][0] * a.x + E[q][1] * a.y + E[q][2] * b.x + E[q][3] * b.y);
          K[(p0 + q) * LD + i] = rnd<F32>(v);
        }
      }
      if (!inP) {  // run duty
        const double2 c01 = *reinterpret_cast<const double2 *>(Ki + p0), c23 = *reinterpret_cast<const double2 *>(Ki + p0 + 2);
        const double c0 = c01.x, c1 = c01.y, c2 = c23.x, c3 = c23.y;
        const double w0 = rnd<F32>(c0 * E[0][0] + c1 * E[1][0] + c2 * E[2][0] + c3 * E[3][0]);
        const double w1 = rnd<F32>(c0 * E[0][1] + c1 * E[1][1] + c2 * E[2][1] + c3 * E[3][1]);
        const double w2 = rnd<F32>(c0 * E[0][2] + c1 * E[1][2] + c2 * E[2][2] + c3 * E[3][2]);
        const double w3 = rnd<F32>(c0 * E[0][3] + c1 * E[1][3] + c2 * E[2][3] + c3 * E[3][3]);
        int j = 0;
        for (; j + 3 < n; j += 4) {  // 4 elements per trip: 8 broadcast LDS.128 (pivot lines) + 2 LDS.128 / 2 STS.128 of the thread's own run
          double2 pa[4], pb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pa[u] = *reinterpret_cast<const double2 *>(prow4 + 4 * (j + u));
            pb[u] = *reinterpret_cast<const double2 *>(prow4 + 4 * (j + u) + 2);
          }
          double2 k01 = *reinterpret_cast<const double2 *>(Ki + j), k23 = *reinterpret_cast<const double2 *>(Ki + j + 2);
          k01.x = rnd<F32>(fma(-w3, pb[0].y, fma(-w2, pb[0].x, fma(-w1, pa[0].y, fma(-w0, pa[0].x, k01.x)))));
          k01.y = rnd<F32>(fma(-w3, pb[1].y, fma(-w2, pb[1].x, fma(-w1, pa[1].y, fma(-w0, pa[1].x, k01.y)))));
          k23.x = rnd<F32>(fma(-w3, pb[2].y, fma(-w2, pb[2].x, fma(-w1, pa[2].y, fma(-w0, pa[2].x, k23.x)))));
          k23.y = rnd<F32>(fma(-w3, pb[3].y, fma(-w2, pb[3].x, fma(-w1, pa[3].y, fma(-w0, pa[3].x, k23.y)))));
          *reinterpret_cast<double2 *>(Ki + j) = k01;
          *reinterpret_cast<double2 *>(Ki + j + 2) = k23;
        }
        for (; j < n; ++j) {
          const double2 a0 = *reinterpret_cast<const double2 *>(prow4 + 4 * j);
          const double2 b0 = *reinterpret_cast<const double2 *>(prow4 + 4 * j + 2);
          Ki[j] = rnd<F32>(fma(-w3, b0.y, fma(-w2, b0.x, fma(-w1, a0.y, fma(-w0, a0.x, Ki[j])))));
        }
        double2 m01, m23;
        m01.x = -w0; m01.y = -w1; m23.x = -w2; m23.y = -w3;
        *reinterpret_cast<double2 *>(Ki + p0) = m01;
        *reinterpret_cast<double2 *>(Ki + p0 + 2) = m23;
      }
    }
    __syncthreads();
  }
  double *prow = prow4;
  for (int p = p0; p < n; ++p) {  // scalar tail
    double pi = 0.0;
    if (i < n) {
      pi = K[p * LD + i];  // pivot line element (p, i)
      prow[i] = pi;
    }
    __syncthreads();
    if (i < n) {
      const double d = rnd<F32>(1.0 / prow[p]);
      if (i != p) {
        const double f = rnd<F32>(Ki[p] * d);
        for (int j = 0; j < n; ++j) Ki[j] = rnd<F32>(fma(-f, prow[j], Ki[j]));
        Ki[p] = -f;
      }
      K[p * LD + i] = (i == p) ? d : rnd<F32>(pi * d);  // rescaled pivot line, element (p, i)
    }
    __syncthreads();
  }
}

// One m8n8k4 FP64 tensor-core MMA: C[8x8] += A[8x4] B[4x8].  Fragments (lane = 4 g + t): A[g][t], B[t][g], C[g][2t], C[g][2t+1].
// DMMA runs at the full FP64 rate of the SM (measured 37 TFLOP/s on B200, the DFMA rate) for 1/8 of the issue slots.
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Register-resident blocked Gauss-Jordan inverse on the FP64 tensor cores, 64 threads, matrices up to 64 x 64.
// The shared-memory version is bound by shared-memory bandwidth (every element is read and written once per four pivots;
// the LSU pipe was the busiest unit of the kernel).  Here the matrix lives in the MMA accumulator fragments of the two
// warps for the whole inversion - warp w owns row blocks w, w+2, w+4, w+6 (8 rows each) and all eight column tiles, 32
// tiles x 2 doubles per thread - and a 4-pivot step is
//   [A] the owners publish the 4 pivot rows (prow4[j][q] = A[p0+q][j]) and the 4 pivot columns (cw[i][q] = A[i][p0+q]);
//   [B] thread i inverts the 4x4 pivot block E = D^-1, forms its multipliers w[i][.] = A[i,P] E (published negated; zero
//       for pivot rows) and the new pivot-row entries of column i, np[i][.] = E A[P,i];
//   [C] trailing update C -= w Pr as one DMMA per tile, then the pivot columns take -w and the pivot rows take np.
// Shared memory is touched only by the O(n) staging of [A]/[B] and the 12 fragment loads per thread of [C].
// Tile geometry: tile rows are contiguous, the 8 columns of tile ct = 2 s + t are {16 s + 2 t + 4 u + e : u < 4, e < 2}
// (two interleaved tiles per 16-column span) so that fragment <-> memory moves with the odd leading dimension are
// bank-conflict free.  Rows / columns >= n are padded with the identity, so n needs no alignment.
// src: symmetric source, element (r, c), r >= c, at src[c * lds + r] (global Hessian scratch or the shared matrix
// itself); dg: added to the diagonal (may be nullptr).  The inverse is written to K[c * LD + r] (full matrix).
// K doubles as staging space while the matrix is in registers.
template <int LD>
__device__ __noinline__ void gj_invert_tc(const double *src, int lds, const double *dg, double *K, int n) {
  const int i = threadIdx.x, lane = i & 31, warp = i >> 5, gid = lane >> 2, tig = lane & 3;
  double *prow4 = K, *cw = K + 256, *nps = K + 512;
  double C[4][8][2];
  const int nrb = (n + 7) >> 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (warp + 2 * k) * 8 + gid;
#pragma unroll
    for (int ct = 0; ct < 8; ++ct) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = (ct >> 1) * 16 + 2 * (ct & 1) + 4 * tig + e;
        double v = (r == c) ? 1.0 : 0.0;
        if (r < n && c < n) {
          v = (r >= c) ? src[c * lds + r] : src[r * lds + c];
          if (r == c && dg != nullptr) v += dg[r];
        }
        C[k][ct][e] = v;
      }
    }
  }
  __syncthreads();  // the source may be K itself
  for (int p0 = 0; p0 < n; p0 += 4) {
    const int rbp = p0 >> 3, kp = rbp >> 1, sp = p0 >> 4, up = (p0 >> 2) & 3;
    const bool rowOwner = (warp == (rbp & 1)) && ((gid >> 2) == ((p0 >> 2) & 1));
    const int qrow = gid & 3;
    // [A] publish pivot rows and pivot columns
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k == kp && rowOwner) {
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) {
          const int c = (ct >> 1) * 16 + 2 * (ct & 1) + 4 * tig;
          prow4[c * 4 + qrow] = C[k][ct][0];
          prow4[(c + 1) * 4 + qrow] = C[k][ct][1];
        }
      }
    }
    if (tig == up) {
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        if (s2 == sp) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = (warp + 2 * k) * 8 + gid;
            *reinterpret_cast<double2 *>(cw + 4 * r) = make_double2(C[k][2 * s2][0], C[k][2 * s2][1]);
            *reinterpret_cast<double2 *>(cw + 4 * r + 2) = make_double2(C[k][2 * s2 + 1][0], C[k][2 * s2 + 1][1]);
          }
        }
      }
    }
    __syncthreads();
    // [B] E = D^-1 by an in-place scalar Gauss-Jordan on the 4x4 block, then multipliers and new pivot-row entries
    {
      double E[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 a = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c));
        const double2 b = *reinterpret_cast<const double2 *>(prow4 + 4 * (p0 + c) + 2);
        E[0][c] = a.x; E[1][c] = a.y; E[2][c] = b.x; E[3][c] = b.y;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const double d = 1.0 / E[p][p];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c != p) E[p][c] *= d;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (r == p) continue;
          const double f = E[r][p];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c != p) E[r][c] = fma(-f, E[p][c], E[r][c]);
          E[r][p] = -f * d;
        }
        E[p][p] = d;
      }
      const double2 pa = *reinterpret_cast<const double2 *>(prow4 + 4 * i);
      const double2 pb = *reinterpret_cast<const double2 *>(prow4 + 4 * i + 2);
      const double2 ca = *reinterpret_cast<const double2 *>(cw + 4 * i);
      const double2 cb = *reinterpret_cast<const double2 *>(cw + 4 * i + 2);
      const bool inP = (i >= p0) && (i < p0 + 4);
      const int rp = i - p0;
      double np[4], w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double epick = (rp == 0) ? E[q][0] : (rp == 1) ? E[q][1] : (rp == 2) ? E[q][2] : E[q][3];
        np[q] = inP ? epick : (E[q][0] * pa.x + E[q][1] * pa.y + E[q][2] * pb.x + E[q][3] * pb.y);
        w[q] = inP ? 0.0 : -(ca.x * E[0][q] + ca.y * E[1][q] + cb.x * E[2][q] + cb.y * E[3][q]);
      }
      *reinterpret_cast<double2 *>(nps + 4 * i) = make_double2(np[0], np[1]);
      *reinterpret_cast<double2 *>(nps + 4 * i + 2) = make_double2(np[2], np[3]);
      *reinterpret_cast<double2 *>(cw + 4 * i) = make_double2(w[0], w[1]);
      *reinterpret_cast<double2 *>(cw + 4 * i + 2) = make_double2(w[2], w[3]);
    }
    __syncthreads();
    // [C] trailing update on the tensor cores, then pivot columns and pivot rows
    {
      double af[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) af[k] = cw[((warp + 2 * k) * 8 + gid) * 4 + tig];
#pragma unroll
      for (int ct = 0; ct < 8; ++ct) {
        const int J0 = (ct >> 1) * 16 + 2 * (ct & 1);
        if (J0 < n) {
          const double b = prow4[(J0 + 4 * (gid >> 1) + (gid & 1)) * 4 + tig];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (warp + 2 * k < nrb) dmma_8x8x4(C[k][ct][0], C[k][ct][1], af[k], b);
        }
      }
      if (tig == up) {
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          if (s2 == sp) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = (warp + 2 * k) * 8 + gid;
              const double2 wa = *reinterpret_cast<const double2 *>(cw + 4 * r);
              const double2 wb = *reinterpret_cast<const double2 *>(cw + 4 * r + 2);
              C[k][2 * s2][0] = wa.x; C[k][2 * s2][1] = wa.y;
              C[k][2 * s2 + 1][0] = wb.x; C[k][2 * s2 + 1][1] = wb.y;
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (k == kp && rowOwner) {
#pragma unroll
          for (int ct = 0; ct < 8; ++ct) {
            const int c = (ct >> 1) * 16 + 2 * (ct & 1) + 4 * tig;
            C[k][ct][0] = nps[c * 4 + qrow];
            C[k][ct][1] = nps[(c + 1) * 4 + qrow];
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (warp + 2 * k) * 8 + gid;
#pragma unroll
    for (int ct = 0; ct < 8; ++ct) {
      const int c = (ct >> 1) * 16 + 2 * (ct & 1) + 4 * tig;
      K[c * LD + r] = C[k][ct][0];
      K[(c + 1) * LD + r] = C[k][ct][1];
    }
  }
  __syncthreads();
}

// In-place blocked (4 columns per step) right-looking Cholesky factorisation of the SPD matrix held in the lower triangle
// of K (element (i, j) at K[j * LD + i], thread i owns row i).  On return the strict lower triangle holds L and the
// diagonal holds 1 / L_ii.  n is padded to a multiple of 4 with identity rows, so there is no scalar tail.  A third of
// the flops of the explicit Gauss-Jordan inverse spent only on the triangle that is used: the polish needs two solves
// per round, not an inverse.   lst: [4 * blockDim] staging of the current block column, dst: [16] diagonal block.
template <int LD>
__device__ __noinline__ void chol_factor(double *K, int n, double *lst, double *dst) {
  const int i = threadIdx.x;
  const int n4 = (n + 3) & ~3;
  if (i >= n && i < n4) {
    for (int j = 0; j < i; ++j) K[j * LD + i] = 0.0;
    K[i * LD + i] = 1.0;
  }
  for (int p0 = 0; p0 < n4; p0 += 4) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const bool mine = (i >= p0) && (i < n4);
    if (mine) {
      a0 = K[p0 * LD + i];
      a1 = K[(p0 + 1) * LD + i];
      a2 = K[(p0 + 2) * LD + i];
      a3 = K[(p0 + 3) * LD + i];
      if (i < p0 + 4) {
        double *d = dst + 4 * (i - p0);
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
      }
    }
    __syncthreads();
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    if (mine) {
      const double r00 = fast_rsqrt(dst[0]);
      const double l10 = dst[4] * r00, l20 = dst[8] * r00, l30 = dst[12] * r00;
      const double r11 = fast_rsqrt(dst[5] - l10 * l10);
      const double l21 = (dst[9] - l20 * l10) * r11, l31 = (dst[13] - l30 * l10) * r11;
      const double r22 = fast_rsqrt(dst[10] - l20 * l20 - l21 * l21);
      const double l32 = (dst[14] - l30 * l20 - l31 * l21) * r22;
      const double r33 = fast_rsqrt(dst[15] - l30 * l30 - l31 * l31 - l32 * l32);
      if (i >= p0 + 4) {
        x0 = a0 * r00;
        x1 = (a1 - x0 * l10) * r11;
        x2 = (a2 - x0 * l20 - x1 * l21) * r22;
        x3 = (a3 - x0 * l30 - x1 * l31 - x2 * l32) * r33;
        K[p0 * LD + i] = x0;
        K[(p0 + 1) * LD + i] = x1;
        K[(p0 + 2) * LD + i] = x2;
        K[(p0 + 3) * LD + i] = x3;
        *reinterpret_cast<double2 *>(lst + 4 * i) = make_double2(x0, x1);
        *reinterpret_cast<double2 *>(lst + 4 * i + 2) = make_double2(x2, x3);
      } else {
        const int r = i - p0;
        if (r == 0) {
          K[p0 * LD + i] = r00;
        } else if (r == 1) {
          K[p0 * LD + i] = l10;
          K[(p0 + 1) * LD + i] = r11;
        } else if (r == 2) {
          K[p0 * LD + i] = l20;
          K[(p0 + 1) * LD + i] = l21;
          K[(p0 + 2) * LD + i] = r22;
        } else {
          K[p0 * LD + i] = l30;
          K[(p0 + 1) * LD + i] = l31;
          K[(p0 + 2) * LD + i] = l32;
          K[(p0 + 3) * LD + i] = r33;
        }
      }
    }
    __syncthreads();
    if (i >= p0 + 4 && i < n4) {  // trailing update of the own row, columns p0+4 .. i
      double *Ki = K + i;
      int j = p0 + 4;
      for (; j + 3 <= i; j += 4) {
        double2 la[4], lb[4];
        double kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          la[u] = *reinterpret_cast<const double2 *>(lst + 4 * (j + u));
          lb[u] = *reinterpret_cast<const double2 *>(lst + 4 * (j + u) + 2);
          kk[u] = Ki[(j + u) * LD];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          Ki[(j + u) * LD] = fma(-x3, lb[u].y, fma(-x2, lb[u].x, fma(-x1, la[u].y, fma(-x0, la[u].x, kk[u]))));
      }
      for (; j <= i; ++j) {
        const double2 la = *reinterpret_cast<const double2 *>(lst + 4 * j);
        const double2 lb = *reinterpret_cast<const double2 *>(lst + 4 * j + 2);
        Ki[j * LD] = fma(-x3, lb.y, fma(-x2, lb.x, fma(-x1, la.y, fma(-x0, la.x, Ki[j * LD]))));
      }
    }
  }
  __syncthreads();
}

// x = (L L^T)^-1 b with the factor of chol_factor: blocked forward and backward substitution, 4 unknowns and one barrier
// per step (the 4x4 diagonal solve is repeated by every thread from broadcast loads).  b and the result are per thread
// (i < n).  yst: [8] double-buffered staging of the block right-hand side.
template <int LD>
__device__ __noinline__ double chol_solve(const double *K, int n, double b, double *yst) {
  const int i = threadIdx.x;
  const int n4 = (n + 3) & ~3;
  double acc = (i < n) ? b : 0.0;
  double yi = 0.0;
  int buf = 0;
  for (int p0 = 0; p0 < n4; p0 += 4, buf ^= 4) {
    if (i >= p0 && i < p0 + 4) yst[buf + (i - p0)] = acc;
    __syncthreads();
    if (i >= p0 && i < n4) {
      const double *c0 = K + p0 * LD + p0, *c1 = c0 + LD + 1, *c2 = c1 + LD + 1, *c3 = c2 + LD + 1;
      const double y0 = yst[buf] * c0[0];
      const double y1 = (yst[buf + 1] - c0[1] * y0) * c1[0];
      const double y2 = (yst[buf + 2] - c0[2] * y0 - c1[1] * y1) * c2[0];
      const double y3 = (yst[buf + 3] - c0[3] * y0 - c1[2] * y1 - c2[1] * y2) * c3[0];
      if (i >= p0 + 4) {
        acc -= K[p0 * LD + i] * y0 + K[(p0 + 1) * LD + i] * y1 + K[(p0 + 2) * LD + i] * y2 + K[(p0 + 3) * LD + i] * y3;
      } else {
        const int r = i - p0;
        yi = (r == 0) ? y0 : (r == 1) ? y1 : (r == 2) ? y2 : y3;
      }
    }
  }
  acc = yi;
  double xi = 0.0;
  for (int p0 = n4 - 4; p0 >= 0; p0 -= 4, buf ^= 4) {
    if (i >= p0 && i < p0 + 4) yst[buf + (i - p0)] = acc;
    __syncthreads();
    if (i < p0 + 4) {
      const double *c0 = K + p0 * LD + p0, *c1 = c0 + LD + 1, *c2 = c1 + LD + 1, *c3 = c2 + LD + 1;
      const double x3 = yst[buf + 3] * c3[0];
      const double x2 = (yst[buf + 2] - c2[1] * x3) * c2[0];
      const double x1 = (yst[buf + 1] - c1[1] * x2 - c1[2] * x3) * c1[0];
      const double x0 = (yst[buf] - c0[1] * x1 - c0[2] * x2 - c0[3] * x3) * c0[0];
      if (i < p0) {  // L[p0 + q][i] = K[i * LD + p0 + q]
        const double *col = K + i * LD + p0;
        acc -= col[0] * x0 + col[1] * x1 + col[2] * x2 + col[3] * x3;
      } else {
        const int r = i - p0;
        xi = (r == 0) ? x0 : (r == 1) ? x1 : (r == 2) ? x2 : x3;
      }
    }
  }
  __syncthreads();
  return xi;
}

template <int LD>
__device__ __forceinline__ double matvec_smem(const double *K, int n, const double *v) {
  // K is symmetric here (an inverse): thread i reads its contiguous run K[i * LD + j] with 128-bit loads instead of the strided column
  const double *Ki = K + (size_t)threadIdx.x * LD;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int j = 0;
  for (; j + 3 < n; j += 4) {
    const double2 v01 = *reinterpret_cast<const double2 *>(v + j);
    const double2 v23 = *reinterpret_cast<const double2 *>(v + j + 2);
    const double2 k01 = *reinterpret_cast<const double2 *>(Ki + j);
    const double2 k23 = *reinterpret_cast<const double2 *>(Ki + j + 2);
    a0 = fma(k01.x, v01.x, a0);
    a1 = fma(k01.y, v01.y, a1);
    a2 = fma(k23.x, v23.x, a2);
    a3 = fma(k23.y, v23.y, a3);
  }
  for (; j < n; ++j) a0 = fma(Ki[j], v[j], a0);
  return (a0 + a1) + (a2 + a3);
}

// y_i = sum_j Hs(i,j) v_j with Hs in global memory (column-major, leading dimension BS; coalesced over i).  The scratch is
// L2-resident, so the loop is latency bound: 12 independent loads are issued before the first FMA.
template <int BS>
__device__ __forceinline__ double matvec_glob(const double *__restrict__ Hs, int n, const double *v) {
  const double *Hi = Hs + threadIdx.x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int j = 0;
  for (; j + 11 < n; j += 12) {
    double h[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) h[u] = __ldcg(Hi + (size_t)(j + u) * BS);
#pragma unroll
    for (int u = 0; u < 12; u += 4) {
      a0 = fma(h[u], v[j + u], a0);
      a1 = fma(h[u + 1], v[j + u + 1], a1);
      a2 = fma(h[u + 2], v[j + u + 2], a2);
      a3 = fma(h[u + 3], v[j + u + 3], a3);
    }
  }
  for (; j + 3 < n; j += 4) {
    const double h0 = __ldcg(Hi + (size_t)j * BS), h1 = __ldcg(Hi + (size_t)(j + 1) * BS), h2 = __ldcg(Hi + (size_t)(j + 2) * BS),
                 h3 = __ldcg(Hi + (size_t)(j + 3) * BS);
    a0 = fma(h0, v[j], a0);
    a1 = fma(h1, v[j + 1], a1);
    a2 = fma(h2, v[j + 2], a2);
    a3 = fma(h3, v[j + 3], a3);
  }
  for (; j < n; ++j) a0 = fma(__ldcg(Hi + (size_t)j * BS), v[j], a0);
  return (a0 + a1) + (a2 + a3);
}

// S2(j,l) = sum_{t=m}^{N-1} (t-j)(t-l), m = max(j,l); cnt = N - m
__device__ __forceinline__ void horizon_sums(int N, int j, int l, double &cnt, double &S2) {
  const int m = j > l ? j : l;
  const int c = N - m;
  // sum t, sum t^2 for t = m..N-1
  const long long a = m, b = N - 1;
  const long long st = (b * (b + 1) - (a - 1) * a) / 2;
  const long long st2 = (b * (b + 1) * (2 * b + 1) - (a - 1) * a * (2 * a - 1)) / 6;
  cnt = (double)c;
  S2 = (double)(st2 - (long long)(j + l) * st + (long long)j * l * c);
}

// ---------------------------------------------------------------------------------------------------------------
// BS = threads per block (multiple of 32), NV = largest reduced problem the instantiation accepts (even, <= BS)
// PREC: 0 = FP64 (the product); 1 = FP32 K^-1 and ADMM iterations on FP64 problem data, FP64 polish ("FP32 factor + FP64
// refinement"); 2 = FP32 everywhere including the assembly, the polish and the outputs (error measurement only: b200qp_mpc_config.precision)
template <int BS, int NV, int PREC = 0>
__global__ void __launch_bounds__(BS, (BS == 64 && !B200QP_TC_GJ) ? 6 : 1) mpc_admm_kernel(const MpcParams P, const MpcBuffers B) {
  constexpr bool R1 = PREC == 1 || PREC == 2, R2 = PREC == 2;  // PREC 3: FP64 with the per-phase cycle counters (its own instantiation,
                                                                // so that the product kernel carries neither the counters nor their branches)
  constexpr int LD = NV + 2;  // even (128-bit accesses to a thread's run) and = 14 or 2 mod 16 for NV = 60, 96, 128, 156 (conflict-free)
  constexpr int MAXF = NV / 3 + 1;
  extern __shared__ __align__(16) double smem[];
  double *K = smem;                 // [NV * LD]
  double *s_g = K + NV * LD;        // [BS] gradient
  double *s_c = s_g + BS;           // [BS] particular solution of the polish
  double *s_v = s_c + BS;           // [BS] rhs / x broadcast buffer
  double *s_xt = s_v + BS;          // [BS]
  double *s_r = s_xt + BS;          // [BS] H x + g
  double *s_w = s_r + BS;           // [BS] reduced solution
  double *prow = s_v;               // [4 * BS] staged pivot rows of the Gauss-Jordan sweep: s_v..s_w are idle during an inversion
  double *red = s_w + BS;           // [4 * 32]
  // Stage quantities are only needed while K is not (assembly before the first factorisation, outputs after the solve), so
  // they alias the K region: one more resident block per SM.  Bw is parked in the per-block global scratch in between.
  double *s_Bw = K;                        // [MAXH * 36] Bw_kf, row-major 3x3 per (k,f); zero for swing feet
  double *s_qe = s_Bw + MAXH * 36;         // [(MAXH + 1) * 12] Q e_k (free-response error), later reused for suffix sums
  double *s_lw = s_qe + (MAXH + 1) * 12;   // [(MAXH + 2) * 3] adjoint, angular-velocity part, per stage
  double *s_lv = s_lw + (MAXH + 2) * 3;    // [(MAXH + 2) * 3] adjoint, linear-velocity part
  double *s_bw = s_lv + (MAXH + 2) * 3;    // [MAXH * 6] per-stage input effect on (omega, v) for the roll-out
  double *s_X = s_bw + MAXH * 6;           // [(MAXH + 1) * 13]
  static_assert(MAXH * 36 + (MAXH + 1) * 12 + 2 * (MAXH + 2) * 3 + MAXH * 6 + (MAXH + 1) * 13 <= NV * LD, "alias region");
  __shared__ double s_yawd[MAXH + 1];
  __shared__ double s_x0[13];
  __shared__ double s_Rt[9];   // Rt(psi0) row-major
  __shared__ double s_G2[9];   // dt^2 Rt^T Qth Rt
  __shared__ double s_fZ[MAXF * 9];
  __shared__ double s_fc[MAXF * 3];
  __shared__ short s_fidx[MAXH * 4];  // (k,f) -> free foot index or -1
  __shared__ unsigned char s_fk[MAXF], s_ff[MAXF], s_fnz[MAXF];
  __shared__ short s_foff[MAXF + 1];
  __shared__ short s_rfoot[BS];
  __shared__ unsigned char s_rk[BS];
  __shared__ unsigned short s_actl[MAXF], s_actu[MAXF];
  __shared__ int s_slot, s_nf, s_changed;
  __shared__ double s_psi0;

  const int tid = threadIdx.x;
  const int N = P.N;
  const double dt = P.dt, mu = P.mu;
  double *Hs = B.hscratch + (size_t)blockIdx.x * (BS * BS + MAXH * 36);
  double *Bw_park = Hs + BS * BS;
  double lo[7], hi[7];
  foot_bounds(P.fmin, P.fmax, lo, hi);

  for (;;) {
    __syncthreads();
    if (tid == 0) s_slot = atomicAdd(B.bin_next, 1);
    __syncthreads();
    const int slot = s_slot;
    if (slot >= *B.bin_count) break;
    const int prob = B.bin_list[slot];
    // Stage the record (2.3 KB at N = 10) and its contact bytes in shared memory with one coalesced pass: the scattered
    // per-thread reads of the assembly then never leave the SM.  The copy lives in the K region behind the stage arrays.
    constexpr int STAGE_DOUBLES = MAXH * 36 + (MAXH + 1) * 12 + 2 * (MAXH + 2) * 3 + MAXH * 6 + (MAXH + 1) * 13;
    constexpr int REC_DOUBLES = 26 + 13 * (MAXH + 1) + 12 * MAXH + (4 * MAXH + 7) / 8;
    static_assert(STAGE_DOUBLES + REC_DOUBLES <= NV * LD, "record staging area");
    double *s_rec = K + STAGE_DOUBLES;  // behind the stage arrays, dead before the first factorisation
    if (B.ready != nullptr) {  // the H2D copy of this problem's chunk may still be in flight (copy engine, own stream)
      if (tid == 0) {
        const volatile int *flag = B.ready + prob / B.ready_chunk;
        // bounded: a chunk copy that failed, or that cannot run beside the kernel (no free copy engine, a tool that
        // serialises streams), must not hang the grid - after ~2 s the problem is reported as failed instead
        const long long t0 = clock64();
        int landed = 1;
        while (*flag == 0) {
          __nanosleep(500);
          if (clock64() - t0 > 4000000000LL) {
            landed = 0;
            break;
          }
        }
        __threadfence_system();  // acquire: the record reads below must not be satisfied by anything older than the flag
        s_changed = landed;
      }
      __syncthreads();
      if (s_changed == 0) {
        if (tid == 0) {
          b200qp_solver_info inf = {B200QP_ACADOS_QP_FAILURE, 0, 0, 0, {0, 0, 0, 0}};
          B.info[prob] = inf;
        }
        continue;
      }
    }
    {
      const double *grec = B.in + (size_t)prob * P.in_stride;
      const uint8_t *gct = B.contact + (size_t)prob * N * 4;
      for (int e = tid; e < P.in_stride; e += BS) s_rec[e] = __ldcg(grec + e);
      uint8_t *sct = reinterpret_cast<uint8_t *>(s_rec + P.in_stride);
      for (int e = tid; e < 4 * N; e += BS) sct[e] = gct[e];
    }
    __syncthreads();
    const double *rec = s_rec;
    const uint8_t *ct = reinterpret_cast<const uint8_t *>(s_rec + P.in_stride);
    const double mass = rec[13];

    constexpr bool timing = PREC == 3;
    long long cyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = timing ? clock64() : 0;
    // ---- phase 0: per-stage quantities -------------------------------------------------------------------
    if (tid <= N) {
      double R[3][3];
      quat_to_rot(rec + 26 + 13 * tid, R);
      s_yawd[tid] = atan2(R[1][0], R[0][0]);  // yaw_from_quaternion (quaternion_operations.hpp:136-141)
    }
    if (tid == 32) {
      double R[3][3], e[3];
      quat_to_rot(rec, R);
      rot_to_euler(R, e);
      s_x0[0] = e[0]; s_x0[1] = e[1]; s_x0[2] = e[2];
      for (int c = 0; c < 3; ++c) {
        s_x0[3 + c] = rec[4 + c] + rec[23 + c];
        s_x0[6 + c] = rec[7 + c];
        s_x0[9 + c] = rec[10 + c];
      }
      s_x0[12] = GRAVITY_CONSTANT;
      const double psi0 = atan2(R[1][0], R[0][0]);
      const double cs = cos(psi0), sn = sin(psi0);
      // MPC::GetR (mpc.cpp:46-58)
      s_Rt[0] = cs; s_Rt[1] = sn; s_Rt[2] = 0.0;
      s_Rt[3] = -sn; s_Rt[4] = cs; s_Rt[5] = 0.0;
      s_Rt[6] = 0.0; s_Rt[7] = 0.0; s_Rt[8] = 1.0;
      // G2 = dt^2 Rt^T diag(qth) Rt
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
          double s = 0.0;
          for (int d = 0; d < 3; ++d) s += s_Rt[3 * d + a] * P.w[d] * s_Rt[3 * d + b];
          s_G2[3 * a + b] = dt * dt * s;
        }
      s_psi0 = psi0;
    }
    // free-variable list (warp 0): ordered by (stage, leg)
    if (tid < 32) {
      int base = 0;
      for (int e0 = 0; e0 < 4 * N; e0 += 32) {
        const int e = e0 + tid;
        const bool on = (e < 4 * N) && (ct[e] != 0);
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        const int pos = base + __popc(bal & ((1u << tid) - 1u));
        if (e < 4 * N) s_fidx[e] = on ? (short)pos : (short)-1;
        if (on && pos < MAXF) {
          s_fk[pos] = (unsigned char)(e >> 2);
          s_ff[pos] = (unsigned char)(e & 3);
        }
        base += __popc(bal);
      }
      if (tid == 0) s_nf = base;
    }
    __syncthreads();
    const int nf = s_nf;
    const int n = 3 * nf;
    if (tid == 0) {  // sequential yaw unwrap (mpc.cpp:334-344)
      double before = s_psi0;
      for (int k = 0; k <= N; ++k) {
        double y = s_yawd[k];
        const double diff = y - before;
        if (diff < -M_PI) y += 2 * M_PI;
        else if (diff > M_PI) y -= 2 * M_PI;
        before = y;
        s_yawd[k] = y;
      }
    }
    __syncthreads();
    if (n > NV) {  // cannot happen after binning; defensive
      if (tid == 0) {
        b200qp_solver_info inf = {B200QP_ACADOS_QP_FAILURE, 0, 0, n, {0, 0, 0, 0}};
        B.info[prob] = inf;
      }
      continue;
    }
    if (tid < N) {
      // MPC::GetB rotational block (mpc.cpp:33-43): Bw = Iw^-1 skew(r) dt, Iw = Rt I Rt^T (mpc.cpp:60-64)
      const int k = tid;
      const double psi = s_yawd[k];
      const double cs = cos(psi), sn = sin(psi);
      const double Rt[3][3] = {{cs, sn, 0.0}, {-sn, cs, 0.0}, {0.0, 0.0, 1.0}};
      double Ib[3][3], T[3][3], Iw[3][3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Ib[r][c] = rec[14 + 3 * c + r];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[r][c] = Rt[r][0] * Ib[0][c] + Rt[r][1] * Ib[1][c] + Rt[r][2] * Ib[2][c];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Iw[r][c] = T[r][0] * Rt[c][0] + T[r][1] * Rt[c][1] + T[r][2] * Rt[c][2];
      const double c00 = Iw[1][1] * Iw[2][2] - Iw[1][2] * Iw[2][1];
      const double c01 = Iw[1][2] * Iw[2][0] - Iw[1][0] * Iw[2][2];
      const double c02 = Iw[1][0] * Iw[2][1] - Iw[1][1] * Iw[2][0];
      const double det = Iw[0][0] * c00 + Iw[0][1] * c01 + Iw[0][2] * c02;
      const double id = 1.0 / det;
      double Ii[3][3];
      Ii[0][0] = c00 * id;
      Ii[0][1] = (Iw[0][2] * Iw[2][1] - Iw[0][1] * Iw[2][2]) * id;
      Ii[0][2] = (Iw[0][1] * Iw[1][2] - Iw[0][2] * Iw[1][1]) * id;
      Ii[1][0] = c01 * id;
      Ii[1][1] = (Iw[0][0] * Iw[2][2] - Iw[0][2] * Iw[2][0]) * id;
      Ii[1][2] = (Iw[0][2] * Iw[1][0] - Iw[0][0] * Iw[1][2]) * id;
      Ii[2][0] = c02 * id;
      Ii[2][1] = (Iw[0][1] * Iw[2][0] - Iw[0][0] * Iw[2][1]) * id;
      Ii[2][2] = (Iw[0][0] * Iw[1][1] - Iw[0][1] * Iw[1][0]) * id;
      const double *dp = rec + 26 + 13 * k + 4;  // desired position
      const double *fp = rec + 26 + 13 * (N + 1) + 12 * k;
      for (int f = 0; f < 4; ++f) {
        double *Bw = s_Bw + (k * 4 + f) * 9;
        if (ct[k * 4 + f]) {
          const double rx = fp[3 * f + 0] - (dp[0] + rec[23]);
          const double ry = fp[3 * f + 1] - (dp[1] + rec[24]);
          const double rz = fp[3 * f + 2] - (dp[2] + rec[25]);
          const double S[3][3] = {{0.0, -rz, ry}, {rz, 0.0, -rx}, {-ry, rx, 0.0}};
          for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
              Bw[3 * r + c] = (Ii[r][0] * S[0][c] + Ii[r][1] * S[1][c] + Ii[r][2] * S[2][c]) * dt;
        } else {
          for (int t = 0; t < 9; ++t) Bw[t] = 0.0;
        }
      }
    }
    // free-response error e_k = A^k x0 - xref_k (closed form), k = 1..N, times Q  (mpc.cpp:372-381)
    if (tid >= 32 && tid - 32 < N) {
      const int k = tid - 32 + 1;
      double R[3][3], e[3];
      quat_to_rot(rec + 26 + 13 * k, R);
      rot_to_euler(R, e);
      e[2] = s_yawd[k];
      const double *st = rec + 26 + 13 * k;
      const double kd = (double)k;
      const double w0 = s_x0[6], w1 = s_x0[7], w2 = s_x0[8];
      double *qe = s_qe + 12 * k;
      for (int c = 0; c < 3; ++c) {
        const double th = s_x0[c] + kd * dt * (s_Rt[3 * c] * w0 + s_Rt[3 * c + 1] * w1 + s_Rt[3 * c + 2] * w2);
        qe[c] = P.w[c] * (th - e[c]);
        double pp = s_x0[3 + c] + kd * dt * s_x0[9 + c];
        double vv = s_x0[9 + c];
        if (c == 2) {
          pp -= 0.5 * kd * (kd - 1.0) * dt * dt * s_x0[12];
          vv -= kd * dt * s_x0[12];
        }
        qe[3 + c] = P.w[3 + c] * (pp - (st[4 + c] + rec[23 + c]));
        qe[6 + c] = P.w[6 + c] * (s_x0[6 + c] - st[7 + c]);
        qe[9 + c] = P.w[9 + c] * (vv - st[10 + c]);
      }
    }
    __syncthreads();
    // adjoint recursion (component-wise suffix sums), lambda_k for k = 1..N
    if (tid < 6) {
      // tid 0..2: theta/omega chain component, 3..5: p/v chain component
      const int c = tid % 3;
      const bool lin = tid >= 3;
      double sA = 0.0;  // suffix sum of Q e^theta (or Q e^p)
      double ssA = 0.0; // sum_{i>k} sA_i
      double sB = 0.0;  // suffix sum of Q e^omega (or Q e^v)
      for (int k = N; k >= 1; --k) {
        ssA += sA;  // now sum over i > k of sA_i
        sA += s_qe[12 * k + (lin ? 3 : 0) + c];
        sB += s_qe[12 * k + (lin ? 9 : 6) + c];
        if (lin) {
          s_lv[3 * k + c] = sB + dt * ssA;
        } else {
          s_lw[3 * k + c] = sB;       // Rt^T (dt ssTheta) added below
          s_qe[12 * k + c] = ssA;      // reuse slot: ssTheta_k[c]
        }
      }
    }
    __syncthreads();
    if (tid < N) {
      const int k = tid + 1;
      const double a0 = s_qe[12 * k], a1 = s_qe[12 * k + 1], a2 = s_qe[12 * k + 2];
      for (int c = 0; c < 3; ++c) s_lw[3 * k + c] += dt * (s_Rt[c] * a0 + s_Rt[3 + c] * a1 + s_Rt[6 + c] * a2);
    }
    __syncthreads();

    PH(0);
    // ---- phase 1: condensed reduced Hessian and gradient (K1) ------------------------------------------------
    const bool act = tid < n;
    const int fa = act ? tid / 3 : 0;
    const int ca = tid - 3 * fa;
    double gi = 0.0;
    if (act) {
      const int j = s_fk[fa], f = s_ff[fa];
      const double *Bwa = s_Bw + (j * 4 + f) * 9;
      double pa[3], pb[3];
      for (int e = 0; e < 3; ++e) {
        pa[e] = Bwa[3 * e + ca] * P.w[6 + e];
        pb[e] = Bwa[ca] * s_G2[e] + Bwa[3 + ca] * s_G2[3 + e] + Bwa[6 + ca] * s_G2[6 + e];
      }
      const double dm = dt / mass;
      const double lv = dm * dm * P.w[9 + ca], lp = dm * dm * dt * dt * P.w[3 + ca];
      for (int b = 0; b < nf; ++b) {
        const int l = s_fk[b], f2 = s_ff[b];
        double cnt, S2;
        horizon_sums(N, j, l, cnt, S2);
        const double v0 = cnt * pa[0] + S2 * pb[0], v1 = cnt * pa[1] + S2 * pb[1], v2 = cnt * pa[2] + S2 * pb[2];
        const double *Bwb = s_Bw + (l * 4 + f2) * 9;
        double h[3];
        for (int cc = 0; cc < 3; ++cc) h[cc] = v0 * Bwb[cc] + v1 * Bwb[3 + cc] + v2 * Bwb[6 + cc];
        h[ca] += cnt * lv + S2 * lp;
        if (b == fa) h[ca] += P.alpha;
        for (int cc = 0; cc < 3; ++cc) Hs[(3 * b + cc) * BS + tid] = rnd<R2>(h[cc]);
      }
      gi = Bwa[ca] * s_lw[3 * (j + 1)] + Bwa[3 + ca] * s_lw[3 * (j + 1) + 1] + Bwa[6 + ca] * s_lw[3 * (j + 1) + 2] +
           dm * s_lv[3 * (j + 1) + ca];
      gi = rnd<R2>(gi);
      s_g[tid] = gi;
    }
    for (int e = tid; e < N * 36; e += BS) Bw_park[e] = s_Bw[e];
    __syncthreads();
    if (prob == B.dbg_index && B.dbg_H != nullptr) {
      if (act) {
        for (int jx = 0; jx < n; ++jx) B.dbg_H[(size_t)jx * n + tid] = Hs[jx * BS + tid];
        B.dbg_g[tid] = gi;
      }
    }

    PH(1);
    // ---- phase 2: ADMM (K2a) -------------------------------------------------------------------------------------
    const double Ei = (ca == 2) ? (1.0 + 4.0 * mu * mu) : 3.0;  // diag(A^T A)
    double rho = P.rho0;
    const double sigma = P.sigma, relax = P.relax;
    double x = 0.0;
    double z[7], y[7];
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      z[r] = fmin(fmax(0.0, lo[r]), hi[r]);
      y[r] = 0.0;
    }
    // warm / hot start from the previous solution of this problem slot
    bool warm = false;
    if (B.warm_mode != 0 && n > 0 && B.warm_valid[prob] != 0) {
      warm = true;
      if (act) {
        const int kf = s_fk[fa] * 4 + s_ff[fa];
        const double *wx = B.warm_x + ((size_t)prob * N * 4 + kf) * 3;
        const double *wy = B.warm_y + ((size_t)prob * N * 4 + kf) * 7;
        const double f0 = wx[0], f1 = wx[1], f2 = wx[2];
        x = (ca == 0) ? f0 : (ca == 1) ? f1 : f2;
        double ax[7];
        foot_rows_eval(f0, f1, f2, mu, ax);
#pragma unroll
        for (int r = 0; r < 7; ++r) {
          z[r] = fmin(fmax(ax[r], lo[r]), hi[r]);
          y[r] = wy[r];
        }
      }
    }
    const bool hot = warm && B.warm_mode >= 2;
    int iters = 0, rounds_total = 0, status = B200QP_ACADOS_MAXITER;
    double eps = P.eps;
    double kkt_stat = 0.0, kkt_ineq = 0.0, kkt_comp = 0.0;
    double ysol[7] = {0, 0, 0, 0, 0, 0, 0};
    double xsol = 0.0;
    bool need_factor = true;
    bool have_cand = false;

    if (n == 0) status = B200QP_ACADOS_SUCCESS;
#pragma unroll 1
    for (int attempt = 0; attempt < 5 && status != B200QP_ACADOS_SUCCESS && n > 0; ++attempt) {
      // (re)build K^-1 from the Hessian scratch when needed
      bool admm_done = hot && attempt == 0;  // hot start: the previous active set goes straight to the polish
#pragma unroll 1
      while (!admm_done) {
        if (need_factor) {
          __syncthreads();
          if constexpr (BS == 64 && NV == 64 && B200QP_TC_GJ) {
            // K = H + sigma I + rho A^T A inverted in the tensor-core accumulators straight from the Hessian scratch
            s_xt[tid] = sigma + rho * Ei;
            __syncthreads();
            gj_invert_tc<LD>(Hs, BS, s_xt, K, n);
          } else {
            if (act) {
              int jx = 0;
              for (; jx + 7 < n; jx += 8) {
                double h[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) h[u] = __ldcg(Hs + (size_t)(jx + u) * BS + tid);
#pragma unroll
                for (int u = 0; u < 8; ++u) K[(jx + u) * LD + tid] = rnd<R1>(h[u]);
              }
              for (; jx < n; ++jx) K[jx * LD + tid] = rnd<R1>(__ldcg(Hs + (size_t)jx * BS + tid));
              K[tid * LD + tid] = rnd<R1>(K[tid * LD + tid] + (sigma + rho * Ei));
            }
            __syncthreads();
            if (B200QP_GJ_SYM) gj_invert_sym<LD, R1>(K, n, prow);
            else gj_invert<LD, R1>(K, n, prow);
          }
          need_factor = false;
          PH(2);
        }
        const double rinv = 1.0 / rho;
        int it_block = 0;
        const int block_len = (warm && iters < 15 && P.check_every > 5) ? 5 : P.check_every;  // warm: look early and often
#pragma unroll 1
        for (; it_block < block_len && iters < P.max_iter; ++it_block, ++iters) {
          if (act) {
            double w[7];
#pragma unroll
            for (int r = 0; r < 7; ++r) w[r] = rho * z[r] - y[r];
            s_v[tid] = rnd<R1>(sigma * x - gi + foot_at(ca, w, mu));
          }
          __syncthreads();
          double xt = 0.0;
          if (act) {
            xt = rnd<R1>(matvec_smem<LD>(K, n, s_v));
            s_xt[tid] = xt;
          }
          __syncthreads();
          if (act) {
            double zt[7];
            foot_rows_eval(s_xt[3 * fa], s_xt[3 * fa + 1], s_xt[3 * fa + 2], mu, zt);
            x = rnd<R1>(relax * xt + (1.0 - relax) * x);
#pragma unroll
            for (int r = 0; r < 7; ++r) {
              const double zr = rnd<R1>(relax * zt[r] + (1.0 - relax) * z[r]);
              const double zn = rnd<R1>(fmin(fmax(zr + y[r] * rinv, lo[r]), hi[r]));
              y[r] = rnd<R1>(y[r] + rho * (zr - zn));
              z[r] = zn;
            }
          }
        }
        PH(3);
        // residuals (OSQP termination test) on the true H
        __syncthreads();
        if (act) s_v[tid] = x;
        __syncthreads();
        double nrm[4] = {0.0, 0.0, 0.0, 0.0};
        bool bad = false;
        if (act) {
          double ax[7];
          foot_rows_eval(s_v[3 * fa], s_v[3 * fa + 1], s_v[3 * fa + 2], mu, ax);
          const double hx = matvec_glob<BS>(Hs, n, s_v);
          const double aty = foot_at(ca, y, mu);
#pragma unroll
          for (int r = 0; r < 7; ++r) {
            nrm[0] = fmax(nrm[0], fabs(ax[r] - z[r]));
            nrm[2] = fmax(nrm[2], absmax(ax[r], z[r]));
          }
          const double rd = hx + gi + aty;
          nrm[1] = fabs(rd);
          nrm[3] = fmax(absmax(hx, aty), fabs(gi));
          bad = !(fabs(rd) < BIG) || !(fabs(x) < BIG);
        }
        const int anybad = __syncthreads_or(bad ? 1 : 0);
        block_max<4>(nrm, red);
        if (anybad) {
          status = B200QP_ACADOS_NAN_DETECTED;
          admm_done = true;
          break;
        }
        const bool conv = (nrm[0] <= eps * (1.0 + nrm[2])) && (nrm[1] <= eps * (1.0 + nrm[3]));
        if (conv || iters >= P.max_iter) {
          admm_done = true;
        } else {
          // residual-balancing rho update (OSQP adaptive rho), refactor only on a factor-5 change
          const double num = nrm[0] / fmax(nrm[2], 1e-30), den = nrm[1] / fmax(nrm[3], 1e-30);
          double nr = rho * sqrt(num / fmax(den, 1e-30));
          nr = fmin(fmax(nr, 1e-6), 1e6);
          if (nr > 5.0 * rho || nr < 0.2 * rho) {
            rho = nr;
            need_factor = true;
          }
        }
      }
      if (status == B200QP_ACADOS_NAN_DETECTED) break;

      PH(4);
      // ---- phase 3: polish = primal-dual active-set sweeps in the null space of the active rows --------------
      if (act && ca == 2) {
        unsigned al = 0, au = 0;
#pragma unroll
        for (int r = 0; r < 7; ++r) {  // OSQP polish rule: lower-active if z - l < -y, upper-active if u - z < y
          if (lo[r] > -BIG && (z[r] - lo[r]) < -y[r]) al |= 1u << r;
          else if (hi[r] < BIG && (hi[r] - z[r]) < y[r]) au |= 1u << r;
        }
        s_actl[fa] = (unsigned short)al;
        s_actu[fa] = (unsigned short)au;
      }
      bool accepted = false;
#pragma unroll 1
      for (int round = 0; round < P.polish_rounds && !accepted; ++round) {
        ++rounds_total;
        __syncthreads();
        if (act && ca == 2) {
          s_fnz[fa] = (unsigned char)foot_nullspace(s_actl[fa], s_actu[fa], mu, lo, hi, s_fc + 3 * fa, s_fZ + 9 * fa);
        }
        __syncthreads();
        if (tid == 0) {
          int off = 0;
          for (int f = 0; f < nf; ++f) {
            s_foff[f] = (short)off;
            for (int q = 0; q < s_fnz[f]; ++q) {
              s_rfoot[off + q] = (short)f;
              s_rk[off + q] = (unsigned char)q;
            }
            off += s_fnz[f];
          }
          s_foff[nf] = (short)off;
          s_changed = 0;
        }
        if (act) s_c[tid] = s_fc[tid];
        __syncthreads();
        const int nr = s_foff[nf];
        PH(5);
        constexpr bool TCP = (BS == 64 && NV == 64 && B200QP_TC_GJ);
        // hc = H c + g
        if (act) s_r[tid] = matvec_glob<BS>(Hs, n, s_c) + gi;
        // reduced Hessian Hr = Z^T H Z into K
        double vp[3] = {0, 0, 0};
        int ap = 0;
        if (tid < nr) {
          ap = s_rfoot[tid];
          const int kq = s_rk[tid];
          vp[0] = s_fZ[9 * ap + 3 * kq];
          vp[1] = s_fZ[9 * ap + 3 * kq + 1];
          vp[2] = s_fZ[9 * ap + 3 * kq + 2];
          // u_c = sum_r vp[r] H(3 ap + r, c) for every column c is shared by the reduced columns of one foot: walk the feet,
          // 9 L2 loads per foot issued together, then up to 3 reduced entries
          for (int bq = 0; bq < nf; ++bq) {
            const int nzb = s_fnz[bq];
            if (B200QP_CHOL_POLISH && !TCP && s_foff[bq] > tid) break;  // Cholesky only needs the lower triangle
            if (nzb == 0) continue;
            double hh[9];
#pragma unroll
            for (int c2 = 0; c2 < 3; ++c2) {
              const double *col = Hs + (size_t)(3 * bq + c2) * BS + 3 * ap;
              hh[3 * c2] = __ldcg(col);
              hh[3 * c2 + 1] = __ldcg(col + 1);
              hh[3 * c2 + 2] = __ldcg(col + 2);
            }
            const double u0 = vp[0] * hh[0] + vp[1] * hh[1] + vp[2] * hh[2];
            const double u1 = vp[0] * hh[3] + vp[1] * hh[4] + vp[2] * hh[5];
            const double u2 = vp[0] * hh[6] + vp[1] * hh[7] + vp[2] * hh[8];
            const int q0 = s_foff[bq];
            for (int kq = 0; kq < nzb; ++kq) {
              const double *vq = s_fZ + 9 * bq + 3 * kq;
              K[(q0 + kq) * LD + tid] = rnd<R2>(vq[0] * u0 + vq[1] * u1 + vq[2] * u2);
            }
          }
        }
        __syncthreads();
        double grp = 0.0;
        if (tid < nr) grp = vp[0] * s_r[3 * ap] + vp[1] * s_r[3 * ap + 1] + vp[2] * s_r[3 * ap + 2];
        __syncthreads();  // s_r is part of the pivot-row staging area (prow = s_v..s_w) the inversion below writes first
        PH(6);
        constexpr bool TC = (BS == 64 && NV == 64 && B200QP_TC_GJ);
        double wred = 0.0;
        if constexpr (TC) {
          gj_invert_tc<LD>(K, LD, nullptr, K, nr);
          PH(7);
          if (tid < nr) s_v[tid] = -grp;
          __syncthreads();
          if (tid < nr) wred = matvec_smem<LD>(K, nr, s_v);
        } else if constexpr (B200QP_CHOL_POLISH) {
          chol_factor<LD>(K, nr, prow, red + 16);
          PH(7);
          wred = chol_solve<LD>(K, nr, -grp, red);
        } else {
          if (B200QP_GJ_SYM) gj_invert_sym<LD, R2>(K, nr, prow);
          else gj_invert<LD, R2>(K, nr, prow);
          PH(7);
          if (tid < nr) s_v[tid] = rnd<R2>(-grp);
          __syncthreads();
          if (tid < nr) wred = rnd<R2>(matvec_smem<LD>(K, nr, s_v));
        }
        double xn = 0.0;
#pragma unroll 1
        for (int refine = 0; refine < 2; ++refine) {
          __syncthreads();
          if (tid < nr) s_w[tid] = wred;
          __syncthreads();
          if (act) {
            xn = s_c[tid];
            const int o = s_foff[fa];
            for (int q = 0; q < s_fnz[fa]; ++q) xn += s_fZ[9 * fa + 3 * q + ca] * s_w[o + q];
            xn = rnd<R2>(xn);
            s_xt[tid] = xn;
          }
          __syncthreads();
          if (act) s_r[tid] = rnd<R2>(matvec_glob<BS>(Hs, n, s_xt) + gi);  // r = H x + g
          __syncthreads();
          if (refine == 0) {  // one step of iterative refinement on the reduced system
            const double rr = (tid < nr) ? -(vp[0] * s_r[3 * ap] + vp[1] * s_r[3 * ap + 1] + vp[2] * s_r[3 * ap + 2]) : 0.0;
            if constexpr (TC || !B200QP_CHOL_POLISH) {
              if (tid < nr) s_v[tid] = rnd<R2>(rr);
              __syncthreads();
              if (tid < nr) wred = rnd<R2>(wred + matvec_smem<LD>(K, nr, s_v));
            } else {
              wred += chol_solve<LD>(K, nr, rr, red);
            }
          }
        }
        PH(8);
        // per-foot optimality check: feasibility, NNLS duals over the geometrically tight rows, new active set
        double chk[3] = {0.0, 0.0, 0.0};  // stat, viol, comp
        bool bad = false;
        double ynew[7] = {0, 0, 0, 0, 0, 0, 0};
        if (act && ca == 2) {
          const double fx = s_xt[3 * fa], fy = s_xt[3 * fa + 1], fz = s_xt[3 * fa + 2];
          double ax[7];
          foot_rows_eval(fx, fy, fz, mu, ax);
          const double sc = fmax(1.0, fmax(fabs(fx), fmax(fabs(fy), fabs(fz))));
          const double ttol = 1e-9 * sc;
          unsigned nl = 0, nu = 0, tl = 0, tu = 0;  // violated rows (lower / upper side), tight rows
#pragma unroll
          for (int r = 0; r < 7; ++r) {
            if (lo[r] > -BIG) {
              const double s = ax[r] - lo[r];
              if (s < -ttol) nl |= 1u << r;
              chk[1] = fmax(chk[1], -s);
              if (fabs(s) <= ttol) tl |= 1u << r;
            }
            if (hi[r] < BIG) {
              const double s = hi[r] - ax[r];
              if (s < -ttol) nu |= 1u << r;
              chk[1] = fmax(chk[1], -s);
              if (fabs(s) <= ttol) tu |= 1u << r;
            }
          }
          const double rx = s_r[3 * fa], ry = s_r[3 * fa + 1], rz = s_r[3 * fa + 2];
          if (((tl | tu) & 3u) == 0u) {
            // Fast path (no x/y box row tight - always the case for mu < 1): sign-feasible multipliers in closed form.
            //   r_x + a3 - a4 = 0,  r_y + b5 - b6 = 0,  r_z - mu (a3 + a4 + b5 + b6) - c_l + c_u = 0,  all >= 0 and
            //   non-zero only on tight rows; the minimal-sum split of the cone multipliers is the feasible one.
            const double a3 = (tu & 8u) ? fmax(-rx, 0.0) : 0.0, a4 = (tl & 16u) ? fmax(rx, 0.0) : 0.0;
            const double b5 = (tu & 32u) ? fmax(-ry, 0.0) : 0.0, b6 = (tl & 64u) ? fmax(ry, 0.0) : 0.0;
            const double t = rz - mu * ((a3 + a4) + (b5 + b6));
            const double cl = (tl & 4u) ? fmax(t, 0.0) : 0.0, cu = (tu & 4u) ? fmax(-t, 0.0) : 0.0;
            chk[0] = fmax(fabs(rx + a3 - a4), fmax(fabs(ry + b5 - b6), fabs(t - cl + cu)));
            ynew[2] = cu - cl;
            ynew[3] = a3; ynew[4] = -a4; ynew[5] = b5; ynew[6] = -b6;
            if (a3 > 0.0) nu |= 8u;
            if (a4 > 0.0) nl |= 16u;
            if (b5 > 0.0) nu |= 32u;
            if (b6 > 0.0) nl |= 64u;
            if (cl > 0.0) nl |= 4u;
            if (cu > 0.0) nu |= 4u;
            chk[2] = fmax(fmax(a3 * fabs(ax[3]), a4 * fabs(ax[4])), fmax(fmax(b5 * fabs(ax[5]), b6 * fabs(ax[6])),
                          fmax(cl * fabs(ax[2] - lo[2]), cu * fabs(hi[2] - ax[2]))));
            bad = !(chk[0] < BIG) || !(sc < BIG);
          } else {  // general case: Lawson-Hanson NNLS over the tight rows
            double Mc[14][3];
            int tagr[14];
            int m = 0;
            for (int r = 0; r < 7; ++r) {
              double a[3];
              foot_row_vec(r, mu, a);
              if ((tl >> r) & 1u) {
                Mc[m][0] = -a[0]; Mc[m][1] = -a[1]; Mc[m][2] = -a[2];
                tagr[m++] = -(r + 1);
              }
              if ((tu >> r) & 1u) {
                Mc[m][0] = a[0]; Mc[m][1] = a[1]; Mc[m][2] = a[2];
                tagr[m++] = (r + 1);
              }
            }
            const double b3[3] = {-rx, -ry, -rz};
            double coef[14], res[3] = {b3[0], b3[1], b3[2]};
            if (m > 0) nnls3(Mc, m, b3, coef, res);
            chk[0] = fmax(fabs(res[0]), fmax(fabs(res[1]), fabs(res[2])));
            bad = !(chk[0] < BIG) || !(sc < BIG);
            for (int q = 0; q < m; ++q)
              if (coef[q] > 0.0) {
                const int r = (tagr[q] > 0 ? tagr[q] : -tagr[q]) - 1;
                if (tagr[q] > 0) {
                  nu |= 1u << r;
                  ynew[r] += coef[q];
                  chk[2] = fmax(chk[2], coef[q] * fabs(hi[r] - ax[r]));
                } else {
                  nl |= 1u << r;
                  ynew[r] -= coef[q];
                  chk[2] = fmax(chk[2], coef[q] * fabs(ax[r] - lo[r]));
                }
              }
          }
          if (nl != s_actl[fa] || nu != s_actu[fa]) s_changed = 1;
          s_actl[fa] = (unsigned short)nl;
          s_actu[fa] = (unsigned short)nu;
        }
        const int anybad = __syncthreads_or(bad ? 1 : 0);
        block_max<3>(chk, red);
        if (anybad) break;
        if (chk[0] <= P.kkt_tol && chk[1] <= P.kkt_tol && chk[2] <= P.kkt_tol) {
          // Within tolerance.  An exact active set gives stationarity ~1e-13; a residual just under the tolerance means a
          // multiplier was clipped at zero, i.e. the set is one constraint off and the forces can still differ by 1e-5
          // relative along the directions only alpha penalises - keep the point as a candidate and let the repaired set have
          // one more round.
          if (!have_cand || chk[0] < kkt_stat) {
            have_cand = true;
            kkt_stat = chk[0];
            kkt_ineq = chk[1];
            kkt_comp = chk[2];
            xsol = xn;
#pragma unroll
            for (int r = 0; r < 7; ++r) ysol[r] = ynew[r];
          }
          if (chk[0] <= 1e-3 * P.kkt_tol || !s_changed) accepted = true;
        } else if (!s_changed) {
          break;  // active set is a fixed point but not optimal to tolerance: go back to ADMM
        }
      }
      if (!accepted && have_cand) accepted = true;
      if (accepted) {
        status = B200QP_ACADOS_SUCCESS;
      } else {
        if (iters >= P.max_iter) break;
        if (!(hot && attempt == 0)) eps *= 0.1;
        need_factor = true;  // K was overwritten by the polish
      }
    }

    PH(9);
    // ---- phase 4: outputs -------------------------------------------------------------------------------------
    __syncthreads();
    if (status == B200QP_ACADOS_SUCCESS) {
      if (act) s_xt[tid] = xsol;
      for (int e = tid; e < N * 36; e += BS) s_Bw[e] = Bw_park[e];
      __syncthreads();
      if (B.lam_box != nullptr) {
        double *lb = B.lam_box + (size_t)prob * N * 12;
        double *lg = B.lam_g + (size_t)prob * N * 16;
        for (int e = tid; e < N * 12; e += BS)
          if (s_fidx[e / 3] < 0) lb[e] = 0.0;
        for (int e = tid; e < N * 16; e += BS)
          if (s_fidx[e / 4] < 0) lg[e] = 0.0;
        if (act && ca == 2) {
          const int kf = s_fk[fa] * 4 + s_ff[fa];
          lb[kf * 3 + 0] = ysol[0];
          lb[kf * 3 + 1] = ysol[1];
          lb[kf * 3 + 2] = ysol[2];
          lg[kf * 4 + 0] = ysol[3];
          lg[kf * 4 + 1] = ysol[4];
          lg[kf * 4 + 2] = ysol[5];
          lg[kf * 4 + 3] = ysol[6];
        }
      }
      // per-stage input effect: d omega = sum_f Bw f, d v = dt/m sum_f f
      if (tid < N) {
        double dw[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
        for (int f = 0; f < 4; ++f) {
          const int fi = s_fidx[tid * 4 + f];
          if (fi < 0) continue;
          const double *Bw = s_Bw + (tid * 4 + f) * 9;
          const double u0 = s_xt[3 * fi], u1 = s_xt[3 * fi + 1], u2 = s_xt[3 * fi + 2];
          for (int r = 0; r < 3; ++r) dw[r] += Bw[3 * r] * u0 + Bw[3 * r + 1] * u1 + Bw[3 * r + 2] * u2;
          dv[0] += u0; dv[1] += u1; dv[2] += u2;
        }
        for (int r = 0; r < 3; ++r) {
          s_bw[6 * tid + r] = dw[r];
          s_bw[6 * tid + 3 + r] = dv[r] * (dt / mass);
        }
      }
      __syncthreads();
      if (tid == 0) {  // x_{k+1} = A x_k + B_k u_k (mpc.cpp:11-44 structure), N short steps
        double xs[13];
        for (int c = 0; c < 13; ++c) {
          xs[c] = s_x0[c];
          s_X[c] = xs[c];
        }
        for (int k = 0; k < N; ++k) {
          double nx[13];
          for (int c = 0; c < 3; ++c) {
            nx[c] = xs[c] + dt * (s_Rt[3 * c] * xs[6] + s_Rt[3 * c + 1] * xs[7] + s_Rt[3 * c + 2] * xs[8]);
            nx[3 + c] = xs[3 + c] + dt * xs[9 + c];
            nx[6 + c] = xs[6 + c] + s_bw[6 * k + c];
            nx[9 + c] = xs[9 + c] + s_bw[6 * k + 3 + c];
          }
          nx[11] -= dt * xs[12];
          nx[12] = xs[12];
          for (int c = 0; c < 13; ++c) {
            xs[c] = nx[c];
            s_X[13 * (k + 1) + c] = nx[c];
          }
        }
      }
      __syncthreads();
      double *xo = B.xpred + (size_t)prob * (N + 1) * 13;
      bool nanout = false;
      for (int e = tid; e < (N + 1) * 13; e += BS) {
        const double v = s_X[e];
        nanout |= !(fabs(v) < BIG);
      }
      if (act) nanout |= !(fabs(xsol) < BIG);
      if (__syncthreads_or(nanout ? 1 : 0)) {
        status = B200QP_NAN_AFTER_SUCCESS;  // mpc.cpp:447-455
      } else {  // outputs are only touched for a solved, finite problem (the buffers may be the caller's own host memory)
        double *fo = B.forces + (size_t)prob * N * 12;
        for (int e = tid; e < N * 12; e += BS) {
          const int kf = e / 3, c = e - 3 * kf;
          const int fi = s_fidx[kf];
          fo[e] = (fi >= 0) ? s_xt[3 * fi + c] : 0.0;
        }
        for (int e = tid; e < (N + 1) * 13; e += BS) xo[e] = s_X[e];
      }
    }
    if (B.warm_mode != 0) {  // remember this solution for the next call on the same slot
      if (status == B200QP_ACADOS_SUCCESS) {
        double *wx = B.warm_x + (size_t)prob * N * 12;
        double *wy = B.warm_y + (size_t)prob * N * 28;
        for (int e = tid; e < N * 12; e += BS) {
          const int fi = s_fidx[e / 3];
          wx[e] = (fi >= 0) ? s_xt[3 * fi + (e - 3 * (e / 3))] : 0.0;
        }
        for (int e = tid; e < N * 28; e += BS)
          if (s_fidx[e / 7] < 0) wy[e] = 0.0;
        if (act && ca == 2) {
          double *w7 = wy + (s_fk[fa] * 4 + s_ff[fa]) * 7;
#pragma unroll
          for (int r = 0; r < 7; ++r) w7[r] = ysol[r];
        }
      }
      if (tid == 0) B.warm_valid[prob] = (status == B200QP_ACADOS_SUCCESS) ? 1 : 0;
    }
    if (timing && tid == 0)
      for (int q = 0; q < 10; ++q) B.dbg_cycles[(size_t)prob * 10 + q] = cyc[q];
    if (tid == 0) {
      b200qp_solver_info inf;
      inf.status = status;
      inf.iterations = iters;
      inf.polish_rounds = rounds_total;
      inf.n_free = n;
      inf.kkt[0] = kkt_stat;
      inf.kkt[1] = 0.0;  // dynamics hold by construction of the roll-out
      inf.kkt[2] = kkt_ineq;
      inf.kkt[3] = kkt_comp;
      if (!(B.refine != 0 && status != B200QP_ACADOS_SUCCESS)) B.info[prob] = inf;  // a failed second opinion leaves the first answer's record
      if (B.retry_list != nullptr && (status == B200QP_ACADOS_MAXITER || status == B200QP_ACADOS_MINSTEP))
        B.retry_list[atomicAdd(B.retry_count, 1)] = prob;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// binning by reduced problem size
__global__ void mpc_bin_kernel(int n, int N, const uint8_t *__restrict__ contact, int *bin_count, int *bin_lists,
                               b200qp_solver_info *info, int overflow_bin) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint8_t *ct = contact + (size_t)p * N * 4;
  int cnt = 0;
  for (int e = 0; e < 4 * N; ++e) cnt += ct[e] != 0;
  const int nv = 3 * cnt;
  int bin;
  if (nv <= 60) bin = 0;
  else if (nv <= 96) bin = 1;
  else if (nv <= 128) bin = 2;
  else if (nv <= 156) bin = 3;
  else bin = 4;
  if (bin == 4 && !overflow_bin) {
    b200qp_solver_info inf = {B200QP_ACADOS_QP_FAILURE, 0, 0, nv, {0, 0, 0, 0}};
    info[p] = inf;
    return;
  }
  const int pos = atomicAdd(bin_count + bin, 1);
  bin_lists[(size_t)bin * n + pos] = p;
}

// the same binning for a list of problem indices (second opinion for the interior-point path); sizes the condensed kernel cannot take are dropped
__global__ void mpc_bin_list_kernel(const int *__restrict__ src, const int *__restrict__ src_count, int n, int N,
                                    const uint8_t *__restrict__ contact, int *bin_count, int *bin_lists) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= *src_count) return;
  const int p = src[k];
  const uint8_t *ct = contact + (size_t)p * N * 4;
  int cnt = 0;
  for (int e = 0; e < 4 * N; ++e) cnt += ct[e] != 0;
  const int nv = 3 * cnt;
  const int bin = nv <= 60 ? 0 : nv <= 96 ? 1 : nv <= 128 ? 2 : nv <= 156 ? 3 : 4;
  if (bin == 4) return;
  bin_lists[(size_t)bin * n + atomicAdd(bin_count + bin, 1)] = p;
}

template <int BS, int NV>
static size_t admm_smem_bytes() {
  return sizeof(double) * ((size_t)NV * (NV + 2) + 6 * BS + 4 * 32);
}

template <int BS, int NV, int PREC = 0>
static cudaError_t launch_one(const MpcParams &P, MpcBuffers B, int grid, cudaStream_t st) {
  const size_t sm = admm_smem_bytes<BS, NV>();
  cudaError_t e = cudaFuncSetAttribute(mpc_admm_kernel<BS, NV, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  mpc_admm_kernel<BS, NV, PREC><<<grid, BS, sm, st>>>(P, B);
  return cudaGetLastError();
}

template <int BS, int NV>
static int occupancy_one() {
  const size_t sm = admm_smem_bytes<BS, NV>();
  int occ = 0;
  if (cudaFuncSetAttribute(mpc_admm_kernel<BS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mpc_admm_kernel<BS, NV>, BS, sm) != cudaSuccess) {
    cudaGetLastError();  // clear
    return 1;
  }
  return occ < 1 ? 1 : occ;
}
// Instantiations (threads, max reduced size): trot at N = 10 is exactly n = 60 -> 6 resident blocks per SM; n = 96 is trot / pace /
// bound at N = 16 and three-leg stance gaits at N = 10.
int admm_occupancy(int bin) {
  return bin == 0 ? occupancy_one<64, 60>() : bin == 1 ? occupancy_one<96, 96>() : bin == 2 ? occupancy_one<128, 128>()
                                                                                           : occupancy_one<160, 156>();
}
int admm_block_size(int bin) { return bin == 0 ? 64 : bin == 1 ? 96 : bin == 2 ? 128 : 160; }

cudaError_t launch_mpc_bin(int n, int N, const uint8_t *contact, int *bin_count, int *bin_lists,
                           b200qp_solver_info *info, int overflow_bin, cudaStream_t st) {
  mpc_bin_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, N, contact, bin_count, bin_lists, info, overflow_bin);
  return cudaGetLastError();
}

cudaError_t launch_mpc_bin_list(const int *src, const int *src_count, int n, int N, const uint8_t *contact, int *bin_count, int *bin_lists,
                                cudaStream_t st) {
  mpc_bin_list_kernel<<<(n + 127) / 128, 128, 0, st>>>(src, src_count, n, N, contact, bin_count, bin_lists);
  return cudaGetLastError();
}

cudaError_t launch_mpc_admm(int bin, const MpcParams &P, const MpcBuffers &B, int grid, cudaStream_t st) {
  if (P.precision == 1) {  // FP32 error study (trot-sized and mid-sized bins only)
    if (bin == 0) return launch_one<64, 60, 1>(P, B, grid, st);
    if (bin == 1) return launch_one<96, 96, 1>(P, B, grid, st);
  } else if (P.precision == 2) {
    if (bin == 0) return launch_one<64, 60, 2>(P, B, grid, st);
    if (bin == 1) return launch_one<96, 96, 2>(P, B, grid, st);
  }
  if (B.dbg_cycles != nullptr) {  // phase accounting (scripts/phase_cycles.py)
    if (bin == 0) return launch_one<64, 60, 3>(P, B, grid, st);
    if (bin == 1) return launch_one<96, 96, 3>(P, B, grid, st);
    if (bin == 2) return launch_one<128, 128, 3>(P, B, grid, st);
    return launch_one<160, 156, 3>(P, B, grid, st);
  }
  if (bin == 0) return launch_one<64, 60>(P, B, grid, st);
  if (bin == 1) return launch_one<96, 96>(P, B, grid, st);
  if (bin == 2) return launch_one<128, 128>(P, B, grid, st);
  return launch_one<160, 156>(P, B, grid, st);
}

}  // namespace b200qp
