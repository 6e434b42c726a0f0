/*
 * mpc_ref.c - C restatement of the reference's CPU path for the MPC QP (TEST / MEASUREMENT INFRASTRUCTURE ONLY).
 *
 * PINNING: the QP DATA this file builds (1) is checked against the compiled reference (oracle/_ref: mpc.cpp built from
 * /root/reference against stand-in headers; tests/test_ref.py: A, x0, bounds, D byte-equal, B <= 4e-17).  The SOLVER
 * arithmetic (2)-(3) is third party in the reference (acados v0.3.6 -> HPIPM condensing + OSQP 0.6.3;
 * /root/reference/docker/Dockerfile:74-76), absent from /root/reference and not buildable here, and the reference ships no
 * fixtures for it: that part is "parity unpinned" in the strict sense and anchored on uniqueness of the optimum + the
 * independent KKT checkers.  This file restates
 *   (1) the QP data exactly as ws/src/controllers/src/mit_controller/mpc.cpp:11-115,332-398 builds it,
 *   (2) full condensing (what HPIPM's d_cond_qp computes: X = Phi x0 + Gamma U, H = Gamma'QGamma + R, g),
 *   (3) the published OSQP algorithm (Stellato et al., "OSQP: an operator splitting solver for quadratic programs",
 *       Alg. 1 + sec. 5.2 adaptive rho + sec. 4 polish) on the reference's own formulation: 12N variables, swing legs
 *       pinned through lbu = ubu = 0 (mpc.cpp:110-113), 12N box rows + 16N pyramid rows (mpc.cpp:66-92),
 *       rho_eq = 1e3 rho on equality rows, sigma = 1e-6, alpha = 1.6, check_termination = 25.
 * It is the `cpu_baseline` / `--impl reference` arm of bench.py (one problem per thread, OpenMP over the batch) and a
 * second, independent solver for the parity tests.  It is never linked into or called from the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NX 13
#define NU 12
#define MAXN 20
#define MAXV (NU * MAXN)
#define MAXM (28 * MAXN)
#define INF_B 99999999.0 /* mpc.hpp:19 */
#define INF_T 9.0e7

typedef struct {
  int N;
  double dt, alpha, w[12], mu, fmin, fmax;
  double rho, sigma, relax, eps_abs, eps_rel, kkt_tol;
  int max_iter, check_every, polish, adaptive_rho;
} ref_params;

typedef struct {
  int status; /* 0 solved to kkt_tol, 2 max iter, 1 nan */
  int iters, n_fact, polished;
  double kkt[4];
} ref_info;

/* ---- small helpers ---------------------------------------------------------------------------------------------- */
static void quat_to_rot(const double *q, double R[3][3]) { /* Eigen toRotationMatrix, (x,y,z,w) */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}
static void rot_to_euler(double R[3][3], double e[3]) { /* quaternion_operations.hpp:106-124 */
  const double eps = 2.220446049250313e-16;
  double m20 = R[2][0], yaw, roll, pitch = -asin(m20);
  if (1.0 - eps <= m20 && m20 <= 1.0 + eps) { yaw = 0; roll = atan2(-R[0][1], -R[0][2]); }
  else if (m20 <= -1.0 + eps) { yaw = 0; roll = atan2(R[0][1], R[0][2]); }
  else { yaw = atan2(R[1][0], R[0][0]); roll = atan2(R[2][1], R[2][2]); }
  e[0] = roll; e[1] = pitch; e[2] = yaw;
}
static void inv3(double A[3][3], double Ai[3][3]) {
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double id = 1.0 / (A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02);
  Ai[0][0] = c00 * id; Ai[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id; Ai[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
  Ai[1][0] = c01 * id; Ai[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id; Ai[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  Ai[2][0] = c02 * id; Ai[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id; Ai[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
}

/* in-place lower Cholesky of the n x n row-major matrix; returns 0 on success */
static int chol(double *A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 0)) return 1;
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * ld + j];
      const double *ai = A + i * ld, *aj = A + j * ld;
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[i * ld + j] = s / d;
    }
  }
  return 0;
}
static void chol_solve(const double *L, int n, int ld, double *b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    const double *li = L + i * ld;
    for (int k = 0; k < i; ++k) s -= li[k] * b[k];
    b[i] = s / li[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    const double *li = L + i * ld;
    const double bi = b[i] / li[i];
    b[i] = bi;
    for (int k = 0; k < i; ++k) b[k] -= li[k] * bi;
  }
}

/* ---- per-problem workspace ---------------------------------------------------------------------------------------- */
typedef struct {
  double A[NX * NX];          /* row-major */
  double B[MAXN][NX * NU];    /* row-major */
  double xref[MAXN + 1][NX], q[MAXN + 1][NX], x0[NX]; /* q_k = -Q xref_k (mpc.cpp:380) */
  double lbu[MAXN][NU], ubu[MAXN][NU];
  double H[MAXV * MAXV], g[MAXV], K[MAXV * MAXV];
  double Gam[MAXN][MAXN][NX * NU]; /* Gamma block (i,j): state i+1 <- input j */
  double lo[MAXM], hi[MAXM], rho_v[MAXM];
  double x[MAXV], z[MAXM], y[MAXM], xt[MAXV], zt[MAXM], rhs[MAXV], tmp[MAXM];
  double xp[MAXV], yp[MAXM];
  uint8_t al[MAXM], au[MAXM];
} ref_ws;

/* A x for the inequality operator: rows [0,12N) identity, rows 12N + 16k + 4f + r = D (mpc.cpp:66-71) */
static void apply_A(int N, double mu, const double *x, double *ax) {
  const int n = NU * N;
  for (int i = 0; i < n; ++i) ax[i] = x[i];
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      const double *u = x + NU * k + 3 * f;
      double *o = ax + n + 16 * k + 4 * f;
      o[0] = u[0] - mu * u[2];
      o[1] = u[0] + mu * u[2];
      o[2] = u[1] - mu * u[2];
      o[3] = u[1] + mu * u[2];
    }
}
static void apply_AT(int N, double mu, const double *y, double *aty) {
  const int n = NU * N;
  for (int i = 0; i < n; ++i) aty[i] = y[i];
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      const double *o = y + n + 16 * k + 4 * f;
      double *u = aty + NU * k + 3 * f;
      u[0] += o[0] + o[1];
      u[1] += o[2] + o[3];
      u[2] += mu * (-o[0] + o[1] - o[2] + o[3]);
    }
}
/* K = H + sigma I + A' diag(rho_v) A  (lower triangle is enough for chol) */
static void build_K(int N, double mu, double sigma, const double *H, const double *rho_v, double *K) {
  const int n = NU * N;
  memcpy(K, H, sizeof(double) * (size_t)n * n);
  for (int i = 0; i < n; ++i) K[i * n + i] += sigma + rho_v[i];
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      const double *r = rho_v + n + 16 * k + 4 * f;
      const int b = NU * k + 3 * f;
      K[b * n + b] += r[0] + r[1];
      K[(b + 1) * n + b + 1] += r[2] + r[3];
      K[(b + 2) * n + b + 2] += mu * mu * (r[0] + r[1] + r[2] + r[3]);
      const double xz = mu * (-r[0] + r[1]), yz = mu * (-r[2] + r[3]);
      K[(b + 2) * n + b] += xz; K[b * n + b + 2] += xz;
      K[(b + 2) * n + b + 1] += yz; K[(b + 1) * n + b + 2] += yz;
    }
}

/* (1) QP data, mpc.cpp:332-398 */
static void build_qp(const ref_params *P, const double *rec, const uint8_t *ct, ref_ws *W) {
  const int N = P->N;
  const double dt = P->dt;
  double R[3][3], e[3];
  quat_to_rot(rec, R);
  double psi_avg = atan2(R[1][0], R[0][0]);
  double yaw_before = psi_avg;
  const double *com = rec + 23;
  const double m = rec[13];
  double Ib[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ib[r][c] = rec[14 + 3 * c + r];
  /* MPC::GetA mpc.cpp:11-18 with GetR :46-58 */
  memset(W->A, 0, sizeof(W->A));
  for (int i = 0; i < NX; ++i) W->A[i * NX + i] = 1.0;
  for (int i = 0; i < 3; ++i) W->A[(3 + i) * NX + 9 + i] = dt;
  W->A[11 * NX + 12] = -dt;
  W->A[0 * NX + 6] = cos(psi_avg) * dt; W->A[0 * NX + 7] = sin(psi_avg) * dt;
  W->A[1 * NX + 6] = -sin(psi_avg) * dt; W->A[1 * NX + 7] = cos(psi_avg) * dt;
  W->A[2 * NX + 8] = dt;
  for (int k = 0; k <= N; ++k) {
    const double *st = rec + 26 + 13 * k;
    quat_to_rot(st, R);
    double desired_yaw = atan2(R[1][0], R[0][0]);
    double diff = desired_yaw - yaw_before;
    if (diff < -M_PI) desired_yaw += 2 * M_PI; else if (diff > M_PI) desired_yaw -= 2 * M_PI;
    yaw_before = desired_yaw;
    if (k < N) {
      double *B = W->B[k];
      memset(B, 0, sizeof(double) * NX * NU);
      for (int f = 0; f < 4; ++f) for (int i = 0; i < 3; ++i) B[(9 + i) * NU + 3 * f + i] = dt / m;
      double cs = cos(desired_yaw), sn = sin(desired_yaw);
      double Rt[3][3] = {{cs, sn, 0}, {-sn, cs, 0}, {0, 0, 1}}, T[3][3], Iw[3][3], Ii[3][3];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) T[r][c] = Rt[r][0] * Ib[0][c] + Rt[r][1] * Ib[1][c] + Rt[r][2] * Ib[2][c];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Iw[r][c] = T[r][0] * Rt[c][0] + T[r][1] * Rt[c][1] + T[r][2] * Rt[c][2];
      inv3(Iw, Ii);
      const double *fp = rec + 26 + 13 * (N + 1) + 12 * k;
      for (int f = 0; f < 4; ++f) {
        const int on = ct[4 * k + f] != 0;
        if (on) {
          double rx = fp[3 * f] - (st[4] + com[0]), ry = fp[3 * f + 1] - (st[5] + com[1]), rz = fp[3 * f + 2] - (st[6] + com[2]);
          double S[3][3] = {{0, -rz, ry}, {rz, 0, -rx}, {-ry, rx, 0}};
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
            B[(6 + r) * NU + 3 * f + c] = (Ii[r][0] * S[0][c] + Ii[r][1] * S[1][c] + Ii[r][2] * S[2][c]) * dt;
        }
        for (int i = 0; i < 3; ++i) { /* MPC::GetUBounds mpc.cpp:94-115 */
          W->lbu[k][3 * f + i] = on ? (i == 2 ? P->fmin : -P->fmax) : 0.0;
          W->ubu[k][3 * f + i] = on ? P->fmax : 0.0;
        }
      }
    }
    quat_to_rot(st, R);
    rot_to_euler(R, e);
    double *t = W->xref[k];
    t[0] = e[0]; t[1] = e[1]; t[2] = desired_yaw;
    for (int i = 0; i < 3; ++i) { t[3 + i] = st[4 + i] + com[i]; t[6 + i] = st[7 + i]; t[9 + i] = st[10 + i]; }
    t[12] = 0;
    for (int i = 0; i < NX; ++i) W->q[k][i] = (i < 12) ? -(P->w[i] * t[i]) : 0.0; /* qs_[k] = -(Q_ * target_state), mpc.cpp:380 */
  }
  quat_to_rot(rec, R);
  rot_to_euler(R, e);
  for (int i = 0; i < 3; ++i) { W->x0[i] = e[i]; W->x0[3 + i] = rec[4 + i] + com[i]; W->x0[6 + i] = rec[7 + i]; W->x0[9 + i] = rec[10 + i]; }
  W->x0[12] = 9.8067; /* mpc.hpp:22 */
}

/* (2) full condensing */
static void condense(const ref_params *P, ref_ws *W) {
  const int N = P->N, n = NU * N;
  double Q[NX];
  for (int i = 0; i < 12; ++i) Q[i] = P->w[i];
  Q[12] = 0;
  /* Gamma(i,j) = A^(i-j) B_j, i >= j (state index i+1) */
  for (int j = 0; j < N; ++j) {
    memcpy(W->Gam[j][j], W->B[j], sizeof(double) * NX * NU);
    for (int i = j + 1; i < N; ++i) {
      const double *prev = W->Gam[i - 1][j];
      double *cur = W->Gam[i][j];
      for (int r = 0; r < NX; ++r)
        for (int c = 0; c < NU; ++c) {
          double s = 0;
          for (int t = 0; t < NX; ++t) s += W->A[r * NX + t] * prev[t * NU + c];
          cur[r * NU + c] = s;
        }
    }
  }
  /* free response error and g = Gamma' Q (Phi x0 - xref) */
  double xs[NX], e[MAXN][NX];
  memcpy(xs, W->x0, sizeof(xs));
  for (int i = 0; i < N; ++i) {
    double nx[NX];
    for (int r = 0; r < NX; ++r) { double s = 0; for (int t = 0; t < NX; ++t) s += W->A[r * NX + t] * xs[t]; nx[r] = s; }
    memcpy(xs, nx, sizeof(xs));
    for (int r = 0; r < NX; ++r) e[i][r] = Q[r] * xs[r] + W->q[i + 1][r];
  }
  for (int j = 0; j < N; ++j)
    for (int c = 0; c < NU; ++c) {
      double s = 0;
      for (int i = j; i < N; ++i) { const double *G = W->Gam[i][j]; for (int r = 0; r < NX; ++r) s += G[r * NU + c] * e[i][r]; }
      W->g[NU * j + c] = s;
    }
  for (int j = 0; j < N; ++j)
    for (int l = 0; l <= j; ++l)
      for (int a = 0; a < NU; ++a)
        for (int b = 0; b < NU; ++b) {
          double s = 0;
          for (int i = j; i < N; ++i) {
            const double *Gj = W->Gam[i][j], *Gl = W->Gam[i][l];
            for (int r = 0; r < 12; ++r) s += Gj[r * NU + a] * Q[r] * Gl[r * NU + b];
          }
          W->H[(NU * j + a) * n + NU * l + b] = s;
          W->H[(NU * l + b) * n + NU * j + a] = s;
        }
  for (int i = 0; i < n; ++i) W->H[i * n + i] += P->alpha;
  const int mrows = 28 * N;
  for (int k = 0; k < N; ++k) {
    for (int i = 0; i < NU; ++i) { W->lo[NU * k + i] = W->lbu[k][i]; W->hi[NU * k + i] = W->ubu[k][i]; }
    for (int f = 0; f < 4; ++f) { /* MPC::GetCbounds mpc.cpp:73-92 */
      double *l = W->lo + n + 16 * k + 4 * f, *h = W->hi + n + 16 * k + 4 * f;
      l[0] = -INF_B; h[0] = 0; l[1] = 0; h[1] = INF_B; l[2] = -INF_B; h[2] = 0; l[3] = 0; h[3] = INF_B;
    }
  }
  for (int i = 0; i < mrows; ++i) { if (W->lo[i] <= -INF_T) W->lo[i] = -HUGE_VAL; if (W->hi[i] >= INF_T) W->hi[i] = HUGE_VAL; }
}

static void set_rho(const ref_params *P, ref_ws *W, double rho) {
  const int m = 28 * P->N;
  for (int i = 0; i < m; ++i) W->rho_v[i] = (W->lo[i] == W->hi[i]) ? 1e3 * rho : rho;
}
static double infnorm(const double *v, int n) { double s = 0; for (int i = 0; i < n; ++i) { double a = fabs(v[i]); if (a > s) s = a; } return s; }

/* ---- polish (OSQP sec. 4 idea: solve the equality-constrained QP on the guessed active set, then verify) -------------
 * Implemented in the null space of the active rows: every inequality row touches one foot only, so the active rows of a
 * foot are reduced (3 columns, rref) to a particular solution + null-space basis; the reduced Hessian Z'HZ is solved by
 * Cholesky with one refinement step; duals are recovered per foot by non-negative least squares over the tight rows and
 * the active set is repaired (primal-dual active-set sweeps) until feasible and sign-correct. */
static void foot_row(int r, double mu, double a[3]) { /* 0..2 boxes x,y,z; 3..6 = D rows (mpc.cpp:66-71) */
  a[0] = (r == 0 || r == 3 || r == 4); a[1] = (r == 1 || r == 5 || r == 6);
  a[2] = (r == 2) ? 1.0 : (r == 3 || r == 5) ? -mu : (r == 4 || r == 6) ? mu : 0.0;
}
static int foot_nullspace(unsigned al, unsigned au, double mu, const double lo[7], const double hi[7], double cf[3], double Z[9]) {
  double M[7][4];
  int m = 0, piv[3], np = 0;
  for (int r = 0; r < 7; ++r)
    if (((al | au) >> r) & 1u) { foot_row(r, mu, M[m]); M[m][3] = ((al >> r) & 1u) ? lo[r] : hi[r]; ++m; }
  for (int col = 0; col < 3 && np < m; ++col) {
    int p = np;
    for (int q = np + 1; q < m; ++q) if (fabs(M[q][col]) > fabs(M[p][col])) p = q;
    if (fabs(M[p][col]) < 1e-12) continue;
    for (int t = 0; t < 4; ++t) { double tmp = M[np][t]; M[np][t] = M[p][t]; M[p][t] = tmp; }
    double inv = 1.0 / M[np][col];
    for (int t = 0; t < 4; ++t) M[np][t] *= inv;
    for (int q = 0; q < m; ++q) if (q != np) { double f = M[q][col]; for (int t = 0; t < 4; ++t) M[q][t] -= f * M[np][t]; }
    piv[np++] = col;
  }
  cf[0] = cf[1] = cf[2] = 0;
  for (int i = 0; i < np; ++i) cf[piv[i]] = M[i][3];
  int nz = 0;
  for (int col = 0; col < 3; ++col) {
    int isp = 0;
    for (int i = 0; i < np; ++i) isp |= piv[i] == col;
    if (isp) continue;
    double *v = Z + 3 * nz++;
    v[0] = v[1] = v[2] = 0; v[col] = 1;
    for (int i = 0; i < np; ++i) v[piv[i]] = -M[i][col];
  }
  return nz;
}
/* NNLS on 3 x m by enumeration of column subsets of size <= 3 (m <= 8): exact for these tiny problems */
static double nnls_enum(double (*Mc)[3], int m, const double b[3], double *coef) {
  double best = fmax(fabs(b[0]), fmax(fabs(b[1]), fabs(b[2])));
  for (int j = 0; j < m; ++j) coef[j] = 0;
  for (unsigned mask = 1; mask < (1u << m); ++mask) {
    int idx[3], k = 0, pc = __builtin_popcount(mask);
    if (pc > 3) continue;
    for (int j = 0; j < m; ++j) if ((mask >> j) & 1u) idx[k++] = j;
    double G[3][4];
    for (int i = 0; i < k; ++i) {
      for (int j = 0; j < k; ++j) G[i][j] = Mc[idx[i]][0] * Mc[idx[j]][0] + Mc[idx[i]][1] * Mc[idx[j]][1] + Mc[idx[i]][2] * Mc[idx[j]][2];
      G[i][3] = Mc[idx[i]][0] * b[0] + Mc[idx[i]][1] * b[1] + Mc[idx[i]][2] * b[2];
    }
    int ok = 1;
    for (int c = 0; c < k && ok; ++c) {
      int p = c;
      for (int q = c + 1; q < k; ++q) if (fabs(G[q][c]) > fabs(G[p][c])) p = q;
      if (fabs(G[p][c]) < 1e-13) { ok = 0; break; }
      for (int t = 0; t < 4; ++t) { double tmp = G[c][t]; G[c][t] = G[p][t]; G[p][t] = tmp; }
      for (int q = 0; q < k; ++q) if (q != c) { double f = G[q][c] / G[c][c]; for (int t = c; t < 4; ++t) G[q][t] -= f * G[c][t]; }
    }
    if (!ok) continue;
    double s[3], res[3] = {b[0], b[1], b[2]};
    for (int i = 0; i < k; ++i) { s[i] = G[i][3] / G[i][i]; if (!(s[i] >= 0)) ok = 0; }
    if (!ok) continue;
    for (int i = 0; i < k; ++i) for (int t = 0; t < 3; ++t) res[t] -= s[i] * Mc[idx[i]][t];
    double r = fmax(fabs(res[0]), fmax(fabs(res[1]), fabs(res[2])));
    if (r < best) { best = r; for (int j = 0; j < m; ++j) coef[j] = 0; for (int i = 0; i < k; ++i) coef[idx[i]] = s[i]; }
  }
  return best;
}

static int polish(const ref_params *P, ref_ws *W, double *kkt) {
  const int N = P->N, n = NU * N;
  const double mu = P->mu;
  double lo[7] = {-P->fmax, -P->fmax, P->fmin, -HUGE_VAL, 0, -HUGE_VAL, 0}, hi[7] = {P->fmax, P->fmax, P->fmax, 0, HUGE_VAL, 0, HUGE_VAL};
  /* stance feet (k,f) and their rows in the full numbering: boxes 12k+3f+c, pyramids 12N+16k+4f+r */
  int fk[4 * MAXN], nf = 0;
  for (int e = 0; e < 4 * N; ++e) if (W->lo[3 * e] != W->hi[3 * e] || W->lo[3 * e + 2] != W->hi[3 * e + 2]) fk[nf++] = e;
  unsigned al[4 * MAXN], au[4 * MAXN];
  for (int a = 0; a < nf; ++a) {
    const int e = fk[a];
    al[a] = au[a] = 0;
    for (int r = 0; r < 7; ++r) {
      const int row = r < 3 ? 3 * e + r : n + 4 * e + (r - 3);
      if (lo[r] > -HUGE_VAL && (W->z[row] - lo[r]) < -W->y[row]) al[a] |= 1u << r;
      else if (hi[r] < HUGE_VAL && (hi[r] - W->z[row]) < W->y[row]) au[a] |= 1u << r;
    }
  }
  static __thread double Zf[4 * MAXN][9], cf[4 * MAXN][3];
  static __thread int nzf[4 * MAXN], off[4 * MAXN + 1];
  double *Hr = W->K; /* reuse */
  for (int round = 0; round < 8; ++round) {
    int nr = 0;
    memset(W->xp, 0, sizeof(double) * n);
    for (int a = 0; a < nf; ++a) {
      nzf[a] = foot_nullspace(al[a], au[a], mu, lo, hi, cf[a], Zf[a]);
      off[a] = nr; nr += nzf[a];
      for (int c = 0; c < 3; ++c) W->xp[3 * fk[a] + c] = cf[a][c];
    }
    off[nf] = nr;
    /* hc = H c + g */
    for (int i = 0; i < n; ++i) { double s = W->g[i]; const double *h = W->H + (size_t)i * n; for (int j = 0; j < n; ++j) s += h[j] * W->xp[j]; W->rhs[i] = s; }
    /* Hr = Z'HZ, gr = Z'hc */
    int pa = 0;
    for (int a = 0; a < nf; ++a) for (int q = 0; q < nzf[a]; ++q, ++pa) {
      const double *vp = Zf[a] + 3 * q;
      int pb = 0;
      for (int b = 0; b < nf; ++b) for (int q2 = 0; q2 < nzf[b]; ++q2, ++pb) {
        if (pb > pa) continue;
        const double *vq = Zf[b] + 3 * q2;
        double s = 0;
        for (int c = 0; c < 3; ++c) for (int c2 = 0; c2 < 3; ++c2) s += vp[c] * W->H[(size_t)(3 * fk[a] + c) * n + 3 * fk[b] + c2] * vq[c2];
        Hr[pa * nr + pb] = s;
      }
      W->tmp[pa] = -(vp[0] * W->rhs[3 * fk[a]] + vp[1] * W->rhs[3 * fk[a] + 1] + vp[2] * W->rhs[3 * fk[a] + 2]);
    }
    if (nr > 0 && chol(Hr, nr, nr)) return 0;
    double wr[MAXV];
    memcpy(wr, W->tmp, sizeof(double) * nr);
    chol_solve(Hr, nr, nr, wr);
    for (int refine = 0; refine < 2; ++refine) {
      for (int a = 0; a < nf; ++a) for (int c = 0; c < 3; ++c) {
        double s = cf[a][c];
        for (int q = 0; q < nzf[a]; ++q) s += Zf[a][3 * q + c] * wr[off[a] + q];
        W->xp[3 * fk[a] + c] = s;
      }
      for (int i = 0; i < n; ++i) { double s = W->g[i]; const double *h = W->H + (size_t)i * n; for (int j = 0; j < n; ++j) s += h[j] * W->xp[j]; W->rhs[i] = s; }
      if (refine == 0 && nr > 0) {
        double d[MAXV];
        int p2 = 0;
        for (int a = 0; a < nf; ++a) for (int q = 0; q < nzf[a]; ++q, ++p2) {
          const double *vp = Zf[a] + 3 * q;
          d[p2] = -(vp[0] * W->rhs[3 * fk[a]] + vp[1] * W->rhs[3 * fk[a] + 1] + vp[2] * W->rhs[3 * fk[a] + 2]);
        }
        chol_solve(Hr, nr, nr, d);
        for (int i = 0; i < nr; ++i) wr[i] += d[i];
      }
    }
    double stat = 0, viol = 0, comp = 0;
    int changed = 0;
    memset(W->yp, 0, sizeof(double) * 28 * N);
    for (int a = 0; a < nf; ++a) {
      const int e = fk[a];
      const double *xf = W->xp + 3 * e;
      double ax[7] = {xf[0], xf[1], xf[2], xf[0] - mu * xf[2], xf[0] + mu * xf[2], xf[1] - mu * xf[2], xf[1] + mu * xf[2]};
      double sc = fmax(1.0, fmax(fabs(xf[0]), fmax(fabs(xf[1]), fabs(xf[2])))), tt = 1e-9 * sc;
      double Mc[14][3], coef[14];
      int tag[14], m = 0;
      unsigned nl = 0, nu = 0;
      for (int r = 0; r < 7; ++r) {
        double av[3];
        foot_row(r, mu, av);
        if (lo[r] > -HUGE_VAL) {
          double s = ax[r] - lo[r];
          if (s < -tt) nl |= 1u << r;
          viol = fmax(viol, -s);
          if (fabs(s) <= tt && m < 8) { Mc[m][0] = -av[0]; Mc[m][1] = -av[1]; Mc[m][2] = -av[2]; tag[m++] = -(r + 1); }
        }
        if (hi[r] < HUGE_VAL) {
          double s = hi[r] - ax[r];
          if (s < -tt) nu |= 1u << r;
          viol = fmax(viol, -s);
          if (fabs(s) <= tt && m < 8) { Mc[m][0] = av[0]; Mc[m][1] = av[1]; Mc[m][2] = av[2]; tag[m++] = r + 1; }
        }
      }
      double b3[3] = {-W->rhs[3 * e], -W->rhs[3 * e + 1], -W->rhs[3 * e + 2]};
      stat = fmax(stat, nnls_enum(Mc, m, b3, coef));
      for (int q = 0; q < m; ++q) if (coef[q] > 0) {
        int r = abs(tag[q]) - 1;
        const int row = r < 3 ? 3 * e + r : n + 4 * e + (r - 3);
        if (tag[q] > 0) { nu |= 1u << r; W->yp[row] += coef[q]; comp = fmax(comp, coef[q] * fabs(hi[r] - ax[r])); }
        else { nl |= 1u << r; W->yp[row] -= coef[q]; comp = fmax(comp, coef[q] * fabs(ax[r] - lo[r])); }
      }
      if (nl != al[a] || nu != au[a]) changed = 1;
      al[a] = nl; au[a] = nu;
    }
    if (stat <= P->kkt_tol && viol <= P->kkt_tol && comp <= P->kkt_tol) {
      kkt[0] = stat; kkt[1] = 0; kkt[2] = viol; kkt[3] = comp;
      return 1;
    }
    if (!changed) return 0;
  }
  return 0;
}

/* (3) OSQP-style ADMM on the condensed QP */
static void solve_ws(const ref_params *P, ref_ws *W, double *forces, double *xpred, ref_info *info);
static void solve_one(const ref_params *P, const double *rec, const uint8_t *ct, ref_ws *W, double *forces, double *xpred,
                      ref_info *info) {
  build_qp(P, rec, ct, W);
  solve_ws(P, W, forces, xpred, info);
}
/* condensing + ADMM + polish on the stage data held in W (A, B_k, q_k, x0, lbu_k, ubu_k) */
static void solve_ws(const ref_params *P, ref_ws *W, double *forces, double *xpred, ref_info *info) {
  const int N = P->N, n = NU * N, m = 28 * N;
  const double mu = P->mu, sigma = P->sigma, al = P->relax;
  condense(P, W);
  double rho = P->rho;
  set_rho(P, W, rho);
  build_K(N, mu, sigma, W->H, W->rho_v, W->K);
  memset(info, 0, sizeof(*info));
  info->status = 2;
  if (chol(W->K, n, n)) { info->status = 1; return; }
  info->n_fact = 1;
  memset(W->x, 0, sizeof(double) * n);
  memset(W->z, 0, sizeof(double) * m);
  memset(W->y, 0, sizeof(double) * m);
  double eps_abs = P->eps_abs, eps_rel = P->eps_rel;
  int it = 0;
  for (int attempt = 0; attempt < 5 && info->status != 0; ++attempt) {
    int conv = 0;
    while (it < P->max_iter && !conv) {
      ++it;
      for (int i = 0; i < m; ++i) W->tmp[i] = W->rho_v[i] * W->z[i] - W->y[i];
      apply_AT(N, mu, W->tmp, W->rhs);
      for (int i = 0; i < n; ++i) W->rhs[i] += sigma * W->x[i] - W->g[i];
      chol_solve(W->K, n, n, W->rhs);
      apply_A(N, mu, W->rhs, W->zt);
      for (int i = 0; i < n; ++i) W->x[i] = al * W->rhs[i] + (1 - al) * W->x[i];
      for (int i = 0; i < m; ++i) {
        const double zr = al * W->zt[i] + (1 - al) * W->z[i];
        double zn = zr + W->y[i] / W->rho_v[i];
        if (zn < W->lo[i]) zn = W->lo[i];
        if (zn > W->hi[i]) zn = W->hi[i];
        W->y[i] += W->rho_v[i] * (zr - zn);
        W->z[i] = zn;
      }
      if (it % P->check_every == 0 || it == P->max_iter) {
        apply_A(N, mu, W->x, W->zt);
        double rp = 0, nax = infnorm(W->zt, m), nz = infnorm(W->z, m);
        for (int i = 0; i < m; ++i) rp = fmax(rp, fabs(W->zt[i] - W->z[i]));
        apply_AT(N, mu, W->y, W->rhs);
        double rd = 0, nhx = 0, naty = infnorm(W->rhs, n), ng = infnorm(W->g, n);
        for (int i = 0; i < n; ++i) {
          double s = 0;
          const double *hi = W->H + (size_t)i * n;
          for (int j = 0; j < n; ++j) s += hi[j] * W->x[j];
          nhx = fmax(nhx, fabs(s));
          rd = fmax(rd, fabs(s + W->g[i] + W->rhs[i]));
        }
        if (!(rd < 1e300)) { info->status = 1; info->iters = it; return; }
        const double np_ = fmax(nax, nz), nd = fmax(fmax(nhx, naty), ng);
        if (rp <= eps_abs + eps_rel * np_ && rd <= eps_abs + eps_rel * nd) conv = 1;
        else if (P->adaptive_rho) {
          double nr = rho * sqrt((rp / fmax(np_, 1e-30)) / fmax(rd / fmax(nd, 1e-30), 1e-30));
          nr = fmin(fmax(nr, 1e-6), 1e6);
          if (nr > 5 * rho || nr < rho / 5) {
            rho = nr;
            set_rho(P, W, rho);
            build_K(N, mu, sigma, W->H, W->rho_v, W->K);
            if (chol(W->K, n, n)) { info->status = 1; return; }
            info->n_fact++;
          }
        }
      }
    }
    info->iters = it;
    if (!P->polish) { if (conv) info->status = 0; break; }
    if (polish(P, W, info->kkt)) {
      info->status = 0;
      info->polished = 1;
      memcpy(W->x, W->xp, sizeof(double) * n);
    } else {
      if (it >= P->max_iter) break;
      eps_abs *= 0.1; eps_rel *= 0.1;
      set_rho(P, W, rho); /* K was overwritten by the polish */
      build_K(N, mu, sigma, W->H, W->rho_v, W->K);
      if (chol(W->K, n, n)) { info->status = 1; return; }
      info->n_fact++;
    }
  }
  if (info->status == 0 || !P->polish) {
    if (forces) memcpy(forces, W->x, sizeof(double) * n);
    if (xpred) { /* roll-out, mpc.cpp:461-466 raw_data */
      double xs[NX];
      memcpy(xs, W->x0, sizeof(xs));
      memcpy(xpred, xs, sizeof(xs));
      for (int k = 0; k < N; ++k) {
        double nx[NX];
        for (int r = 0; r < NX; ++r) {
          double s = 0;
          for (int t = 0; t < NX; ++t) s += W->A[r * NX + t] * xs[t];
          for (int c = 0; c < NU; ++c) s += W->B[k][r * NU + c] * W->x[NU * k + c];
          nx[r] = s;
        }
        memcpy(xs, nx, sizeof(xs));
        memcpy(xpred + NX * (k + 1), xs, sizeof(xs));
      }
    }
  }
}

/* batch entry: one problem per thread. Returns number of problems with status 0. */
int ref_solve_batch(const ref_params *P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                    ref_info *info, int threads) {
  if (P->N < 1 || P->N > MAXN) return -1;
  const int stride = 26 + 13 * (P->N + 1) + 12 * P->N;
  int solved = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel reduction(+ : solved)
  {
    ref_ws *W = (ref_ws *)malloc(sizeof(ref_ws));
#pragma omp for schedule(dynamic, 4)
    for (int p = 0; p < n; ++p) {
      solve_one(P, in + (size_t)p * stride, contact + (size_t)p * P->N * 4, W, forces ? forces + (size_t)p * P->N * NU : 0,
                xpred ? xpred + (size_t)p * (P->N + 1) * NX : 0, info + p);
      solved += info[p].status == 0;
    }
    free(W);
  }
  return solved;
}

/* Solve an OCP-QP handed over as STAGE DATA (column-major, the acados setters' layout): used by oracle/ref_harness.cpp,
 * where the reference's own MPC::GetWrenchSequence (compiled from /root/reference) assembled A, B_k, q_k, x0 and the
 * bounds and calls ocp_qp_solve().  The cost / constraint structure (Q diagonal, R = alpha I, pyramid rows with mu,
 * boxes with fmin / fmax) is checked by the caller and arrives through P. */
int ref_solve_ocp(const ref_params *P, const double *A_cm, const double *const *B_cm, const double *const *q, const double *x0,
                  const double *const *lbu, const double *const *ubu, double *u_out, double *x_out, ref_info *info) {
  if (P->N < 1 || P->N > MAXN) return -1;
  ref_ws *W = (ref_ws *)malloc(sizeof(ref_ws));
  for (int r = 0; r < NX; ++r) for (int c = 0; c < NX; ++c) W->A[r * NX + c] = A_cm[c * NX + r];
  for (int k = 0; k < P->N; ++k) {
    for (int r = 0; r < NX; ++r) for (int c = 0; c < NU; ++c) W->B[k][r * NU + c] = B_cm[k][c * NX + r];
    memcpy(W->lbu[k], lbu[k], sizeof(double) * NU);
    memcpy(W->ubu[k], ubu[k], sizeof(double) * NU);
  }
  for (int k = 0; k <= P->N; ++k) memcpy(W->q[k], q[k], sizeof(double) * NX);
  memcpy(W->x0, x0, sizeof(double) * NX);
  solve_ws(P, W, u_out, x_out, info);
  free(W);
  return info->status;
}


/* ---- independent KKT check of a candidate solution, sparse OCP form (what acados' ocp_qp_inf_norm_residuals reports,
 * mpc.cpp:430-433): stationarity, dynamics, inequality, complementarity inf-norms.  The checker shares nothing with the
 * solvers above but build_qp: the states are rolled out with the explicit A, B_k, the dynamics multipliers are the
 * exact adjoint lam_k = Q x_k + q_k + A' lam_{k+1}, and - when the solver's inequality multipliers are not supplied
 * (lam_box [N][12] > 0 at the upper, < 0 at the lower bound; lam_g [N][16] likewise, the reference's row order) - the
 * sign-feasible multipliers of the rows that are tight within 1e-7 (scaled) are reconstructed per foot by exact NNLS. */
static void kkt_one(const ref_params *P, const double *rec, const uint8_t *ct, const double *U, const double *Xpred,
                    const double *lam_box, const double *lam_g, ref_ws *W, double kkt[4]) {
  const int N = P->N;
  const double mu = P->mu;
  build_qp(P, rec, ct, W);
  static __thread double X[MAXN + 1][NX], lam[MAXN + 2][NX];
  double dyn = 0, stat = 0, ineq = 0, comp = 0;
  memcpy(X[0], W->x0, sizeof(double) * NX);
  for (int k = 0; k < N; ++k)
    for (int r = 0; r < NX; ++r) {
      double s = 0;
      for (int t = 0; t < NX; ++t) s += W->A[r * NX + t] * X[k][t];
      for (int c = 0; c < NU; ++c) s += W->B[k][r * NU + c] * U[NU * k + c];
      X[k + 1][r] = s;
    }
  if (Xpred) { /* the solver's own prediction must satisfy the dynamics: x0 fixed, x_{k+1} = A x_k + B_k u_k */
    for (int r = 0; r < NX; ++r) dyn = fmax(dyn, fabs(Xpred[r] - W->x0[r]));
    for (int k = 0; k < N; ++k)
      for (int r = 0; r < NX; ++r) {
        double s = 0;
        for (int t = 0; t < NX; ++t) s += W->A[r * NX + t] * Xpred[NX * k + t];
        for (int c = 0; c < NU; ++c) s += W->B[k][r * NU + c] * U[NU * k + c];
        dyn = fmax(dyn, fabs(Xpred[NX * (k + 1) + r] - s));
      }
  }
  memset(lam[N + 1], 0, sizeof(double) * NX);
  for (int k = N; k >= 1; --k)
    for (int r = 0; r < NX; ++r) {
      double s = (r < 12 ? P->w[r] : 0.0) * X[k][r] + W->q[k][r];
      if (k < N) for (int t = 0; t < NX; ++t) s += W->A[t * NX + r] * lam[k + 1][t];
      lam[k][r] = s;
    }
  const double lo[7] = {-P->fmax, -P->fmax, P->fmin, -HUGE_VAL, 0, -HUGE_VAL, 0}, hi[7] = {P->fmax, P->fmax, P->fmax, 0, HUGE_VAL, 0, HUGE_VAL};
  for (int k = 0; k < N; ++k)
    for (int f = 0; f < 4; ++f) {
      const double *u = U + NU * k + 3 * f;
      double g[3];
      for (int c = 0; c < 3; ++c) {
        double s = P->alpha * u[c];
        for (int t = 0; t < NX; ++t) s += W->B[k][t * NU + 3 * f + c] * lam[k + 1][t];
        g[c] = s;
      }
      if (!ct[4 * k + f]) { /* lbu = ubu = 0: an equality, any multiplier is admissible */
        ineq = fmax(ineq, fmax(fabs(u[0]), fmax(fabs(u[1]), fabs(u[2]))));
        continue;
      }
      const double ax[7] = {u[0], u[1], u[2], u[0] - mu * u[2], u[0] + mu * u[2], u[1] - mu * u[2], u[1] + mu * u[2]};
      for (int r = 0; r < 7; ++r) {
        if (lo[r] > -HUGE_VAL) ineq = fmax(ineq, lo[r] - ax[r]);
        if (hi[r] < HUGE_VAL) ineq = fmax(ineq, ax[r] - hi[r]);
      }
      if (lam_box) {
        double y[7];
        for (int c = 0; c < 3; ++c) y[c] = lam_box[NU * k + 3 * f + c];
        for (int r = 0; r < 4; ++r) y[3 + r] = lam_g[16 * k + 4 * f + r];
        double rs[3] = {g[0], g[1], g[2]};
        for (int r = 0; r < 7; ++r) {
          double a[3];
          foot_row(r, mu, a);
          for (int c = 0; c < 3; ++c) rs[c] += a[c] * y[r];
          if (y[r] > 0) { if (hi[r] < HUGE_VAL) comp = fmax(comp, y[r] * fabs(hi[r] - ax[r])); else stat = fmax(stat, y[r]); }
          if (y[r] < 0) { if (lo[r] > -HUGE_VAL) comp = fmax(comp, -y[r] * fabs(ax[r] - lo[r])); else stat = fmax(stat, -y[r]); }
        }
        stat = fmax(stat, fmax(fabs(rs[0]), fmax(fabs(rs[1]), fabs(rs[2]))));
      } else {
        const double sc = fmax(1.0, fmax(fabs(u[0]), fmax(fabs(u[1]), fabs(u[2])))), tt = 1e-7 * sc;
        double Mc[14][3], coef[14], slack[14];
        int m = 0;
        for (int r = 0; r < 7; ++r) {
          double a[3];
          foot_row(r, mu, a);
          if (lo[r] > -HUGE_VAL && ax[r] - lo[r] <= tt && m < 8) { Mc[m][0] = -a[0]; Mc[m][1] = -a[1]; Mc[m][2] = -a[2]; slack[m++] = fabs(ax[r] - lo[r]); }
          if (hi[r] < HUGE_VAL && hi[r] - ax[r] <= tt && m < 8) { Mc[m][0] = a[0]; Mc[m][1] = a[1]; Mc[m][2] = a[2]; slack[m++] = fabs(hi[r] - ax[r]); }
        }
        const double b3[3] = {-g[0], -g[1], -g[2]};
        stat = fmax(stat, nnls_enum(Mc, m, b3, coef));
        for (int q = 0; q < m; ++q) comp = fmax(comp, coef[q] * slack[q]);
      }
    }
  kkt[0] = stat; kkt[1] = dyn; kkt[2] = fmax(ineq, 0.0); kkt[3] = comp;
}

/* batch entry; xpred, lam_box, lam_g may be NULL.  kkt [n][4]. */
int ref_kkt_batch(const ref_params *P, int n, const double *in, const uint8_t *contact, const double *forces, const double *xpred,
                  const double *lam_box, const double *lam_g, double *kkt, int threads) {
  if (P->N < 1 || P->N > MAXN) return -1;
  const int stride = 26 + 13 * (P->N + 1) + 12 * P->N;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    ref_ws *W = (ref_ws *)malloc(sizeof(ref_ws));
#pragma omp for schedule(dynamic, 16)
    for (int p = 0; p < n; ++p)
      kkt_one(P, in + (size_t)p * stride, contact + (size_t)p * P->N * 4, forces + (size_t)p * P->N * NU,
              xpred ? xpred + (size_t)p * (P->N + 1) * NX : 0, lam_box ? lam_box + (size_t)p * P->N * NU : 0,
              lam_g ? lam_g + (size_t)p * P->N * 16 : 0, W, kkt + 4 * (size_t)p);
    free(W);
  }
  return 0;
}


/* ---- (4) Riccati-based primal-dual interior-point method on the SPARSE OCP-QP: the restatement of the reference's
 * default plan PARTIAL_CONDENSING_HPIPM (config/mit_controller_sim_go2.yaml:24) at cond_N = N.  HPIPM's algorithm
 * (Frison & Diehl, "HPIPM: a high-performance quadratic programming framework for model predictive control", 2020) is
 * a Mehrotra predictor-corrector IPM whose Newton systems are solved by a backward Riccati recursion; its modes differ in
 * tolerances / iteration caps / refinement (mpc.cpp:226-234 only passes the mode string; the constants are upstream's and
 * not verifiable here).  This port uses dense nx = 13, nu = 12 stage matrices as the reference formulates them; forces of
 * swing feet (lbu = ubu = 0, mpc.cpp:110-113) are eliminated.  `tol` is the residual target, `max_iter` the cap:
 *     SPEED-like  : tol 1e-6,  max_iter 15        BALANCE-like: tol 1e-8, max_iter 30        parity: tol 1e-11, max_iter 60
 * CPU baseline of bench.py for the sparse configuration; never linked into the product. */
typedef struct {
  double K[MAXN][NU * NX], L[MAXN][NU * NU], kff[MAXN][NU], P[NX * NX], p[NX];
  double X[MAXN + 1][NX], lam[MAXN + 2][NX], U[MAXN][NU], dU[MAXN][NU];
  double s[MAXN][4][10], z[MAXN][4][10], ds[MAXN][4][10], dz[MAXN][4][10], rp[MAXN][4][10], rd[MAXN][NU], rt[MAXN][NU];
  int idx[MAXN][NU], nu[MAXN];
} ric_ws;

static void foot_G(const double *f, double mu, double g[10]) {
  g[0] = f[0]; g[1] = -f[0]; g[2] = f[1]; g[3] = -f[1]; g[4] = f[2]; g[5] = -f[2];
  g[6] = f[0] - mu * f[2]; g[7] = -f[0] - mu * f[2]; g[8] = f[1] - mu * f[2]; g[9] = -f[1] - mu * f[2];
}
static void foot_GT(const double *v, double mu, double o[3]) {
  o[0] = (v[0] - v[1]) + (v[6] - v[7]);
  o[1] = (v[2] - v[3]) + (v[8] - v[9]);
  o[2] = (v[4] - v[5]) - mu * ((v[6] + v[7]) + (v[8] + v[9]));
}

static int riccati_ipm_one(const ref_params *P, const double *rec, const uint8_t *ct, ref_ws *W, ric_ws *R, double tol, int max_iter,
                           double *forces, double *xpred, ref_info *info) {
  const int N = P->N;
  const double mu = P->mu;
  build_qp(P, rec, ct, W);
  double Q[NX];
  for (int i = 0; i < 12; ++i) Q[i] = P->w[i];
  Q[12] = 0;
  const double hh[10] = {P->fmax, P->fmax, P->fmax, P->fmax, P->fmax, -P->fmin, 0, 0, 0, 0};
  int mtot = 0;
  for (int k = 0; k < N; ++k) {
    int ns = ct[4 * k] + ct[4 * k + 1] + ct[4 * k + 2] + ct[4 * k + 3];
    R->nu[k] = 0;
    for (int j = 0; j < NU; ++j) {
      R->U[k][j] = 0;
      if (ct[4 * k + j / 3]) {
        R->idx[k][R->nu[k]++] = j;
        if (j % 3 == 2) { /* weight-sharing start, strictly inside the box and the pyramid */
          double v = rec[13] * 9.8067 / ns, span = P->fmax - P->fmin;
          R->U[k][j] = fmin(fmax(v, P->fmin + 0.05 * span), P->fmax - 0.05 * span);
        }
      }
    }
    for (int f = 0; f < 4; ++f) if (ct[4 * k + f]) {
      double g[10];
      foot_G(R->U[k] + 3 * f, mu, g);
      for (int r = 0; r < 10; ++r) { R->s[k][f][r] = fmax(hh[r] - g[r], 1e-2); R->z[k][f][r] = 1.0 / R->s[k][f][r]; }
      mtot += 10;
    }
  }
  memset(info, 0, sizeof(*info));
  info->status = 2;
  int it;
  for (it = 0; it <= max_iter; ++it) {
    /* roll-out and exact adjoint */
    memcpy(R->X[0], W->x0, sizeof(double) * NX);
    for (int k = 0; k < N; ++k)
      for (int r = 0; r < NX; ++r) {
        double a = 0;
        for (int t = 0; t < NX; ++t) a += W->A[r * NX + t] * R->X[k][t];
        for (int c = 0; c < NU; ++c) a += W->B[k][r * NU + c] * R->U[k][c];
        R->X[k + 1][r] = a;
      }
    memset(R->lam[N + 1], 0, sizeof(double) * NX);
    for (int k = N; k >= 1; --k)
      for (int r = 0; r < NX; ++r) {
        double a = Q[r] * R->X[k][r] + W->q[k][r];
        if (k < N) for (int t = 0; t < NX; ++t) a += W->A[t * NX + r] * R->lam[k + 1][t];
        R->lam[k][r] = a;
      }
    /* residuals */
    double res = 0, cmax = 0, gap = 0;
    for (int k = 0; k < N; ++k)
      for (int f = 0; f < 4; ++f) {
        if (!ct[4 * k + f]) continue;
        double g[10], gz[3];
        foot_G(R->U[k] + 3 * f, mu, g);
        foot_GT(R->z[k][f], mu, gz);
        for (int c = 0; c < 3; ++c) {
          double a = P->alpha * R->U[k][3 * f + c] + gz[c];
          for (int t = 0; t < NX; ++t) a += W->B[k][t * NU + 3 * f + c] * R->lam[k + 1][t];
          R->rd[k][3 * f + c] = a;
          res = fmax(res, fabs(a));
        }
        for (int r = 0; r < 10; ++r) {
          R->rp[k][f][r] = g[r] + R->s[k][f][r] - hh[r];
          res = fmax(res, fabs(R->rp[k][f][r]));
          cmax = fmax(cmax, R->s[k][f][r] * R->z[k][f][r]);
          gap += R->s[k][f][r] * R->z[k][f][r];
        }
      }
    info->kkt[0] = res; info->kkt[3] = cmax;
    if (!(res < 1e300)) { info->status = 1; break; }
    if ((res <= tol && cmax <= tol) || mtot == 0) { info->status = 0; break; }
    if (it == max_iter) { if (res <= 1e3 * tol && cmax <= 1e3 * tol && res <= 1e-6) info->status = 0; break; }
    const double mu_c = gap / mtot;
    /* Riccati factorisation with R~ = alpha I + G' W G */
    for (int i = 0; i < NX * NX; ++i) R->P[i] = 0;
    for (int i = 0; i < NX; ++i) R->P[i * NX + i] = Q[i];
    for (int k = N - 1; k >= 0; --k) {
      const int nu = R->nu[k];
      const int *ix = R->idx[k];
      double PA[NX * NX], PB[NX * NU], Huu[NU * NU], Hux[NU * NX];
      for (int r = 0; r < NX; ++r) {
        for (int c = 0; c < NX; ++c) { double a = 0; for (int t = 0; t < NX; ++t) a += R->P[r * NX + t] * W->A[t * NX + c]; PA[r * NX + c] = a; }
        for (int j = 0; j < nu; ++j) { double a = 0; for (int t = 0; t < NX; ++t) a += R->P[r * NX + t] * W->B[k][t * NU + ix[j]]; PB[r * NU + j] = a; }
      }
      for (int i = 0; i < nu; ++i) {
        for (int j = 0; j <= i; ++j) { double a = 0; for (int t = 0; t < NX; ++t) a += W->B[k][t * NU + ix[i]] * PB[t * NU + j]; Huu[i * NU + j] = a; }
        for (int c = 0; c < NX; ++c) { double a = 0; for (int t = 0; t < NX; ++t) a += W->B[k][t * NU + ix[i]] * PA[t * NX + c]; Hux[i * NX + c] = a; }
      }
      for (int f = 0; f < 4; ++f) {
        if (!ct[4 * k + f]) continue;
        double w[10];
        for (int r = 0; r < 10; ++r) w[r] = R->z[k][f][r] / R->s[k][f][r];
        int b = 0;
        while (ix[b] != 3 * f) ++b;
        Huu[b * NU + b] += P->alpha + (w[0] + w[1]) + (w[6] + w[7]);
        Huu[(b + 1) * NU + b + 1] += P->alpha + (w[2] + w[3]) + (w[8] + w[9]);
        Huu[(b + 2) * NU + b + 2] += P->alpha + (w[4] + w[5]) + mu * mu * ((w[6] + w[7]) + (w[8] + w[9]));
        Huu[(b + 2) * NU + b] += -mu * (w[6] - w[7]);
        Huu[(b + 2) * NU + b + 1] += -mu * (w[8] - w[9]);
      }
      if (nu > 0 && chol(Huu, nu, NU)) { info->status = 1; goto done; }
      memcpy(R->L[k], Huu, sizeof(Huu));
      /* K = -Huu^-1 Hux (column by column), P <- Q + A'PA + Hux' K */
      for (int c = 0; c < NX; ++c) {
        double col[NU];
        for (int i = 0; i < nu; ++i) col[i] = Hux[i * NX + c];
        chol_solve(Huu, nu, NU, col);
        for (int i = 0; i < nu; ++i) R->K[k][i * NX + c] = -col[i];
      }
      double Pn[NX * NX];
      for (int r = 0; r < NX; ++r)
        for (int c = 0; c < NX; ++c) {
          double a = (r == c) ? Q[r] : 0.0;
          for (int t = 0; t < NX; ++t) a += W->A[t * NX + r] * PA[t * NX + c];
          for (int i = 0; i < nu; ++i) a += Hux[i * NX + r] * R->K[k][i * NX + c];
          Pn[r * NX + c] = a;
        }
      memcpy(R->P, Pn, sizeof(Pn));
    }
    /* predictor / corrector */
    double sigma_mu = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int k = 0; k < N; ++k)
        for (int f = 0; f < 4; ++f) {
          if (!ct[4 * k + f]) continue;
          double t10[10], o[3];
          for (int r = 0; r < 10; ++r) {
            double rc = R->s[k][f][r] * R->z[k][f][r] + (pass ? R->ds[k][f][r] * R->dz[k][f][r] - sigma_mu : 0.0);
            t10[r] = (R->z[k][f][r] * R->rp[k][f][r] - rc) / R->s[k][f][r];
          }
          foot_GT(t10, mu, o);
          for (int c = 0; c < 3; ++c) R->rt[k][3 * f + c] = R->rd[k][3 * f + c] + o[c];
        }
      memset(R->p, 0, sizeof(R->p));
      for (int k = N - 1; k >= 0; --k) { /* hu = r~ + B'p, kff = -Huu^-1 hu, p <- A'p + Hux' kff = A'p - K' hu */
        const int nu = R->nu[k];
        const int *ix = R->idx[k];
        double hu[NU], pn[NX];
        for (int i = 0; i < nu; ++i) { double a = R->rt[k][ix[i]]; for (int t = 0; t < NX; ++t) a += W->B[k][t * NU + ix[i]] * R->p[t]; hu[i] = a; }
        for (int r = 0; r < NX; ++r) {
          double a = 0;
          for (int t = 0; t < NX; ++t) a += W->A[t * NX + r] * R->p[t];
          for (int i = 0; i < nu; ++i) a += R->K[k][i * NX + r] * hu[i];
          pn[r] = a;
        }
        chol_solve(R->L[k], nu, NU, hu);
        for (int i = 0; i < nu; ++i) R->kff[k][i] = -hu[i];
        memcpy(R->p, pn, sizeof(pn));
      }
      double dx[NX] = {0};
      for (int k = 0; k < N; ++k) {
        const int nu = R->nu[k];
        const int *ix = R->idx[k];
        double nx[NX];
        for (int c = 0; c < NU; ++c) R->dU[k][c] = 0;
        for (int i = 0; i < nu; ++i) { double a = R->kff[k][i]; for (int c = 0; c < NX; ++c) a += R->K[k][i * NX + c] * dx[c]; R->dU[k][ix[i]] = a; }
        for (int r = 0; r < NX; ++r) {
          double a = 0;
          for (int t = 0; t < NX; ++t) a += W->A[r * NX + t] * dx[t];
          for (int i = 0; i < nu; ++i) a += W->B[k][r * NU + ix[i]] * R->dU[k][ix[i]];
          nx[r] = a;
        }
        memcpy(dx, nx, sizeof(nx));
      }
      double tmax = 0, g2 = 0, c1 = 0, c2 = 0;
      static __thread double dsn[MAXN][4][10], dzn[MAXN][4][10];
      for (int k = 0; k < N; ++k)
        for (int f = 0; f < 4; ++f) {
          if (!ct[4 * k + f]) continue;
          double gd[10];
          foot_G(R->dU[k] + 3 * f, mu, gd);
          for (int r = 0; r < 10; ++r) {
            double sk = R->s[k][f][r], zk = R->z[k][f][r];
            double rc = sk * zk + (pass ? R->ds[k][f][r] * R->dz[k][f][r] - sigma_mu : 0.0);
            dsn[k][f][r] = -R->rp[k][f][r] - gd[r];
            dzn[k][f][r] = (-rc - zk * dsn[k][f][r]) / sk;
            tmax = fmax(tmax, fmax(-dsn[k][f][r] / sk, -dzn[k][f][r] / zk));
          }
        }
      double a = tmax > 1.0 ? 1.0 / tmax : 1.0;
      if (pass == 0) {
        for (int k = 0; k < N; ++k)
          for (int f = 0; f < 4; ++f) if (ct[4 * k + f])
            for (int r = 0; r < 10; ++r) {
              g2 += (R->s[k][f][r] + a * dsn[k][f][r]) * (R->z[k][f][r] + a * dzn[k][f][r]);
              R->ds[k][f][r] = dsn[k][f][r]; R->dz[k][f][r] = dzn[k][f][r];
            }
        double ratio = (g2 / mtot) / fmax(mu_c, 1e-300);
        sigma_mu = ratio * ratio * ratio * mu_c;
      } else {
        a = fmin(1.0, 0.995 * a);
        if (res < cmax) { /* monotone-gap safeguard against Mehrotra's two-cycle */
          for (int k = 0; k < N; ++k)
            for (int f = 0; f < 4; ++f) if (ct[4 * k + f])
              for (int r = 0; r < 10; ++r) { c1 += R->s[k][f][r] * dzn[k][f][r] + R->z[k][f][r] * dsn[k][f][r]; c2 += dsn[k][f][r] * dzn[k][f][r]; }
          double lin = c1 + 0.1 * gap;
          if (c2 > 0 && lin < 0) a = fmin(a, -lin / c2);
        }
        for (int k = 0; k < N; ++k) {
          for (int c = 0; c < NU; ++c) R->U[k][c] += a * R->dU[k][c];
          for (int f = 0; f < 4; ++f) if (ct[4 * k + f])
            for (int r = 0; r < 10; ++r) { R->s[k][f][r] += a * dsn[k][f][r]; R->z[k][f][r] += a * dzn[k][f][r]; }
        }
      }
    }
  }
done:
  info->iters = it;
  if (info->status == 0) {
    if (forces) for (int k = 0; k < N; ++k) memcpy(forces + NU * k, R->U[k], sizeof(double) * NU);
    if (xpred) for (int k = 0; k <= N; ++k) memcpy(xpred + NX * k, R->X[k], sizeof(double) * NX);
  }
  return info->status;
}

/* batch entry of the Riccati IPM port: one problem per thread.  Returns the number of problems with status 0. */
int ref_riccati_batch(const ref_params *P, int n, const double *in, const uint8_t *contact, double *forces, double *xpred,
                      ref_info *info, double tol, int max_iter, int threads) {
  if (P->N < 1 || P->N > MAXN) return -1;
  const int stride = 26 + 13 * (P->N + 1) + 12 * P->N;
  int solved = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel reduction(+ : solved)
  {
    ref_ws *W = (ref_ws *)malloc(sizeof(ref_ws));
    ric_ws *R = (ric_ws *)malloc(sizeof(ric_ws));
#pragma omp for schedule(dynamic, 4)
    for (int p = 0; p < n; ++p)
      solved += riccati_ipm_one(P, in + (size_t)p * stride, contact + (size_t)p * P->N * 4, W, R, tol, max_iter,
                                forces ? forces + (size_t)p * P->N * NU : 0, xpred ? xpred + (size_t)p * (P->N + 1) * NX : 0, info + p) == 0;
    free(W);
    free(R);
  }
  return solved;
}

/* condensed H (row-major 12N x 12N) and g of one problem, for cross-checks */
int ref_condense(const ref_params *P, const double *rec, const uint8_t *ct, double *H, double *g) {
  ref_ws *W = (ref_ws *)malloc(sizeof(ref_ws));
  build_qp(P, rec, ct, W);
  condense(P, W);
  const int n = NU * P->N;
  memcpy(H, W->H, sizeof(double) * (size_t)n * n);
  memcpy(g, W->g, sizeof(double) * n);
  free(W);
  return n;
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
