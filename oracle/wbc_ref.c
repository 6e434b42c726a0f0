/*
 * wbc_ref.c - C restatement of a dense QP solve for the whole-body-control task-space QP (TEST / MEASUREMENT
 * INFRASTRUCTURE ONLY; CPU baseline of bench.py for BASELINE.json configs[3]).
 *
 * The reference hands the TSID QP to a `wbc::QPSolver` plugin of ARC-OPT (wbc_arc_opt.cpp:51-74,297; default
 * EiquadprogSolver, alternatives ProxQP / qpOASES / OSQP / qpSWIFT / HPIPM).  Those are third-party libraries that are not
 * in /root/reference and cannot be built here; this file restates a textbook dense primal-dual interior-point method
 * (Mehrotra predictor-corrector, Nocedal & Wright ch. 16.6; the same method, safeguards and stopping rule as the
 * conservative attempt of the device kernel wbc_qp.cu - step fraction 0.995, end-game centring floor 0.05 - in scalar C;
 * the device kernel's aggressive first attempt is a speed measure that this file does not need) for
 *     min 1/2 x'Hx + g'x   s.t.  A[0:neq] x = b,   lo <= A[neq:] x <= hi        (|bound| >= 1e19: absent)
 * with the layouts of ARC-OPT's wbc::QuadraticProgram as they cross include/b200qp.h (column-major H and A).
 * One problem per thread, OpenMP over the batch.  Never linked into the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXQ 48
#define MAXE 32
#define MAXI 128 /* one-sided rows */

typedef struct { int status, iters; double kkt[4]; } wbc_info;

/* Cholesky with the pivot floor of the device kernel (1e-13 x the original diagonal entry) */
static void cholq(double *A, int n, int ld) {
  double d0[MAXQ];
  for (int j = 0; j < n; ++j) d0[j] = A[j * ld + j];
  for (int j = 0; j < n; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    const double flo = 1e-13 * d0[j];
    d = sqrt(d > flo ? d : flo);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s / d;
    }
  }
}
static void cholq_solve(const double *L, int n, int ld, double *b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * ld + k] * b[k]; b[i] = s / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * b[k]; b[i] = s / L[i * ld + i]; }
}

typedef struct {
  int nq, neq, m;
  double Ha[MAXQ * MAXQ], Ae[MAXE][MAXQ], G[MAXI][MAXQ];
  double Phi[MAXQ * MAXQ], Y[MAXE][MAXQ], S[MAXE * MAXE];
} wbc_ws;

/* factor the reduced Newton system for row weights w:  Phi = Ha + G'WG = LL',  Y = Phi^-1 Ae',  S = Ae Y = Ls Ls' */
static void wbc_factor(wbc_ws *W, const double *w) {
  const int nq = W->nq, neq = W->neq, m = W->m;
  for (int j = 0; j < nq; ++j)
    for (int k = 0; k <= j; ++k) {
      double a = W->Ha[j * MAXQ + k];
      for (int i = 0; i < m; ++i) a += W->G[i][j] * w[i] * W->G[i][k];
      W->Phi[j * MAXQ + k] = a;
    }
  cholq(W->Phi, nq, MAXQ);
  for (int i = 0; i < neq; ++i) { memcpy(W->Y[i], W->Ae[i], sizeof(double) * nq); cholq_solve(W->Phi, nq, MAXQ, W->Y[i]); }
  for (int i = 0; i < neq; ++i)
    for (int k = 0; k <= i; ++k) { double a = 0; for (int j = 0; j < nq; ++j) a += W->Ae[i][j] * W->Y[k][j]; W->S[i * MAXE + k] = a; }
  if (neq) cholq(W->S, neq, MAXE);
}
/* [Phi Ae'; Ae 0][dx; dnu] = [-rt; -re] */
static void wbc_solve(const wbc_ws *W, const double *rt, const double *re, double *dx, double *dnu) {
  const int nq = W->nq, neq = W->neq;
  double t[MAXQ];
  for (int j = 0; j < nq; ++j) t[j] = -rt[j];
  cholq_solve(W->Phi, nq, MAXQ, t);
  for (int i = 0; i < neq; ++i) { double a = re[i]; for (int j = 0; j < nq; ++j) a += W->Ae[i][j] * t[j]; dnu[i] = a; }
  if (neq) cholq_solve(W->S, neq, MAXE, dnu);
  for (int j = 0; j < nq; ++j) { double a = t[j]; for (int i = 0; i < neq; ++i) a -= W->Y[i][j] * dnu[i]; dx[j] = a; }
}

/* The method of the device kernel (dfki_quad_b200/csrc/wbc_qp.cu), one problem: augmented-Lagrangian form of the equalities
 * (H + rho Ae'Ae), CVXOPT-style starting point, Mehrotra predictor-corrector with separate primal / dual step lengths,
 * end-game centring and the two-cycle guard, best iterate kept; success = residuals <= tol (target 1e-4 tol). */
static int wbc_one(int nq, int neq, int nin, const double *H, const double *g, const double *A, const double *lo, const double *hi,
                   double tol, int max_iter, double *xout, wbc_info *info) {
  const int nc = neq + nin;
  const double rho_eq = 10.0, target = 1e-4 * tol;
  wbc_ws *W = (wbc_ws *)malloc(sizeof(wbc_ws));
  double h[MAXI], be[MAXE], ga[MAXQ];
  int m = 0;
  for (int i = 0; i < neq; ++i) { for (int j = 0; j < nq; ++j) W->Ae[i][j] = A[j * nc + i]; be[i] = lo[i]; }
  for (int i = neq; i < nc; ++i) {
    if (hi[i] < 1e19) { for (int j = 0; j < nq; ++j) W->G[m][j] = A[j * nc + i]; h[m++] = hi[i]; }
    if (lo[i] > -1e19) { for (int j = 0; j < nq; ++j) W->G[m][j] = -A[j * nc + i]; h[m++] = -lo[i]; }
  }
  W->nq = nq; W->neq = neq; W->m = m;
  for (int j = 0; j < nq; ++j) {
    for (int k = 0; k < nq; ++k) { double a = H[k * nq + j]; for (int i = 0; i < neq; ++i) a += rho_eq * W->Ae[i][j] * W->Ae[i][k]; W->Ha[j * MAXQ + k] = a; }
    double a = g[j];
    for (int i = 0; i < neq; ++i) a -= rho_eq * W->Ae[i][j] * be[i];
    ga[j] = a;
  }
  double x[MAXQ], nu[MAXE], s[MAXI], z[MAXI], w[MAXI], bx[MAXQ], rd[MAXQ], re[MAXE], rp[MAXI];
  memset(info, 0, sizeof(*info));
  info->status = 2;
  /* starting point: W = I solve, then shift s and z into the interior */
  {
    double rt[MAXQ];
    for (int i = 0; i < m; ++i) w[i] = 1.0;
    wbc_factor(W, w);
    for (int j = 0; j < nq; ++j) { double a = ga[j]; for (int i = 0; i < m; ++i) a -= W->G[i][j] * h[i]; rt[j] = a; }
    for (int i = 0; i < neq; ++i) re[i] = -be[i];
    wbc_solve(W, rt, re, x, nu);
    double smin = 1e300, zmin = 1e300;
    for (int i = 0; i < m; ++i) { double a = -h[i]; for (int j = 0; j < nq; ++j) a += W->G[i][j] * x[j]; z[i] = a; smin = fmin(smin, -a); zmin = fmin(zmin, a); }
    const double sh = smin > 0 ? 0.0 : 1.0 - smin, zh = zmin > 0 ? 0.0 : 1.0 - zmin;
    for (int i = 0; i < m; ++i) { s[i] = -z[i] + sh; z[i] = z[i] + zh; }
  }
  double best = 1e300, bk[4] = {0, 0, 0, 0}, mu_prev = 1e300;
  int it, stall = 0, damp = 0;
  for (it = 1; it <= max_iter + 1; ++it) {
    double kstat = 0, keq = 0, kin = 0, rpm = 0, gap = 0, cmax = 0;
    for (int j = 0; j < nq; ++j) {
      double a = ga[j];
      for (int k = 0; k < nq; ++k) a += W->Ha[j * MAXQ + k] * x[k];
      for (int i = 0; i < neq; ++i) a += W->Ae[i][j] * nu[i];
      for (int i = 0; i < m; ++i) a += W->G[i][j] * z[i];
      rd[j] = a; kstat = fmax(kstat, fabs(a));
    }
    for (int i = 0; i < neq; ++i) { double a = -be[i]; for (int j = 0; j < nq; ++j) a += W->Ae[i][j] * x[j]; re[i] = a; keq = fmax(keq, fabs(a)); }
    for (int i = 0; i < m; ++i) {
      double y = -h[i];
      for (int j = 0; j < nq; ++j) y += W->G[i][j] * x[j];
      rp[i] = y + s[i]; rpm = fmax(rpm, fabs(rp[i])); kin = fmax(kin, y);
      gap += s[i] * z[i]; cmax = fmax(cmax, s[i] * z[i]);
    }
    const double rmax = fmax(fmax(kstat, keq), fmax(rpm, cmax));
    if (!(rmax < 1e300)) { info->status = 1; break; }
    if (rmax <= target) { info->status = 0; memcpy(bx, x, sizeof(double) * nq); bk[0] = kstat; bk[1] = keq; bk[2] = fmax(kin, 0); bk[3] = cmax; break; }
    if (rmax < 0.9 * best) { best = rmax; stall = 0; memcpy(bx, x, sizeof(double) * nq); bk[0] = kstat; bk[1] = keq; bk[2] = fmax(kin, 0); bk[3] = cmax; }
    else ++stall;
    if (it > max_iter || (stall >= 2 && best <= tol)) { info->status = best <= tol ? 0 : 2; break; }
    const double mu_c = m ? gap / m : 0.0;
    if (mu_c > 0.9 * mu_prev) damp = 2;
    mu_prev = mu_c;
    for (int i = 0; i < m; ++i) w[i] = z[i] / s[i];
    wbc_factor(W, w);
    double dsa[MAXI], dza[MAXI], sigma_mu = 0;
    for (int pass = 0; pass < 2; ++pass) {
      double rt[MAXQ], dx[MAXQ], dnu[MAXE], ds[MAXI], dz[MAXI];
      for (int j = 0; j < nq; ++j) rt[j] = rd[j];
      for (int i = 0; i < m; ++i) {
        const double rc = s[i] * z[i] + (pass ? dsa[i] * dza[i] - sigma_mu : 0.0);
        const double c = (z[i] * rp[i] - rc) / s[i];
        for (int j = 0; j < nq; ++j) rt[j] += W->G[i][j] * c;
      }
      wbc_solve(W, rt, re, dx, dnu);
      if (pass == 1 && mu_c < 1e-3) { /* one step of iterative refinement, as on the device */
        double r1[MAXQ], r2[MAXE], q[MAXI], dx2[MAXQ], dnu2[MAXE];
        for (int i = 0; i < m; ++i) { double a = 0; for (int j = 0; j < nq; ++j) a += W->G[i][j] * dx[j]; q[i] = a * w[i]; }
        for (int j = 0; j < nq; ++j) {
          double a = rt[j];
          for (int k = 0; k < nq; ++k) a += W->Ha[j * MAXQ + k] * dx[k];
          for (int i = 0; i < m; ++i) a += W->G[i][j] * q[i];
          for (int i = 0; i < neq; ++i) a += W->Ae[i][j] * dnu[i];
          r1[j] = a;
        }
        for (int i = 0; i < neq; ++i) { double a = re[i]; for (int j = 0; j < nq; ++j) a += W->Ae[i][j] * dx[j]; r2[i] = a; }
        wbc_solve(W, r1, r2, dx2, dnu2);
        for (int j = 0; j < nq; ++j) dx[j] += dx2[j];
        for (int i = 0; i < neq; ++i) dnu[i] += dnu2[i];
      }
      double ap = 1.0, ad = 1.0;
      for (int i = 0; i < m; ++i) {
        double gd = 0; for (int j = 0; j < nq; ++j) gd += W->G[i][j] * dx[j];
        const double rc = s[i] * z[i] + (pass ? dsa[i] * dza[i] - sigma_mu : 0.0);
        ds[i] = -rp[i] - gd;
        dz[i] = (-rc - z[i] * ds[i]) / s[i];
        if (ds[i] < 0) ap = fmin(ap, -s[i] / ds[i]);
        if (dz[i] < 0) ad = fmin(ad, -z[i] / dz[i]);
      }
      if (pass == 0) {
        double g2 = 0;
        for (int i = 0; i < m; ++i) { g2 += (s[i] + ap * ds[i]) * (z[i] + ad * dz[i]); dsa[i] = ds[i]; dza[i] = dz[i]; }
        const double ratio = m ? (g2 / m) / fmax(mu_c, 1e-300) : 0;
        sigma_mu = ratio * ratio * ratio * mu_c;
        if (mu_c < 1e-4) sigma_mu = fmax(sigma_mu, 0.05 * mu_c);
        if (damp > 0) sigma_mu = fmax(sigma_mu, 0.2 * mu_c);
      } else {
        ap = fmin(1.0, 0.995 * ap); ad = fmin(1.0, 0.995 * ad);
        for (int j = 0; j < nq; ++j) x[j] += ap * dx[j];
        for (int i = 0; i < neq; ++i) nu[i] += ad * dnu[i];
        for (int i = 0; i < m; ++i) { s[i] += ap * ds[i]; z[i] += ad * dz[i]; }
        if (damp > 0) --damp;
      }
    }
  }
  info->iters = it > 0 ? it - 1 : 0;
  for (int k = 0; k < 4; ++k) info->kkt[k] = bk[k];
  if (info->status == 0) memcpy(xout, bx, sizeof(double) * nq);
  const int status = info->status;
  free(W);
  return status;
}

int wbc_ref_batch(int n, int nq, int neq, int nin, const double *H, const double *g, const double *A, const double *lo, const double *hi,
                  double tol, int max_iter, double *x, wbc_info *info, int threads) {
  if (nq > MAXQ || neq > MAXE || 2 * nin > MAXI) return -1;
  const int nc = neq + nin;
  int solved = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : solved)
  for (int p = 0; p < n; ++p)
    solved += wbc_one(nq, neq, nin, H + (size_t)p * nq * nq, g + (size_t)p * nq, A + (size_t)p * nc * nq, lo + (size_t)p * nc,
                      hi + (size_t)p * nc, tol, max_iter, x + (size_t)p * nq, info + p) == 0;
  return solved;
}
