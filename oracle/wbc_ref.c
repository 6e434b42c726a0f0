/*
 * wbc_ref.c - C restatement of a dense QP solve for the whole-body-control task-space QP (TEST / MEASUREMENT
 * INFRASTRUCTURE ONLY; CPU baseline of bench.py for BASELINE.json configs[3]).
 *
 * The reference hands the TSID QP to a `wbc::QPSolver` plugin of ARC-OPT (wbc_arc_opt.cpp:51-74,297; default
 * EiquadprogSolver, alternatives ProxQP / qpOASES / OSQP / qpSWIFT / HPIPM).  Those are third-party libraries that are not
 * in /root/reference and cannot be built here; this file restates a textbook dense primal-dual interior-point method
 * (Mehrotra predictor-corrector, Nocedal & Wright ch. 16.6) for
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

static int cholq(double *A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 0)) return 1;
    d = sqrt(d);
    A[j * ld + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s / d;
    }
  }
  return 0;
}
static void cholq_solve(const double *L, int n, int ld, double *b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * ld + k] * b[k]; b[i] = s / L[i * ld + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * b[k]; b[i] = s / L[i * ld + i]; }
}

static int wbc_one(int nq, int neq, int nin, const double *H, const double *g, const double *A, const double *lo, const double *hi,
                   double tol, int max_iter, double *xout, wbc_info *info) {
  const int nc = neq + nin;
  double G[MAXI][MAXQ], h[MAXI], Ae[MAXE][MAXQ], be[MAXE];
  int m = 0;
  for (int i = 0; i < neq; ++i) { for (int j = 0; j < nq; ++j) Ae[i][j] = A[j * nc + i]; be[i] = lo[i]; }
  for (int i = neq; i < nc; ++i) {
    if (hi[i] < 1e19) { for (int j = 0; j < nq; ++j) G[m][j] = A[j * nc + i]; h[m++] = hi[i]; }
    if (lo[i] > -1e19) { for (int j = 0; j < nq; ++j) G[m][j] = -A[j * nc + i]; h[m++] = -lo[i]; }
  }
  double x[MAXQ] = {0}, nu[MAXE] = {0}, s[MAXI], z[MAXI];
  for (int i = 0; i < m; ++i) { s[i] = fmax(h[i], 1.0); z[i] = 1.0; }
  memset(info, 0, sizeof(*info));
  info->status = 2;
  int it;
  for (it = 0; it <= max_iter; ++it) {
    double rd[MAXQ], re[MAXE], rp[MAXI], res = 0, gap = 0, cmax = 0;
    for (int j = 0; j < nq; ++j) {
      double a = g[j];
      for (int k = 0; k < nq; ++k) a += H[k * nq + j] * x[k];
      for (int i = 0; i < neq; ++i) a += Ae[i][j] * nu[i];
      for (int i = 0; i < m; ++i) a += G[i][j] * z[i];
      rd[j] = a; res = fmax(res, fabs(a));
    }
    for (int i = 0; i < neq; ++i) { double a = -be[i]; for (int j = 0; j < nq; ++j) a += Ae[i][j] * x[j]; re[i] = a; res = fmax(res, fabs(a)); }
    for (int i = 0; i < m; ++i) { double a = s[i] - h[i]; for (int j = 0; j < nq; ++j) a += G[i][j] * x[j]; rp[i] = a; res = fmax(res, fabs(a)); gap += s[i] * z[i]; cmax = fmax(cmax, s[i] * z[i]); }
    info->kkt[0] = res; info->kkt[3] = cmax;
    if (!(res < 1e300)) { info->status = 1; break; }
    if (res <= tol && cmax <= tol) { info->status = 0; break; }
    if (it == max_iter) break;
    const double mu_c = m ? gap / m : 0.0;
    /* Phi = H + G'WG (+ tiny proximal term), S = Ae Phi^-1 Ae' */
    double Phi[MAXQ * MAXQ], Y[MAXE][MAXQ], S[MAXE * MAXE];
    for (int j = 0; j < nq; ++j) for (int k = 0; k <= j; ++k) {
      double a = H[k * nq + j];
      for (int i = 0; i < m; ++i) a += G[i][j] * (z[i] / s[i]) * G[i][k];
      Phi[j * MAXQ + k] = a + (j == k ? 1e-9 : 0.0);
    }
    if (cholq(Phi, nq, MAXQ)) { info->status = (res <= 1e-7 && cmax <= 1e-7) ? 0 : 1; break; }
    for (int i = 0; i < neq; ++i) { memcpy(Y[i], Ae[i], sizeof(double) * nq); cholq_solve(Phi, nq, MAXQ, Y[i]); }
    for (int i = 0; i < neq; ++i) for (int k = 0; k <= i; ++k) { double a = 0; for (int j = 0; j < nq; ++j) a += Ae[i][j] * Y[k][j]; S[i * MAXE + k] = a + (i == k ? 1e-12 : 0.0); }
    if (neq && cholq(S, neq, MAXE)) { info->status = (res <= 1e-7 && cmax <= 1e-7) ? 0 : 1; break; } /* end game: accept at 1e-7 */
    double dxa[MAXQ], dsa[MAXI], dza[MAXI], sigma_mu = 0;
    for (int pass = 0; pass < 2; ++pass) {
      double rt[MAXQ], dx[MAXQ], dnu[MAXE], t[MAXQ];
      for (int j = 0; j < nq; ++j) rt[j] = rd[j];
      for (int i = 0; i < m; ++i) {
        double rc = s[i] * z[i] + (pass ? dsa[i] * dza[i] - sigma_mu : 0.0);
        double c = (z[i] * rp[i] - rc) / s[i];
        for (int j = 0; j < nq; ++j) rt[j] += G[i][j] * c;
      }
      /* [Phi Ae'; Ae 0][dx; dnu] = [-rt; -re] */
      for (int j = 0; j < nq; ++j) t[j] = -rt[j];
      cholq_solve(Phi, nq, MAXQ, t);
      for (int i = 0; i < neq; ++i) { double a = re[i]; for (int j = 0; j < nq; ++j) a += Ae[i][j] * t[j]; dnu[i] = a; }
      if (neq) cholq_solve(S, neq, MAXE, dnu);
      for (int j = 0; j < nq; ++j) { double a = t[j]; for (int i = 0; i < neq; ++i) a -= Y[i][j] * dnu[i]; dx[j] = a; }
      double ds[MAXI], dz[MAXI], tmax = 0;
      for (int i = 0; i < m; ++i) {
        double gd = 0; for (int j = 0; j < nq; ++j) gd += G[i][j] * dx[j];
        double rc = s[i] * z[i] + (pass ? dsa[i] * dza[i] - sigma_mu : 0.0);
        ds[i] = -rp[i] - gd;
        dz[i] = (-rc - z[i] * ds[i]) / s[i];
        tmax = fmax(tmax, fmax(-ds[i] / s[i], -dz[i] / z[i]));
      }
      double a = tmax > 1 ? 1 / tmax : 1;
      if (pass == 0) {
        double g2 = 0;
        for (int i = 0; i < m; ++i) { g2 += (s[i] + a * ds[i]) * (z[i] + a * dz[i]); dsa[i] = ds[i]; dza[i] = dz[i]; }
        double ratio = m ? (g2 / m) / fmax(mu_c, 1e-300) : 0;
        sigma_mu = ratio * ratio * ratio * mu_c;
        memcpy(dxa, dx, sizeof(double) * nq);
      } else {
        a = fmin(1.0, 0.995 * a);
        for (int j = 0; j < nq; ++j) x[j] += a * dx[j];
        for (int i = 0; i < neq; ++i) nu[i] += a * dnu[i];
        for (int i = 0; i < m; ++i) { s[i] += a * ds[i]; z[i] += a * dz[i]; }
      }
    }
  }
  info->iters = it;
  if (info->status == 0) memcpy(xout, x, sizeof(double) * nq);
  return info->status;
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
