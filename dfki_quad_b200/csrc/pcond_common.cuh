// Shared device helpers of the block-condensed Riccati interior-point kernels (mpc_pcond.cu: one thread block per problem,
// mpc_pcw.cu: one warp per problem).  See mpc_pcond.cu for the formulation.
#pragma once
#include "common.cuh"

namespace b200qp {
namespace pc {
constexpr double BIGP = 1.0e300;

template <int V, class Op>
__device__ __forceinline__ void blk_reduce(double (&v)[V], double *red, Op op) {
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
struct PMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };
struct PMin { __device__ double operator()(double a, double b) const { return fmin(a, b); } };
struct PSum { __device__ double operator()(double a, double b) const { return a + b; } };

// G f for one foot; rows: +x -x +y -y +z -z  x-mu z  -x-mu z  y-mu z  -y-mu z   (boxes mpc.cpp:103-109, pyramid :66-92)
__device__ __forceinline__ void Gf(double fx, double fy, double fz, double mu, double (&g)[10]) {
  g[0] = fx; g[1] = -fx; g[2] = fy; g[3] = -fy; g[4] = fz; g[5] = -fz;
  g[6] = fx - mu * fz; g[7] = -fx - mu * fz; g[8] = fy - mu * fz; g[9] = -fy - mu * fz;
}
__device__ __forceinline__ void GTv(const double (&v)[10], double mu, double (&o)[3]) {
  o[0] = (v[0] - v[1]) + (v[6] - v[7]);
  o[1] = (v[2] - v[3]) + (v[8] - v[9]);
  o[2] = (v[4] - v[5]) - mu * ((v[6] + v[7]) + (v[8] + v[9]));
}

// offset of column a in the packed factor of a block with na rows (column a holds rows a .. na-1)
__device__ __forceinline__ int fcol(int a, int na) { return a * na - (a * (a - 1)) / 2; }

// Blocked (4 pivots per step) right-looking elimination of the first n4 columns of the symmetric matrix held in the lower
// triangle of K (element (i, j) at K[j * ld + i], thread i owns row i, i < na; ld odd => conflict-free rows and columns).
// K = [Huu Hux' hu; . Hxx hx; . . *] is the AUGMENTED matrix of one block: after n4 pivots the first n4 columns hold the
// Cholesky factor L of Huu (diagonal: 1 / L_ii), the rows below it Y' = (L^-1 Hux)' and y' = (L^-1 hu)', and the trailing
// 13 x 13 block the Schur complement [Hxx - Y'Y, hx - Y'y] = cost-to-go (P, p) of the block's first state - the Riccati
// step and the predictor's backward vector sweep in one factorisation, no explicit inverse.  Pivots are floored at 1e-13
// of the original diagonal (the interior-point end game drives Huu towards singularity along closed directions).
// lst: [4 * blockDim] staging of the block column, dst: [20] diagonal block + floors.
static __device__ __noinline__ void chol_aug(double *K, int ld, int n4, int na, double *lst, double *dst) {
  const int i = threadIdx.x;
  const double flo = (i < n4) ? 1e-13 * K[i * ld + i] : 0.0;
  for (int p0 = 0; p0 < n4; p0 += 4) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const bool mine = (i >= p0) && (i < na);
    if (mine) {
      a0 = K[p0 * ld + i];
      a1 = K[(p0 + 1) * ld + i];
      a2 = K[(p0 + 2) * ld + i];
      a3 = K[(p0 + 3) * ld + i];
      if (i < p0 + 4) {
        double *d = dst + 5 * (i - p0);
        d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
        d[4] = flo;
      }
    }
    __syncthreads();
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    if (mine) {
      const double r00 = fast_rsqrt(fmax(dst[0], dst[4]));
      const double l10 = dst[5] * r00, l20 = dst[10] * r00, l30 = dst[15] * r00;
      const double r11 = fast_rsqrt(fmax(dst[6] - l10 * l10, dst[9]));
      const double l21 = (dst[11] - l20 * l10) * r11, l31 = (dst[16] - l30 * l10) * r11;
      const double r22 = fast_rsqrt(fmax(dst[12] - l20 * l20 - l21 * l21, dst[14]));
      const double l32 = (dst[17] - l30 * l20 - l31 * l21) * r22;
      const double r33 = fast_rsqrt(fmax(dst[18] - l30 * l30 - l31 * l31 - l32 * l32, dst[19]));
      if (i >= p0 + 4) {
        x0 = a0 * r00;
        x1 = (a1 - x0 * l10) * r11;
        x2 = (a2 - x0 * l20 - x1 * l21) * r22;
        x3 = (a3 - x0 * l30 - x1 * l31 - x2 * l32) * r33;
        K[p0 * ld + i] = x0;
        K[(p0 + 1) * ld + i] = x1;
        K[(p0 + 2) * ld + i] = x2;
        K[(p0 + 3) * ld + i] = x3;
        *reinterpret_cast<double2 *>(lst + 4 * i) = make_double2(x0, x1);
        *reinterpret_cast<double2 *>(lst + 4 * i + 2) = make_double2(x2, x3);
      } else {
        const int r = i - p0;
        if (r == 0) {
          K[p0 * ld + i] = r00;
        } else if (r == 1) {
          K[p0 * ld + i] = l10;
          K[(p0 + 1) * ld + i] = r11;
        } else if (r == 2) {
          K[p0 * ld + i] = l20;
          K[(p0 + 1) * ld + i] = l21;
          K[(p0 + 2) * ld + i] = r22;
        } else {
          K[p0 * ld + i] = l30;
          K[(p0 + 1) * ld + i] = l31;
          K[(p0 + 2) * ld + i] = l32;
          K[(p0 + 3) * ld + i] = r33;
        }
      }
    }
    __syncthreads();
    if (i >= p0 + 4 && i < na) {  // trailing update of the own row, columns p0+4 .. i
      double *Ki = K + i;
      int j = p0 + 4;
      for (; j + 3 <= i; j += 4) {
        double2 la[4], lb[4];
        double kk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          la[u] = *reinterpret_cast<const double2 *>(lst + 4 * (j + u));
          lb[u] = *reinterpret_cast<const double2 *>(lst + 4 * (j + u) + 2);
          kk[u] = Ki[(j + u) * ld];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          Ki[(j + u) * ld] = fma(-x3, lb[u].y, fma(-x2, lb[u].x, fma(-x1, la[u].y, fma(-x0, la[u].x, kk[u]))));
      }
      for (; j <= i; ++j) {
        const double2 la = *reinterpret_cast<const double2 *>(lst + 4 * j);
        const double2 lb = *reinterpret_cast<const double2 *>(lst + 4 * j + 2);
        Ki[j * ld] = fma(-x3, lb.y, fma(-x2, lb.x, fma(-x1, la.y, fma(-x0, la.x, Ki[j * ld]))));
      }
    }
  }
  __syncthreads();
}

// Forward substitution with the packed factor F of a block (n4 pivots, nr rows in play): thread i < nr enters with acc = rhs_i;
// rows i < n4 leave with y_i = (L^-1 rhs)_i, rows n4 <= i < nr with acc_i - (Y'y)_i - which is the costate update
// p <- Abar'p - Y'y when the caller loaded (Abar'p) into those rows.  One barrier per 4 unknowns.  yst: [8].
__device__ __forceinline__ double tri_fwd(const double *F, int n4, int na, int nr, double acc, double *yst) {
  const int i = threadIdx.x;
  double yi = acc;
  int buf = 0;
  for (int p0 = 0; p0 < n4; p0 += 4, buf ^= 4) {
    if (i >= p0 && i < p0 + 4) yst[buf + (i - p0)] = acc;
    __syncthreads();
    if (i >= p0 && i < nr) {
      const double *c0 = F + fcol(p0, na), *c1 = F + fcol(p0 + 1, na), *c2 = F + fcol(p0 + 2, na), *c3 = F + fcol(p0 + 3, na);
      const double y0 = yst[buf] * c0[0];
      const double y1 = (yst[buf + 1] - c0[1] * y0) * c1[0];
      const double y2 = (yst[buf + 2] - c0[2] * y0 - c1[1] * y1) * c2[0];
      const double y3 = (yst[buf + 3] - c0[3] * y0 - c1[2] * y1 - c2[1] * y2) * c3[0];
      if (i >= p0 + 4) {
        acc -= c0[i - p0] * y0 + c1[i - p0 - 1] * y1 + c2[i - p0 - 2] * y2 + c3[i - p0 - 3] * y3;
        yi = acc;
      } else {
        const int r = i - p0;
        yi = (r == 0) ? y0 : (r == 1) ? y1 : (r == 2) ? y2 : y3;
      }
    }
  }
  __syncthreads();
  return yi;
}
// Back substitution L' x = t: thread a < n4 enters with acc = t_a and leaves with x_a.
__device__ __forceinline__ double tri_bwd(const double *F, int n4, int na, double acc, double *yst) {
  const int a = threadIdx.x;
  double xi = 0.0;
  int buf = 0;
  const double *ca = F + fcol(a < n4 ? a : 0, na) - a;  // L[i][a] = ca[i]
  for (int p0 = n4 - 4; p0 >= 0; p0 -= 4, buf ^= 4) {
    if (a >= p0 && a < p0 + 4) yst[buf + (a - p0)] = acc;
    __syncthreads();
    if (a < p0 + 4) {
      const double *c0 = F + fcol(p0, na), *c1 = F + fcol(p0 + 1, na), *c2 = F + fcol(p0 + 2, na), *c3 = F + fcol(p0 + 3, na);
      const double x3 = yst[buf + 3] * c3[0];
      const double x2 = (yst[buf + 2] - c2[1] * x3) * c2[0];
      const double x1 = (yst[buf + 1] - c1[1] * x2 - c1[2] * x3) * c1[0];
      const double x0 = (yst[buf] - c0[1] * x1 - c0[2] * x2 - c0[3] * x3) * c0[0];
      if (a < p0) {
        acc -= ca[p0] * x0 + ca[p0 + 1] * x1 + ca[p0 + 2] * x2 + ca[p0 + 3] * x3;
      } else {
        const int r = a - p0;
        xi = (r == 0) ? x0 : (r == 1) ? x1 : (r == 2) ? x2 : x3;
      }
    }
  }
  __syncthreads();
  return xi;
}
}  // namespace pc
}  // namespace b200qp
