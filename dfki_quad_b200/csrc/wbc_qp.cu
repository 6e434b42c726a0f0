// Kernel K3: batched dense QP for the whole-body-control task-space problem, one thread block per QP, FP64.
//     min 1/2 x'Hx + g'x   s.t.   Aeq x = beq,   lo <= Ain x <= hi          (nq <= 48, neq <= 32, nin <= 64)
// Replaces the `wbc::QPSolver` plugin call inside `wbc_scene_->solve(qp)`
// (ws/src/controllers/src/mit_controller/wbc_arc_opt.cpp:51-74,297): Eiquadprog / ProxQP / qpOASES / OSQP / qpSWIFT /
// HPIPM behind ARC-OPT.  Same Mehrotra predictor-corrector machinery as the MPC Riccati kernel (K2b), with the Newton
// system reduced by block elimination instead of a Riccati sweep:
//     Phi = H + Ain' W Ain = L L',   Y = L^-1 Aeq',   S = Y'Y = Ls Ls'
//     dnu = S^-1 (re + Y'u1),  dx = L^-T (u1 - Y dnu),   u1 = -L^-1 r~
// Cholesky factors and explicit triangular inverses (not Phi^-1: see K2b) live in shared memory; thread r owns the
// (s, z) pairs of inequality row r (lower and upper side) in registers.
#include "common.cuh"

namespace b200qp {

struct WbcParams {
  int nq, neq, nin;
  double kkt_tol;
  int max_iter;
  double tau, sig0;  // first attempt: fraction of the step to the boundary that is taken, end-game centring floor
};

namespace {
constexpr double WBIG = 1.0e300;
constexpr double WINF = 1.0e19;  // |bound| >= this: side absent

template <class Op>
__device__ __forceinline__ double wblock_reduce(double v, double *red, Op op) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double m = red[0];
  for (int w = 1; w < nw; ++w) m = op(m, red[w]);
  return m;
}
// min over the block of two values at once (one pair of barriers instead of two); red must hold 16 doubles
__device__ __forceinline__ void wblock_min2(double &a, double &b, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmin(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
  __syncthreads();
  if (lane == 0) {
    red[warp] = a;
    red[8 + warp] = b;
  }
  __syncthreads();
  double ma = red[0], mb = red[8];
  for (int w = 1; w < nw; ++w) {
    ma = fmin(ma, red[w]);
    mb = fmin(mb, red[8 + w]);
  }
  a = ma;
  b = mb;
}
// five maxima (m[0..4]) and one sum (m[5]) over the block with one pair of barriers; the kernel runs 64 threads, so red's 16 doubles
// hold the 6 partial results of both warps
__device__ __forceinline__ void wblock_max5_sum1(double (&m)[6], double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 5; ++k) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
    m[5] += __shfl_xor_sync(0xffffffffu, m[5], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) red[warp * 6 + k] = m[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 5; ++k) m[k] = fmax(red[k], red[6 + k]);
  m[5] = red[5] + red[11];
}
__device__ __forceinline__ void wwarp_min2(double &a, double &b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = fmin(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmin(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
}
__device__ __forceinline__ double wwarp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// e-th entry of a lower triangle stored by rows: e = i (i + 1) / 2 + j, 0 <= j <= i
__device__ __forceinline__ void tri_index(int e, int &i, int &j) {
  int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
  if (r * (r + 1) / 2 > e) --r;
  else if ((r + 1) * (r + 2) / 2 <= e) ++r;
  i = r;
  j = e - r * (r + 1) / 2;
}
struct WSum { __device__ double operator()(double a, double b) const { return a + b; } };

// In-place lower Cholesky of the n x n row-major matrix A (ld = n) by the whole block, then A <- L^-1 (lower triangle, row-major).
__device__ void chol_and_inverse(double *A, int n, double *diag0) {
  const int tid = threadIdx.x, BS = blockDim.x;
  if (tid < n) diag0[tid] = A[tid * n + tid];
  __syncthreads();
  for (int p = 0; p < n; ++p) {
    const double flo = 1e-13 * diag0[p];
    const double piv0 = A[p * n + p];
    const double piv = piv0 > flo ? piv0 : flo;
    const double rd = fast_rsqrt(piv);  // reciprocal square root once per pivot: no FP64 divisions in the column scaling / L^-1
    __syncthreads();
    for (int i = p + tid; i < n; i += BS) A[i * n + p] = (i == p) ? piv * rd : A[i * n + p] * rd;
    if (tid == 0) diag0[p] = rd;  // from here on diag0[p] = 1 / L[p][p]
    __syncthreads();
    // trailing update on an 8 x 8 thread grid (no integer division by the shrinking remainder)
    for (int i = p + 1 + (tid >> 3); i < n; i += 8) {
      const double lip = A[i * n + p];
      for (int j = p + 1 + (tid & 7); j <= i; j += 8) A[i * n + j] -= lip * A[j * n + p];
    }
    __syncthreads();
  }
  // L^-1 IN PLACE (the factor is not needed afterwards, and a separate array would cost an n x n block of shared memory = two
  // resident blocks per SM): columns from the last to the first, X[i][j] = -X[j][j] * sum_{k=j+1..i} X[i][k] L[k][j], thread i
  // owns row i; column j of L is only overwritten after every row has read it.
  for (int j = n - 1; j >= 0; --j) {
    double t = 0.0;
    if (tid > j && tid < n) {
      for (int k = j + 1; k <= tid; ++k) t = fma(A[tid * n + k], A[k * n + j], t);
      t *= -diag0[j];
    }
    __syncthreads();
    if (tid > j && tid < n) A[tid * n + j] = t;
    else if (tid == j) A[j * n + j] = diag0[j];
    __syncthreads();
  }
}

// The same for a compile-time size N <= 32.  Two alternatives to the block version, selectable per half for A/B builds:
//   bit 0  Cholesky by ONE warp (lane i owns row i, the scaled pivot column goes through shared memory, __syncwarp between steps):
//          measured SLOWER (884 k vs 975 k QPs/s) - the 8 x 8 trailing update of the block version uses both warps;
//   bit 1  L^-1 with lane j owning COLUMN j: its own forward substitution L x = e_j with x in registers (fully unrolled, the L entries
//          are warp-uniform broadcast loads), so the columns need no synchronisation at all; written back in place afterwards:
//          measured faster (1010 k/s), this is what ships.
//   (a third one, the Cholesky with row i in the registers of lane i and the multipliers moved by 64-bit warp shuffles, ran at 565 k/s:
//   435 dependent shuffle + FMA pairs per factorisation - the same lesson as in the block Riccati kernel - and is not kept.)
//   (a blocked Cholesky - four columns per step, the 4 x 4 diagonal block factored redundantly in registers, two barriers per step instead
//   of three per column - is exactly as fast as the column version, 1005 vs 1014 k/s: the barriers are not what the step waits for.)
// Called by every thread of the block.
#ifndef B200QP_WBC_VARIANT
#define B200QP_WBC_VARIANT 2  // Phi (30 x 30). bit 0: warp Cholesky, bit 1: column-parallel inverse; A/B builds
#endif
#ifndef B200QP_WBC_VARIANT_S
#define B200QP_WBC_VARIANT_S 2  // S (18 x 18)
#endif
template <int N, int V>
__device__ __noinline__ void chol_and_inverse_warp(double *A, double *dinv) {
  static_assert(N <= 32, "one lane per row / column");
  const int tid = threadIdx.x, BS = blockDim.x;
  if (V & 1) {
    if (threadIdx.x < 32) {
      const int i = threadIdx.x;
      const double d0 = (i < N) ? A[i * N + i] : 1.0;
#pragma unroll 1
      for (int p = 0; p < N; ++p) {
        const double piv0 = A[p * N + p];
        const double flo = 1e-13 * __shfl_sync(0xffffffffu, d0, p);
        const double piv = piv0 > flo ? piv0 : flo;
        const double rd = fast_rsqrt(piv);
        __syncwarp();  // every lane has read the pivot before lane p overwrites it
        double lip = 0.0;
        if (i >= p && i < N) {
          lip = (i == p) ? piv * rd : A[i * N + p] * rd;
          A[i * N + p] = lip;
        }
        if (i == p) dinv[p] = rd;
        __syncwarp();
        if (i > p && i < N) {
          double *Ai = A + i * N;
          int j = p + 1;
          for (; j + 1 <= i; j += 2) {  // two independent elements per trip
            const double l0 = A[j * N + p], l1 = A[(j + 1) * N + p];
            const double a0 = Ai[j], a1 = Ai[j + 1];
            Ai[j] = fma(-lip, l0, a0);
            Ai[j + 1] = fma(-lip, l1, a1);
          }
          if (j <= i) Ai[j] = fma(-lip, A[j * N + p], Ai[j]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
  } else {  // block Cholesky (as chol_and_inverse)
    if (tid < N) dinv[tid] = A[tid * N + tid];
    __syncthreads();
    for (int p = 0; p < N; ++p) {
      const double flo = 1e-13 * dinv[p];
      const double piv0 = A[p * N + p];
      const double piv = piv0 > flo ? piv0 : flo;
      const double rd = fast_rsqrt(piv);
      // two barriers per column: the diagonal of L is never read again (the inverse takes 1 / L_pp from dinv), so it is not written
      // and the pivot can be read without a barrier of its own; dinv[p] changes from "original diagonal" to 1 / L_pp only after
      // every thread has passed the barrier behind its read
      for (int i = p + 1 + tid; i < N; i += BS) A[i * N + p] *= rd;
      __syncthreads();
      if (tid == 0) dinv[p] = rd;
      for (int i = p + 1 + (tid >> 3); i < N; i += 8) {
        const double lip = A[i * N + p];
        for (int j = p + 1 + (tid & 7); j <= i; j += 8) A[i * N + j] -= lip * A[j * N + p];
      }
      __syncthreads();
    }
  }
  if (V & 2) {
    if (threadIdx.x < 32) {
      const int j = threadIdx.x;
      double x[N];
#pragma unroll
      for (int r = 0; r < N; ++r) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < r; k += 2) {
          s0 = fma(A[r * N + k], x[k], s0);
          s1 = fma(A[r * N + k + 1], x[k + 1], s1);
        }
        if (r & 1) s0 = fma(A[r * N + r - 1], x[r - 1], s0);
        const double di = dinv[r];
        x[r] = (r == j) ? di : ((r > j) ? -(s0 + s1) * di : 0.0);
      }
      __syncwarp();  // every lane is done reading L
      if (j < N) {
#pragma unroll
        for (int r = 0; r < N; ++r)
          if (r >= j) A[r * N + j] = x[r];
      }
    }
    __syncthreads();
  } else {
    for (int j = N - 1; j >= 0; --j) {
      double t = 0.0;
      if (tid > j && tid < N) {
        for (int k = j + 1; k <= tid; ++k) t = fma(A[tid * N + k], A[k * N + j], t);
        t *= -dinv[j];
      }
      __syncthreads();
      if (tid > j && tid < N) A[tid * N + j] = t;
      else if (tid == j) A[j * N + j] = dinv[j];
      __syncthreads();
    }
  }
}

// dx, dnu from r~ (srt) and re (sre):  u1 = -L^-1 r~,  dnu = S^-1 (re + Y'u1),  dx = L^-T (u1 - Y dnu)
// W1: called by warp 0 alone (every vector fits its 32 lanes' strides), __syncwarp instead of block barriers
template <bool W1>
__device__ __forceinline__ void wsync() {
  if constexpr (W1) __syncwarp();
  else __syncthreads();
}
template <bool W1>
__device__ __noinline__ void kkt_solve(int nq, int me, const double *sLi, const double *sY, const double *sSi, const double *srt,
                          const double *sre, double *su1, double *sv, double *sdx, double *sdnu) {
  const int tid = threadIdx.x, BS = W1 ? 32 : blockDim.x;
  for (int i = tid; i < nq; i += BS) {
    double a = 0.0;
    for (int k = 0; k <= i; ++k) a = fma(sLi[i * nq + k], srt[k], a);
    su1[i] = -a;
  }
  wsync<W1>();
  for (int r = tid; r < me; r += BS) {
    double a = sre[r];
    for (int i = 0; i < nq; ++i) a = fma(sY[i * me + r], su1[i], a);
    sv[r] = a;
  }
  wsync<W1>();
  for (int r = tid; r < me; r += BS) {
    double a = 0.0;
    for (int k = 0; k <= r; ++k) a = fma(sSi[r * me + k], sv[k], a);
    sdnu[r] = a;
  }
  wsync<W1>();
  double dn = 0.0;
  if (tid < me)
    for (int k = tid; k < me; ++k) dn = fma(sSi[k * me + tid], sdnu[k], dn);
  wsync<W1>();
  if (tid < me) sdnu[tid] = dn;
  wsync<W1>();
  for (int i = tid; i < nq; i += BS) {
    double a = su1[i];
    for (int r = 0; r < me; ++r) a = fma(-sY[i * me + r], sdnu[r], a);
    sv[i] = a;
  }
  wsync<W1>();
  for (int i = tid; i < nq; i += BS) {
    double a = 0.0;
    for (int k = i; k < nq; ++k) a = fma(sLi[k * nq + i], sv[k], a);
    sdx[i] = a;
  }
  wsync<W1>();
}
}  // namespace

// NQ_ / ME_ / MI_ > 0 fix the problem dimensions at compile time (Go2 reduced TSID: 30 / 18 / 28): the element loops divide by
// them (a runtime integer division is ~25 instructions) and their inner products unroll.  0 = take the dimension from P.
template <int NQ_, int ME_, int MI_>
__global__ void __launch_bounds__(64, 5) wbc_qp_kernel(const WbcParams P, int nprob, const double *__restrict__ Hg,
                                                    const double *__restrict__ gg, const double *__restrict__ Ag,
                                                    const double *__restrict__ log, const double *__restrict__ hig,
                                                    double *__restrict__ xout, b200qp_solver_info *__restrict__ info,
                                                    int *next) {
  extern __shared__ __align__(16) double sm[];
  const int nq = NQ_ > 0 ? NQ_ : P.nq, me = NQ_ > 0 ? ME_ : P.neq, mi = NQ_ > 0 ? MI_ : P.nin, nc = me + mi;
  double *sH = sm;                  // [nq*nq] row-major (symmetric)
  double *sAe = sH + nq * nq;       // [me*nq] row-major
  double *sAi = sAe + me * nq;      // [mi*nq]
  double *sPhi = sAi + mi * nq;     // [nq*nq] Phi -> L
  double *sLi = sPhi;               // L^-1 overwrites the factor in place
  double *sY = sPhi + nq * nq;      // [nq*me]  Y = L^-1 Aeq'
  double *sS = sY + nq * me;        // [me*me]  S -> Ls -> Ls^-1
  double *sSi = sS;
  double *sx = sS + me * me;        // [nq]
  double *sg = sx + nq;             // [nq]
  double *snu = sg + nq;            // [me]
  double *sbe = snu + me;           // [me]
  double *srd = sbe + me;           // [nq]
  double *sre = srd + nq;           // [me]
  double *srt = sre + me;           // [nq]  r~
  double *su1 = srt + nq;           // [nq]
  double *sv = su1 + nq;            // [max(nq,me)] scratch
  double *sdx = sv + (nq > me ? nq : me);  // [nq]
  double *sdnu = sdx + nq;          // [me]
  double *sw = sdnu + me;           // [mi] row weights / row scalars
  double *sd0 = sw + mi;            // [nq] diag scratch
  double *sbx = sd0 + nq;           // [nq] best iterate
  double *sdx2 = sbx + nq;          // [nq] refinement step
  double *sdnu2 = sdx2 + nq;        // [me]
  double *sre2 = sdnu2 + me;        // [me]
  double *red = sre2 + me;          // [16]
  __shared__ int sSlot;
  const int tid = threadIdx.x, BS = blockDim.x;
  constexpr bool ONEW = (NQ_ > 0 && NQ_ <= 32 && ME_ <= 32 && MI_ <= 32);
  const int BSW = ONEW ? 32 : BS;

  for (;;) {
    __syncthreads();
    if (tid == 0) sSlot = atomicAdd(next, 1);
    __syncthreads();
    const int prob = sSlot;
    if (prob >= nprob) break;
    const double *Hp = Hg + (size_t)prob * nq * nq, *Ap = Ag + (size_t)prob * nc * nq;
    // H column-major == row-major for the symmetric Hessian; A column-major nc x nq -> rows
    for (int e = tid; e < nq * nq; e += BS) sH[e] = Hp[e];
    for (int e = tid; e < nc * nq; e += BS) {
      const int col = e / nc, r = e - col * nc;
      const double v = Ap[e];
      if (r < me) sAe[r * nq + col] = v;
      else sAi[(r - me) * nq + col] = v;
    }
    for (int e = tid; e < nq; e += BS) {
      sg[e] = gg[(size_t)prob * nq + e];
      sx[e] = 0.0;
    }
    for (int e = tid; e < me; e += BS) {
      sbe[e] = log[(size_t)prob * nc + e];
      snu[e] = 0.0;
    }
    // inequality row of this thread
    const bool row = tid < mi;
    double lo = -WBIG, hi = WBIG;
    if (row) {
      lo = log[(size_t)prob * nc + me + tid];
      hi = hig[(size_t)prob * nc + me + tid];
    }
    const bool hasl = row && lo > -WINF, hasu = row && hi < WINF;
    double sl = hasl ? fmax(-lo, 1.0) : 1.0, su = hasu ? fmax(hi, 1.0) : 1.0;  // x = 0 start: slack = max(bound distance, 1)
    double zl = hasl ? 1.0 : 0.0, zu = hasu ? 1.0 : 0.0;
    const double mtot = fmax(wblock_reduce((hasl ? 1.0 : 0.0) + (hasu ? 1.0 : 0.0), red, WSum()), 1.0);
    __syncthreads();
    // augmented-Lagrangian form of the same QP: H += rho Aeq'Aeq, g -= rho Aeq'beq keeps the minimiser and makes Phi
    // positive definite along directions that only the equalities determine (stance-leg joint accelerations)
    {
      const double rho_eq = 10.0;
      for (int e = tid; e < nq * nq; e += BS) {
        const int i = e / nq, j = e - i * nq;
        double a = 0.0;
        for (int r = 0; r < me; ++r) a = fma(sAe[r * nq + i], sAe[r * nq + j], a);
        sH[e] += rho_eq * a;
      }
      for (int i = tid; i < nq; i += BS) {
        double a = 0.0;
        for (int r = 0; r < me; ++r) a = fma(sAe[r * nq + i], sbe[r], a);
        sg[i] -= rho_eq * a;
      }
    }
    __syncthreads();

    int status = B200QP_ACADOS_MAXITER, iter = 0, iter_first = 0;
#ifdef B200QP_WBC_TIMING
    long long wcyc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, wlast = clock64();
#define WPHW(idx) do { if (tid == 0) { const long long t__ = clock64(); wcyc[idx] += t__ - wlast; wlast = t__; } } while (0)
#else
#define WPHW(idx) do { } while (0)
#endif
    WPHW(0);
    double k_stat = 0.0, k_eq = 0.0, k_ineq = 0.0, k_comp = 0.0, bk[4] = {0.0, 0.0, 0.0, 0.0};
    const double target = 1e-4 * P.kkt_tol;
    // Two attempts. The first is aggressive: P.tau (0.9999) of the way to the boundary and an end-game centring floor of only P.sig0
    // (0.01) - 7.5 instead of 10.1 iterations - but about one QP in 3000 then stalls next to a face or at the 1e-6 stationarity
    // floor. It is given up after 20 iterations and the QP is solved again from the start with the conservative 0.995 / 0.05, which
    // has converged on every QP seen (everything the second attempt needs is still in shared memory).
    double tau = P.tau;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
    status = B200QP_ACADOS_MAXITER;
    double best = WBIG;
    int stall = 0;
    double rpl = 0.0, rpu = 0.0, mu_c = 1.0, mu_prev = WBIG;
    int damp = 0;
    // the first attempt gives up early (converging QPs need at most 19 iterations): a straggler at the end of the work queue costs the whole launch its time
    const int cap = (attempt == 0 && tau > 0.995 && P.max_iter > 20) ? 20 : P.max_iter;
    for (iter = 0; iter <= cap + 1; ++iter) {
      const bool init = (iter == 0);  // iteration 0 computes the starting point (least-squares point + shifts)
      if (!init) {
      // residuals (warp 0 alone where everything fits its lanes, then one block barrier hands the six norms to both warps) --------
      double y = 0.0, lm = 0.0, le = 0.0;
      bool bad = false;
      if (!ONEW || tid < 32) {
        if (row)
          for (int j = 0; j < nq; ++j) y = fma(sAi[tid * nq + j], sx[j], y);
        rpl = hasl ? (-y + sl + lo) : 0.0;
        rpu = hasu ? (y + su - hi) : 0.0;
        if (row) sw[tid] = zu - zl;
        for (int e = tid; e < me; e += BSW) {
          double a = -sbe[e];
          for (int j = 0; j < nq; ++j) a = fma(sAe[e * nq + j], sx[j], a);
          sre[e] = a;
        }
        wsync<ONEW>();
        for (int i = tid; i < nq; i += BSW) {
          double a = sg[i];
          for (int j = 0; j < nq; ++j) a = fma(sH[i * nq + j], sx[j], a);
          for (int r = 0; r < me; ++r) a = fma(sAe[r * nq + i], snu[r], a);
          for (int r = 0; r < mi; ++r) a = fma(sAi[r * nq + i], sw[r], a);
          srd[i] = a;
          lm = fmax(lm, fabs(a));
        }
        for (int e = tid; e < me; e += BSW) le = fmax(le, fabs(sre[e]));
        bad = !(lm < WBIG) || !(fabs(y) < WBIG);
      }
      double m6[6] = {lm, le, fmax(fabs(rpl), fabs(rpu)), fmax(hasl ? sl * zl : 0.0, hasu ? su * zu : 0.0),
                      fmax(fmax(hasl ? lo - y : 0.0, hasu ? y - hi : 0.0), 0.0), (hasl ? sl * zl : 0.0) + (hasu ? su * zu : 0.0)};
      int anybad;
      if constexpr (ONEW) {
        if (tid < 32) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 5; ++k) m6[k] = fmax(m6[k], __shfl_xor_sync(0xffffffffu, m6[k], o));
            m6[5] += __shfl_xor_sync(0xffffffffu, m6[5], o);
          }
          const int ab = __any_sync(0xffffffffu, bad);
          if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) red[k] = m6[k];
            red[6] = ab ? 1.0 : 0.0;
          }
        }
        __syncthreads();  // also publishes srd / sre; red is not written again before the barriers of the factorisation
#pragma unroll
        for (int k = 0; k < 6; ++k) m6[k] = red[k];
        anybad = red[6] != 0.0;
      } else {
        anybad = __syncthreads_or(bad ? 1 : 0);
        wblock_max5_sum1(m6, red);
      }
      k_stat = m6[0];
      k_eq = m6[1];
      const double rpm = m6[2];
      k_comp = m6[3];
      k_ineq = m6[4];
      const double gap = m6[5];
      const double rmax = anybad ? WBIG : fmax(fmax(k_stat, k_eq), fmax(rpm, k_comp));
      if (rmax <= target) {
        status = B200QP_ACADOS_SUCCESS;
        for (int i = tid; i < nq; i += BS) sbx[i] = sx[i];
        break;
      }
      if (rmax < 0.9 * best) {  // remember the best iterate: the endgame hits the round-off floor of large-scale data
        best = rmax;
        stall = 0;
        bk[0] = k_stat; bk[1] = k_eq; bk[2] = k_ineq; bk[3] = k_comp;
        for (int i = tid; i < nq; i += BS) sbx[i] = sx[i];
      } else {
        ++stall;
      }
      if (anybad || iter > cap || (stall >= 2 && best <= P.kkt_tol)) {
        if (best <= P.kkt_tol) status = B200QP_ACADOS_SUCCESS;
        else status = anybad ? B200QP_ACADOS_NAN_DETECTED : B200QP_ACADOS_MAXITER;
        k_stat = bk[0]; k_eq = bk[1]; k_ineq = bk[2]; k_comp = bk[3];
        break;
      }
      mu_c = gap / mtot;
      // Two-cycle guard: Mehrotra's centring can alternate between a long and a short step without reducing the gap (seen on
      // ~1 of 1000 QPs: mu 90 <-> 570 for 40 iterations).  When the gap did not fall by 10 %, centre the next two steps.
      if (mu_c > 0.9 * mu_prev) damp = 2;
      mu_prev = mu_c;
      }  // !init

      WPHW(1);
      // Phi = H + Ain' W Ain ------------------------------------------------------------------------------------------
      if (row) sw[tid] = init ? ((hasl ? 1.0 : 0.0) + (hasu ? 1.0 : 0.0)) : ((hasl ? zl / sl : 0.0) + (hasu ? zu / su : 0.0));
      __syncthreads();
      if constexpr (NQ_ > 0 && MI_ * NQ_ <= NQ_ * ME_ + ME_ * ME_) {
        // the rows scaled once (W Ain, into the not yet needed Y / S area), then only the 465 lower-triangle entries are walked
        // (the square walk skipped every second slot and left half the lanes idle); same roundings as the generic loop below
        double *sT = sY;
        for (int e = tid; e < mi * nq; e += BS) sT[e] = sAi[e] * sw[e / nq];
        __syncthreads();
        for (int e = tid; e < nq * (nq + 1) / 2; e += BS) {
          int i, j;
          tri_index(e, i, j);
          double a = sH[i * nq + j];
          for (int r = 0; r < mi; ++r) a = fma(sT[r * nq + i], sAi[r * nq + j], a);
          sPhi[i * nq + j] = a;
        }
      } else {
        for (int e = tid; e < nq * nq; e += BS) {
          const int i = e / nq, j = e - i * nq;
          if (j > i) continue;
          double a = sH[e];
          for (int r = 0; r < mi; ++r) a = fma(sAi[r * nq + i] * sw[r], sAi[r * nq + j], a);
          sPhi[e] = a;
        }
      }
      __syncthreads();
      WPHW(2);
      if constexpr (NQ_ > 0 && NQ_ <= 32) chol_and_inverse_warp<(NQ_ > 0 ? NQ_ : 1), B200QP_WBC_VARIANT>(sPhi, sd0);
      else chol_and_inverse(sPhi, nq, sd0);
      WPHW(3);
      for (int e = tid; e < nq * me; e += BS) {  // Y[i][r] = sum_{k<=i} Li[i][k] Aeq[r][k]
        const int i = e / me, r = e - i * me;
        double a = 0.0;
        for (int k = 0; k <= i; ++k) a = fma(sLi[i * nq + k], sAe[r * nq + k], a);
        sY[e] = a;
      }
      __syncthreads();
      for (int e = tid; e < me * (me + 1) / 2; e += BS) {  // S = Y'Y, lower triangle
        int r, c;
        tri_index(e, r, c);
        double a = 0.0;
        for (int i = 0; i < nq; ++i) a = fma(sY[i * me + r], sY[i * me + c], a);
        sS[r * me + c] = a;
      }
      __syncthreads();
      WPHW(4);
      if constexpr (NQ_ > 0 && ME_ <= 32) chol_and_inverse_warp<(ME_ > 0 ? ME_ : 1), B200QP_WBC_VARIANT_S>(sS, sd0);
      else chol_and_inverse(sS, me, sd0);
      WPHW(5);

      if (init) {  // starting point as in CVXOPT's coneqp: solve with W = I, then shift s and z into the interior
        if (row) sw[tid] = (hasu ? hi : 0.0) - (hasl ? lo : 0.0);
        for (int e = tid; e < me; e += BS) sre[e] = -sbe[e];
        __syncthreads();
        for (int i = tid; i < nq; i += BS) {
          double a = sg[i];
          for (int r = 0; r < mi; ++r) a = fma(-sAi[r * nq + i], sw[r], a);
          srt[i] = a;
        }
        __syncthreads();
        kkt_solve<false>(nq, me, sLi, sY, sSi, srt, sre, su1, sv, sdx, sdnu);
        for (int i = tid; i < nq; i += BS) sx[i] = sdx[i];
        for (int r = tid; r < me; r += BS) snu[r] = sdnu[r];
        __syncthreads();
        double y0 = 0.0;
        if (row)
          for (int j = 0; j < nq; ++j) y0 = fma(sAi[tid * nq + j], sx[j], y0);
        const double zu0 = y0 - hi, zl0 = -y0 + lo;  // z = G x - h, s = -z
        double smin = fmin(hasu ? -zu0 : WBIG, hasl ? -zl0 : WBIG), zmin = fmin(hasu ? zu0 : WBIG, hasl ? zl0 : WBIG);
        wblock_min2(smin, zmin, red);
        const double sh = smin > 0.0 ? 0.0 : 1.0 - smin, zh = zmin > 0.0 ? 0.0 : 1.0 - zmin;
        su = hasu ? -zu0 + sh : 1.0;
        sl = hasl ? -zl0 + sh : 1.0;
        zu = hasu ? zu0 + zh : 0.0;
        zl = hasl ? zl0 + zh : 0.0;
        __syncthreads();
        continue;
      }
      WPHW(6);
      // predictor / corrector ---------------------------------------------------------------------------------------------
      // With every row and vector inside one warp's reach (Go2: 28 rows, 30 + 18 unknowns) the whole section runs on warp 0 with warp
      // barriers and shuffle reductions; warp 1, which would only wait at ~30 block barriers, waits once at the end.
      double dsl = 0.0, dsu = 0.0, dzl = 0.0, dzu = 0.0, sigma_mu = 0.0;
      if (!ONEW || tid < 32)
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const double rcl = (pass == 0) ? sl * zl : (sl * zl + dsl * dzl - sigma_mu);
        const double rcu = (pass == 0) ? su * zu : (su * zu + dsu * dzu - sigma_mu);
        if (row) sw[tid] = (hasu ? (zu * rpu - rcu) / su : 0.0) - (hasl ? (zl * rpl - rcl) / sl : 0.0);
        wsync<ONEW>();
        for (int i = tid; i < nq; i += BSW) {  // r~ = rd + Ain' t
          double a = srd[i];
          for (int r = 0; r < mi; ++r) a = fma(sAi[r * nq + i], sw[r], a);
          srt[i] = a;
        }
        wsync<ONEW>();
        kkt_solve<ONEW>(nq, me, sLi, sY, sSi, srt, sre, su1, sv, sdx, sdnu);
        if (pass == 1 && mu_c < 1e-3) {
          // one step of iterative refinement on the Newton system (Phi dx + Aeq'dnu = -r~, Aeq dx = -re): with W up to
          // 1e12 the factored solve alone stalls the stationarity residual at the 1e-6 level on large-scale data
          double q = 0.0;
          if (row) {
            for (int j = 0; j < nq; ++j) q = fma(sAi[tid * nq + j], sdx[j], q);
            sw[tid] = q * ((hasl ? zl / sl : 0.0) + (hasu ? zu / su : 0.0));
          }
          wsync<ONEW>();
          for (int i = tid; i < nq; i += BSW) {
            double a = srt[i];
            for (int j = 0; j < nq; ++j) a = fma(sH[i * nq + j], sdx[j], a);
            for (int r = 0; r < mi; ++r) a = fma(sAi[r * nq + i], sw[r], a);
            for (int r = 0; r < me; ++r) a = fma(sAe[r * nq + i], sdnu[r], a);
            sd0[i] = a;  // = -(residual of the first block row)
          }
          for (int e = tid; e < me; e += BSW) {
            double a = sre[e];
            for (int j = 0; j < nq; ++j) a = fma(sAe[e * nq + j], sdx[j], a);
            sre2[e] = a;
          }
          wsync<ONEW>();
          kkt_solve<ONEW>(nq, me, sLi, sY, sSi, sd0, sre2, su1, sv, sdx2, sdnu2);
          for (int i = tid; i < nq; i += BSW) sdx[i] += sdx2[i];
          for (int e = tid; e < me; e += BSW) sdnu[e] += sdnu2[e];
          wsync<ONEW>();
        }
        double dy = 0.0;
        if (row)
          for (int j = 0; j < nq; ++j) dy = fma(sAi[tid * nq + j], sdx[j], dy);
        const double nsl = hasl ? (-rpl + dy) : 0.0, nsu = hasu ? (-rpu - dy) : 0.0;
        const double nzl = hasl ? (-rcl - zl * nsl) / sl : 0.0, nzu = hasu ? (-rcu - zu * nsu) / su : 0.0;
        // primal and dual step lengths are limited separately (s by the primal direction, z by the dual one): a common step lets
        // one blocking multiplier stall the primal progress for the rest of the solve (the ~0.1 % of QPs that ran into max_iter)
        double stp = 1.0, stq = 1.0;
        if (hasl) {
          if (nsl < 0.0) stp = fmin(stp, -sl / nsl);
          if (nzl < 0.0) stq = fmin(stq, -zl / nzl);
        }
        if (hasu) {
          if (nsu < 0.0) stp = fmin(stp, -su / nsu);
          if (nzu < 0.0) stq = fmin(stq, -zu / nzu);
        }
        if constexpr (ONEW) wwarp_min2(stp, stq);
        else wblock_min2(stp, stq, red);
        if (pass == 0) {
          const double g2l = (hasl ? (sl + stp * nsl) * (zl + stq * nzl) : 0.0) + (hasu ? (su + stp * nsu) * (zu + stq * nzu) : 0.0);
          double g2;
          if constexpr (ONEW) g2 = wwarp_sum(g2l);
          else g2 = wblock_reduce(g2l, red, WSum());
          const double ratio = (g2 / mtot) / fmax(mu_c, 1e-300);
          sigma_mu = ratio * ratio * ratio * mu_c;
          // end game: keep some centring once the gap is small - with sigma -> 0 single pairs run ahead of the path (s z << mu), W
          // reaches 1e16 and the factorisation of Phi loses the stationarity residual at the 1e-6 level
          if (mu_c < 1e-4) sigma_mu = fmax(sigma_mu, (tau > 0.995 ? P.sig0 : 0.05) * mu_c);
          if (damp > 0) sigma_mu = fmax(sigma_mu, 0.2 * mu_c);
          dsl = nsl; dsu = nsu; dzl = nzl; dzu = nzu;
        } else {
          const double a = fmin(1.0, tau * stp), ad = fmin(1.0, tau * stq);
          sl += a * nsl; su += a * nsu; zl += ad * nzl; zu += ad * nzu;
          for (int i = tid; i < nq; i += BSW) sx[i] += a * sdx[i];
          for (int r = tid; r < me; r += BSW) snu[r] += ad * sdnu[r];
          if (damp > 0) --damp;
        }
        wsync<ONEW>();
      }
      if constexpr (ONEW) __syncthreads();
      WPHW(7);
    }
    __syncthreads();
    if (status == B200QP_ACADOS_SUCCESS || tau <= 0.995) break;
    tau = 0.995;
    iter_first = iter;
    }  // attempt
    __syncthreads();
    if (status == B200QP_ACADOS_SUCCESS)
      for (int i = tid; i < nq; i += BS) xout[(size_t)prob * nq + i] = sbx[i];
#ifdef B200QP_WBC_TIMING
    if (tid == 0) for (int q = 0; q < 10; ++q) xout[(size_t)prob * nq + q] = (double)wcyc[q];
#endif
    if (tid == 0) {
      b200qp_solver_info inf;
      inf.status = status;
      inf.iterations = (iter > 0 ? iter - 1 : 0) + iter_first;
      inf.polish_rounds = 0;
      inf.n_free = nq;
      inf.kkt[0] = k_stat;
      inf.kkt[1] = k_eq;
      inf.kkt[2] = k_ineq;
      inf.kkt[3] = k_comp;
      info[prob] = inf;
    }
  }
}

size_t wbc_smem_bytes(const WbcParams &P) {
  const size_t nq = P.nq, me = P.neq, mi = P.nin;
  const size_t d = nq * nq * 2 + me * nq + mi * nq + nq * me + me * me + 10 * nq + 7 * me + (nq > me ? nq : me) + mi + 16;
  return d * sizeof(double) + 64;
}

template <int NQ_, int ME_, int MI_>
static int wbc_occupancy(size_t smb) {
  int occ = 0;
  if (cudaFuncSetAttribute(wbc_qp_kernel<NQ_, ME_, MI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wbc_qp_kernel<NQ_, ME_, MI_>, 64, smb) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return occ < 1 ? 1 : occ;
}

int wbc_grid(const WbcParams &P, int num_sms) {
  const size_t smb = wbc_smem_bytes(P);
  const int occ = (P.nq == 30 && P.neq == 18 && P.nin == 28) ? wbc_occupancy<30, 18, 28>(smb) : wbc_occupancy<0, 0, 0>(smb);
  return occ < 0 ? -1 : num_sms * occ;
}

cudaError_t launch_wbc_qp(const WbcParams &P, int n, const double *H, const double *g, const double *A, const double *lo,
                          const double *hi, double *x, b200qp_solver_info *info, int *next, int grid, cudaStream_t st) {
  if (grid > n) grid = n;
  if (P.nq == 30 && P.neq == 18 && P.nin == 28)
    wbc_qp_kernel<30, 18, 28><<<grid, 64, wbc_smem_bytes(P), st>>>(P, n, H, g, A, lo, hi, x, info, next);
  else
    wbc_qp_kernel<0, 0, 0><<<grid, 64, wbc_smem_bytes(P), st>>>(P, n, H, g, A, lo, hi, x, info, next);
  return cudaGetLastError();
}

}  // namespace b200qp
