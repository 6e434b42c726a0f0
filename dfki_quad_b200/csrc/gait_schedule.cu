// Kernel K4: gait schedule -> contact mask -> swing time -> force-box indexing, bit-exact with the strict IEEE-754
// evaluation of the reference expressions.  Compiled with -fmad=false (no contraction) and no fast-math.
//   Gait::update_sequence        ws/src/controllers/src/mit_controller/gait.cpp:67-112
//   MPC::GetUBounds              ws/src/controllers/src/mit_controller/mpc.cpp:94-115
// One thread per (robot, leg); the loop over steps is sequential because the early-contact override
// (gait.cpp:88-99) carries state from step to step.
#include "common.cuh"

namespace b200qp {

__global__ void gait_schedule_kernel(int n, int steps, double dt, const double *__restrict__ period,
                                     const double *__restrict__ duty, const double *__restrict__ offset,
                                     const double *__restrict__ phase, const uint8_t *__restrict__ measured,
                                     double fmin, double fmax, uint8_t *__restrict__ contact,
                                     double *__restrict__ swing_time, double *__restrict__ lbu,
                                     double *__restrict__ ubu) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * n) return;
  const int p = t >> 2, leg = t & 3;
  const double T = period[p], d = duty[t], off = offset[t], ph = phase[p];
  const bool meas = measured != nullptr && measured[t] != 0;
  const bool has_meas = measured != nullptr;
  bool early = false;
  const double base = (ph + 1.0) - off;  // (phase + 1.0 - phase_offset_[leg]), left to right
  for (int i = 0; i < steps; ++i) {
    const double leg_phase = fmod(base + ((double)(unsigned)i * dt) / T, 1.0);  // gait.cpp:82
    bool c = leg_phase < d;                                                      // gait.cpp:83
    if (i == 0 && has_meas) {                                                    // gait.cpp:85-90
      if (meas && !c && (leg_phase > ((d + 1.0) / 2.0))) early = true;
    }
    if (c) early = false;  // gait.cpp:92-96
    else if (early) c = true;
    const size_t o = ((size_t)p * steps + i) * 4 + leg;
    contact[o] = c ? 1 : 0;
    if (swing_time != nullptr) swing_time[o] = (!c && d > 0) ? (1 - leg_phase) * T : 0.0;  // gait.cpp:98-103
    if (lbu != nullptr) {  // mpc.cpp:103-113
      const size_t b = ((size_t)p * steps + i) * 12 + 3 * leg;
      if (c) {
        lbu[b] = -fmax; lbu[b + 1] = -fmax; lbu[b + 2] = fmin;
        ubu[b] = fmax;  ubu[b + 1] = fmax;  ubu[b + 2] = fmax;
      } else {
        lbu[b] = 0.0; lbu[b + 1] = 0.0; lbu[b + 2] = 0.0;
        ubu[b] = 0.0; ubu[b + 1] = 0.0; ubu[b + 2] = 0.0;
      }
    }
  }
}

cudaError_t launch_gait_schedule(int n, int steps, double dt, const double *period, const double *duty,
                                 const double *offset, const double *phase, const uint8_t *measured, double fmin,
                                 double fmax, uint8_t *contact, double *swing_time, double *lbu, double *ubu,
                                 cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int threads = 128;
  gait_schedule_kernel<<<(4 * n + threads - 1) / threads, threads, 0, st>>>(n, steps, dt, period, duty, offset, phase,
                                                                           measured, fmin, fmax, contact, swing_time,
                                                                           lbu, ubu);
  return cudaGetLastError();
}

}  // namespace b200qp

// ---- FP64 FMA peak microbenchmark (roofline denominator for the on-chip solve kernels; MEASURED_PEAKS.json has no
// FP64 entry).  8 independent DFMA chains per thread, 1024 threads per block, 8 blocks per SM.
namespace b200qp {
__global__ void fp64_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
cudaError_t fp64_peak(int num_sms, double *tflops) {
  const int blocks = num_sms * 2, threads = 1024, iters = 20000;
  double *d = nullptr;
  cudaError_t e = cudaMalloc(&d, (size_t)blocks * threads * sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads>>>(d, 100, 1.0);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double t = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (t > best) best = t;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return e;
}
}  // namespace b200qp
