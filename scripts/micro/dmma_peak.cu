// Microbenchmark: FP64 throughput of mma.sync.m8n8k4.f64 (DMMA) vs plain DFMA on this GPU.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters) {
  double c[CHAINS][2];
  for (int k = 0; k < CHAINS; ++k) c[k][0] = c[k][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) dmma(c[k][0], c[k][1], a, b);
  }
  double s = 0;
  for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters) {
  double c[CHAINS];
  for (int k = 0; k < CHAINS; ++k) c[k] = k;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) c[k] = fma(c[k], a, b);
  }
  double s = 0;
  for (int k = 0; k < CHAINS; ++k) s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps = 1; warps <= 16; warps *= 2) {
    const int threads = warps * 32, blocks = sms;
    float m1 = time_ms([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters); });
    float m4 = time_ms([&] { dmma_kernel<4><<<blocks, threads>>>(out, iters); });
    float m8 = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters); });
    float f8 = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters); });
    float f1 = time_ms([&] { dfma_kernel<1><<<blocks, threads>>>(out, iters); });
    auto tf_mma = [&](int ch, float ms) { return 2.0 * 256 * ch * (double)iters * warps * blocks / (ms * 1e-3) / 1e12; };
    auto tf_fma = [&](int ch, float ms) { return 2.0 * 32 * ch * (double)iters * warps * blocks / (ms * 1e-3) / 1e12; };
    // latency of a dependent chain in cycles (1 chain, 1 warp/SMSP cases are the interesting ones)
    printf("warps/SM %2d | DMMA TF/s: 1 chain %6.2f  4 chains %6.2f  8 chains %6.2f | DFMA TF/s: 1 chain %6.2f  8 chains %6.2f | "
           "dep. DMMA %.1f clk  dep. DFMA %.1f clk\n",
           warps, tf_mma(1, m1), tf_mma(4, m4), tf_mma(8, m8), tf_fma(1, f1), tf_fma(8, f8),
           m1 * 1e-3 * p.clockRate * 1e3 / iters, f1 * 1e-3 * p.clockRate * 1e3 / iters);
  }
  return 0;
}
