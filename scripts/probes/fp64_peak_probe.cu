// Independent FP64 peak probe (not part of the library): 12 independent DFMA chains per thread, 1024 threads per SM
// resident, CUDA-event timing, best of 7.  Cross-checks zm_fp64_peak_flops (8 chains, 2048 threads per SM).
// build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak scripts/probes/fp64_peak_probe.cu && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) k(double* out, int iters, double m, double c) {
  double a[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) a[j] = 1.0 + 1e-3 * j + 1e-9 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 12; ++j) a[j] = fma(a[j], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 12; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 2, threads = 512, iters = 40000;
  double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, threads>>>(d, 1000, 0.999999, 1e-9);
  float best = 1e30f;
  for (int r = 0; r < 7; ++r) {
    cudaEventRecord(e0); k<<<blocks, threads>>>(d, iters, 0.999999, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  printf("{\"sms\": %d, \"fp64_tflops_probe\": %.3f, \"ms\": %.4f}\n", sms, 2.0 * 12.0 * iters * blocks * threads / (best * 1e-3) / 1e12, best);
  return cudaGetLastError() != cudaSuccess;
}
