// FP64 pipe peak of the GPU (DFMA microbenchmark) -- the denominator for "FP64 pipe utilisation" of the
// instruction-bound kernels (north_star).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/tools/fp64_peak.cu -o /tmp/fp64_peak && /tmp/fp64_peak
// Every thread runs NACC independent fma chains (enough ILP to cover the DFMA latency); the grid fills every SM
// with resident warps.  Reports TFLOP/s (2 flops per DFMA) for several occupancies and the DFMA issue rate per SM
// and clock derived from the measured SM clock (clock64 ticks / elapsed time).
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b, long long* cycles) {
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; k++) acc[k] = (double)(threadIdx.x + k);
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = fma(acc[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NACC; k++) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NACC>
static void run(int sms, int blocks_per_sm, int threads, int iters) {
  double* out;
  long long* cyc;
  const int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads);
  cudaMalloc(&cyc, sizeof(long long));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_dfma<NACC><<<blocks, threads>>>(out, iters / 10, 0.999999, 1e-9, cyc);  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  long long hc = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_dfma<NACC><<<blocks, threads>>>(out, iters, 0.999999, 1e-9, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) { best = ms; cudaMemcpy(&hc, cyc, sizeof hc, cudaMemcpyDeviceToHost); }
  }
  const double fmas = (double)blocks * threads * (double)iters * NACC;
  const double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
  const double mhz = hc / (best * 1e-3) / 1e6;  // block 0's loop spans ~ the whole kernel
  const double per_sm_clk = fmas / sms / ((double)hc);
  printf("{\"nacc\": %d, \"blocks_per_sm\": %d, \"threads\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f, "
         "\"sm_mhz_est\": %.0f, \"dfma_per_sm_per_clk\": %.1f}\n",
         NACC, blocks_per_sm, threads, blocks_per_sm * threads / 32, best, tflops, mhz, per_sm_clk);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", p.name, p.multiProcessorCount, p.major, p.minor);
  const int sms = p.multiProcessorCount;
  run<8>(sms, 1, 128, 200000);
  run<8>(sms, 2, 256, 100000);
  run<8>(sms, 4, 256, 50000);
  run<8>(sms, 8, 256, 25000);
  run<4>(sms, 8, 256, 50000);
  run<16>(sms, 4, 256, 25000);
  run<2>(sms, 8, 256, 100000);
  run<1>(sms, 8, 256, 200000);
  return 0;
}
