// Instruction-cache microbenchmark for B200 (sm_100a): every warp loops over a straight-line body
// of N FFMA instructions (16 B each).  Reports time per warp-instruction per SM sub-partition
// (ideal: 1 cycle with 8 warps per sub-partition) for aligned warps and for warps that are spread
// over the body (each warp starts after a different delay), i.e. the access pattern of a kernel
// whose warps sweep a large code region at different times.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o icache tools/ubench/icache.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(1024, 1) body(float* out, int reps, int desync) {
  float a0 = threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  if (desync) {
    const long long t0 = clock64(), wait = (long long)desync * (threadIdx.x >> 5);
    while (clock64() - t0 < wait) {
    }
  }
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int i = 0; i < N / 8; i++) {
      a0 = fmaf(a0, 1.0001f, 0.5f);
      a1 = fmaf(a1, 1.0002f, 0.5f);
      a2 = fmaf(a2, 1.0003f, 0.5f);
      a3 = fmaf(a3, 1.0004f, 0.5f);
      a4 = fmaf(a4, 1.0005f, 0.5f);
      a5 = fmaf(a5, 1.0006f, 0.5f);
      a6 = fmaf(a6, 1.0007f, 0.5f);
      a7 = fmaf(a7, 1.0008f, 0.5f);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int N>
void run(float* out, int sms, double ghz, int desync) {
  const long long total = 1ll << 24;  // warp-instructions per warp
  const int reps = (int)(total / N);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  body<N><<<sms, 1024>>>(out, 16, desync);
  cudaEventRecord(e0);
  body<N><<<sms, 1024>>>(out, reps, desync);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * ghz * 1e9 / ((double)reps * N * 8.0);
  printf("body %5d instr = %5.1f KB  desync %6d : %8.2f ms  %.3f cycles per warp-instruction per sub-partition (at %.2f GHz)\n", N, N * 16 / 1024.0, desync, ms, cyc, ghz);
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double ghz = khz * 1e-6;
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int only = argc > 1 ? atoi(argv[1]) : 0;
  for (int desync : {0, 3000}) {
    if (!only || only == 512) run<512>(out, sms, ghz, desync);
    if (!only || only == 1024) run<1024>(out, sms, ghz, desync);
    if (!only || only == 1536) run<1536>(out, sms, ghz, desync);
    if (!only || only == 1792) run<1792>(out, sms, ghz, desync);
    if (!only || only == 2048) run<2048>(out, sms, ghz, desync);
    if (!only || only == 2560) run<2560>(out, sms, ghz, desync);
    if (!only || only == 3072) run<3072>(out, sms, ghz, desync);
    if (!only || only == 4096) run<4096>(out, sms, ghz, desync);
    if (!only || only == 8192) run<8192>(out, sms, ghz, desync);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
