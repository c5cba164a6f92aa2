// Random-gather rate of the B200 L2 for the voxel fetch of the Woodcock loop: every lane reads ONE byte at a random
// offset of an L2-resident array (62.5 MB = Catphan 500^3 at 4 bits per voxel; 31 MB; 125 MB > one L2 partition set),
// each access a different 32-byte sector, `ILP` independent loads in flight per lane, the next address depending on
// the loaded byte only when DEP=1 (the transport kernel's chain is dependent per photon but 32 warps/SM overlap).
// Reports sectors/s and GB/s of sector traffic: the denominator for the kernel's measured lts sector rate.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2gather tools/ubench/l2gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
  return x;
}

template <int ILP, int DEP>
__global__ void __launch_bounds__(512, 2) gather(const unsigned char* __restrict__ vol, unsigned n_bytes, int iters, unsigned* out) {
  unsigned s[ILP], acc = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s[k] = mix((blockIdx.x * blockDim.x + threadIdx.x) * ILP + k + 1u);
  for (int i = 0; i < iters; i++) {
    unsigned v[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = __ldg(vol + (s[k] % n_bytes));
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      acc += v[k];
      s[k] = mix(s[k] + 0x9e3779b9u + (DEP ? v[k] : 0u));
    }
  }
  if (acc == 0xffffffffu) out[0] = acc;
}

template <int ILP, int DEP>
static void run(const unsigned char* vol, unsigned n_bytes, int sms, unsigned* out) {
  const int iters = 2048 / ILP * 4, grid = sms * 2, block = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  gather<ILP, DEP><<<grid, block>>>(vol, n_bytes, 64, out);
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    gather<ILP, DEP><<<grid, block>>>(vol, n_bytes, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double n = (double)grid * block * iters * ILP;
  printf("{\"array_MB\": %.1f, \"loads_in_flight_per_lane\": %d, \"dependent\": %d, \"gathers\": %.0f, \"ms\": %.3f, \"sectors_per_s\": %.4g, \"sector_GB_per_s\": %.1f}\n",
         n_bytes / 1e6, ILP, DEP, n, best, n / (best * 1e-3), 32.0 * n / (best * 1e-3) / 1e9);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned* out;
  cudaMalloc(&out, 64);
  for (unsigned mb : {31u, 62u, 125u, 445u}) {
    const unsigned n_bytes = mb * 1000000u + 500000u;
    unsigned char* vol;
    cudaMalloc(&vol, n_bytes);
    cudaMemset(vol, 1, n_bytes);
    run<1, 1>(vol, n_bytes, sms, out);
    run<2, 1>(vol, n_bytes, sms, out);
    run<4, 0>(vol, n_bytes, sms, out);
    run<8, 0>(vol, n_bytes, sms, out);
    cudaFree(vol);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
