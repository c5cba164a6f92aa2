// u64 atomic-add throughput of B200 (sm_100a) for the detector tally (K:545-548: atomicAdd on a u64 image, here
// fire-and-forget ATOMG/RED to the L2-resident 45 MB image).  Patterns:
//   random   every lane adds to a uniformly random pixel of the 4 x 1848 x 768 image (scattered photons)
//   primary  every lane adds to a random pixel of ONE plane inside a 1024-column band (the half-fan primary footprint;
//            the air scan is 100 % this)
//   hot      all lanes of a warp hit the same 32-pixel row segment (worst-case contention on neighbouring addresses)
// Reports atomics/s; the air scan's tally rate is compared against `primary`.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomics tools/ubench/atomics.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mix(unsigned x) {
  x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
  return x;
}

template <int PATTERN>
__global__ void __launch_bounds__(512, 2) tally(unsigned long long* __restrict__ image, int nx, int nz, int per_thread) {
  unsigned s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1u);
  const unsigned npix = (unsigned)(nx * nz);
  for (int i = 0; i < per_thread; i++) {
    s = mix(s + 0x9e3779b9u);
    size_t idx;
    if (PATTERN == 0) {
      idx = (size_t)(s % (4u * npix));
    } else if (PATTERN == 1) {
      const unsigned ix = 412u + (s % 1024u), iz = (s >> 12) % (unsigned)nz;
      idx = (size_t)ix + (size_t)iz * nx;
    } else {
      const unsigned w = mix((blockIdx.x * blockDim.x + threadIdx.x) >> 5) + i;
      idx = (size_t)((w % (npix / 32u)) * 32u + (threadIdx.x & 31u));
    }
    atomicAdd(image + idx, (unsigned long long)(5000000u + (s & 0xffffu)));
  }
}

template <int PATTERN>
static void run(const char* name, unsigned long long* image, int sms, int nx, int nz) {
  const int per_thread = 4096, grid = sms * 2, block = 512;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  tally<PATTERN><<<grid, block>>>(image, nx, nz, 64);
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    tally<PATTERN><<<grid, block>>>(image, nx, nz, per_thread);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double n = (double)grid * block * per_thread;
  printf("{\"pattern\": \"%s\", \"atomics\": %.0f, \"ms\": %.3f, \"atomics_per_s\": %.4g, \"GB_per_s_8B\": %.1f}\n", name, n, best, n / (best * 1e-3), 8.0 * n / (best * 1e-3) / 1e9);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int nx = 1848, nz = 768;
  unsigned long long* image;
  cudaMalloc(&image, sizeof(unsigned long long) * 4 * nx * nz);
  cudaMemset(image, 0, sizeof(unsigned long long) * 4 * nx * nz);
  run<0>("random_4_planes", image, sms, nx, nz);
  run<1>("primary_band_1_plane", image, sms, nx, nz);
  run<2>("warp_row_segment", image, sms, nx, nz);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
