// Kernels and their launch.  This file is compiled TWICE into libmcgpu_b200.so:
//   launch_exact.o : -fmad=false, no fast-math  -> namespace mcgpu,      mcgpu_launch_exact  (default; bit-exact tallies)
//   launch_fast.o  : -use_fast_math             -> namespace mcgpu_fast, mcgpu_launch_fast   (opt-in; the flags the reference ships with,
//                                                   docker/compile.sh:36; statistically equivalent results)
#include "device_internal.h"

#ifdef MCGPU_FAST_MATH
#define MCGPU_LAUNCH_NAME mcgpu_launch_fast
#else
#define MCGPU_LAUNCH_NAME mcgpu_launch_exact
#endif

using namespace MCGPU_NS;

static int pow_mod_host(long long a, unsigned long long n, long long m) {
  long long y = 1, z = a % m;
  while (n) {
    if (n & 1ull) y = (y * z) % m;
    z = (z * z) % m;
    n >>= 1;
  }
  return (int)y;
}

namespace {
struct LaunchArgs {
  mcgpu_device* d;
  const mcgpu_view* view;
  const mcgpu_launch* l;
  long long n_streams;
  int g1, g2;
  SceneDev scene;  // the device's scene with the image pointer of the launch's slot
};

// The product kernel: persistent grid, one or two CTAs per SM, shared memory split between the Compton scratch and the photon pool.
template <int BITS, bool DOSE, int ROT>
int launch_wavefront(const LaunchArgs& a, char* err, size_t errlen) {
  mcgpu_device* d = a.d;
  const int pal = (BITS == 4 || BITS == 8) ? d->scene.palette_size : 0;
  const int wblock = d->wf_block;
  int smem_sm = 0, per_sm = MCGPU_WF_MAX_BLOCK / wblock, pool = 0;
  CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d->ordinal));
  int rows = d->wf_rows;
  if (rows == 0) {  // 32 scratch rows when the full pool still fits next to them
    const long long b32 = (long long)smem_sm / per_sm - 1024 - (long long)wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, 0, wblock / 32, 32).total;
    rows = b32 >= (long long)(sizeof(float) * MCGPU_WF_STRIDE) * (2 * wblock) ? 32 : 16;
  }
  const size_t fixed = wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, 0, wblock / 32, rows).total;
  for (; per_sm >= 1; per_sm--) {
    const long long budget = (long long)smem_sm / per_sm - 1024 - (long long)fixed;
    pool = budget > 0 ? (int)(budget / (long long)(sizeof(float) * MCGPU_WF_STRIDE)) & ~31 : 0;
    if (pool > 2 * wblock) pool = 2 * wblock;
    if (pool > MCGPU_WF_MAX_POOL) pool = MCGPU_WF_MAX_POOL;
    if (pool >= wblock) break;
  }
  if (pool < 64) {
    snprintf(err, errlen, "device %d: not enough shared memory for the wavefront kernel", d->ordinal);
    return -1;
  }
  const size_t wsmem = wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, pool, wblock / 32, rows).total;
  CK(cudaFuncSetAttribute(transport_wavefront<BITS, DOSE, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transport_wavefront<BITS, DOSE, ROT>, wblock, wsmem));
  long long pgrid = (long long)d->sm_count * (per_sm > 0 ? per_sm : 1);
  const long long useful = (a.n_streams + pool - 1) / pool;
  if (pgrid > useful) pgrid = useful;
  CK(cudaMemsetAsync(d->d_stream_counter, 0, sizeof(unsigned long long), d->stream));  // [1], the error flag, is sticky until read
  transport_wavefront<BITS, DOSE, ROT><<<(unsigned)pgrid, wblock, wsmem, d->stream>>>(
      a.scene, *a.view, a.l->stream_begin, a.l->stream_end, a.l->histories_per_thread, a.l->seed_input, a.g1, a.g2, d->d_stream_counter, d->w_threshold, pool, pal, rows,
      reinterpret_cast<int*>(d->d_stream_counter + 1));
  return 0;
}

#ifdef MCGPU_AB_KERNELS
template <int BITS, bool DOSE, int ROT>
int launch_regroup(const LaunchArgs& a, size_t smem, char* err, size_t errlen) {
  mcgpu_device* d = a.d;
  const int block = MCGPU_REGROUP_BLOCK;
  int per_sm = 0;
  smem += sizeof(float) * (MCGPU_REGROUP_BLOCK / 32) * MCGPU_SCRATCH_ROWS * regroup_scratch_stride(d->scene.max_shells) + 8;
  CK(cudaFuncSetAttribute(transport_regroup<BITS, DOSE, ROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transport_regroup<BITS, DOSE, ROT>, block, smem));
  long long pgrid = (long long)d->sm_count * (per_sm > 0 ? per_sm : 1);
  const long long grid = (a.n_streams + block - 1) / block;
  if (pgrid > grid) pgrid = grid;
  CK(cudaMemsetAsync(d->d_stream_counter, 0, sizeof(unsigned long long), d->stream));
  transport_regroup<BITS, DOSE, ROT><<<(unsigned)pgrid, block, smem, d->stream>>>(a.scene, *a.view, a.l->stream_begin, a.l->stream_end, a.l->histories_per_thread,
                                                                                   a.l->seed_input, a.g1, a.g2, d->d_stream_counter, d->w_threshold);
  return 0;
}
#endif

template <int BITS>
int launch_bits(const LaunchArgs& a, char* err, size_t errlen) {
  mcgpu_device* d = a.d;
  const bool dose = d->scene.materials_dose || d->scene.voxels_edep;
  if (d->kernel_generation == 3) {
    if (dose) return launch_wavefront<BITS, true, -1>(a, err, errlen);
    return a.view->rotation_flag == 1 ? launch_wavefront<BITS, false, 1>(a, err, errlen) : launch_wavefront<BITS, false, 0>(a, err, errlen);
  }
#ifdef MCGPU_AB_KERNELS
  size_t smem = ((sizeof(SharedTables) + 15) & ~size_t(15)) + sizeof(float4) * d->scene.num_slots * MCGPU_MAX_SHELLS;
  if (BITS == 4 || BITS == 8) smem += sizeof(float2) * d->scene.palette_size;
  if (d->kernel_generation == 2) {
    if (dose) return launch_regroup<BITS, true, -1>(a, smem, err, errlen);
    return a.view->rotation_flag == 1 ? launch_regroup<BITS, false, 1>(a, smem, err, errlen) : launch_regroup<BITS, false, 0>(a, smem, err, errlen);
  }
  if (dose) {  // the reference-structured kernel has no dose path: refuse instead of silently dropping the tallies
    snprintf(err, errlen, "device %d: MCGPU_KERNEL=1 does not implement the dose tallies", d->ordinal);
    return -1;
  }
  const int block = 128;
  const long long grid = (a.n_streams + block - 1) / block;
  CK(cudaFuncSetAttribute(transport_streams<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  transport_streams<BITS><<<(unsigned)grid, block, smem, d->stream>>>(a.scene, *a.view, a.l->stream_begin, a.l->stream_end, a.l->histories_per_thread, a.l->seed_input, a.g1,
                                                                      a.g2);
  return 0;
#else
  snprintf(err, errlen, "device %d: MCGPU_KERNEL=%d asks for an A/B kernel generation this library was not built with (make AB=1)", d->ordinal, d->kernel_generation);
  return -1;
#endif
}
}  // namespace

extern "C" int MCGPU_LAUNCH_NAME(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  if (!d->d_image) {
    snprintf(err, errlen, "device %d: nothing uploaded", d->ordinal);
    return -1;
  }
  unsigned long long* image = l->image_slot ? d->d_image_alt : d->d_image;
  if (!image) {
    snprintf(err, errlen, "device %d: image slot %d is not allocated", d->ordinal, l->image_slot);
    return -1;
  }
  if (l->zero_image) CK(cudaMemsetAsync(image, 0, sizeof(unsigned long long) * d->image_words, d->stream));
  LaunchArgs a;
  a.d = d, a.view = view, a.l = l;
  a.scene = d->scene;
  a.scene.image = image;
  a.n_streams = l->stream_end - l->stream_begin;
  CK(cudaEventRecord(d->ev0, d->stream));
  if (a.n_streams > 0) {
    const unsigned long long leap = (unsigned long long)(l->histories_per_thread * 256);
    a.g1 = pow_mod_host(40014, leap, 2147483563LL), a.g2 = pow_mod_host(40692, leap, 2147483399LL);
    int rc;
    switch (d->voxel_bits) {
      case 4: rc = launch_bits<4>(a, err, errlen); break;
      case 8: rc = launch_bits<8>(a, err, errlen); break;
      case 16: rc = launch_bits<16>(a, err, errlen); break;
      default: rc = launch_bits<64>(a, err, errlen); break;
    }
    if (rc != 0) return rc;
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(d->ev1, d->stream));
  d->timed = 1;
  return 0;
}

#ifndef MCGPU_FAST_MATH
// ---- device self tests of the arithmetic shortcuts of transport.cuh (exhaustive over their whole domain) ------------------
namespace {
__global__ void selftest_log_uniform(unsigned long long* mismatches) {  // every value RANECU can return: i2 in [1, 2147483562]
  unsigned long long bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; i <= 2147483562ll; i += (long long)gridDim.x * blockDim.x) {
    const float a = __int2float_rn((int)i) * 4.65661305739e-10f;
    bad += __float_as_uint(log_uniform(a)) != __float_as_uint(logf(a));
  }
  if (bad) atomicAdd(mismatches, bad);
}
__global__ void selftest_rsqrt_normal(unsigned long long* mismatches) {  // every positive normal float
  unsigned long long bad = 0;
  for (unsigned long long b = 0x00800000ull + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= 0x7f7fffffull; b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)b);
    bad += __float_as_uint(rsqrt_normal(x)) != __float_as_uint(rsqrtf(x));
  }
  if (bad) atomicAdd(mismatches, bad);
}
__global__ void selftest_outside_box(const SceneDev sc, unsigned long long* mismatches) {  // every non-NaN float on each axis
  unsigned long long bad = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= 0xffffffffull; b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)b);
    if (x != x) continue;
    for (int k = 0; k < 3; k++) {
      Photon p;
      p.x = p.y = p.z = 0.5f * fminf(sc.bbox[0], fminf(sc.bbox[1], sc.bbox[2]));  // the other two axes inside
      (k == 0 ? p.x : k == 1 ? p.y : p.z) = x;
      bad += outside_box(sc, p) != (locate_voxel(sc, p) < 0);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

extern "C" int mcgpu_dev_selftest(struct mcgpu_device* d, const char* name, unsigned long long* mismatches, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  unsigned long long* counter = NULL;
  CK(cudaMalloc((void**)&counter, sizeof *counter));
  CK(cudaMemset(counter, 0, sizeof *counter));
  const int grid = d->sm_count * 8, block = 256;
  if (!strcmp(name, "log_uniform"))
    selftest_log_uniform<<<grid, block, 0, d->stream>>>(counter);
  else if (!strcmp(name, "rsqrt_normal"))
    selftest_rsqrt_normal<<<grid, block, 0, d->stream>>>(counter);
  else if (!strcmp(name, "outside_box")) {
    if (!d->d_image) {
      cudaFree(counter);
      snprintf(err, errlen, "selftest outside_box needs a loaded geometry");
      return -1;
    }
    selftest_outside_box<<<grid, block, 0, d->stream>>>(d->scene, counter);
  } else {
    cudaFree(counter);
    snprintf(err, errlen, "unknown self test '%s' (log_uniform, rsqrt_normal, outside_box)", name);
    return -1;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
  if (e == cudaSuccess) e = cudaMemcpy(mismatches, counter, sizeof *counter, cudaMemcpyDeviceToHost);
  cudaFree(counter);
  CK(e);
  return 0;
}
#endif
