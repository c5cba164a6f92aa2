// Internals shared by device.cu (device management) and launch.cu (kernels + launch, built twice:
// exact arithmetic and fast-math).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "transport.cuh"
#include "wavefront.cuh"
#ifdef MCGPU_AB_KERNELS  // generations 1 and 2, for A/B measurements only (make AB=1)
#include "regroup.cuh"
#include "streams.cuh"
#endif

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      snprintf(err, errlen, "CUDA failure %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, #call); \
      return -1;                                                                                    \
    }                                                                                               \
  } while (0)

struct mcgpu_device {
  int ordinal;
  int sm_count;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  McgpuSceneDev scene;
  int voxel_bits;
  size_t image_words;
  // owned allocations
  void* d_volume;
  float2* d_palette;
  mcgpu_mfp_record* d_mfp;
  float2* d_woodcock;
  float4* d_ray_xpab;
  uchar2* d_ray_itl_itu;
  float4* d_cmp_shells;
  mcgpu_spectrum* d_spectrum;
  unsigned long long* d_image;
  unsigned long long* d_peer_stage;  // used when peer access is unavailable
  unsigned long long* d_materials_dose;  // [25][2] or NULL
  unsigned long long* d_voxels_edep;     // [roi][2] or NULL
  long long dose_roi_voxels;
  unsigned long long* d_stream_counter;  // [0] next stream of the running launch (persistent kernels), [1] kernel error flag (wavefront)
  int kernel_generation;                 // 3 = block wavefront with work queues (default), 2 = regrouping persistent warps, 1 = one thread per stream (reference structure); 1 and 2 for A/B
  int w_threshold;
  int wf_rows;                           // wavefront kernel: scratch rows per warp (16, 32; 0 = choose by shared-memory budget)
  int wf_block;                          // wavefront kernel: threads per CTA (512: two CTAs per SM, 1024: one)
  int fast_math;                         // 0 = bit-exact arithmetic (default), 1 = the reference's shipped -use_fast_math flags
  uint64_t* h_stage;
  // pipelined scan (mcgpu_dev_pipeline_*): a second image and a copy stream, so that the device->host copy of one
  // projection overlaps the transport of the next; slot 0 is d_image
  unsigned long long* d_image_alt;
  uint64_t* h_pinned[2];
  cudaStream_t copy_stream;
  cudaEvent_t p_ev0[2], p_ev1[2], p_copied[2];
  int* h_flag;  // pinned copy of the kernel error flag
  int pipeline_on;
  void* post_ws;                         // workspace of the post-processing kernels (postprocess.cu)
  size_t post_ws_bytes;
  int timed;
};


extern "C" int mcgpu_launch_exact(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen);
extern "C" int mcgpu_launch_fast(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen);
