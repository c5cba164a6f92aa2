// CUDA layer of libmcgpu_b200: device memory, uploads, peer reduction; the kernels are in launch.cu.
// The C host (csrc/host/api.c) drives it through the extern "C" functions declared in
// csrc/host/mcgpu_host.h.  Built for sm_100a only, with -fmad=false and no fast-math (see
// transport.cuh for why).
#include "device_internal.h"


// dst[i] += src[i]; src may live on a peer GPU (NVLink load) -- integer sums commute, so the
// result is independent of how the streams were split.
__global__ void accumulate_u64(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// ------------------------------------------------------------------------------------------
extern "C" int mcgpu_dev_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" struct mcgpu_device* mcgpu_dev_open(int ordinal, char* err, size_t errlen) {
  cudaDeviceProp prop;
  if (cudaSetDevice(ordinal) != cudaSuccess || cudaGetDeviceProperties(&prop, ordinal) != cudaSuccess) {
    snprintf(err, errlen, "cannot open CUDA device %d: %s", ordinal, cudaGetErrorString(cudaGetLastError()));
    return NULL;
  }
  if (prop.major != 10) {
    snprintf(err, errlen, "device %d is sm_%d%d; this engine is built for sm_100a (B200) only", ordinal, prop.major, prop.minor);
    return NULL;
  }
  mcgpu_device* d = (mcgpu_device*)calloc(1, sizeof(mcgpu_device));
  if (!d) return NULL;
  d->ordinal = ordinal;
  d->sm_count = prop.multiProcessorCount;
  {  // tuning / A-B switches, read ONCE per device handle; the defaults are the product path.  The arithmetic mode
     // is not among them: only mcgpu_set_fast_math selects it, so mcgpu_get_info always reports what runs.
    const char* k = getenv("MCGPU_KERNEL");
    const char* t = getenv("MCGPU_W_THRESHOLD");
    d->kernel_generation = (k && atoi(k) == 1) ? 1 : (k && atoi(k) == 2) ? 2 : 3;
    d->w_threshold = t ? atoi(t) : (d->kernel_generation == 3 ? 12 : 8);
    d->wf_rows = getenv("MCGPU_WF_ROWS") ? atoi(getenv("MCGPU_WF_ROWS")) : 0;
    if (d->wf_rows != 16 && d->wf_rows != 32) d->wf_rows = 0;
    d->wf_block = (getenv("MCGPU_WF_BLOCK") && atoi(getenv("MCGPU_WF_BLOCK")) == 1024) ? 1024 : 512;
    if (d->w_threshold < 1) d->w_threshold = 1;
    if (d->w_threshold > 32) d->w_threshold = 32;
  }
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&d->ev0) != cudaSuccess || cudaEventCreate(&d->ev1) != cudaSuccess) {
    snprintf(err, errlen, "cannot create stream/events on device %d: %s", ordinal, cudaGetErrorString(cudaGetLastError()));
    free(d);
    return NULL;
  }
  return d;
}

static void free_scene_allocs(mcgpu_device* d) {
  cudaFree(d->d_volume), cudaFree(d->d_palette), cudaFree(d->d_mfp), cudaFree(d->d_woodcock);
  cudaFree(d->d_ray_xpab), cudaFree(d->d_ray_itl_itu), cudaFree(d->d_cmp_shells), cudaFree(d->d_spectrum);
  cudaFree(d->d_image), cudaFree(d->d_peer_stage), cudaFree(d->d_stream_counter);
  cudaFree(d->d_materials_dose), cudaFree(d->d_voxels_edep);
  d->d_materials_dose = NULL, d->d_voxels_edep = NULL;
  d->d_stream_counter = NULL;
  if (d->h_stage) cudaFreeHost(d->h_stage);
  d->d_volume = NULL, d->d_palette = NULL, d->d_mfp = NULL, d->d_woodcock = NULL, d->d_ray_xpab = NULL;
  d->d_ray_itl_itu = NULL, d->d_cmp_shells = NULL, d->d_spectrum = NULL, d->d_image = NULL, d->d_peer_stage = NULL, d->h_stage = NULL;
}

extern "C" void mcgpu_dev_close(struct mcgpu_device* d) {
  if (!d) return;
  cudaSetDevice(d->ordinal);
  cudaFree(d->post_ws);
  d->post_ws = NULL, d->post_ws_bytes = 0;
  cudaStreamSynchronize(d->stream);
  free_scene_allocs(d);
  cudaEventDestroy(d->ev0), cudaEventDestroy(d->ev1);
  cudaStreamDestroy(d->stream);
  free(d);
}

extern "C" int mcgpu_dev_ordinal(const struct mcgpu_device* d) { return d ? d->ordinal : -1; }
extern "C" void* mcgpu_dev_image_ptr(struct mcgpu_device* d) { return d ? (void*)d->d_image : NULL; }

template <class T>
static int upload(T** dst, const void* src, size_t bytes, char* err, size_t errlen) {
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc((void**)dst, bytes));
  if (src) CK(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int mcgpu_dev_upload(struct mcgpu_device* d, const mcgpu_scene* s, const mcgpu_volume* v, const mcgpu_spectrum* spc, int npix_total, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  CK(cudaStreamSynchronize(d->stream));
  free_scene_allocs(d);
  const size_t ns = (size_t)s->num_slots;
  if (upload(&d->d_volume, v->packed, v->packed_bytes, err, errlen)) return -1;
  if (upload(&d->d_palette, s->palette, sizeof(float2) * (size_t)s->palette_size, err, errlen)) return -1;
  if (upload(&d->d_mfp, s->mfp, sizeof(mcgpu_mfp_record) * ns * s->num_values, err, errlen)) return -1;
  if (upload(&d->d_woodcock, s->woodcock, sizeof(float2) * (size_t)s->num_values, err, errlen)) return -1;
  if (upload(&d->d_ray_xpab, s->ray_xpab, sizeof(float4) * ns * MCGPU_NP_RAYLEIGH, err, errlen)) return -1;
  if (upload(&d->d_ray_itl_itu, s->ray_itl_itu, sizeof(uchar2) * ns * MCGPU_NP_RAYLEIGH, err, errlen)) return -1;
  if (upload(&d->d_cmp_shells, s->cmp_shells, sizeof(float4) * ns * MCGPU_MAX_SHELLS, err, errlen)) return -1;
  if (upload(&d->d_spectrum, spc, sizeof(mcgpu_spectrum), err, errlen)) return -1;
  d->image_words = (size_t)4 * npix_total;
  if (upload(&d->d_image, NULL, sizeof(unsigned long long) * d->image_words, err, errlen)) return -1;
  CK(cudaMemset(d->d_image, 0, sizeof(unsigned long long) * d->image_words));
  CK(cudaMalloc((void**)&d->d_stream_counter, 2 * sizeof(unsigned long long)));
  CK(cudaMemset(d->d_stream_counter, 0, 2 * sizeof(unsigned long long)));
  d->dose_roi_voxels = 0;
  if (s->tally_material_dose) {
    CK(cudaMalloc((void**)&d->d_materials_dose, sizeof(unsigned long long) * 2 * MCGPU_MAX_MATERIALS));
    CK(cudaMemset(d->d_materials_dose, 0, sizeof(unsigned long long) * 2 * MCGPU_MAX_MATERIALS));
  }
  if (s->tally_voxel_dose) {
    d->dose_roi_voxels = s->dose_roi_voxels;
    CK(cudaMalloc((void**)&d->d_voxels_edep, sizeof(unsigned long long) * 2 * (size_t)s->dose_roi_voxels));
    CK(cudaMemset(d->d_voxels_edep, 0, sizeof(unsigned long long) * 2 * (size_t)s->dose_roi_voxels));
  }
  McgpuSceneDev& sc = d->scene;
  memset(&sc, 0, sizeof sc);
  sc.volume = d->d_volume;
  sc.palette = d->d_palette;
  sc.mfp = d->d_mfp;
  sc.woodcock = d->d_woodcock;
  sc.ray_xpab = d->d_ray_xpab;
  sc.ray_itl_itu = d->d_ray_itl_itu;
  sc.cmp_shells = d->d_cmp_shells;
  sc.spectrum = d->d_spectrum;
  sc.image = d->d_image;
  for (int k = 0; k < MCGPU_MAX_MATERIALS; k++) sc.cmp_noscco[k] = s->cmp_noscco[k];
  sc.num_slots = s->num_slots;
  sc.palette_size = s->palette_size;
  sc.num_values = s->num_values;
  sc.max_shells = 1;
  for (int k = 0; k < s->num_slots; k++)
    if (s->cmp_noscco[k] > sc.max_shells) sc.max_shells = s->cmp_noscco[k];
  sc.nvx = v->nx, sc.nvy = v->ny, sc.nvz = v->nz;
  for (int k = 0; k < 3; k++) {
    sc.inv_voxel[k] = v->inv_voxel_size[k];
    sc.bbox[k] = v->size_bbox[k];
  }
  sc.e0 = s->e0;
  sc.ide = s->ide;
  sc.materials_dose = d->d_materials_dose;
  sc.voxels_edep = d->d_voxels_edep;
  for (int k = 0; k < 6; k++) sc.dose_roi[k] = s->dose_roi[k];
  for (int k = 0; k < MCGPU_MAX_MATERIALS; k++) sc.material_of_slot[k] = s->material_of_slot[k];
  d->voxel_bits = s->voxel_bits;
  return 0;
}

extern "C" int mcgpu_dev_launch(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen) {
  if ((size_t)4 * (size_t)view->total_num_pixels != d->image_words) {  // a pose of another detector would tally past the end of d_image
    snprintf(err, errlen, "device %d: the projection has %d pixels but the uploaded image holds %zu; load the materials again after a new input", d->ordinal,
             view->total_num_pixels, d->image_words / 4);
    return -1;
  }
  return d->fast_math ? mcgpu_launch_fast(d, view, l, err, errlen) : mcgpu_launch_exact(d, view, l, err, errlen);
}

extern "C" void mcgpu_dev_set_fast_math(struct mcgpu_device* d, int on) {
  if (d) d->fast_math = on != 0;
}

extern "C" int mcgpu_dev_sync(struct mcgpu_device* d, float* kernel_ms, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  CK(cudaStreamSynchronize(d->stream));
  if (kernel_ms) {
    *kernel_ms = 0.f;
    if (d->timed) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
  }
  if (d->kernel_generation == 3 && d->d_stream_counter) {  // the wavefront kernel reports a lost context instead of hanging
    unsigned long long flag = 0;
    CK(cudaMemcpy(&flag, d->d_stream_counter + 1, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) {
      snprintf(err, errlen, "device %d: wavefront transport kernel watchdog fired (code %llu)", d->ordinal, flag & 0xffffffffull);
      return -1;
    }
  }
  return 0;
}

extern "C" int mcgpu_dev_fetch(struct mcgpu_device* d, uint64_t* host, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  CK(cudaMemcpyAsync(host, d->d_image, sizeof(unsigned long long) * d->image_words, cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}

extern "C" int mcgpu_dev_accumulate_peer(struct mcgpu_device* dst, struct mcgpu_device* src, char* err, size_t errlen) {
  int can = 0;
  CK(cudaSetDevice(dst->ordinal));
  CK(cudaDeviceCanAccessPeer(&can, dst->ordinal, src->ordinal));
  const size_t n = dst->image_words;
  const unsigned long long* from = src->d_image;
  if (can) {
    cudaError_t e = cudaDeviceEnablePeerAccess(src->ordinal, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
    cudaGetLastError();
  }
  if (!can) {  // no NVLink/P2P path: stage through a device-to-device copy
    if (!dst->d_peer_stage) CK(cudaMalloc((void**)&dst->d_peer_stage, sizeof(unsigned long long) * n));
    CK(cudaMemcpyPeerAsync(dst->d_peer_stage, dst->ordinal, src->d_image, src->ordinal, sizeof(unsigned long long) * n, dst->stream));
    from = dst->d_peer_stage;
  }
  accumulate_u64<<<dst->sm_count * 4, 256, 0, dst->stream>>>(dst->d_image, from, n);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(dst->stream));
  return 0;
}

extern "C" int mcgpu_dev_reset_dose(struct mcgpu_device* d, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  if (d->d_materials_dose) CK(cudaMemsetAsync(d->d_materials_dose, 0, sizeof(unsigned long long) * 2 * MCGPU_MAX_MATERIALS, d->stream));
  if (d->d_voxels_edep) CK(cudaMemsetAsync(d->d_voxels_edep, 0, sizeof(unsigned long long) * 2 * (size_t)d->dose_roi_voxels, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}

extern "C" int mcgpu_dev_add_dose(struct mcgpu_device* d, uint64_t* materials, uint64_t* voxels, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  CK(cudaStreamSynchronize(d->stream));
  if (materials && d->d_materials_dose) {
    unsigned long long tmp[2 * MCGPU_MAX_MATERIALS];
    CK(cudaMemcpy(tmp, d->d_materials_dose, sizeof tmp, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 2 * MCGPU_MAX_MATERIALS; k++) materials[k] += tmp[k];
  }
  if (voxels && d->d_voxels_edep) {
    const size_t n = 2 * (size_t)d->dose_roi_voxels;
    unsigned long long* tmp = (unsigned long long*)malloc(sizeof(unsigned long long) * n);
    if (!tmp) {
      snprintf(err, errlen, "out of memory fetching the voxel dose");
      return -1;
    }
    cudaError_t e = cudaMemcpy(tmp, d->d_voxels_edep, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
      for (size_t k = 0; k < n; k++) voxels[k] += tmp[k];
    free(tmp);
    CK(e);
  }
  return 0;
}
