// CUDA layer of libmcgpu_b200: device memory, uploads, peer reduction; the kernels are in launch.cu.
// The C host (csrc/host/api.c) drives it through the extern "C" functions declared in
// csrc/host/mcgpu_host.h.  Built for sm_100a only, with -fmad=false and no fast-math (see
// transport.cuh for why).
#include "device_internal.h"


// History-split reduction (the reference's MPI_Reduce of the partial images, MC-GPU_v1.3.cu:1019): the u64 tallies of the
// devices that shared one projection's streams are summed on the first device.  Integer sums commute, so the result
// does not depend on how the streams were split.  Two implementations behind mcgpu_dev_reduce:
//   * ncclReduce(ncclUint64, ncclSum) over NVLink/NVSwitch with one communicator per device (ncclCommInitAll), all
//     ranks driven from this process inside one ncclGroup -- the default when libnccl can be loaded;
//   * accumulate_peers_u64: ONE kernel on the first device that reads every peer's image through NVLink peer
//     mappings and adds them in a single pass over the destination (instead of n-1 passes) -- when NCCL is not
//     available or MCGPU_REDUCE=peer.
#define MCGPU_MAX_PEERS 16
struct PeerImages {
  const unsigned long long* src[MCGPU_MAX_PEERS];
  int n;
};

__global__ void __launch_bounds__(256) accumulate_peers_u64(unsigned long long* __restrict__ dst, const PeerImages peers, size_t n2) {
  // two tallies per thread and iteration: 16-byte loads over NVLink, one read-modify-write of the local image
  ulonglong2* d2 = reinterpret_cast<ulonglong2*>(dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    ulonglong2 acc = d2[i];
#pragma unroll 1
    for (int k = 0; k < peers.n; k++) {
      const ulonglong2 v = reinterpret_cast<const ulonglong2*>(peers.src[k])[i];
      acc.x += v.x, acc.y += v.y;
    }
    d2[i] = acc;
  }
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
    d->w_threshold = t ? atoi(t) : (d->kernel_generation == 3 ? 16 : 8);  // r02e: 16 is 0.3-0.7 % ahead of 12 on every workload
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
  mcgpu_dev_pipeline_end(d);
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
  CK(cudaMalloc((void**)&d->d_stream_counter, 16 * sizeof(unsigned long long)));  // [2..12]: MCGPU_WF_STATS diagnostics
  CK(cudaMemset(d->d_stream_counter, 0, 16 * sizeof(unsigned long long)));
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
    {  // transport.cuh: locate_voxel_fast.  A box thinner than 2 EPS (0.3 um) contains nothing.
      const float eps = 0.000015f, top = v->size_bbox[k] - eps;
      unsigned be, bt;
      memcpy(&be, &eps, 4), memcpy(&bt, &top, 4);
      sc.box_hi[k] = ((bt > be && bt < 0x7f800000u) ? bt : be) - be;
    }
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
#ifdef MCGPU_WF_STATS
  if (d->d_stream_counter) {
    unsigned long long s[11];
    CK(cudaMemcpy(s, d->d_stream_counter + 2, sizeof s, cudaMemcpyDeviceToHost));
    CK(cudaMemset(d->d_stream_counter + 2, 0, sizeof s));
    const char* qn[4] = {"W", "N", "C", "R"};
    for (int q = 0; q < 4; q++) fprintf(stderr, "wf_stats %s: %llu batches, %.2f contexts per batch\n", qn[q], s[q], s[q] ? (double)s[4 + q] / s[q] : 0.0);
    fprintf(stderr, "wf_stats tracking: %.2f steps per batch, %.2f lanes per step; idle polls %llu\n", s[0] ? (double)s[8] / s[0] : 0.0, s[8] ? (double)s[9] / s[8] : 0.0, s[10]);
  }
#endif
  if (d->kernel_generation == 3 && d->d_stream_counter) {  // the wavefront kernel reports a lost context instead of hanging
    unsigned long long flag = 0;
    CK(cudaMemcpy(&flag, d->d_stream_counter + 1, sizeof flag, cudaMemcpyDeviceToHost));
    if (flag) {
      CK(cudaMemset(d->d_stream_counter + 1, 0, sizeof flag));
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

// ---- reducer -------------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: libnccl is opened at run time, the library has no link-time dependency on it

struct NcclApi {
  void* handle;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
};

static const NcclApi* nccl_api(void) {
  static NcclApi api;
  static int state = 0;  // 0 untried, 1 loaded, -1 unavailable
  if (state == 0) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (size_t i = 0; i < sizeof names / sizeof names[0] && !api.handle; i++) api.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    state = -1;
    if (api.handle) {
      api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.handle, "ncclCommInitAll");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
      api.Reduce = (decltype(api.Reduce))dlsym(api.handle, "ncclReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
      if (api.CommInitAll && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Reduce && api.GetErrorString) state = 1;
    }
  }
  return state == 1 ? &api : NULL;
}

struct mcgpu_reducer {
  int n;
  int mode;  // 1 = ncclReduce, 2 = one-pass peer kernel, 3 = staged copies (no peer access)
  struct mcgpu_device* dev[MCGPU_MAX_PEERS];
  ncclComm_t comm[MCGPU_MAX_PEERS];
  cudaEvent_t ready[MCGPU_MAX_PEERS];  // peer kernel: the source images are complete
  cudaEvent_t t0, t1;
};

extern "C" void mcgpu_dev_reducer_free(struct mcgpu_reducer* r) {
  if (!r) return;
  const NcclApi* api = nccl_api();
  for (int k = 0; k < r->n; k++) {
    cudaSetDevice(r->dev[k]->ordinal);
    if (r->mode == 1 && api && r->comm[k]) api->CommDestroy(r->comm[k]);
    if (r->ready[k]) cudaEventDestroy(r->ready[k]);
  }
  if (r->n > 0) {
    cudaSetDevice(r->dev[0]->ordinal);
    if (r->t0) cudaEventDestroy(r->t0);
    if (r->t1) cudaEventDestroy(r->t1);
  }
  free(r);
}

extern "C" const char* mcgpu_dev_reducer_kind(const struct mcgpu_reducer* r) {
  return !r ? "none" : r->mode == 1 ? "ncclReduce" : r->mode == 2 ? "peer-kernel" : "staged-copy";
}

// Set up the reduction of the images of devs[0..n) onto devs[0].
extern "C" struct mcgpu_reducer* mcgpu_dev_reducer_create(struct mcgpu_device** devs, int n, char* err, size_t errlen) {
  if (n < 2 || n > MCGPU_MAX_PEERS) {
    snprintf(err, errlen, "reducer: %d devices (2..%d supported)", n, MCGPU_MAX_PEERS);
    return NULL;
  }
  mcgpu_reducer* r = (mcgpu_reducer*)calloc(1, sizeof *r);
  if (!r) return NULL;
  r->n = n;
  for (int k = 0; k < n; k++) r->dev[k] = devs[k];
  // MCGPU_REDUCE = "peer" | "nccl".  Default: the peer kernel when every device can map device 0's peers (NVLink/NVSwitch boxes),
  // NCCL otherwise.  Both move the same bytes at the same speed (0.42 ms for 7 x 45 MB on 8 x B200), but ncclCommInitAll costs
  // ~2.5 s per process on 8 GPUs, which the air scan -- ONE projection per process, run before every run-mc -- would pay in full.
  const char* want = getenv("MCGPU_REDUCE");
  bool all_peers = true;
  for (int k = 1; k < n; k++) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devs[0]->ordinal, devs[k]->ordinal) != cudaSuccess || !can) all_peers = false;
  }
  cudaGetLastError();
  const bool try_nccl = want ? !strcmp(want, "nccl") : !all_peers;
  const NcclApi* api = try_nccl ? nccl_api() : NULL;
  if (api) {
    int ids[MCGPU_MAX_PEERS];
    for (int k = 0; k < n; k++) ids[k] = devs[k]->ordinal;
    const ncclResult_t rc = api->CommInitAll(r->comm, n, ids);
    if (rc == ncclSuccess)
      r->mode = 1;
    else if (want && !strcmp(want, "nccl")) {
      snprintf(err, errlen, "reducer: ncclCommInitAll failed: %s", api->GetErrorString(rc));
      free(r);
      return NULL;
    }
  } else if (want && !strcmp(want, "nccl")) {
    snprintf(err, errlen, "reducer: MCGPU_REDUCE=nccl but libnccl.so.2 cannot be loaded");
    free(r);
    return NULL;
  }
  if (r->mode == 0) {  // peer mappings for the one-pass kernel
    r->mode = 2;
    cudaSetDevice(devs[0]->ordinal);
    for (int k = 1; k < n; k++) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devs[0]->ordinal, devs[k]->ordinal);
      if (can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(devs[k]->ordinal, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
        cudaGetLastError();
      }
      if (!can) r->mode = 3;
    }
  }
  for (int k = 0; k < n; k++) {
    cudaSetDevice(devs[k]->ordinal);
    if (cudaEventCreateWithFlags(&r->ready[k], cudaEventDisableTiming) != cudaSuccess) {
      snprintf(err, errlen, "reducer: cannot create events");
      mcgpu_dev_reducer_free(r);
      return NULL;
    }
  }
  cudaSetDevice(devs[0]->ordinal);
  if (cudaEventCreate(&r->t0) != cudaSuccess || cudaEventCreate(&r->t1) != cudaSuccess) {
    snprintf(err, errlen, "reducer: cannot create events");
    mcgpu_dev_reducer_free(r);
    return NULL;
  }
  return r;
}

// Sum the images of all devices of the reducer onto the first one.  Enqueued behind whatever is in the devices' streams
// (the transport kernels); returns after the sum is complete, *reduce_ms = device time of the reduction on device 0.
extern "C" int mcgpu_dev_reduce(struct mcgpu_reducer* r, float* reduce_ms, char* err, size_t errlen) {
  mcgpu_device* root = r->dev[0];
  const size_t n = root->image_words;
  for (int k = 1; k < r->n; k++)
    if (r->dev[k]->image_words != n) {
      snprintf(err, errlen, "reducer: device images differ in size");
      return -1;
    }
  if (r->mode == 1) {
    const NcclApi* api = nccl_api();
    // the root's clock starts when ITS transport kernel is done; the collective itself waits for the slowest device
    CK(cudaSetDevice(root->ordinal));
    CK(cudaEventRecord(r->t0, root->stream));
    ncclResult_t rc = api->GroupStart();
    for (int k = 0; k < r->n && rc == ncclSuccess; k++)
      rc = api->Reduce(r->dev[k]->d_image, r->dev[k]->d_image, n, ncclUint64, ncclSum, 0, r->comm[k], r->dev[k]->stream);
    const ncclResult_t rc2 = api->GroupEnd();
    if (rc != ncclSuccess || rc2 != ncclSuccess) {
      snprintf(err, errlen, "reducer: ncclReduce failed: %s", api->GetErrorString(rc != ncclSuccess ? rc : rc2));
      return -1;
    }
    CK(cudaSetDevice(root->ordinal));
    CK(cudaEventRecord(r->t1, root->stream));
    for (int k = 0; k < r->n; k++) {
      CK(cudaSetDevice(r->dev[k]->ordinal));
      CK(cudaStreamSynchronize(r->dev[k]->stream));
    }
  } else {
    for (int k = 1; k < r->n; k++) {  // the root's stream waits until every peer image is complete
      CK(cudaSetDevice(r->dev[k]->ordinal));
      CK(cudaEventRecord(r->ready[k], r->dev[k]->stream));
    }
    CK(cudaSetDevice(root->ordinal));
    for (int k = 1; k < r->n; k++) CK(cudaStreamWaitEvent(root->stream, r->ready[k], 0));
    CK(cudaEventRecord(r->t0, root->stream));
    if (r->mode == 2) {
      PeerImages peers;
      peers.n = r->n - 1;
      for (int k = 1; k < r->n; k++) peers.src[k - 1] = r->dev[k]->d_image;
      accumulate_peers_u64<<<root->sm_count * 8, 256, 0, root->stream>>>(root->d_image, peers, n / 2);  // image_words = 4*Npix: even
      CK(cudaGetLastError());
    } else {  // no peer access: stage each image through a device-to-device copy
      if (!root->d_peer_stage) CK(cudaMalloc((void**)&root->d_peer_stage, sizeof(unsigned long long) * n));
      for (int k = 1; k < r->n; k++) {
        PeerImages peers;
        peers.n = 1, peers.src[0] = root->d_peer_stage;
        CK(cudaMemcpyPeerAsync(root->d_peer_stage, root->ordinal, r->dev[k]->d_image, r->dev[k]->ordinal, sizeof(unsigned long long) * n, root->stream));
        accumulate_peers_u64<<<root->sm_count * 8, 256, 0, root->stream>>>(root->d_image, peers, n / 2);
        CK(cudaGetLastError());
      }
    }
    CK(cudaEventRecord(r->t1, root->stream));
    CK(cudaStreamSynchronize(root->stream));
  }
  if (reduce_ms) {
    CK(cudaSetDevice(root->ordinal));
    CK(cudaEventSynchronize(r->t1));
    CK(cudaEventElapsedTime(reduce_ms, r->t0, r->t1));
  }
  return 0;
}


// ---- pipelined scan ------------------------------------------------------------------------------------------
// The reference's loop serialises kernel, device->host copy and report (H:861-1040).  Here a device owns two images and
// two pinned host buffers: while projection p is transported into one image, the other one (projection p-1) travels to
// the host on a second stream and is formatted by the host thread (api.c: scan_thread), so the GPU never waits.
extern "C" void mcgpu_dev_pipeline_end(struct mcgpu_device* d) {
  if (!d || !d->pipeline_on) return;
  cudaSetDevice(d->ordinal);
  cudaStreamSynchronize(d->stream);
  if (d->copy_stream) cudaStreamSynchronize(d->copy_stream), cudaStreamDestroy(d->copy_stream);
  cudaFree(d->d_image_alt);
  for (int k = 0; k < 2; k++) {
    if (d->h_pinned[k]) cudaFreeHost(d->h_pinned[k]);
    if (d->p_ev0[k]) cudaEventDestroy(d->p_ev0[k]);
    if (d->p_ev1[k]) cudaEventDestroy(d->p_ev1[k]);
    if (d->p_copied[k]) cudaEventDestroy(d->p_copied[k]);
    d->h_pinned[k] = NULL, d->p_ev0[k] = NULL, d->p_ev1[k] = NULL, d->p_copied[k] = NULL;
  }
  if (d->h_flag) cudaFreeHost(d->h_flag);
  d->h_flag = NULL, d->d_image_alt = NULL, d->copy_stream = NULL, d->pipeline_on = 0;
}

extern "C" int mcgpu_dev_pipeline_begin(struct mcgpu_device* d, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  if (d->pipeline_on) return 0;
  if (!d->d_image) {
    snprintf(err, errlen, "device %d: nothing uploaded", d->ordinal);
    return -1;
  }
  d->pipeline_on = 1;
  const size_t bytes = sizeof(unsigned long long) * d->image_words;
  CK(cudaMalloc((void**)&d->d_image_alt, bytes));
  CK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
  CK(cudaHostAlloc((void**)&d->h_flag, 2 * sizeof(unsigned long long), cudaHostAllocDefault));
  for (int k = 0; k < 2; k++) {
    CK(cudaHostAlloc((void**)&d->h_pinned[k], bytes, cudaHostAllocDefault));
    CK(cudaEventCreate(&d->p_ev0[k]));
    CK(cudaEventCreate(&d->p_ev1[k]));
    CK(cudaEventCreateWithFlags(&d->p_copied[k], cudaEventDisableTiming));
    CK(cudaEventRecord(d->p_copied[k], d->copy_stream));  // "nothing pending" for the first use of the slot
  }
  return 0;
}

extern "C" int mcgpu_dev_pipeline_launch(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen) {
  const int slot = l->image_slot & 1;
  CK(cudaSetDevice(d->ordinal));
  if (!d->pipeline_on) {
    snprintf(err, errlen, "device %d: pipeline not started", d->ordinal);
    return -1;
  }
  CK(cudaStreamWaitEvent(d->stream, d->p_copied[slot], 0));  // the slot's previous image has left the device
  CK(cudaEventRecord(d->p_ev0[slot], d->stream));
  if (mcgpu_dev_launch(d, view, l, err, errlen) != 0) return -1;
  // mcgpu_dev_launch bracketed the kernel with ev0/ev1 on d->stream; per-slot events keep the time of THIS projection
  CK(cudaEventRecord(d->p_ev1[slot], d->stream));
  CK(cudaStreamWaitEvent(d->copy_stream, d->p_ev1[slot], 0));
  CK(cudaMemcpyAsync(d->h_pinned[slot], slot ? d->d_image_alt : d->d_image, sizeof(unsigned long long) * d->image_words, cudaMemcpyDeviceToHost, d->copy_stream));
  CK(cudaMemcpyAsync(d->h_flag, d->d_stream_counter + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, d->copy_stream));
  CK(cudaEventRecord(d->p_copied[slot], d->copy_stream));
  return 0;
}

extern "C" int mcgpu_dev_pipeline_wait(struct mcgpu_device* d, int slot, float* kernel_ms, uint64_t** host, char* err, size_t errlen) {
  slot &= 1;
  CK(cudaSetDevice(d->ordinal));
  CK(cudaEventSynchronize(d->p_copied[slot]));
  if (kernel_ms) CK(cudaEventElapsedTime(kernel_ms, d->p_ev0[slot], d->p_ev1[slot]));
  if (d->h_flag && d->h_flag[0]) {
    snprintf(err, errlen, "device %d: wavefront transport kernel watchdog fired (code %d)", d->ordinal, d->h_flag[0]);
    return -1;
  }
  if (host) *host = d->h_pinned[slot];
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
