// CUDA layer of libmcgpu_b200: device memory, uploads, the transport kernel and its launch.
// The C host (csrc/host/api.c) drives it through the extern "C" functions declared in
// csrc/host/mcgpu_host.h.  Built for sm_100a only, with -fmad=false and no fast-math (see
// transport.cuh for why).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include "regroup.cuh"
#include "transport.cuh"

using namespace mcgpu;

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
  SceneDev scene;
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
  unsigned long long* d_stream_counter;  // next stream of the running launch (regrouping kernel)
  int kernel_generation;                 // 2 = regrouping persistent warps (default), 1 = one thread per stream (reference structure, for A/B)
  int w_threshold;
  uint64_t* h_stage;
  int timed;
};

// ------------------------------------------------------------------------------------------
// Transport kernel, generation 1: one RANECU stream (= one thread of the reference grid) per
// thread, histories of a stream run back to back (K:206-382).  BITS selects the voxel packing.
template <int BITS>
__global__ void __launch_bounds__(128) transport_streams(const SceneDev sc, const __grid_constant__ mcgpu_view vw, long long stream_begin, long long stream_end,
                                                         int histories_per_thread, int seed_input, int g1, int g2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SharedTables& st = *reinterpret_cast<SharedTables*>(smem_raw);
  float4* sh_shells = reinterpret_cast<float4*>(smem_raw + ((sizeof(SharedTables) + 15) & ~size_t(15)));
  float2* sh_palette = reinterpret_cast<float2*>(sh_shells + sc.num_slots * MCGPU_MAX_SHELLS);

  for (int i = threadIdx.x; i < MCGPU_MAX_ENERGY_BINS; i += blockDim.x) {
    st.espc[i] = sc.spectrum->espc[i];
    st.cutoff[i] = sc.spectrum->cutoff[i];
    st.alias[i] = sc.spectrum->alias[i];
  }
  if (threadIdx.x == 0) st.num_bins = sc.spectrum->num_bins;
  for (int i = threadIdx.x; i < sc.num_slots * MCGPU_MAX_SHELLS; i += blockDim.x) sh_shells[i] = sc.cmp_shells[i];
  if (BITS == 4 || BITS == 8)
    for (int i = threadIdx.x; i < sc.palette_size; i += blockDim.x) sh_palette[i] = sc.palette[i];
  __syncthreads();

  const long long stream = stream_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (stream >= stream_end) return;

  Ranecu rng;
  ranecu_init(rng, stream, seed_input, g1, g2);
  RnLocal rn;

  for (int h = histories_per_thread; h > 0; h--) {
    Photon p;
    const bool enters = emit_photon(sc, vw, st, rng, p);
    int scatter_state = 0;
    int index = __float2int_rd((p.E - sc.e0) * sc.ide);  // K:220
    float mfp_woodcock;
    {
      const float2 w = __ldg(&sc.woodcock[index]);
      mfp_woodcock = w.x + p.E * w.y;
    }
    int slot_old = -1;
    mcgpu_mfp_record rec;
    rec.ax = rec.ay = rec.az = rec.bx = rec.by = rec.bz = rec.pmax_next = rec.pad = 0.f;

    if (enters) {
      for (;;) {  // interaction loop (K:237-375)
        int absvox, slot;
        float prob, randno, mfp_density;
        do {  // delta-tracking steps until a real interaction or escape (K:249-279)
          const float step = -(mfp_woodcock)*logf(rng.uniform());
          p.x += step * p.u;
          p.y += step * p.v;
          p.z += step * p.w;
          absvox = locate_voxel(sc, p);
          if (absvox < 0) break;
          const float2 md = fetch_voxel<BITS>(sc, sh_palette, absvox);
          slot = __float_as_int(md.y);
          if (slot != slot_old) {
            const float4* r4 = reinterpret_cast<const float4*>(&sc.mfp[(size_t)index * sc.num_slots + slot]);
            const float4 lo = __ldg(r4), hi = __ldg(r4 + 1);
            rec.ax = lo.x, rec.ay = lo.y, rec.az = lo.z, rec.bx = lo.w;
            rec.by = hi.x, rec.bz = hi.y, rec.pmax_next = hi.z;
            slot_old = slot;
          }
          mfp_density = mfp_woodcock * md.x;
          prob = 1.0f - mfp_density * (rec.ax + p.E * rec.bx);
          randno = rng.uniform();
        } while (randno < prob);
        if (absvox < 0) break;

        prob += mfp_density * (rec.ay + p.E * rec.by);
        if (randno < prob) {  // Compton (K:290-326)
          const float e_before = p.E;
          const double costh = sample_compton(p.E, sh_shells + slot * MCGPU_MAX_SHELLS, sc.cmp_noscco[slot], rng, rn);
          deposit_energy(sc, p, slot, -1.0f * (p.E - e_before));
          deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
          index = __float2int_rd((p.E - sc.e0) * sc.ide);
          if (index > -1) {
            const float2 w = __ldg(&sc.woodcock[index]);
            mfp_woodcock = w.x + p.E * w.y;
            slot_old = -2;
            scatter_state = (scatter_state == 0) ? 1 : 3;
          }
        } else {
          prob += mfp_density * (rec.az + p.E * rec.bz);
          if (randno < prob) {  // Rayleigh (K:329-347)
            const double costh = sample_rayleigh(sc, p.E, slot, rec.pmax_next, rng);
            deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
            scatter_state = (scatter_state == 0) ? 2 : 3;
          } else {
            deposit_energy(sc, p, slot, p.E);
            index = -11;  // photoelectric absorption (K:348-353)
          }
        }
        if (index < 0) break;
      }
    }
    if (index > -1) tally_photon(sc, vw, p, scatter_state);
  }
}

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
  CK(cudaMalloc((void**)&d->d_stream_counter, sizeof(unsigned long long)));
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
  {  // tuning / A-B switches (documented in DESIGN.md); the defaults are the product path
    const char* k = getenv("MCGPU_KERNEL");
    const char* t = getenv("MCGPU_W_THRESHOLD");
    d->kernel_generation = (k && atoi(k) == 1) ? 1 : 2;
    d->w_threshold = t ? atoi(t) : 8;
    if (d->w_threshold < 1) d->w_threshold = 1;
    if (d->w_threshold > 32) d->w_threshold = 32;
  }

  SceneDev& sc = d->scene;
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

static int pow_mod_host(long long a, unsigned long long n, long long m) {
  long long y = 1, z = a % m;
  while (n) {
    if (n & 1ull) y = (y * z) % m;
    z = (z * z) % m;
    n >>= 1;
  }
  return (int)y;
}

extern "C" int mcgpu_dev_launch(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  if (!d->d_image) {
    snprintf(err, errlen, "device %d: nothing uploaded", d->ordinal);
    return -1;
  }
  if (l->zero_image) CK(cudaMemsetAsync(d->d_image, 0, sizeof(unsigned long long) * d->image_words, d->stream));
  const long long n_streams = l->stream_end - l->stream_begin;
  CK(cudaEventRecord(d->ev0, d->stream));
  if (n_streams > 0) {
    const int block = 128;
    const long long grid = (n_streams + block - 1) / block;
    const unsigned long long leap = (unsigned long long)(l->histories_per_thread * 256);
    const int g1 = pow_mod_host(40014, leap, 2147483563LL), g2 = pow_mod_host(40692, leap, 2147483399LL);
    size_t smem = ((sizeof(SharedTables) + 15) & ~size_t(15)) + sizeof(float4) * d->scene.num_slots * MCGPU_MAX_SHELLS;
    if (d->voxel_bits == 4 || d->voxel_bits == 8) smem += sizeof(float2) * d->scene.palette_size;
#define LAUNCH_REGROUP(B)                                                                                                                \
  {                                                                                                                                      \
    int per_sm = 0;                                                                                                                      \
    CK(cudaFuncSetAttribute(transport_regroup<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transport_regroup<B>, block, smem));                                   \
    long long pgrid = (long long)d->sm_count * (per_sm > 0 ? per_sm : 1);                                                                \
    if (pgrid > grid) pgrid = grid;                                                                                                      \
    CK(cudaMemsetAsync(d->d_stream_counter, 0, sizeof(unsigned long long), d->stream));                                                  \
    transport_regroup<B><<<(unsigned)pgrid, block, smem, d->stream>>>(d->scene, *view, l->stream_begin, l->stream_end,               \
                                                                          l->histories_per_thread, l->seed_input, g1, g2,                \
                                                                          d->d_stream_counter, d->w_threshold);                          \
  }
#define LAUNCH(B)                                                                                                                        \
  if (d->kernel_generation == 1) {                                                                                                       \
    CK(cudaFuncSetAttribute(transport_streams<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                              \
    transport_streams<B><<<(unsigned)grid, block, smem, d->stream>>>(d->scene, *view, l->stream_begin, l->stream_end, l->histories_per_thread, \
                                                                     l->seed_input, g1, g2);                                            \
  } else {                                                                                                                               \
    smem += sizeof(float) * (MCGPU_REGROUP_BLOCK / 32) * MCGPU_SCRATCH_ROWS * regroup_scratch_stride(d->scene.max_shells) + 8;           \
    LAUNCH_REGROUP(B)                                                                                                                    \
  }
    switch (d->voxel_bits) {
      case 4: LAUNCH(4) break;
      case 8: LAUNCH(8) break;
      case 16: LAUNCH(16) break;
      default: LAUNCH(64) break;
    }
#undef LAUNCH
#undef LAUNCH_REGROUP
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(d->ev1, d->stream));
  d->timed = 1;
  return 0;
}

extern "C" int mcgpu_dev_sync(struct mcgpu_device* d, float* kernel_ms, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  CK(cudaStreamSynchronize(d->stream));
  if (kernel_ms) {
    *kernel_ms = 0.f;
    if (d->timed) CK(cudaEventElapsedTime(kernel_ms, d->ev0, d->ev1));
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
