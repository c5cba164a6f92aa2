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

namespace MCGPU_NS {
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
          const double costh = sample_compton(p.E, sh_shells + slot * MCGPU_MAX_SHELLS, sc.cmp_noscco[slot], rng, rn);
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
            index = -11;  // photoelectric absorption (K:348-353)
          }
        }
        if (index < 0) break;
      }
    }
    if (index > -1) tally_photon(sc, vw, p, scatter_state);
  }
}

}  // namespace MCGPU_NS

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

extern "C" int MCGPU_LAUNCH_NAME(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen) {
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
#define LAUNCH_REGROUP_D(B, DOSE_, ROT_)                                                                                                           \
  {                                                                                                                                      \
    int per_sm = 0;                                                                                                                      \
    CK(cudaFuncSetAttribute(transport_regroup<B, DOSE_, ROT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transport_regroup<B, DOSE_, ROT_>, block, smem));                                   \
    long long pgrid = (long long)d->sm_count * (per_sm > 0 ? per_sm : 1);                                                                \
    if (pgrid > grid) pgrid = grid;                                                                                                      \
    CK(cudaMemsetAsync(d->d_stream_counter, 0, sizeof(unsigned long long), d->stream));                                                  \
    transport_regroup<B, DOSE_, ROT_><<<(unsigned)pgrid, block, smem, d->stream>>>(d->scene, *view, l->stream_begin, l->stream_end,               \
                                                                          l->histories_per_thread, l->seed_input, g1, g2,                \
                                                                          d->d_stream_counter, d->w_threshold);                          \
  }
#define LAUNCH_WAVEFRONT_D(B, DOSE_, ROT_)                                                                                                \
  {                                                                                                                                      \
    const int pal = (B == 4 || B == 8) ? d->scene.palette_size : 0;                                                                      \
    const int wblock = d->wf_block;                                                                                                      \
    int smem_sm = 0, per_sm = MCGPU_WF_MAX_BLOCK / wblock, pool = 0;                                                                     \
    CK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, d->ordinal));                                       \
    int rows = d->wf_rows;                                                                                                               \
    if (rows == 0) { /* 32 scratch rows when the full pool still fits next to them */                                                       \
      const long long b32 = (long long)smem_sm / per_sm - 1024 - (long long)wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, 0, wblock / 32, 32).total; \
      rows = b32 >= (long long)(sizeof(float) * MCGPU_WF_STRIDE) * (2 * wblock) ? 32 : 16;                                               \
    }                                                                                                                                    \
    const size_t fixed = wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, 0, wblock / 32, rows).total;                     \
    for (; per_sm >= 1; per_sm--) {                                                                                                      \
      const long long budget = (long long)smem_sm / per_sm - 1024 - (long long)fixed;                                                    \
      pool = budget > 0 ? (int)(budget / (long long)(sizeof(float) * MCGPU_WF_STRIDE)) & ~31 : 0;                                        \
      if (pool > 2 * wblock) pool = 2 * wblock;                                                                                          \
      if (pool > MCGPU_WF_MAX_POOL) pool = MCGPU_WF_MAX_POOL;                                                                            \
      if (pool >= wblock) break;                                                                                                         \
    }                                                                                                                                    \
    if (pool < 64) {                                                                                                                     \
      snprintf(err, errlen, "device %d: not enough shared memory for the wavefront kernel", d->ordinal);                                 \
      return -1;                                                                                                                         \
    }                                                                                                                                    \
    const size_t wsmem = wavefront_layout(d->scene.num_slots, d->scene.max_shells, pal, pool, wblock / 32, rows).total;                        \
    CK(cudaFuncSetAttribute(transport_wavefront<B, DOSE_, ROT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));              \
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, transport_wavefront<B, DOSE_, ROT_>, wblock, wsmem));                      \
    long long pgrid = (long long)d->sm_count * (per_sm > 0 ? per_sm : 1);                                                                \
    const long long useful = (n_streams + pool - 1) / pool;                                                                              \
    if (pgrid > useful) pgrid = useful;                                                                                                  \
    CK(cudaMemsetAsync(d->d_stream_counter, 0, 2 * sizeof(unsigned long long), d->stream));                                              \
    transport_wavefront<B, DOSE_, ROT_><<<(unsigned)pgrid, wblock, wsmem, d->stream>>>(                                                  \
        d->scene, *view, l->stream_begin, l->stream_end, l->histories_per_thread, l->seed_input, g1, g2, d->d_stream_counter,            \
        d->w_threshold, pool, pal, rows, reinterpret_cast<int*>(d->d_stream_counter + 1));                                        \
  }
#define LAUNCH(B)                                                                                                                        \
  if (d->kernel_generation == 1) {                                                                                                       \
    CK(cudaFuncSetAttribute(transport_streams<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                              \
    transport_streams<B><<<(unsigned)grid, block, smem, d->stream>>>(d->scene, *view, l->stream_begin, l->stream_end, l->histories_per_thread, \
                                                                     l->seed_input, g1, g2);                                            \
  } else if (d->kernel_generation == 3) {                                                                                                \
    if (d->scene.materials_dose || d->scene.voxels_edep) {                                                                               \
      LAUNCH_WAVEFRONT_D(B, true, -1)                                                                                                    \
    } else {                                                                                                                             \
      if (view->rotation_flag == 1) { LAUNCH_WAVEFRONT_D(B, false, 1) } else { LAUNCH_WAVEFRONT_D(B, false, 0) }                         \
    }                                                                                                                                    \
  } else {                                                                                                                               \
    smem += sizeof(float) * (MCGPU_REGROUP_BLOCK / 32) * MCGPU_SCRATCH_ROWS * regroup_scratch_stride(d->scene.max_shells) + 8;           \
    if (d->scene.materials_dose || d->scene.voxels_edep) {                                                                               \
      LAUNCH_REGROUP_D(B, true, -1)                                                                                                        \
    } else {                                                                                                                             \
      if (view->rotation_flag == 1) { LAUNCH_REGROUP_D(B, false, 1) } else { LAUNCH_REGROUP_D(B, false, 0) }                                                                                                        \
    }                                                                                                                                    \
  }
    switch (d->voxel_bits) {
      case 4: LAUNCH(4) break;
      case 8: LAUNCH(8) break;
      case 16: LAUNCH(16) break;
      default: LAUNCH(64) break;
    }
#undef LAUNCH
#undef LAUNCH_WAVEFRONT_D
#undef LAUNCH_REGROUP_D
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(d->ev1, d->stream));
  d->timed = 1;
  return 0;
}

