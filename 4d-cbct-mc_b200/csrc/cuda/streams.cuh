// Transport kernel, generation 1: the reference's structure (one thread = one RANECU stream, nested loops).
// A/B ONLY: compiled when the library is built with `make AB=1` (-DMCGPU_AB_KERNELS); the product is wavefront.cuh.
#pragma once
#include "transport.cuh"

namespace MCGPU_NS {

// per-thread shell-weight scratch kept in local memory (the reference's rn[MAX_SHELLS], K:1290)
struct RnLocal {
  float v[MCGPU_MAX_SHELLS];
  __device__ __forceinline__ void set(int i, float x) { v[i] = x; }
  __device__ __forceinline__ float get(int i) const { return v[i]; }
};


// GCOa (K:1287-1515): Compton with Doppler broadening (relativistic impulse approximation,
// analytical one-electron profiles).  Updates E, returns the polar cosine.  `rn` is per-thread
// scratch for the shell weights.
template <class RnStore>
__device__ __forceinline__ double sample_compton(float& E, const float4* __restrict__ shells, int nosc, Ranecu& rng, RnStore& rn) {
  float s, s0, af, tau, pzomc = 0.0f;
  double cdt1, costh;
  const float ek = E * 1.956951306108245e-6f;
  const float ek2 = ek * 2.f + 1.f;
  const float ek3 = ek * ek;
  const float taumin = 1.f / ek2;
  const float a1 = logf(ek2);

  s0 = 0.0f;
  for (int i = 0; i < nosc; i++) {
    const float4 sh = shells[i];
    float t = sh.y;
    if (t < E) {
      const float aux = E * (E - t) * 2.f;
      pzomc = compton_pz(sh.z, aux, t);
      if (pzomc > 0.0f)
        t = (0.707106781186545f + pzomc * 1.4142135623731f) * (0.707106781186545f + pzomc * 1.4142135623731f);
      else
        t = (0.707106781186545f - pzomc * 1.4142135623731f) * (0.707106781186545f - pzomc * 1.4142135623731f);
      t = 0.5f * expf(0.5f - t);
      if (pzomc > 0.0f) t = 1.0f - t;
      s0 += sh.x * t;
    }
  }

  do {
    if (rng.uniform() * (a1 + 2. * ek * (ek + 1.f) * taumin * taumin) < a1)
      tau = powf(taumin, rng.uniform());
    else
      tau = sqrtf(1.f + rng.uniform() * (taumin * taumin - 1.f));
    cdt1 = (double)(1.f - tau) / (((double)tau) * ((double)E) * 1.956951306108245e-6);
    if (cdt1 > 2.0) cdt1 = 1.99999999;
    s = 0.0f;
    for (int i = 0; i < nosc; i++) {
      const float4 sh = shells[i];
      float t = sh.y;
      if (t < E) {
        const float aux = E * (E - t) * ((float)cdt1);
        if ((aux > 1.0e-12f) || (t > 1.0e-12f))
          pzomc = compton_pz(sh.z, aux, t);
        else
          pzomc = 0.002f;
        t = pzomc * 1.4142135623731f;
        if (pzomc > 0.0f)
          t = 0.5f - (t + 0.70710678118654502f) * (t + 0.70710678118654502f);
        else
          t = 0.5f - (0.70710678118654502f - t) * (0.70710678118654502f - t);
        t = 0.5f * expf(t);
        if (pzomc > 0.0f) t = 1.0f - t;
        s += sh.x * t;
        rn.set(i, t);
      }
    }
  } while ((rng.uniform() * s0) > (s * (1.0f + tau * ((ek3 - ek2 - 1.0f) + tau * (ek2 + tau * ek3))) / (ek3 * tau * (tau * tau + 1.0f))));

  costh = 1.0 - cdt1;

  for (;;) {
    float t = s * rng.uniform();
    float pac = 0.0f;
    int ishell = nosc - 1;
    for (int i = 0; i < (nosc - 1); i++) {
      pac += shells[i].x * rn.get(i);
      if (pac > t) {
        ishell = i;
        break;
      }
    }
    t = rng.uniform() * rn.get(ishell);
    const float fj0 = shells[ishell].z;
    if (t < 0.5f)
      pzomc = (0.70710678118654502f - sqrtf(0.5f - logf(t + t))) / (fj0 * 1.4142135623731f);
    else
      pzomc = (sqrtf(0.5f - logf(2.0f - 2.0f * t)) - 0.70710678118654502f) / (fj0 * 1.4142135623731f);
    if (pzomc < -1.0f) continue;
    t = tau * (tau - costh * 2.f) + 1.f;  // evaluated in double, stored as float (K:1441)
    if (t > 1.0e-20f)
      af = sqrtf(t) * (tau * (tau - ((float)costh)) / t + 1.f);
    else
      af = 0.00200f;
    if (af > 0.0f)
      t = af * 0.2f + 1.f;
    else
      t = 1.f - af * 0.2f;
    const float pz_lo = (pzomc < 0.2f) ? pzomc : 0.2f;
    const float pz_cl = (pz_lo > -0.2f) ? pz_lo : -0.2f;
    if (rng.uniform() * t < (af * pz_cl + 1.f)) break;
  }

  {
    float t = pzomc * pzomc;
    const float b1 = 1.f - t * tau * tau;
    const float b2 = 1.f - t * tau * ((float)costh);
    float root = sqrtf(fabsf(b2 * b2 - b1 * (1.0f - t)));
    if (pzomc < 0.0f) root *= -1.0f;
    t = (tau / b1) * (b2 + root);
    if (t > 1.0f) t = 1.0f;
    E *= t;
  }
  return costh;
}

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
