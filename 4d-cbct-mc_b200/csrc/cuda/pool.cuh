// Transport kernel, generation 3: persistent warps over a per-lane POOL of photon contexts.
//
// Generation 2 (regroup.cuh) executes one event type at a time for the lanes waiting for it, but a
// warp only owns 32 photons, so while the pending events accumulate the delta-tracking phase runs
// with ~40 % of its lanes (ncu, profiles/r01_v2*: 13 of 32 lanes per instruction).  Here every lane
// owns P photon contexts (P RANECU streams), parked in shared memory as a structure of arrays
// [field][context][lane] (bank = lane: conflict-free whichever context a lane picks).  Each phase
// loads the fields it needs for ONE context of the lane that is waiting for that phase, and stores
// what it changed.  A lane is idle in the delta-tracking phase only if none of its P contexts is in
// flight, and an event type is run when enough lanes hold a context waiting for it, so both the
// steps and the events run on well-filled warps.  Per stream nothing changes: its histories, its
// random numbers and every float operation are executed in the reference's order, so the tallies
// stay bit-identical (checked against generations 1 and 2 and the reference CUDA source).
//
// Context states:  W in flight | C Compton, needs S0 | CT Compton, needs a tau trial | R Rayleigh |
//                  T tally then next history | N next history (or next stream) | F finished
#pragma once
#include "transport.cuh"

namespace mcgpu {

enum CtxField : int { F_X = 0, F_Y, F_Z, F_U, F_V, F_W, F_E, F_S1, F_S2, F_MFPW, F_INDEX, F_SLOT, F_AX, F_BX, F_HIST, F_S0, F_COUNT };
enum CtxState : int { CS_W = 0, CS_C = 1, CS_CT = 2, CS_R = 3, CS_T = 4, CS_N = 5, CS_F = 6 };

#define MCGPU_POOL_BLOCK 128

struct PoolTuning {
  int th_w;  // below this many lanes able to step, every pending event type is run
  int th_n;  // run tally / next-history when this many lanes wait for it
  int th_c;  // same for Compton
  int th_r;  // same for Rayleigh
};

__host__ __device__ inline int pool_scratch_stride(int max_shells) { return max_shells | 1; }
__host__ __device__ inline size_t pool_smem_bytes(int num_slots, int max_shells, int palette_size, int contexts, bool palette_in_smem) {
  size_t b = (sizeof(SharedTables) + 15) & ~size_t(15);
  b += sizeof(float4) * (size_t)num_slots * MCGPU_MAX_SHELLS;
  b += sizeof(float) * (MCGPU_POOL_BLOCK / 32) * 32 * (size_t)pool_scratch_stride(max_shells);
  b += sizeof(unsigned) * (size_t)MCGPU_POOL_BLOCK * F_COUNT * contexts;
  b = (b + 7) & ~size_t(7);
  if (palette_in_smem) b += sizeof(float2) * (size_t)palette_size;
  return b;
}

template <int P>
__device__ __forceinline__ int find_ctx(unsigned states, int st) {
#pragma unroll
  for (int p = 0; p < P; p++)
    if ((int)((states >> (4 * p)) & 15u) == st) return p;
  return -1;
}
__device__ __forceinline__ unsigned with_state(unsigned states, int p, int st) { return (states & ~(15u << (4 * p))) | ((unsigned)st << (4 * p)); }

template <int BITS, int P>
__global__ void __launch_bounds__(MCGPU_POOL_BLOCK)
    transport_pool(const SceneDev sc, const __grid_constant__ mcgpu_view vw, long long stream_begin, long long stream_end, int histories_per_thread, int seed_input,
                   int g1, int g2, unsigned long long* __restrict__ stream_counter, const PoolTuning tune) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SharedTables& st = *reinterpret_cast<SharedTables*>(smem_raw);
  float4* sh_shells = reinterpret_cast<float4*>(smem_raw + ((sizeof(SharedTables) + 15) & ~size_t(15)));
  float* sh_scratch = reinterpret_cast<float*>(sh_shells + sc.num_slots * MCGPU_MAX_SHELLS);
  const int stride = pool_scratch_stride(sc.max_shells);
  unsigned* sh_ctx = reinterpret_cast<unsigned*>(sh_scratch + (MCGPU_POOL_BLOCK / 32) * 32 * stride);
  float2* sh_palette = reinterpret_cast<float2*>((reinterpret_cast<size_t>(sh_ctx + MCGPU_POOL_BLOCK * F_COUNT * P) + 7) & ~size_t(7));

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

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int warp = threadIdx.x >> 5;
  float* wbuf = sh_scratch + warp * 32 * stride;          // shell-term scratch of this warp [32][stride]
  unsigned* cw = sh_ctx + warp * (F_COUNT * P * 32) + lane;  // this lane's column of the warp's context pool
  const long long n_streams = stream_end - stream_begin;

#define CTX_U(f, p) cw[((f) * P + (p)) * 32]
#define CTX_F(f, p) (reinterpret_cast<float*>(cw))[((f) * P + (p)) * 32]

  unsigned states = 0;
#pragma unroll
  for (int p = 0; p < P; p++) {
    states = with_state(states, p, CS_N);
    CTX_U(F_HIST, p) = 0u;  // no histories left: the context fetches a stream first
  }

  for (;;) {
    const bool has_w = find_ctx<P>(states, CS_W) >= 0;
    const bool has_n = find_ctx<P>(states, CS_T) >= 0 || find_ctx<P>(states, CS_N) >= 0;
    const bool has_c = find_ctx<P>(states, CS_C) >= 0 || find_ctx<P>(states, CS_CT) >= 0;
    const bool has_r = find_ctx<P>(states, CS_R) >= 0;
    const int c_w = __popc(__ballot_sync(0xffffffffu, has_w));
    const int c_n = __popc(__ballot_sync(0xffffffffu, has_n));
    const int c_c = __popc(__ballot_sync(0xffffffffu, has_c));
    const int c_r = __popc(__ballot_sync(0xffffffffu, has_r));
    if ((c_w | c_n | c_c | c_r) == 0) break;  // every context of the warp is finished
    const bool starving = c_w < tune.th_w;

    // ------------------------------------------------------------------ T / N: tally, next history (or next stream), source
    if (c_n && (c_n >= tune.th_n || starving)) {
      int sel = find_ctx<P>(states, CS_T);
      if (sel >= 0) {  // K:377-381
        Photon p;
        p.x = CTX_F(F_X, sel), p.y = CTX_F(F_Y, sel), p.z = CTX_F(F_Z, sel);
        p.u = CTX_F(F_U, sel), p.v = CTX_F(F_V, sel), p.w = CTX_F(F_W, sel);
        p.E = CTX_F(F_E, sel);
        tally_photon(sc, vw, p, (int)(CTX_U(F_HIST, sel) & 3u));
      } else {
        sel = find_ctx<P>(states, CS_N);
      }
      bool active = sel >= 0;
      int hist_left = active ? (int)(CTX_U(F_HIST, sel) >> 2) : 1;
      Ranecu rng;
      rng.s1 = rng.s2 = 1;
      const bool want = active && hist_left == 0;
      const unsigned m_i = __ballot_sync(0xffffffffu, want);
      if (m_i) {  // next stream of the launch (K:198): one atomic per warp
        unsigned long long base = 0;
        const int leader = __ffs(m_i) - 1;
        if ((int)lane == leader) base = atomicAdd(stream_counter, (unsigned long long)__popc(m_i));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (want) {
          const long long s = (long long)base + __popc(m_i & lt_mask);
          if (s < n_streams) {
            ranecu_init(rng, stream_begin + s, seed_input, g1, g2);
            hist_left = histories_per_thread;
          } else {
            states = with_state(states, sel, CS_F);
            active = false;
          }
        }
      }
      if (active) {  // K:210-234
        if (!want) rng.s1 = (int)CTX_U(F_S1, sel), rng.s2 = (int)CTX_U(F_S2, sel);
        hist_left--;
        Photon p;
        const bool enters = emit_photon(sc, vw, st, rng, p);
        const int index = __float2int_rd((p.E - sc.e0) * sc.ide);
        const float2 w = __ldg(&sc.woodcock[index]);
        CTX_F(F_X, sel) = p.x, CTX_F(F_Y, sel) = p.y, CTX_F(F_Z, sel) = p.z;
        CTX_F(F_U, sel) = p.u, CTX_F(F_V, sel) = p.v, CTX_F(F_W, sel) = p.w;
        CTX_F(F_E, sel) = p.E;
        CTX_U(F_S1, sel) = (unsigned)rng.s1, CTX_U(F_S2, sel) = (unsigned)rng.s2;
        CTX_F(F_MFPW, sel) = w.x + p.E * w.y;
        CTX_U(F_INDEX, sel) = (unsigned)index;
        CTX_U(F_SLOT, sel) = (unsigned)-1;
        CTX_U(F_HIST, sel) = (unsigned)hist_left << 2;  // scatter state 0
        states = with_state(states, sel, enters ? CS_W : CS_T);  // a primary that misses the voxels can still hit the detector
      }
    }

    // ------------------------------------------------------------------ C / CT: Compton
    if (c_c && (c_c >= tune.th_c || starving)) {
      {  // S0 for fresh events (K:1315-1339)
        const int sel = find_ctx<P>(states, CS_C);
        const unsigned m_c = __ballot_sync(0xffffffffu, sel >= 0);
        if (m_c) {
          const float E = sel >= 0 ? CTX_F(F_E, sel) : 0.f;
          const int slot = sel >= 0 ? (int)CTX_U(F_SLOT, sel) : 0;
          coop_shell_terms<0>(m_c, E, slot, 2.f, sh_shells, sc, wbuf, stride, lane);
          if (sel >= 0) {
            CTX_F(F_S0, sel) = compton_ordered_sum<false>(sc.cmp_noscco[slot], wbuf + __popc(m_c & lt_mask) * stride);
            states = with_state(states, sel, CS_CT);
          }
          __syncwarp();
        }
      }
      {  // one tau trial (K:1342-1403); on acceptance the rest of GCOa and the deflection
        const int sel = find_ctx<P>(states, CS_CT);
        const unsigned m_ct = __ballot_sync(0xffffffffu, sel >= 0);
        if (m_ct) {
          float E = 1.f, tau = 1.f;
          int slot = 0;
          double cdt1 = 0.0;
          Ranecu rng;
          rng.s1 = rng.s2 = 1;
          if (sel >= 0) {
            E = CTX_F(F_E, sel);
            slot = (int)CTX_U(F_SLOT, sel);
            rng.s1 = (int)CTX_U(F_S1, sel), rng.s2 = (int)CTX_U(F_S2, sel);
          }
          const ComptonKin kin(E);
          if (sel >= 0) cdt1 = compton_propose_tau(kin, E, rng, tau);
          coop_shell_terms<1>(m_ct, E, slot, (float)cdt1, sh_shells, sc, wbuf, stride, lane);
          if (sel >= 0) {
            const float4* shells = sh_shells + slot * MCGPU_MAX_SHELLS;
            const int nosc = sc.cmp_noscco[slot];
            float* row = wbuf + __popc(m_ct & lt_mask) * stride;
            const float s = compton_ordered_sum<true>(nosc, row);
            if (compton_accept(kin, CTX_F(F_S0, sel), s, tau, rng)) {
              const double costh = compton_finish(E, s, tau, cdt1, shells, nosc, row, rng);
              Photon p;
              p.u = CTX_F(F_U, sel), p.v = CTX_F(F_V, sel), p.w = CTX_F(F_W, sel);
              deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
              CTX_F(F_U, sel) = p.u, CTX_F(F_V, sel) = p.v, CTX_F(F_W, sel) = p.w;
              CTX_F(F_E, sel) = E;
              const int index = __float2int_rd((E - sc.e0) * sc.ide);
              if (index > -1) {
                const float2 w = __ldg(&sc.woodcock[index]);
                CTX_F(F_MFPW, sel) = w.x + E * w.y;
                CTX_U(F_INDEX, sel) = (unsigned)index;
                CTX_U(F_SLOT, sel) = (unsigned)-2;
                const unsigned h = CTX_U(F_HIST, sel);
                CTX_U(F_HIST, sel) = (h & ~3u) | (((h & 3u) == 0u) ? 1u : 3u);
                states = with_state(states, sel, CS_W);
              } else {
                states = with_state(states, sel, CS_N);  // below the tabulated energies: absorbed (K:311, K:372)
              }
            }
            CTX_U(F_S1, sel) = (unsigned)rng.s1, CTX_U(F_S2, sel) = (unsigned)rng.s2;
          }
          __syncwarp();
        }
      }
    }

    // ------------------------------------------------------------------ R: Rayleigh (K:329-347)
    if (c_r && (c_r >= tune.th_r || starving)) {
      const int sel = find_ctx<P>(states, CS_R);
      if (sel >= 0) {
        const float E = CTX_F(F_E, sel);
        const int slot = (int)CTX_U(F_SLOT, sel);
        const int index = (int)CTX_U(F_INDEX, sel);
        Ranecu rng;
        rng.s1 = (int)CTX_U(F_S1, sel), rng.s2 = (int)CTX_U(F_S2, sel);
        const float pmax = __ldg(&sc.mfp[(size_t)index * sc.num_slots + slot].pmax_next);
        const double costh = sample_rayleigh(sc, E, slot, pmax, rng);
        Photon p;
        p.u = CTX_F(F_U, sel), p.v = CTX_F(F_V, sel), p.w = CTX_F(F_W, sel);
        deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
        CTX_F(F_U, sel) = p.u, CTX_F(F_V, sel) = p.v, CTX_F(F_W, sel) = p.w;
        CTX_U(F_S1, sel) = (unsigned)rng.s1, CTX_U(F_S2, sel) = (unsigned)rng.s2;
        const unsigned h = CTX_U(F_HIST, sel);
        CTX_U(F_HIST, sel) = (h & ~3u) | (((h & 3u) == 0u) ? 2u : 3u);
        states = with_state(states, sel, CS_W);
      }
    }

    // ------------------------------------------------------------------ W: one delta-tracking step (K:249-279)
    {
      const int sel = find_ctx<P>(states, CS_W);
      if (sel >= 0) {
        Photon p;
        p.x = CTX_F(F_X, sel), p.y = CTX_F(F_Y, sel), p.z = CTX_F(F_Z, sel);
        p.u = CTX_F(F_U, sel), p.v = CTX_F(F_V, sel), p.w = CTX_F(F_W, sel);
        p.E = CTX_F(F_E, sel);
        Ranecu rng;
        rng.s1 = (int)CTX_U(F_S1, sel), rng.s2 = (int)CTX_U(F_S2, sel);
        const float mfp_woodcock = CTX_F(F_MFPW, sel);
        const float step = -(mfp_woodcock)*logf(rng.uniform());
        p.x += step * p.u;
        p.y += step * p.v;
        p.z += step * p.w;
        CTX_F(F_X, sel) = p.x, CTX_F(F_Y, sel) = p.y, CTX_F(F_Z, sel) = p.z;
        const int absvox = locate_voxel(sc, p);
        if (absvox < 0) {
          states = with_state(states, sel, CS_T);  // escaped with index > -1: goes to the detector
        } else {
          const float2 md = fetch_voxel<BITS>(sc, sh_palette, absvox);
          const int slot = __float_as_int(md.y);
          float ax, bx;
          float4 lo, hi;
          bool have_rec = false;
          if (slot != (int)CTX_U(F_SLOT, sel)) {
            const float4* r4 = reinterpret_cast<const float4*>(&sc.mfp[(size_t)CTX_U(F_INDEX, sel) * sc.num_slots + slot]);
            lo = __ldg(r4), hi = __ldg(r4 + 1);
            ax = lo.x, bx = lo.w;
            CTX_U(F_SLOT, sel) = (unsigned)slot;
            CTX_F(F_AX, sel) = ax, CTX_F(F_BX, sel) = bx;
            have_rec = true;
          } else {
            ax = CTX_F(F_AX, sel), bx = CTX_F(F_BX, sel);
          }
          const float mfp_density = mfp_woodcock * md.x;
          float prob = 1.0f - mfp_density * (ax + p.E * bx);
          const float randno = rng.uniform();
          if (!(randno < prob)) {  // real interaction: classify now (K:289-353), sample with the other lanes later
            if (!have_rec) {
              const float4* r4 = reinterpret_cast<const float4*>(&sc.mfp[(size_t)CTX_U(F_INDEX, sel) * sc.num_slots + slot]);
              lo = __ldg(r4), hi = __ldg(r4 + 1);
            }
            prob += mfp_density * (lo.y + p.E * hi.x);  // Compton: a.y + E*b.y
            int next = CS_C;
            if (!(randno < prob)) {
              prob += mfp_density * (lo.z + p.E * hi.y);  // Rayleigh: a.z + E*b.z
              next = (randno < prob) ? CS_R : CS_N;       // else photoelectric absorption: history over
            }
            states = with_state(states, sel, next);
          }
        }
        CTX_U(F_S1, sel) = (unsigned)rng.s1, CTX_U(F_S2, sel) = (unsigned)rng.s2;
      }
    }
  }
#undef CTX_U
#undef CTX_F
}

}  // namespace mcgpu
