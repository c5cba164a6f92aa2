// Transport kernel, generation 2: persistent warps that regroup photons by event.
// A/B ONLY: compiled when the library is built with `make AB=1` (-DMCGPU_AB_KERNELS); the product is wavefront.cuh.
//
// Why: the reference's structure (one thread = one stream, nested variable-length loops:
// histories > interactions > delta-tracking steps, rejection loops over up to 34 shells inside
// Compton) leaves, on B200, 2.3 of 32 lanes active per issued warp instruction (ncu:
// smsp__thread_inst_executed_per_inst_executed = 2.31, profiles/r01_v1_thorax_ncu_summary.txt):
// every nesting level multiplies the divergence loss while the issue slots are ~80 % busy issuing
// mostly-empty instructions.
//
// How: each lane still owns one RANECU stream and runs that stream's histories strictly in order
// (so every float of every trajectory, and therefore every integer tally, is unchanged), but the
// nested loops are flattened into a per-lane state machine and the warp executes one EVENT TYPE at
// a time for all lanes that are waiting for it:
//   W   delta-tracking step (the dominant unit, ~100 instructions)       -> W | C | R | T | N
//   C   Compton, fresh: S0 = incoherent scattering function at theta=pi   -> CT
//   CT  Compton, one tau trial (propose, S(tau), accept/reject)           -> CT | W | N
//   R   Rayleigh interaction                                              -> W
//   T   tally on the detector                                             -> N
//   N   next history of the stream (source sampling)                      -> W | T
//   I   next stream from the global counter (RANECU jump-ahead)           -> N | F(inished)
// Lanes that leave W wait (masked) while the rest keep stepping; when fewer than `w_threshold`
// lanes are still in W the pending events are executed type by type, which turns (almost) all
// lanes back to W.  A rejected Compton trial is not looped on the spot: the lane stays in CT and
// takes its next trial together with the Compton lanes of the next event phase.  The per-shell
// terms of the Compton sums (rsqrtf + expf per shell, up to 34 shells for tissue) are evaluated
// cooperatively by all 32 lanes (coop_shell_terms) and added by the owner lane in shell order, so
// the heavy part runs on full warps whatever the number of photons in Compton.  Streams are handed
// out dynamically (warp-aggregated atomic on a global counter): the grid is persistent, SM count x
// resident CTAs, and finished lanes refill.
#pragma once
#include "transport.cuh"

#ifndef MCGPU_NS
#define MCGPU_NS mcgpu
#endif
namespace MCGPU_NS {

#define MCGPU_REGROUP_BLOCK 128

#define MCGPU_SCRATCH_ROWS 16  // photons per cooperative Compton call; further lanes wait for the next event phase

// Warp-cooperative evaluation of the shell terms of every photon whose lane is in `mask`: the
// (photon, shell) pairs are spread over all 32 lanes (G lanes per photon, G the largest power of
// two with G*popc(mask) <= 32), results go to the warp's scratch row of the photon's rank.  The
// owner lane then adds fco*term in shell order, exactly like the sequential loop of the reference,
// so the sum is bit-identical while the expensive part (rsqrtf, expf) runs on full warps.
__device__ __forceinline__ void coop_shell_terms(unsigned mask, float E, int slot, float factor, bool trial, const float4* __restrict__ sh_shells, const SceneDev& sc,
                                                 float* __restrict__ wbuf, int stride, unsigned lane) {
  const int n = __popc(mask);
  int G = 32;
  while (G * n > 32) G >>= 1;
  const int g = (int)lane / G, sub = (int)lane % G;
  const bool helper = g < n;
  const int owner = helper ? (int)__fns(mask, 0, g + 1) : 0;
  const float oE = __shfl_sync(0xffffffffu, E, owner);
  const int oslot = __shfl_sync(0xffffffffu, slot, owner);
  const float ofac = __shfl_sync(0xffffffffu, factor, owner);
  const bool otrial = __shfl_sync(0xffffffffu, (int)trial, owner) != 0;
  if (helper) {
    const int nosc = sc.cmp_noscco[oslot];
    const float4* sh = sh_shells + oslot * MCGPU_MAX_SHELLS;
#pragma unroll 1
    for (int i = sub; i < nosc; i += G) {
      const float4 s4 = sh[i];
      wbuf[g * stride + i] = s4.x * compton_shell_term(s4, oE, ofac, otrial);
    }
  }
  __syncwarp();
}

// Ordered sum over the shells (the `s0 +=` / `s +=` chain of K:1337, K:1399) of the weighted terms the
// helpers left in `row`.  With KEEP the running sums replace the terms: they are the reference's
// `pac` values of the target-shell search (K:1414-1422), which adds the same numbers in the same order.
template <bool KEEP>
__device__ __forceinline__ float compton_ordered_sum(int nosc, float* __restrict__ row) {
  float s = 0.0f;
#pragma unroll 1
  for (int i = 0; i < nosc; i++) {
    s += row[i];
    if (KEEP) row[i] = s;
  }
  return s;
}

// keep the lowest MCGPU_SCRATCH_ROWS set bits of a lane mask
__device__ __forceinline__ unsigned limit_rows(unsigned m) {
  while (__popc(m) > MCGPU_SCRATCH_ROWS) m &= ~(0x80000000u >> __clz(m));
  return m;
}

template <int BITS, bool DOSE, int ROT>
__global__ void __launch_bounds__(MCGPU_REGROUP_BLOCK, 8)
    transport_regroup(const SceneDev sc, const __grid_constant__ mcgpu_view vw, long long stream_begin, long long stream_end, int histories_per_thread, int seed_input,
                      int g1, int g2, unsigned long long* __restrict__ stream_counter, int w_threshold) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SharedTables& st = *reinterpret_cast<SharedTables*>(smem_raw);
  float4* sh_shells = reinterpret_cast<float4*>(smem_raw + ((sizeof(SharedTables) + 15) & ~size_t(15)));
  float* sh_scratch = reinterpret_cast<float*>(sh_shells + sc.num_slots * MCGPU_MAX_SHELLS);
  const int stride = regroup_scratch_stride(sc.max_shells);
  float2* sh_palette = reinterpret_cast<float2*>(sh_scratch + (MCGPU_REGROUP_BLOCK / 32) * MCGPU_SCRATCH_ROWS * stride + ((MCGPU_REGROUP_BLOCK / 32) * MCGPU_SCRATCH_ROWS * stride & 1));

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
  float* wbuf = sh_scratch + (threadIdx.x >> 5) * MCGPU_SCRATCH_ROWS * stride;  // this warp's shell-term scratch [rows][stride]
  const long long n_streams = stream_end - stream_begin;

  // per-lane photon / stream state (registers)
  Photon p;
  Ranecu rng;
  mcgpu_mfp_record rec;
  float mfp_woodcock = 0.f, mfp_density = 0.f, s0 = 0.f;
  int index = 0, slot = 0, slot_old = -1, scatter_state = 0, hist_left = 0;
  int state = ST_I;
  p.x = p.y = p.z = p.u = p.v = p.w = p.E = 0.f;
  rng.s1 = rng.s2 = 1;
  rec.ax = rec.ay = rec.az = rec.bx = rec.by = rec.bz = rec.pmax_next = rec.pad = 0.f;

  for (;;) {
    const unsigned m_w = __ballot_sync(MCGPU_FULL_MASK, state == ST_W);
    const unsigned m_f = __ballot_sync(MCGPU_FULL_MASK, state == ST_F);
    if (m_f == MCGPU_FULL_MASK) break;
    const int n_w = __popc(m_w);
    const int n_pending = 32 - n_w - __popc(m_f);

    if (n_w >= w_threshold || n_pending == 0) {
      // ---------------------------------------------------------------- W: one delta-tracking step (K:249-279)
      if (state == ST_W) {
        const float step = -(mfp_woodcock)*logf(rng.uniform());
        p.x += step * p.u;
        p.y += step * p.v;
        p.z += step * p.w;
        const int absvox = locate_voxel(sc, p);
        if (absvox < 0) {
          state = ST_T;  // escaped with index > -1: goes to the detector
        } else {
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
          float prob = 1.0f - mfp_density * (rec.ax + p.E * rec.bx);
          const float randno = rng.uniform();
          if (!(randno < prob)) {  // real interaction: classify now (K:289-353), sample later with the other lanes
            prob += mfp_density * (rec.ay + p.E * rec.by);
            if (randno < prob) {
              state = ST_C;
            } else {
              prob += mfp_density * (rec.az + p.E * rec.bz);
              state = (randno < prob) ? ST_R : ST_N;  // else: photoelectric absorption, history over
              if (DOSE && state == ST_N) deposit_energy(sc, p, slot, p.E);  // K:351: all of E is deposited
            }
          }
        }
      }
    } else {
      // ---------------------------------------------------------------- T: detector tally (K:377-381)
      if (state == ST_T) {
        tally_photon<ROT>(sc, vw, p, scatter_state);
        state = ST_N;
      }
      // ---------------------------------------------------------------- I: next stream of the launch (K:198)
      {
        const bool want = (state == ST_N && hist_left == 0) || state == ST_I;
        const unsigned m_i = __ballot_sync(MCGPU_FULL_MASK, want);
        if (m_i) {
          unsigned long long base = 0;
          const int leader = __ffs(m_i) - 1;
          if ((int)lane == leader) base = atomicAdd(stream_counter, (unsigned long long)__popc(m_i));
          base = __shfl_sync(MCGPU_FULL_MASK, base, leader);
          if (want) {
            const long long s = (long long)base + __popc(m_i & lt_mask);
            if (s < n_streams) {
              ranecu_init(rng, stream_begin + s, seed_input, g1, g2);
              hist_left = histories_per_thread;
              state = ST_N;
            } else {
              state = ST_F;
            }
          }
        }
      }
      // ---------------------------------------------------------------- N: next history (K:210-234)
      if (state == ST_N) {
        hist_left--;
        const bool enters = emit_photon<ROT>(sc, vw, st, rng, p);
        scatter_state = 0;
        index = __float2int_rd((p.E - sc.e0) * sc.ide);
        const float2 w = __ldg(&sc.woodcock[index]);
        mfp_woodcock = w.x + p.E * w.y;
        slot_old = -1;
        state = enters ? ST_W : ST_T;  // a primary that misses the voxels can still hit the detector (K:240-241)
      }
      // ---------------------------------------------------------------- C: Compton, S0 for fresh events (K:1315-1339)
      double costh = 0.0;
      bool deflect_pending = false;
      {
        const unsigned m_c = limit_rows(__ballot_sync(MCGPU_FULL_MASK, state == ST_C));
        if (m_c) {
          coop_shell_terms(m_c, p.E, slot, 2.f, false, sh_shells, sc, wbuf, stride, lane);
          if ((m_c >> lane) & 1u) {
            s0 = compton_ordered_sum<false>(sc.cmp_noscco[slot], wbuf + __popc(m_c & lt_mask) * stride);
            state = ST_CT;
          }
          __syncwarp();
        }
      }
      // ---------------------------------------------------------------- CT: one tau trial per lane (K:1342-1403), rest of GCOa if accepted
      {
        const unsigned m_ct = limit_rows(__ballot_sync(MCGPU_FULL_MASK, state == ST_CT));
        if (m_ct) {
          const bool mine = (m_ct >> lane) & 1u;
          const ComptonKin kin(p.E);
          float tau = 1.f;
          double cdt1 = 0.0;
          if (mine) cdt1 = compton_propose_tau(kin, p.E, rng, tau);
          coop_shell_terms(m_ct, p.E, slot, (float)cdt1, true, sh_shells, sc, wbuf, stride, lane);
          if (mine) {
            const int nosc = sc.cmp_noscco[slot];
            float* row = wbuf + __popc(m_ct & lt_mask) * stride;
            const float s = compton_ordered_sum<true>(nosc, row);
            if (compton_accept(kin, s0, s, tau, rng)) {
              const float e_before = p.E;
              costh = compton_finish(p.E, s, tau, cdt1, sh_shells + slot * MCGPU_MAX_SHELLS, nosc, row, rng);
              if (DOSE) deposit_energy(sc, p, slot, -1.0f * (p.E - e_before));  // K:296-301, 359
              deflect_pending = true;
            }
          }
          __syncwarp();
        }
      }
      // ---------------------------------------------------------------- R: Rayleigh (K:329-347)
      if (state == ST_R) {
        costh = sample_rayleigh(sc, p.E, slot, rec.pmax_next, rng);
        deflect_pending = true;
      }
      // ---------------------------------------------------------------- new direction for both kinds of scattering (K:299, K:339)
      if (deflect_pending) {
        deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
        if (state == ST_R) {
          scatter_state = (scatter_state == 0) ? 2 : 3;
          state = ST_W;
        } else {
          index = __float2int_rd((p.E - sc.e0) * sc.ide);
          if (index > -1) {
            const float2 w = __ldg(&sc.woodcock[index]);
            mfp_woodcock = w.x + p.E * w.y;
            slot_old = -2;
            scatter_state = (scatter_state == 0) ? 1 : 3;
            state = ST_W;
          } else {
            state = ST_N;  // below the tabulated energies: absorbed (K:311, K:372)
          }
        }
      }
    }
  }
}

}  // namespace MCGPU_NS
