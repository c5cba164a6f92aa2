// Transport kernel, generation 3: block-level wavefront with typed work queues.
//
// Why: generation 2 (regroup.cuh) regroups the 32 photons a warp owns, so at any time only ~half
// of them wait for the same event: 14.5-16 of 32 lanes are active per issued instruction
// (profiles/r01_final_*_ncu_summary.txt) and the serial parts of Compton run at 4-5 lanes.  The
// kernel is issue-bound, so idle lanes are the loss that is left.
//
// How: a CTA of 16 warps owns a POOL of photon contexts in shared memory (2 per thread; a context
// = one RANECU stream and the photon it is tracking, 13 words) and five queues of context ids, one
// per kind of work:
//   Q_W  delta-tracking steps                       Q_N  tally / next history (source)
//   Q_C  Compton (S0 for fresh events + one tau trial)   Q_R  Rayleigh      Q_I  next RANECU stream
// A warp repeatedly pops up to 32 ids from ONE queue, loads those contexts into registers, runs
// that kind of work for all lanes (the same code as generation 2's phases), stores the contexts and
// pushes every id to the queue of its new state.  Because a queue collects the photons of 512
// threads, batches are (nearly) full whatever the event mix.  A context is held by one warp at a
// time and its stream is advanced strictly in order, so every float of every trajectory and hence
// every integer tally is identical to generations 1 and 2 and to the reference (tested).
//
// Queues are rings of 16-bit ids in shared memory: `tail` is reserved with one warp-aggregated atomicAdd per push,
// entries are published individually (an entry is EMPTY until written), `avail` counts published entries and is
// claimed with atomicCAS by the popping warp, `head` gives it its ring positions.  No locks; the only waits are on
// an entry that is reserved but not yet written, and they are bounded (a watchdog raises the launch's error flag
// instead of hanging).  A leaner protocol without `avail` (pop = one CAS on `head` against `tail`, two atomics per
// exchange instead of four; git fa86281) was measured: +0.5 % on Catphan and thorax but -9 % on the air
// scan, whose photons change queue twice per history -- popping warps then claim entries that are reserved but not
// written yet and burn issue slots waiting for them (r02i, profiles/r02_experiments.txt).
//
// Code size is a first-class constraint here: with exact arithmetic the kernel was bound by instruction-
// cache misses until its hot code fitted the SM's 32 KB instruction cache.  Hence ONE
// scheduling policy (sticky: a CTA drains the queue it is on), ONE context load/store site for every kind
// of batch, ONE cooperative shell-term call site for S0 and S(tau), rolled loops.  Measure
// sm__icc_request_hit_rate / gcc__cache_requests_type_instruction (tools/icc_probe.sh) after any change.
#pragma once
#include "transport.cuh"

namespace MCGPU_NS {

#define MCGPU_WF_MAX_BLOCK 1024
#define MCGPU_WF_EMPTY 0xffffu
#define MCGPU_WF_FIELDS 13
#define MCGPU_WF_STRIDE 13  // words per context in the pool: odd, so contexts spread over the shared-memory banks
#define MCGPU_WF_MAX_POOL 2048
#ifndef MCGPU_WF_CHAIN_MIN
#define MCGPU_WF_CHAIN_MIN 24  // lanes of a batch that must want the same next kind of work for the warp to chain into it (33: never)
#endif

// Q_I: contexts whose RANECU stream is used up (or not assigned yet).  Re-initialising a generator costs ~700 instructions; done
// inside a tally/source batch it ran for the one or two lanes that needed it while the rest of the warp waited (2.6 % of Catphan's
// instructions at 1.2 lanes, r02l); collected in their own queue, 32 generators are re-initialised together.
enum WfQueue : int { Q_W = 0, Q_N = 1, Q_C = 2, Q_R = 3, Q_I = 4, Q_COUNT = 5 };
// F_MFPW: the Woodcock mean free path at the photon's energy (K:246-247), fetched when the photon enters the tracking state so
// that a tracking batch starts without a dependent L2 access
enum WfField : int { F_X = 0, F_Y, F_Z, F_U, F_V, F_W, F_E, F_S1, F_S2, F_S0, F_HIST, F_META, F_MFPW };

struct WfControl {
  unsigned head[Q_COUNT];  // next ring position to pop
  unsigned tail[Q_COUNT];  // next ring position to push (reserved by atomicAdd)
  int avail[Q_COUNT];      // entries published and not yet claimed
  int live;          // contexts that still have work (not finished)
  int active_warps;  // warps that have not retired
  int phase;         // queue the CTA is draining (sticky scheduling)
  int pad[1];
};

// shared-memory carve-up, used by the kernel and by the host to size the launch
struct WfLayout {
  size_t shells, scratch, palette, control, rings, pool, total;
  int stride, ring;
};
// `rows` = photons per cooperative Compton pass (16 or 32 scratch rows per warp); the rings hold 64 ids per warp
// (>= the pool, which is at most 2 contexts per thread; a power of two)
__host__ __device__ inline WfLayout wavefront_layout(int num_slots, int max_shells, int palette_entries, int pool_size, int warps, int rows) {
  WfLayout L;
  L.stride = wavefront_scratch_stride(max_shells);
  L.ring = warps <= 16 ? 1024 : 2048;
  L.shells = (sizeof(SharedTables) + 15) & ~size_t(15);
  L.scratch = L.shells + sizeof(float4) * num_slots * MCGPU_MAX_SHELLS;
  L.palette = (L.scratch + sizeof(float) * warps * rows * L.stride + 15) & ~size_t(15);
  L.control = (L.palette + sizeof(float2) * palette_entries + 15) & ~size_t(15);
  L.rings = L.control + sizeof(WfControl);
  L.pool = (L.rings + sizeof(unsigned short) * Q_COUNT * L.ring + 15) & ~size_t(15);
  L.total = L.pool + sizeof(float) * MCGPU_WF_STRIDE * pool_size;
  return L;
}

__device__ __forceinline__ int wf_pack_meta(int state, int scatter_state, int slot) { return state | (scatter_state << 3) | (slot << 8); }

template <int BITS, bool DOSE, int ROT>
__global__ void __launch_bounds__(MCGPU_WF_MAX_BLOCK, 1)
    transport_wavefront(const SceneDev sc, const __grid_constant__ mcgpu_view vw, long long stream_begin, long long stream_end, int histories_per_thread, int seed_input,
                        int g1, int g2, unsigned long long* __restrict__ stream_counter, int w_threshold, int pool_size, int palette_entries, int rows, int* __restrict__ error_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const WfLayout L = wavefront_layout(sc.num_slots, sc.max_shells, palette_entries, pool_size, (int)(blockDim.x >> 5), rows);
  const int ring = L.ring, ring_mask = L.ring - 1;
  SharedTables& st = *reinterpret_cast<SharedTables*>(smem_raw);
  float4* sh_shells = reinterpret_cast<float4*>(smem_raw + L.shells);
  float* sh_scratch = reinterpret_cast<float*>(smem_raw + L.scratch);
  float2* sh_palette = reinterpret_cast<float2*>(smem_raw + L.palette);
  WfControl* ctl = reinterpret_cast<WfControl*>(smem_raw + L.control);
  volatile unsigned short* rings = reinterpret_cast<volatile unsigned short*>(smem_raw + L.rings);
  float* pool = reinterpret_cast<float*>(smem_raw + L.pool);
  int* pool_i = reinterpret_cast<int*>(pool);
  const int stride = L.stride;

  for (int i = threadIdx.x; i < MCGPU_MAX_ENERGY_BINS; i += blockDim.x) {
    st.espc[i] = sc.spectrum->espc[i];
    st.cutoff[i] = sc.spectrum->cutoff[i];
    st.alias[i] = sc.spectrum->alias[i];
  }
  if (threadIdx.x == 0) st.num_bins = sc.spectrum->num_bins;
  for (int i = threadIdx.x; i < sc.num_slots * MCGPU_MAX_SHELLS; i += blockDim.x) sh_shells[i] = sc.cmp_shells[i];
  if (BITS == 4 || BITS == 8)
    for (int i = threadIdx.x; i < sc.palette_size; i += blockDim.x) sh_palette[i] = sc.palette[i];
  // every context starts in Q_I asking for a stream
  for (int i = threadIdx.x; i < Q_COUNT * ring; i += blockDim.x) rings[i] = MCGPU_WF_EMPTY;
  __syncthreads();
  for (int i = threadIdx.x; i < pool_size; i += blockDim.x) {
    rings[Q_I * ring + i] = (unsigned short)i;
    for (int k = 0; k < MCGPU_WF_STRIDE; k++) pool_i[i * MCGPU_WF_STRIDE + k] = 0;
    pool_i[i * MCGPU_WF_STRIDE + F_META] = wf_pack_meta(ST_I, 0, 0);
    pool_i[i * MCGPU_WF_STRIDE + F_S1] = 1;
    pool_i[i * MCGPU_WF_STRIDE + F_S2] = 1;
  }
  if (threadIdx.x == 0) {
    for (int t = 0; t < Q_COUNT; t++) ctl->head[t] = 0u, ctl->tail[t] = 0u;
    ctl->tail[Q_I] = (unsigned)pool_size;
    for (int t = 0; t < Q_COUNT; t++) ctl->avail[t] = 0;
    ctl->avail[Q_I] = pool_size;
    ctl->live = pool_size;
    ctl->active_warps = (int)(blockDim.x >> 5);
    ctl->phase = Q_I;
  }
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  float* wbuf = sh_scratch + (threadIdx.x >> 5) * rows * stride;
  const long long n_streams = stream_end - stream_begin;
  volatile int* v_avail = ctl->avail;
  volatile int* v_live = &ctl->live;
  volatile int* v_phase = &ctl->phase;

#define PF(f) pool[pid * MCGPU_WF_STRIDE + (f)]
#define PI(f) pool_i[pid * MCGPU_WF_STRIDE + (f)]

#ifdef MCGPU_WF_STATS  // diagnostics build (make ... XFLAGS=-DMCGPU_WF_STATS): batch sizes per queue, tracking steps, idle polls
  unsigned long long st_pops[Q_COUNT] = {0, 0, 0, 0, 0}, st_lanes[Q_COUNT] = {0, 0, 0, 0, 0}, st_wsteps = 0, st_wlanes = 0, st_idle = 0;
#define WF_STAT(x) x
#else
#define WF_STAT(x)
#endif
  // The batch a warp is working on lives in registers across iterations: after a source batch nearly every lane wants a tracking
  // step, and after a tracking batch in an empty geometry (the air scan) every lane wants the tally -- the warp then CHAINS into that
  // kind of work with the lanes it holds (the others are stored and pushed as usual) instead of pushing 32 ids and popping 32 others.
  Photon p;
  Ranecu rng;
  int state = ST_F, slot = 0, scatter_state = 0, hist_left = 0, pid = 0;
  float s0 = 0.f, mfpw = 0.f;
  bool act = false, unsaved = false;  // unsaved: the lanes kept from the previous iteration hold fields their pool entries do not
  int chain_q = -1;
  for (;;) {
    int q = 0, n = 0;
    if (chain_q >= 0) {
      q = chain_q;
      n = __popc(__ballot_sync(MCGPU_FULL_MASK, act));
    } else {
    // ------------------------------------------------------------------ acquire a batch: up to 32 ids of one queue
    unsigned pos = 0;
    unsaved = false;
    if (lane == 0) {
      int idle = 0;
      for (;;) {
        const int a0 = v_avail[0], a1 = v_avail[1], a2 = v_avail[2], a3 = v_avail[3], a4 = v_avail[4];
        int best = -1, a = 0;
        {  // keep draining the queue the CTA is working on, then move to the fullest one: most warps of
           // the CTA run the same kind of code, which is what the SM's 32 KB instruction cache rewards
          const int ph = *v_phase;
          const int ap = ph == Q_W ? a0 : ph == Q_N ? a1 : ph == Q_C ? a2 : ph == Q_R ? a3 : a4;
          if (a4 >= 32) best = Q_I, a = a4;  // a full batch of used-up streams: those contexts do nothing until they are served
          else if (ap >= 32) best = ph, a = ap;
          else {
            if (a0 > a) best = Q_W, a = a0;
            if (a1 > a) best = Q_N, a = a1;
            if (a2 > a) best = Q_C, a = a2;
            if (a3 > a) best = Q_R, a = a3;
            if (a4 > a) best = Q_I, a = a4;
            if (best >= 0 && best != ph) *v_phase = best;
          }
        }
        if (best < 0) {
          if (*v_live == 0) break;  // all streams of this CTA are done
          if (++idle > 4096) {      // ~1 ms without work: this warp retires, the others finish the tail
            if (atomicSub(&ctl->active_warps, 1) > 1) break;
            atomicAdd(&ctl->active_warps, 1);
            if (idle > (1 << 22)) {  // last warp, contexts alive but nowhere: a context was lost
              atomicExch(error_flag, 1);
              break;
            }
          }
          WF_STAT(st_idle++;)
          __nanosleep(200);
          continue;
        }
        const int take = a < 32 ? a : 32;
        if (atomicCAS(&ctl->avail[best], a, a - take) == a) {
          q = best, n = take;
          pos = atomicAdd(&ctl->head[best], (unsigned)take);
          break;
        }
      }
    }
    {  // one broadcast: queue (3 bits), count (6 bits), ring position (the rings have at most 2048 entries)
      const unsigned packed = __shfl_sync(MCGPU_FULL_MASK, (unsigned)q | ((unsigned)n << 3) | ((pos & (unsigned)ring_mask) << 9), 0);
      q = (int)(packed & 7u), n = (int)((packed >> 3) & 63u), pos = packed >> 9;
    }
    if (n <= 0) break;
    WF_STAT(st_pops[q]++; st_lanes[q] += n;)

    act = (int)lane < n;
    pid = 0;
    if (act) {
      volatile unsigned short* e = rings + q * ring + ((pos + lane) & ring_mask);
      unsigned v;
      int guard = 0;
#pragma unroll 1
      while ((v = *e) == MCGPU_WF_EMPTY) {
        if (++guard > (1 << 24)) break;
      }
      if (v == MCGPU_WF_EMPTY) {
        atomicExch(error_flag, 2);
        act = false;
      } else {
        *e = MCGPU_WF_EMPTY;
        pid = (int)v;
      }
    }
    __syncwarp();
    __threadfence_block();

    state = ST_F, slot = 0, scatter_state = 0, hist_left = 0;
    s0 = 0.f, mfpw = 0.f;
    p.x = p.y = p.z = p.u = p.v = p.w = p.E = 0.f;
    rng.s1 = rng.s2 = 1;
    if (act) {  // one load / store site for every kind of batch keeps the code small (the kernel is instruction-cache bound)
      const int meta = PI(F_META);
      state = meta & 7, scatter_state = (meta >> 3) & 3, slot = meta >> 8;
      rng.s1 = PI(F_S1), rng.s2 = PI(F_S2);
      p.x = PF(F_X), p.y = PF(F_Y), p.z = PF(F_Z), p.u = PF(F_U), p.v = PF(F_V), p.w = PF(F_W), p.E = PF(F_E);
      s0 = PF(F_S0);
      hist_left = PI(F_HIST);
      mfpw = PF(F_MFPW);
    }
    }  // acquired or chained

    if (q == Q_W) {
      // ---------------------------------------------------------------- W: delta-tracking steps (K:249-279)
      mcgpu_mfp_record rec;
      rec.ax = rec.ay = rec.az = rec.bx = rec.by = rec.bz = rec.pmax_next = rec.pad = 0.f;
      float mfp_woodcock = 0.f;
      int index = 0, slot_old = -1;
      if (act) {
        index = __float2int_rd((p.E - sc.e0) * sc.ide);
        mfp_woodcock = mfpw;
        // (fetching the table record of the photon's last material here, ahead of the first step, was measured: -1 % Catphan, -0.4 %
        //  thorax, r02f -- the kernel is bound by issue slots and dependent arithmetic, not by the latency of that access)
      }
      const int thr = min(w_threshold, (n + 1) >> 1);
      do {
        if (state == ST_W) {
          const float step = -(mfp_woodcock)*log_uniform(rng.uniform());
          p.x += step * p.u;
          p.y += step * p.v;
          p.z += step * p.w;
          if (outside_box(sc, p)) {
            state = ST_T;  // escaped with index > -1: goes to the detector
          } else {
            const float2 md = fetch_voxel<BITS>(sc, sh_palette, voxel_index(sc, p));
            slot = __float_as_int(md.y);
            if (slot != slot_old) {
              const float4* r4 = reinterpret_cast<const float4*>(&sc.mfp[(size_t)index * sc.num_slots + slot]);
              const float4 lo = __ldg(r4), hi = __ldg(r4 + 1);
              rec.ax = lo.x, rec.ay = lo.y, rec.az = lo.z, rec.bx = lo.w;
              rec.by = hi.x, rec.bz = hi.y, rec.pmax_next = hi.z;
              slot_old = slot;
            }
            const float mfp_density = mfp_woodcock * md.x;
            float prob = 1.0f - mfp_density * (rec.ax + p.E * rec.bx);
            const float randno = rng.uniform();
            if (!(randno < prob)) {  // real interaction: classify now (K:289-353), sample in a batch of its kind
              prob += mfp_density * (rec.ay + p.E * rec.by);
              if (randno < prob) {
                state = ST_C;
              } else {
                prob += mfp_density * (rec.az + p.E * rec.bz);
                state = (randno < prob) ? ST_R : ST_N;  // else: photoelectric absorption, history over
                if (DOSE && state == ST_N) deposit_energy(sc, p, slot, p.E);  // K:351
              }
            }
          }
        }
        WF_STAT(st_wsteps++; st_wlanes += __popc(__ballot_sync(MCGPU_FULL_MASK, state == ST_W));)
      } while (__popc(__ballot_sync(MCGPU_FULL_MASK, state == ST_W)) >= thr);
    } else if (q == Q_N) {
      // ---------------------------------------------------------------- T / I / N: tally, next stream, next history
      const int thr = min(16, (n + 1) >> 1);
      for (;;) {
        if (state == ST_T) {  // K:377-381
          tally_photon<ROT>(sc, vw, p, scatter_state);
          state = ST_N;
        }
        if (state == ST_N && hist_left == 0) state = ST_I;  // the stream is used up: the next one is assigned in a Q_I batch
        if (state == ST_N) {  // K:210-234
          hist_left--;
          const bool enters = emit_photon<ROT>(sc, vw, st, rng, p);
          scatter_state = 0;
          state = enters ? ST_W : ST_T;  // a primary that misses the voxels can still hit the detector (K:240-241)
          if (enters) {  // K:246-247, for the tracking batches to come
            const float2 wc = __ldg(&sc.woodcock[__float2int_rd((p.E - sc.e0) * sc.ide)]);
            mfpw = wc.x + p.E * wc.y;
          }
        }
        if (__popc(__ballot_sync(MCGPU_FULL_MASK, state == ST_T)) < thr) break;
      }
    } else if (q == Q_I) {
      // ---------------------------------------------------------------- I: next stream of the launch (K:198), RANECU jump-ahead (K:841-894)
      const unsigned m_i = __ballot_sync(MCGPU_FULL_MASK, act);
      unsigned long long base = 0;
      const int leader = __ffs(m_i) - 1;
      if ((int)lane == leader) base = atomicAdd(stream_counter, (unsigned long long)__popc(m_i));
      base = __shfl_sync(MCGPU_FULL_MASK, base, leader);
      if (act) {
        const long long s = (long long)base + __popc(m_i & lt_mask);
        if (s < n_streams) {
          ranecu_init(rng, stream_begin + s, seed_input, g1, g2);
          hist_left = histories_per_thread;
          state = ST_N;
        } else {
          state = ST_F;
        }
      }
    } else {
      // ---------------------------------------------------------------- C / CT / R: one scattering step
      double costh = 0.0;
      bool deflect_pending = false;
      if (q == Q_C) {
        // The scratch holds `rows` photons (16: lanes [0,16) and [16,32) take turns, two helper lanes per photon;
        // 32 when the scene's shell count leaves room: one turn, each lane evaluates its own photon).  One
        // call site serves S0 of the fresh events (pass 0, K:1315-1339) and the S of the trial (pass 1,
        // K:1359-1402): a single copy of the shell-term code in the instruction stream.
        const unsigned live = __ballot_sync(MCGPU_FULL_MASK, act);
        const unsigned fresh = __ballot_sync(MCGPU_FULL_MASK, state == ST_C);
        const ComptonKin kin(p.E);
        float tau = 1.f, s = 0.f;
        double cdt1 = 0.0;
        if (act) cdt1 = compton_propose_tau(kin, p.E, rng, tau);  // S0 draws no random numbers: the order of the stream is kept
        const int nosc = sc.cmp_noscco[slot];
        float* row = wbuf + (lane & (unsigned)(rows - 1)) * stride;
#pragma unroll 1
        for (int h = 0; h < (rows == 32 ? 1 : 2); h++) {
          const unsigned half = rows == 32 ? live : live & (0xffffu << (16 * h));
          if (!half) continue;
          const bool mine = (half >> lane) & 1u;
#pragma unroll 1
          for (int pass = (half & fresh) ? 0 : 1; pass < 2; pass++) {
            const unsigned sel = pass == 0 ? (half & fresh) : half;
            coop_shell_terms_half(h, rows, sel, p.E, slot, pass == 0 ? 2.f : (float)cdt1, pass != 0, sh_shells, sc, wbuf, stride, lane);
            if ((sel >> lane) & 1u) {
              const float sum = compton_ordered_sum_rt(nosc, row, pass != 0);
              if (pass == 0) {
                s0 = sum;
                state = ST_CT;
              } else {
                s = sum;
              }
            }
            __syncwarp();
          }
          if (mine && compton_accept(kin, s0, s, tau, rng)) {  // rest of GCOa (K:1405-1513)
            const float e_before = p.E;
            costh = compton_finish(p.E, s, tau, cdt1, sh_shells + slot * MCGPU_MAX_SHELLS, nosc, row, rng);
            if (DOSE) deposit_energy(sc, p, slot, -1.0f * (p.E - e_before));  // K:296-301, 359
            deflect_pending = true;
          }
          __syncwarp();
        }
      } else if (act) {  // Rayleigh (K:329-347); pmax of the bin above, same table entry the tracking step used
        const int index = __float2int_rd((p.E - sc.e0) * sc.ide);
        const float pmax_next = __ldg(&sc.mfp[(size_t)index * sc.num_slots + slot].pmax_next);
        costh = sample_rayleigh(sc, p.E, slot, pmax_next, rng);
        deflect_pending = true;
      }
      if (deflect_pending) {  // new direction for both kinds of scattering (K:299, K:339)
        deflect(p, costh, 6.28318530717958647693 * rng.uniform_d());
        if (state == ST_R) {
          scatter_state = (scatter_state == 0) ? 2 : 3;
          state = ST_W;
        } else {
          const int index = __float2int_rd((p.E - sc.e0) * sc.ide);
          if (index > -1) {
            scatter_state = (scatter_state == 0) ? 1 : 3;
            state = ST_W;
            const float2 wc = __ldg(&sc.woodcock[index]);  // the energy changed: K:308-309
            mfpw = wc.x + p.E * wc.y;
          } else {
            state = ST_N;  // below the tabulated energies: absorbed (K:311, K:372)
          }
        }
      }
    }

    // ------------------------------------------------------------------ chain, or store what every kind changes and hand the ids on
    int nq = !act || state == ST_F ? -1 : state == ST_W ? Q_W : (state == ST_C || state == ST_CT) ? Q_C : state == ST_R ? Q_R : state == ST_I ? Q_I : Q_N;
    chain_q = -1;
    if (q == Q_N || q == Q_W || q == Q_I) {  // new stream -> source -> tracking, tracking -> tally: keep going with the lanes that want it when they are >= 3/4 of a warp
      const int want = q == Q_N ? Q_W : Q_N;
      if (__popc(__ballot_sync(MCGPU_FULL_MASK, nq == want)) >= MCGPU_WF_CHAIN_MIN) chain_q = want;
    }
    const bool keep = chain_q >= 0 && nq == chain_q;
    {
      const unsigned m_f = __ballot_sync(MCGPU_FULL_MASK, act && state == ST_F);
      if (m_f && lane == 0) atomicSub(&ctl->live, __popc(m_f));
    }
    if (act && !keep) {
      PF(F_X) = p.x, PF(F_Y) = p.y, PF(F_Z) = p.z;
      PI(F_S1) = rng.s1, PI(F_S2) = rng.s2;
      PI(F_META) = wf_pack_meta(state, scatter_state, slot);
      if (q != Q_W || unsaved) {  // a tracking batch moves the photon and draws random numbers; direction, energy, S0 and the history count stay
        PF(F_U) = p.u, PF(F_V) = p.v, PF(F_W) = p.w, PF(F_E) = p.E;
        PF(F_S0) = s0;
        PI(F_HIST) = hist_left;
        PF(F_MFPW) = mfpw;
      }
    }
    if (keep) nq = -1;  // not pushed
    act = keep;
    if (chain_q >= 0) unsaved = true;
    if (__all_sync(MCGPU_FULL_MASK, nq < 0)) continue;  // nothing to hand on
    __threadfence_block();
    {  // lanes bound for the same queue find each other with one match: one reservation per queue, no loop over queues
      const unsigned peers = __match_any_sync(MCGPU_FULL_MASK, nq);
      const int leader = __ffs(peers) - 1, cnt = __popc(peers);
      unsigned base = 0;
      if (nq >= 0 && (int)lane == leader) base = atomicAdd(&ctl->tail[nq], (unsigned)cnt);
      base = __shfl_sync(MCGPU_FULL_MASK, base, leader);
      if (nq >= 0) {
        volatile unsigned short* e = rings + nq * ring + ((base + __popc(peers & lt_mask)) & ring_mask);
        int guard = 0;
#pragma unroll 1
        while (*e != MCGPU_WF_EMPTY) {  // its previous occupant is being taken by another warp right now
          if (++guard > (1 << 24)) {
            atomicExch(error_flag, 3);
            break;
          }
        }
        *e = (unsigned short)pid;
      }
      __threadfence_block();
      __syncwarp();
      if (nq >= 0 && (int)lane == leader) atomicAdd(&ctl->avail[nq], cnt);
    }
  }
#ifdef MCGPU_WF_STATS
  if (lane == 0) {
    unsigned long long* g = stream_counter + 2;
    for (int t = 0; t < 4; t++) atomicAdd(g + t, st_pops[t]), atomicAdd(g + 4 + t, st_lanes[t]);
    atomicAdd(g + 8, st_wsteps), atomicAdd(g + 9, st_wlanes), atomicAdd(g + 10, st_idle);
  }
#endif
#undef WF_STAT
#undef PF
#undef PI
}

}  // namespace MCGPU_NS
