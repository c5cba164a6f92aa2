/*
 * mcgpu_b200.h -- C ABI of libmcgpu_b200.so, the B200-native drop-in for the
 * MC-GPU v1.3 photon-transport path that cbctmc/mc drives.
 *
 * The reference has no library boundary for this path: cbctmc shells out to the
 * executable `MC-GPU_v1.3.x <input.in>` (cbctmc/mc/simulation.py:187-198,
 * cbctmc/docker.py:31-65).  Each entry point below therefore replaces one stage
 * of that executable's `main` (docker/mcgpu/MC-GPU_v1.3.cu:377-1214, "H" below;
 * the kernel file MC-GPU_kernel_v1.3.cu is "K").  The executable shipped by
 * this repo (csrc/host/main.c) is a thin argv -> ABI shim, so `run-mc` keeps
 * working unchanged, and INTEGRATION.md shows the ctypes binding a cbctmc
 * maintainer would add to skip the process boundary.
 *
 * Conventions: plain C types only; every function returning int gives 0 on
 * success or a negative MCGPU_E_* code (the first four mirror the reference's
 * exit codes, H:1255/1287/2823/988); mcgpu_last_error() returns a message for
 * the last failure on that context.  A context is single-threaded from the
 * caller's point of view.  Host image buffers are caller-owned:
 * uint64_t[4 * Nx * Nz], planes = non-scattered, Compton, Rayleigh, multiple
 * scatter (K:545-548), pixel = ix + iz*Nx, value = sum of round(E[eV]*100).
 *
 * There is no CPU fallback: every run entry point fails with MCGPU_E_CUDA when
 * no sm_100 device is usable.
 */
#ifndef MCGPU_B200_H_
#define MCGPU_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCGPU_OK 0
#define MCGPU_E_ARG (-1)    /* missing / unreadable file, bad argument   (reference exit(-1)) */
#define MCGPU_E_PARSE (-2)  /* malformed or inconsistent input            (reference exit(-2)) */
#define MCGPU_E_OUTPUT (-3) /* output file cannot be opened               (reference exit(-3)) */
#define MCGPU_E_NOMEM (-4)  /* host allocation failed                     (reference exit(-4)) */
#define MCGPU_E_CUDA (-5)   /* no usable device / CUDA runtime error */
#define MCGPU_E_STATE (-6)  /* call order violated (e.g. run before load) */

typedef struct mcgpu_ctx mcgpu_ctx;

/* What `main` prints/needs about a loaded simulation (H:501-543, H:823-848). */
typedef struct mcgpu_info {
  int num_projections;          /* H:1538-1542 */
  int num_pixels_x, num_pixels_z;
  int num_voxels_x, num_voxels_y, num_voxels_z;
  int num_materials_used;       /* materials present in the voxels */
  int num_energy_values;        /* rows of the MFP tables (H:2239) */
  int num_spectrum_bins;
  int threads_per_block;        /* .in value */
  int histories_per_thread;     /* current value; grows and sticks when > 65535 blocks (H:825-835) */
  int num_blocks;               /* blocks of the reference launch for the current history count */
  int seed_input;               /* .in seed */
  int enable_specific_angles;
  int num_devices;
  int voxel_bits;               /* packed voxel layout in HBM: 4, 8, 16 (palette index) or 64 (float2) */
  int palette_size;
  unsigned long long requested_histories; /* .in value (H:1293) */
  unsigned long long launched_histories;  /* blocks*tpb*hpt (H:841) */
  float mean_energy_spectrum;   /* eV (H:3575) */
  float e0, ide;                /* MFP energy grid (H:2308, H:2337) */
  int fast_math;                /* 0 = bit-exact arithmetic (default), 1 = reference's shipped fast-math flags */
} mcgpu_info;

typedef void (*mcgpu_progress_cb)(int projection_index, int num_projections, double seconds, void* user);

/* ---- lifecycle -------------------------------------------------------------------------- */

/* Replaces init_CUDA_device's device pick (H:2454-2557).  device_ids==NULL or n_devices<=0
 * means "all visible devices"; ids >= deviceCount fall back to what is visible (Q13).
 * Returns NULL only when out of memory; a context without devices can still parse inputs
 * and build tables (used by the CPU-side tests), but every run call fails. */
mcgpu_ctx* mcgpu_create(const int* device_ids, int n_devices);
void mcgpu_destroy(mcgpu_ctx* ctx);
const char* mcgpu_last_error(const mcgpu_ctx* ctx);
/* Quiet (0, default for library use) or the reference's stdout banners (1, used by main). */
void mcgpu_set_verbose(mcgpu_ctx* ctx, int verbose);

/* ---- input stages ----------------------------------------------------------------------- */

/* read_input + init_energy_spectrum + set_CT_trajectory (H:1240-1895, 3498-3587, 3280-3434). */
int mcgpu_load_input(mcgpu_ctx* ctx, const char* in_path);
/* load_voxels (H:1996-2145).  vox_path==NULL uses the path named in the .in file. */
int mcgpu_load_voxels(mcgpu_ctx* ctx, const char* vox_path);
/* Same stage from memory (x fastest): what a cbctmc binding would call instead of writing a
 * text .vox.gz (cbctmc/mc/geometry.py:579-623).  material = 1-based MC-GPU material number. */
int mcgpu_set_voxels(mcgpu_ctx* ctx, int nx, int ny, int nz, float dx_cm, float dy_cm, float dz_cm,
                     const uint8_t* material, const float* density);
/* load_material (H:2177-2443).  paths==NULL uses the list in the .in file.  Also builds the
 * device-side layouts and uploads everything to every device of the context (H:2612-2690). */
int mcgpu_load_materials(mcgpu_ctx* ctx, const char* const* paths, int n_paths);
/* Overrides of the .in values, applied before the next run (used by benchmarks / bindings). */
int mcgpu_set_histories(mcgpu_ctx* ctx, unsigned long long total_histories);
int mcgpu_set_seed(mcgpu_ctx* ctx, int seed);
/* Arithmetic of the transport kernels.  0 (default): -fmad=false, no fast-math -- tallies bit-identical to
 * the reference CUDA source compiled the same way.  1: the flags the reference ships with
 * (docker/compile.sh:36, -use_fast_math) -- faster, statistically equivalent results only. */
int mcgpu_set_fast_math(mcgpu_ctx* ctx, int on);

/* ---- simulation ------------------------------------------------------------------------- */

/* One iteration of the projection loop (H:667-1056) for projection p (0-based) with the
 * reference's single-rank seed schedule (Q1: the seed of projection p is closed-form in p,
 * H:869 + H:3456-3485): zero the tally, transport, copy the 4 planes into image_host.
 * With several devices in the context the stream (= reference thread id) range is split
 * evenly across them and the integer tallies are summed on device 0 over NVLink. */
int mcgpu_run_projection(mcgpu_ctx* ctx, int p, uint64_t* image_host);
/* Same launch restricted to the streams [stream_begin, stream_end) of projection p on the
 * context's first device -- the unit a multi-process driver shards (one rank per GPU);
 * summing the images of a partition of [0, num_blocks*threads_per_block) is bit-identical
 * to mcgpu_run_projection.  image_host may be NULL to leave the tally on the device. */
int mcgpu_run_streams(mcgpu_ctx* ctx, int p, long long stream_begin, long long stream_end, uint64_t* image_host);
/* Device pointer of the tally of the context's first device (uint64[4*Npix]), for callers
 * that reduce it themselves (torch.distributed / NCCL). */
void* mcgpu_device_image(mcgpu_ctx* ctx);
/* Device time of the transport kernel(s) of the last run call, in ms (CUDA events on the launch stream). */
double mcgpu_last_kernel_ms(const mcgpu_ctx* ctx);

/* History-split runs (several devices, one projection): device time of the reduction of the partial images in the last
 * mcgpu_run_projection (the reference: MPI_Reduce, H:1019), and how it was done: "peer-kernel" (ONE kernel on device 0 reading
 * every peer's image through NVLink peer mappings; the default where all devices are peers), "ncclReduce" (ncclUint64 sum over
 * NVLink/NVSwitch, one communicator per device, libnccl opened at run time; default without full peer access, or with
 * MCGPU_REDUCE=nccl in the environment), "staged-copy" (neither) or "none". */
double mcgpu_last_reduce_ms(const mcgpu_ctx* ctx);
const char* mcgpu_reduce_kind(const mcgpu_ctx* ctx);

/* Where the last mcgpu_run_all spent its time, summed over the devices' host threads: out[0] wall seconds of the loop,
 * [1] device seconds in transport kernels, [2] host seconds waiting for kernel + device->host copy, [3] host seconds
 * formatting and writing the reports, [4] projections simulated, [5] devices.  Returns the number of values written. */
int mcgpu_get_scan_stats(const mcgpu_ctx* ctx, double* out, int n);

/* The whole projection loop: projections are dealt round-robin to the devices (p -> p mod n),
 * each projection written with mcgpu_write_projection_ascii as soon as it is done, the
 * callback invoked in projection order.  With fewer projections than devices the
 * history-split path of mcgpu_run_projection is used instead. */
int mcgpu_run_all(mcgpu_ctx* ctx, mcgpu_progress_cb cb, void* user);

/* report_image (H:2783-2953): '<base>_%010.6fdeg' ASCII file, byte-compatible format. */
int mcgpu_write_projection_ascii(mcgpu_ctx* ctx, int p, const uint64_t* image, double seconds);
/* Optional binary side-file '<base>_%010.6fdeg.raw': little-endian float32 [4][Nz][Nx], same values as the
 * ASCII columns (the reference's own .raw writer is commented out, H:2911-2949).  mcgpu_run_all writes it
 * next to every ASCII file when the environment has MCGPU_WRITE_RAW=1. */
int mcgpu_write_projection_raw(mcgpu_ctx* ctx, int p, const uint64_t* image);
/* ---- projection post-processing on the device -------------------------------------------------
 * What cbctmc computes in NumPy/SciPy after the simulation, from the u64 tallies instead of the text files.
 *
 * mcgpu_post_intensity: the float32 images MCProjection._read_raw (cbctmc/mc/projection.py:36-51) would obtain
 * from the ASCII file of this tally -- the values "%.8lf" prints and np.loadtxt reads back, detector rows flipped,
 * x cropped to crop_x (n_detector_pixels_half_fan[0] = 1024; <= 0: no crop) -- summed like projections_to_itk
 * (projection.py:143-149): total = sum of the 4 planes, unscattered = plane 0, scattered = planes 1..3.  Outputs
 * are host arrays [Nz][crop_x] (any may be NULL); min_positive[3] receives the smallest positive value of each
 * (for the stack-wide `min_non_zero`, projection.py:151).  tally == NULL: use the tally of the projection
 * simulated last on device 0 (no host round trip).  launched_histories = mcgpu_info.launched_histories. */
int mcgpu_post_intensity(mcgpu_ctx* ctx, const uint64_t* tally, unsigned long long launched_histories, int crop_x, float* total, float* unscattered, float* scattered,
                         float* min_positive);
/* scipy.ndimage.gaussian_filter(in[n0][n1] float32, sigma=(sigma0, sigma1)) as normalize_projections applies it to
 * the air image (projection.py:108-111; sigma (10, 10) in simulation.py:241): mode 'reflect', truncate 4.0. */
int mcgpu_post_gaussian(mcgpu_ctx* ctx, const float* in, int n0, int n1, double sigma0, double sigma1, float* out);
/* The kernel mcgpu_post_gaussian uses for one axis: w[k], k = 0..radius, the weight at distance k of
 * scipy.ndimage's _gaussian_kernel1d(sigma, 0, radius = int(4 sigma + 0.5)) -- same formula and summation order, equal to
 * within 2 ulp of float64 (libm exp vs NumPy's SIMD exp); host only.  Returns the radius. */
int mcgpu_gaussian_weights(double sigma, double* w, int capacity);
/* In place on stack[n_images][n0][n1]: zeros -> min_nonzero (projection.py:153), then log(air / p) in float32
 * (normalize_projections, projection.py:119-120). */
int mcgpu_post_normalize(mcgpu_ctx* ctx, const float* air, float* stack, long long n_images, int n0, int n1, float min_nonzero);

/* Name report_image gives the file of projection p; returns strlen or <0. */
int mcgpu_projection_filename(const mcgpu_ctx* ctx, int p, char* out, size_t out_len);

/* ---- dose tallies (optional) ------------------------------------------------------------ */

/* tally_materials_dose / tally_voxel_energy_deposition (K:1547-1563, K:418-443), enabled by the .in's
 * SECTION DOSE DEPOSITION (off in every cbctmc run).  The counters accumulate over run calls, like the
 * reference accumulates them over projections; mcgpu_run_all resets them first and writes the reports
 * at the end.  which = "materials": uint64[25][2] (sum of round(Edep*100), sum of round(Edep^2)) by
 * material number; "voxels": uint64[ROI voxels][2], x fastest inside the ROI.  out==NULL returns the
 * number of words (0 when the tally is off). */
int mcgpu_reset_dose(mcgpu_ctx* ctx);
long long mcgpu_get_dose(mcgpu_ctx* ctx, const char* which, uint64_t* out, size_t cap_words);
/* report_voxels_dose + report_materials_dose (H:2976-3263): '<dose file>', '.raw', '_2sigma.raw'. */
int mcgpu_write_dose_reports(mcgpu_ctx* ctx, double seconds, int projections_simulated);

/* ---- introspection (tests, bindings) ---------------------------------------------------- */

int mcgpu_get_info(const mcgpu_ctx* ctx, mcgpu_info* out);
/* Seed the reference's main would hand to the kernel for projection p (H:869, H:3456-3485). */
int mcgpu_projection_seed(mcgpu_ctx* ctx, int p, int* seed_out);
/* Host copy of a table in the REFERENCE layout, by name, for parity tests:
 *   "woodcock" float2[nE]            (H:2434-2441)   "mfp_a","mfp_b" float3[nE*25] (H:2300-2358)
 *   "rayleigh_xco|pco|aco|bco" float[128*25], "rayleigh_itlco|ituco" uint8[128*25],
 *   "rayleigh_pmax" float[nE*25]     (H:2304, 2381-2394)
 *   "compton_fco|uico|fj0" float[25*40], "compton_noscco" int[25] (H:2415-2426)
 *   "espc","espc_cutoff" float[256], "espc_alias" int16[256]      (H:3498-3587)
 *   "views": mcgpu_view records (source + detector pose) per projection, see csrc/host/mcgpu_host.h
 *   "voxel_material" uint8[Nvox], "voxel_density" float[Nvox], "voxel_packed" (device layout)
 *   "density_max" float[25]          (H:2132)
 * Returns bytes copied (<= cap) or a negative error; out==NULL returns the size. */
long long mcgpu_copy_table(const mcgpu_ctx* ctx, const char* name, void* out, size_t cap);

/* Exhaustive device-side checks of the arithmetic shortcuts the transport kernel takes inside the reference's expressions,
 * each against the CUDA function it replaces, bit for bit over its whole domain; *mismatches must come back 0:
 *   "log_uniform"  logf without its special-case branches, on every value RANECU can return (K:965-1015 -> K:251)
 *   "rsqrt_normal" rsqrtf without its subnormal scaling, on every positive normal float (K:1329, K:1369)
 *   "outside_box"  locate_voxel's six float comparisons (K:1033-1045) as three unsigned ones, on every non-NaN float per
 *                  axis of the loaded geometry (needs load_voxels + load_materials first) */
int mcgpu_device_selftest(mcgpu_ctx* ctx, const char* name, unsigned long long* mismatches);

/* RANECU helpers exposed for known-answer tests (K:841-894, K:965-986, H:3456-3485). */
void mcgpu_ranecu_init_stream(long long stream, int histories_per_thread, int seed_input, int* s1, int* s2);
float mcgpu_ranecu_next(int* s1, int* s2);
int mcgpu_ranecu_advance_projection_seed(int seed, unsigned long long total_histories);
/* Grid rule H:823-841; histories_per_thread is in/out (sticky). */
void mcgpu_grid_rule(unsigned long long requested, int threads_per_block, int* histories_per_thread,
                     int* num_blocks, unsigned long long* launched);

#ifdef __cplusplus
}
#endif
#endif /* MCGPU_B200_H_ */
