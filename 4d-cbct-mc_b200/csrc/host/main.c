/* MC-GPU_v1.3.x -- command-line drop-in: `MC-GPU_v1.3.x <input.in>`.
 *
 * A thin argv -> C-ABI shim around libmcgpu_b200 (include/mcgpu_b200.h) that keeps the process
 * contract cbctmc relies on (cbctmc/mc/simulation.py:187-226):
 *   - exactly one argument, the .in file; exit 0 on success, the reference's negative codes on
 *     failure (docker/mcgpu/MC-GPU_v1.3.cu:1255, 1287, 2823);
 *   - stdout carries "Simulating Projection <i> of <n>" per projection (H:680), which cbctmc
 *     scrapes for its progress bar, and never the substring "error" on success (Q9);
 *   - it tolerates being started N times by `mpirun -n N`: there is no MPI in this engine, one
 *     process drives every visible GPU, so ranks other than 0 (as seen in the launcher's
 *     environment) exit 0 immediately.
 *
 * Built with -DMCGPU_BATCH_MAIN the same file gives MC-GPU_v1.3_batch.x <a.in> <b.in> ...: the 4D case
 * (one input file per respiratory phase, cbctmc/mc/simulation.py:622-692) in ONE process, so that the CUDA
 * context and the device buffers are set up once instead of once per phase, and the next input is
 * parsed and uploaded (into a second context) while the current one is simulated.  Every input is simulated
 * exactly as a separate invocation would (same seeds, same files). */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mcgpu_b200.h"

static int launcher_rank(void) {
  /* MPI launchers only: SLURM_PROCID alone (srun -n4 ... phase_$SLURM_PROCID.in) is one independent run per task */
  const char* names[] = {"OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "MV2_COMM_WORLD_RANK"};
  size_t i;
  for (i = 0; i < sizeof names / sizeof names[0]; i++) {
    const char* v = getenv(names[i]);
    if (v && *v) return atoi(v);
  }
  return 0;
}

static void on_projection(int p, int total, double seconds, void* user) {
  (void)user;
  if (seconds < 0.0) {
    if (total != 1) printf("\n\n\n   << Simulating Projection %d of %d >>\n\n\n", p + 1, total);
  } else
    printf("       projection %d of %d done in %.3f s\n", p + 1, total, seconds);
  fflush(stdout);
}

static int die(mcgpu_ctx* ctx, int rc) {
  printf("\n\n   !!MC-GPU b200 failure (code %d)!! %s\n\n", rc, mcgpu_last_error(ctx));
  fflush(stdout);
  mcgpu_destroy(ctx);
  return rc;
}

static double seconds_since(const struct timespec* t0) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (t.tv_sec - t0->tv_sec) + 1e-9 * (t.tv_nsec - t0->tv_nsec);
}

static int load_one(mcgpu_ctx* ctx, const char* in_path, int quiet) {
  struct timespec t0;
  double t_in, t_vox;
  int rc;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  if (!quiet) printf("\n    -- Reading the input file '%s':\n", in_path);
  if ((rc = mcgpu_load_input(ctx, in_path)) != MCGPU_OK) return rc;
  t_in = seconds_since(&t0);
  if ((rc = mcgpu_load_voxels(ctx, NULL)) != MCGPU_OK) return rc;
  t_vox = seconds_since(&t0);
  rc = mcgpu_load_materials(ctx, NULL, 0);
  if (!quiet && rc == MCGPU_OK) /* the devices are opened in the background meanwhile; the last stage waits for them and uploads */
    printf("       input + spectrum + poses %.3f s, voxels %.3f s, material tables + waiting for the devices + upload %.3f s\n", t_in, t_vox - t_in, seconds_since(&t0) - t_vox);
  return rc;
}

static int simulate_one(mcgpu_ctx* ctx, const struct timespec* t0) {
  struct timespec t1, t2;
  mcgpu_info info;
  double t_init, t_total;
  int rc;
  mcgpu_get_info(ctx, &info);
  printf("              x-ray tracks to simulate = %llu\n", info.requested_histories);
  printf("                   initial random seed = %d\n", info.seed_input);
  printf("                number of pixels image = %dx%d = %d\n", info.num_pixels_x, info.num_pixels_z, info.num_pixels_x * info.num_pixels_z);
  printf("                 number of projections = %d\n", info.num_projections);
  printf("            number of energy bins read = %d\n", info.num_spectrum_bins);
  printf("                  mean energy spectrum = %.3f keV\n", 0.001f * info.mean_energy_spectrum);
  printf("       Number of voxels in the input geometry file: %d x %d x %d\n", info.num_voxels_x, info.num_voxels_y, info.num_voxels_z);
  printf("       Packed voxel layout: %d bits per voxel, %d distinct (material, density) pairs, %d materials in use\n", info.voxel_bits, info.palette_size, info.num_materials_used);
  printf("       Number of energy values in the mean free path database: %d\n", info.num_energy_values);
  printf("       ==> CUDA: %d device(s) in use; executing %d blocks of %d threads, %d histories per thread: %llu histories per projection\n", info.num_devices, info.num_blocks,
         info.threads_per_block, info.histories_per_thread, info.launched_histories);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  t_init = (t1.tv_sec - t0->tv_sec) + 1e-9 * (t1.tv_nsec - t0->tv_nsec);
  printf("\n    -- INITIALIZATION finished: elapsed time = %.3f s. \n\n", t_init);
  printf("\n\n    -- MONTE CARLO LOOP phase.\n\n");
  fflush(stdout);

  if (info.num_devices < 1) return MCGPU_E_CUDA;
  if ((rc = mcgpu_run_all(ctx, on_projection, NULL)) != MCGPU_OK) return rc;

  clock_gettime(CLOCK_MONOTONIC, &t2);
  t_total = (t2.tv_sec - t0->tv_sec) + 1e-9 * (t2.tv_nsec - t0->tv_nsec);
  mcgpu_get_info(ctx, &info);
  printf("\n\n\n    -- SIMULATION FINISHED!\n");
  printf("\n\n       ****** TOTAL SIMULATION PERFORMANCE (including initialization and reporting) ******\n\n");
  printf("          >>> Execution time including initialization, transport and report: %.3f s.\n", t_total);
  printf("          >>> Time spent in the Monte Carlo transport and reporting: %.3f s.\n", t_total - t_init);
  printf("          >>> Total number of simulated x rays:  %llu\n", info.launched_histories * (unsigned long long)info.num_projections);
  {
    double st[6] = {0, 0, 0, 0, 0, 0};
    if (mcgpu_get_scan_stats(ctx, st, 6) == 6 && st[4] > 0.0) /* projection-parallel scans only */
      printf("          >>> Projection loop: %.3f s wall for %.0f projections on %.0f device(s) = %.4f s per projection; per device thread: "
             "%.3f s in transport kernels, %.3f s waiting for kernel + device->host copy, %.3f s formatting and writing reports\n",
             st[0], st[4], st[5], st[0] / st[4], st[1] / st[5], st[2] / st[5], st[3] / st[5]);
  }
  if (t_total > 0.000001)
    printf("          >>> Total speed (including initialization time) [x-rays/s]:  %.2f\n\n", (double)(info.launched_histories * (unsigned long long)info.num_projections) / t_total);
  fflush(stdout);
  return MCGPU_OK;
}

#ifdef MCGPU_BATCH_MAIN
/* the next input is parsed and uploaded by this thread, into the other context, while the current one is simulated */
typedef struct {
  mcgpu_ctx* ctx;
  const char* path;
  int rc;
} prefetch_job;

static void* prefetch_main(void* arg) {
  prefetch_job* j = (prefetch_job*)arg;
  /* silent: cbctmc scrapes stdout of the RUNNING scan for its progress bar; this context's banners would land in the middle */
  mcgpu_set_verbose(j->ctx, 0);
  j->rc = load_one(j->ctx, j->path, 1);
  mcgpu_set_verbose(j->ctx, 1);
  return NULL;
}
#endif

int main(int argc, char** argv) {
  struct timespec t0;
  mcgpu_ctx* ctx;
  time_t now = time(NULL);
  int rc;

  if (launcher_rank() != 0) {
    printf("MC-GPU (mcgpu-b200): MPI rank %d has nothing to do -- rank 0 drives every visible GPU in one process; exiting.\n", launcher_rank());
    return 0;
  }
  clock_gettime(CLOCK_MONOTONIC, &t0);
#ifdef MCGPU_BATCH_MAIN
  if (argc < 2) {
#else
  if (argc != 2) {
#endif
    printf("\n\n   !!read_input!! %s\n\n", argc > 2 ? "Too many input parameters: provide only the input file name." : "Input file name not given as an execution parameter.");
    return -1;
  }
  printf("\n\n     *****************************************************************************\n");
  printf("     ***   MC-GPU v1.3 command-line compatible engine (mcgpu-b200, sm_100a)      ***\n");
  printf("     ***   drop-in for the MC-GPU v1.3 x-ray transport code of A. Badal (FDA)    ***\n");
  printf("     *****************************************************************************\n\n");
  printf("****** Code execution started on: %s\n\n", ctime(&now));
  printf("\n             *** CUDA SIMULATION IN THE GPU ***\n");
  printf("\n    -- INITIALIZATION phase:\n");
  fflush(stdout);

  ctx = mcgpu_create(NULL, 0); /* every visible device; a .in gpu id that is out of range means the same (Q13) */
  printf("       CUDA driver initialised in %.3f s; the devices are being opened in the background\n", seconds_since(&t0));
  if (!ctx) {
    printf("\n\n   !!out of memory creating the context!!\n\n");
    return -4;
  }
  mcgpu_set_verbose(ctx, 1);
#ifndef MCGPU_BATCH_MAIN
  if ((rc = load_one(ctx, argv[1], 0)) != MCGPU_OK) return die(ctx, rc);
  if ((rc = simulate_one(ctx, &t0)) != MCGPU_OK) return die(ctx, rc);
#else
  {
    mcgpu_ctx* both[2];
    int k;
    both[0] = ctx;
    both[1] = argc > 2 ? mcgpu_create(NULL, 0) : NULL;
    if (argc > 2 && !both[1]) {
      printf("\n\n   !!out of memory creating the context!!\n\n");
      mcgpu_destroy(ctx);
      return -4;
    }
    if (both[1]) mcgpu_set_verbose(both[1], 1);
    if ((rc = load_one(both[0], argv[1], 0)) != MCGPU_OK) {
      if (both[1]) mcgpu_destroy(both[1]);
      return die(both[0], rc);
    }
    for (k = 1; k < argc; k++) {
      mcgpu_ctx* cur = both[(k - 1) & 1];
      prefetch_job job;
      pthread_t th;
      int have_thread = 0;
      job.ctx = both[k & 1], job.path = k + 1 < argc ? argv[k + 1] : NULL, job.rc = MCGPU_OK;
      if (job.path) have_thread = pthread_create(&th, NULL, prefetch_main, &job) == 0;
      rc = simulate_one(cur, &t0);
      if (job.path && !have_thread) job.rc = load_one(job.ctx, job.path, 0); /* no thread: load in turn */
      if (have_thread) {
        pthread_join(th, NULL);
        if (job.rc == MCGPU_OK) printf("\n    -- Input file '%s' was read while the previous one was simulated.\n", job.path);
      }
      if (rc != MCGPU_OK || job.rc != MCGPU_OK) {
        mcgpu_ctx* bad = rc != MCGPU_OK ? cur : job.ctx;
        mcgpu_destroy(bad == both[0] ? both[1] : both[0]);
        return die(bad, rc != MCGPU_OK ? rc : job.rc);
      }
      clock_gettime(CLOCK_MONOTONIC, &t0);
    }
    if (both[1]) mcgpu_destroy(both[1]);
  }
#endif
  now = time(NULL);
  printf("\n****** Code execution finished on: %s\n\n", ctime(&now));
  mcgpu_destroy(ctx);
  return 0;
}
