/* C ABI (include/mcgpu_b200.h): context lifecycle, stage ordering, the projection loop and its
 * two multi-GPU partitions.  Mirrors the control flow of the reference's main
 * (docker/mcgpu/MC-GPU_v1.3.cu:377-1214) for the single-rank, fixed-history case (the only
 * reproducible one): per projection -> grid rule (H:823-841), launch with the
 * current seed (H:861), advance the seed by the launched histories (H:869), copy the tally
 * back, report, zero the tally.  Because the seed of projection p and the RANECU state of
 * stream t are closed-form in (p, t), projections can be dealt to different GPUs and a
 * projection's streams can be split across GPUs without changing a single integer tally. */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mcgpu_host.h"

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

typedef struct open_job {
  int ordinal, threaded;
  pthread_t thread;
  struct mcgpu_device* dev;
  char err[256];
} open_job;

static void* open_main(void* arg) {
  open_job* j = (open_job*)arg;
  j->dev = mcgpu_dev_open(j->ordinal, j->err, sizeof j->err);
  return NULL;
}

typedef struct upload_job {
  mcgpu_ctx* ctx;
  int device, threaded, rc;
  pthread_t thread;
  char err[512];
} upload_job;

static void* upload_main(void* arg) {
  upload_job* j = (upload_job*)arg;
  mcgpu_ctx* ctx = j->ctx;
  j->rc = mcgpu_dev_upload(ctx->dev[j->device], &ctx->scene, &ctx->vol, &ctx->spc, ctx->views[0].total_num_pixels, j->err, sizeof j->err);
  return NULL;
}

mcgpu_ctx* mcgpu_create(const int* device_ids, int n_devices) {
  mcgpu_ctx* ctx = (mcgpu_ctx*)calloc(1, sizeof *ctx);
  int visible, i;
  if (!ctx) return NULL;
  visible = mcgpu_dev_count();
  if (visible <= 0) { /* parsing and table building still work; run calls fail loudly */
    snprintf(ctx->err, sizeof ctx->err, "no CUDA device visible");
    return ctx;
  }
  if (!device_ids || n_devices <= 0) n_devices = visible;
  ctx->dev = (struct mcgpu_device**)calloc((size_t)n_devices, sizeof *ctx->dev);
  if (!ctx->dev) {
    free(ctx);
    return NULL;
  }
  { /* CUDA context creation costs ~0.3-1 s per device (more on a cold driver): open the devices concurrently, in the background */
    open_job* jobs = (open_job*)calloc((size_t)n_devices, sizeof *jobs);
    int n_jobs = 0;
    if (!jobs) {
      free(ctx->dev);
      free(ctx);
      return NULL;
    }
    for (i = 0; i < n_devices; i++) {
      int id = device_ids ? device_ids[i] : i, k, dup = 0;
      if (id < 0 || id >= visible) id = n_jobs % visible; /* Q13: out-of-range id -> use what is visible */
      for (k = 0; k < n_jobs; k++) dup |= (jobs[k].ordinal == id);
      if (dup) continue;
      jobs[n_jobs++].ordinal = id;
    }
    /* the threads run on while the caller parses the input files; whoever needs a device first joins them (mcgpu_devices_ready) */
    for (i = 0; i < n_jobs; i++) jobs[i].threaded = pthread_create(&jobs[i].thread, NULL, open_main, &jobs[i]) == 0;
    ctx->opening = jobs;
    ctx->n_opening = n_jobs;
  }
  return ctx;
}

/* Join the background device opens of mcgpu_create (idempotent); devices that could not be opened are dropped. */
void mcgpu_devices_ready(mcgpu_ctx* ctx) {
  open_job* jobs;
  int i;
  if (!ctx || !ctx->opening) return;
  jobs = (open_job*)ctx->opening;
  for (i = 0; i < ctx->n_opening; i++) {
    if (jobs[i].threaded)
      pthread_join(jobs[i].thread, NULL);
    else
      open_main(&jobs[i]);
    if (jobs[i].dev) {
      mcgpu_dev_set_fast_math(jobs[i].dev, ctx->fast_math);
      ctx->dev[ctx->num_devices++] = jobs[i].dev;
    } else
      snprintf(ctx->err, sizeof ctx->err, "%s", jobs[i].err);
  }
  free(jobs);
  ctx->opening = NULL;
  ctx->n_opening = 0;
}

void mcgpu_destroy(mcgpu_ctx* ctx) {
  int i;
  if (!ctx) return;
  mcgpu_devices_ready(ctx);
  mcgpu_dev_reducer_free(ctx->reducer);
  for (i = 0; i < ctx->num_devices; i++) mcgpu_dev_close(ctx->dev[i]);
  free(ctx->dev);
  free(ctx->views);
  mcgpu_free_volume(&ctx->vol);
  mcgpu_free_tables(&ctx->tab);
  mcgpu_free_scene(&ctx->scene);
  free(ctx);
}

const char* mcgpu_last_error(const mcgpu_ctx* ctx) { return ctx ? ctx->err : "null context"; }
void mcgpu_set_verbose(mcgpu_ctx* ctx, int verbose) {
  if (ctx) ctx->verbose = verbose;
}

int mcgpu_load_input(mcgpu_ctx* ctx, const char* in_path) {
  int rc;
  if (!ctx || !in_path) return MCGPU_E_ARG;
  free(ctx->views);
  ctx->views = NULL;
  ctx->have_input = 0;
  /* the devices hold the previous input's spectrum and an image sized for its detector: a new input needs
   * mcgpu_load_materials again (which re-uploads everything) before anything can run */
  ctx->have_tables = 0;
  if ((rc = mcgpu_parse_input(ctx, in_path)) != MCGPU_OK) return rc;
  if ((rc = mcgpu_read_spectrum(ctx, ctx->in.file_espc)) != MCGPU_OK) return rc;
  if ((rc = mcgpu_build_views(ctx)) != MCGPU_OK) return rc;
  ctx->hpt_current = ctx->in.histories_per_thread;
  ctx->hist_current = 0;
  ctx->have_input = 1;
  return MCGPU_OK;
}

int mcgpu_load_voxels(mcgpu_ctx* ctx, const char* vox_path) {
  if (!ctx) return MCGPU_E_ARG;
  if (!vox_path) {
    if (!ctx->have_input) return mcgpu_fail(ctx, MCGPU_E_STATE, "load_voxels: no path given and no input file loaded");
    vox_path = ctx->in.file_voxels;
  }
  return mcgpu_read_voxels(ctx, vox_path);
}

int mcgpu_load_materials(mcgpu_ctx* ctx, const char* const* paths, int n_paths) {
  const char* from_input[MCGPU_MAX_MATERIALS];
  int rc, i;
  if (!ctx) return MCGPU_E_ARG;
  if (!ctx->have_voxels) return mcgpu_fail(ctx, MCGPU_E_STATE, "load_material: the voxels must be loaded first (their densities set the Woodcock majorant)");
  if (!paths) {
    if (!ctx->have_input) return mcgpu_fail(ctx, MCGPU_E_STATE, "load_material: no paths given and no input file loaded");
    for (i = 0; i < MCGPU_MAX_MATERIALS; i++) from_input[i] = ctx->in.file_materials[i];
    paths = from_input;
    n_paths = MCGPU_MAX_MATERIALS;
  }
  if ((rc = mcgpu_read_materials(ctx, paths, n_paths)) != MCGPU_OK) return rc;
  if ((rc = mcgpu_build_scene(ctx)) != MCGPU_OK) return rc;
  ctx->have_tables = 1;
  mcgpu_devices_ready(ctx);
  if (ctx->have_input && ctx->num_devices > 0) { /* every device gets its own copy, uploaded concurrently (H:2612-2690 does one device per MPI rank) */
    upload_job* jobs = (upload_job*)calloc((size_t)ctx->num_devices, sizeof *jobs);
    if (!jobs) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "load_material: out of memory");
    rc = MCGPU_OK;
    for (i = 0; i < ctx->num_devices; i++) {
      jobs[i].ctx = ctx, jobs[i].device = i;
      jobs[i].threaded = ctx->num_devices > 1 && pthread_create(&jobs[i].thread, NULL, upload_main, &jobs[i]) == 0;
    }
    for (i = 0; i < ctx->num_devices; i++) {
      if (jobs[i].threaded)
        pthread_join(jobs[i].thread, NULL);
      else
        upload_main(&jobs[i]);
      if (jobs[i].rc != 0) {
        snprintf(ctx->err, sizeof ctx->err, "%s", jobs[i].err);
        rc = MCGPU_E_CUDA;
      }
    }
    free(jobs);
    if (rc != MCGPU_OK) {
      ctx->have_tables = 0;
      return rc;
    }
  }
  return MCGPU_OK;
}

int mcgpu_set_histories(mcgpu_ctx* ctx, unsigned long long total_histories) {
  if (!ctx || !ctx->have_input) return MCGPU_E_STATE;
  ctx->in.total_histories = total_histories;
  ctx->hpt_current = ctx->in.histories_per_thread;
  ctx->hist_current = 0;
  return MCGPU_OK;
}

int mcgpu_set_fast_math(mcgpu_ctx* ctx, int on) {
  int d;
  if (!ctx) return MCGPU_E_ARG;
  ctx->fast_math = on != 0;
  mcgpu_devices_ready(ctx);
  for (d = 0; d < ctx->num_devices; d++) mcgpu_dev_set_fast_math(ctx->dev[d], ctx->fast_math);
  return MCGPU_OK;
}

int mcgpu_set_seed(mcgpu_ctx* ctx, int seed) {
  if (!ctx || !ctx->have_input) return MCGPU_E_STATE;
  ctx->in.seed_input = seed;
  return MCGPU_OK;
}

static int projection_skipped(const mcgpu_ctx* ctx, int p) { /* H:671-677, Q6 */
  const double a = ctx->in.initial_angle + p * ctx->in.D_angle;
  return (a < ctx->in.angularROI_0) || (a > ctx->in.angularROI_1);
}

/* Seed, histories/thread and grid the reference's loop reaches projection p with (Q1).  Both inputs of the grid rule are
 * sticky across projections: histories_per_thread (H:833) and the history count itself, which the reference overwrites
 * with the launched count (H:841) -- so after a >65535-block correction every later projection runs 65000 blocks too. */
static void schedule_for(const mcgpu_ctx* ctx, int p, int* seed, int* hpt, int* blocks, unsigned long long* launched) {
  unsigned long long hist = ctx->in.total_histories;
  int q;
  *seed = ctx->in.seed_input;
  *hpt = ctx->in.histories_per_thread;
  for (q = 0;; q++) {
    if (q < p && projection_skipped(ctx, q)) continue;
    mcgpu_grid_rule(hist, ctx->in.threads_per_block, hpt, blocks, launched);
    hist = *launched;
    if (q >= p) break;
    *seed = mcgpu_ranecu_advance_projection_seed(*seed, *launched);
  }
}

void mcgpu_current_grid(const mcgpu_ctx* ctx, int* hpt, int* blocks, unsigned long long* launched) {
  *hpt = ctx->hpt_current ? ctx->hpt_current : ctx->in.histories_per_thread;
  mcgpu_grid_rule(ctx->hist_current ? ctx->hist_current : ctx->in.total_histories, ctx->in.threads_per_block, hpt, blocks, launched);
}

int mcgpu_projection_seed(mcgpu_ctx* ctx, int p, int* seed_out) {
  int hpt, blocks;
  unsigned long long launched;
  if (!ctx || !ctx->have_input || p < 0 || p >= ctx->in.num_projections || !seed_out) return MCGPU_E_ARG;
  schedule_for(ctx, p, seed_out, &hpt, &blocks, &launched);
  return MCGPU_OK;
}

static int ready_to_run(mcgpu_ctx* ctx, int p) {
  if (!ctx) return MCGPU_E_ARG;
  mcgpu_devices_ready(ctx);
  if (!ctx->have_input || !ctx->have_voxels || !ctx->have_tables) return mcgpu_fail(ctx, MCGPU_E_STATE, "run: input, voxels and materials must be loaded first");
  if (ctx->num_devices < 1) return mcgpu_fail(ctx, MCGPU_E_CUDA, "run: no usable CUDA device (there is no CPU fallback)");
  if (p < 0 || p >= ctx->in.num_projections) return mcgpu_fail(ctx, MCGPU_E_ARG, "run: projection %d out of range [0,%d)", p, ctx->in.num_projections);
  return MCGPU_OK;
}

/* launch the streams [b,e) of projection p on device d (asynchronous) */
static int launch_on(mcgpu_ctx* ctx, int d, int p, long long b, long long e, int seed, int hpt) {
  mcgpu_launch l;
  l.histories_per_thread = hpt;
  l.seed_input = seed;
  l.threads_per_block = ctx->in.threads_per_block;
  l.stream_begin = b;
  l.stream_end = e;
  l.zero_image = 1;
  l.image_slot = 0;
  return mcgpu_dev_launch(ctx->dev[d], &ctx->views[p], &l, ctx->err, sizeof ctx->err) == 0 ? MCGPU_OK : MCGPU_E_CUDA;
}

int mcgpu_run_streams(mcgpu_ctx* ctx, int p, long long stream_begin, long long stream_end, uint64_t* image_host) {
  int rc, seed, hpt, blocks;
  unsigned long long launched;
  float ms = 0.f;
  if ((rc = ready_to_run(ctx, p)) != MCGPU_OK) return rc;
  schedule_for(ctx, p, &seed, &hpt, &blocks, &launched);
  ctx->hpt_current = hpt;
  ctx->hist_current = launched;
  if (stream_begin < 0 || stream_end > (long long)blocks * ctx->in.threads_per_block || stream_begin > stream_end)
    return mcgpu_fail(ctx, MCGPU_E_ARG, "run_streams: stream range [%lld,%lld) outside [0,%lld)", stream_begin, stream_end, (long long)blocks * ctx->in.threads_per_block);
  if ((rc = launch_on(ctx, 0, p, stream_begin, stream_end, seed, hpt)) != MCGPU_OK) return rc;
  if (mcgpu_dev_sync(ctx->dev[0], &ms, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
  ctx->last_kernel_ms = ms;
  if (image_host && mcgpu_dev_fetch(ctx->dev[0], image_host, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
  return MCGPU_OK;
}

int mcgpu_run_projection(mcgpu_ctx* ctx, int p, uint64_t* image_host) {
  int rc, seed, hpt, blocks, d, n;
  unsigned long long launched;
  float ms_max = 0.f;
  if ((rc = ready_to_run(ctx, p)) != MCGPU_OK) return rc;
  schedule_for(ctx, p, &seed, &hpt, &blocks, &launched);
  ctx->hpt_current = hpt;
  ctx->hist_current = launched;
  n = ctx->num_devices;
  if (n > blocks) n = blocks;
  /* history split: contiguous block ranges of the reference grid, one per device */
  for (d = 0; d < n; d++) {
    const long long b = (long long)blocks * d / n, e = (long long)blocks * (d + 1) / n;
    if ((rc = launch_on(ctx, d, p, b * ctx->in.threads_per_block, e * ctx->in.threads_per_block, seed, hpt)) != MCGPU_OK) return rc;
  }
  for (d = 0; d < n; d++) {
    float ms = 0.f;
    if (mcgpu_dev_sync(ctx->dev[d], &ms, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
    if (ms > ms_max) ms_max = ms;
  }
  ctx->last_kernel_ms = ms_max;
  ctx->last_reduce_ms = 0.0;
  if (n > 1) { /* integer tallies: summing the partial images in any order is bit-identical to one device (the reference: MPI_Reduce, H:1019) */
    float rms = 0.f;
    if (ctx->reducer && ctx->reducer_devices != n) {
      mcgpu_dev_reducer_free(ctx->reducer);
      ctx->reducer = NULL;
    }
    if (!ctx->reducer) {
      ctx->reducer = mcgpu_dev_reducer_create(ctx->dev, n, ctx->err, sizeof ctx->err);
      ctx->reducer_devices = n;
      if (!ctx->reducer) return MCGPU_E_CUDA;
    }
    if (mcgpu_dev_reduce(ctx->reducer, &rms, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
    ctx->last_reduce_ms = rms;
  }
  if (image_host && mcgpu_dev_fetch(ctx->dev[0], image_host, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
  return MCGPU_OK;
}

void* mcgpu_device_image(mcgpu_ctx* ctx) {
  mcgpu_devices_ready(ctx);
  return (ctx && ctx->num_devices > 0) ? mcgpu_dev_image_ptr(ctx->dev[0]) : NULL;
}
double mcgpu_last_kernel_ms(const mcgpu_ctx* ctx) { return ctx ? ctx->last_kernel_ms : 0.0; }
double mcgpu_last_reduce_ms(const mcgpu_ctx* ctx) { return ctx ? ctx->last_reduce_ms : 0.0; }
const char* mcgpu_reduce_kind(const mcgpu_ctx* ctx) { return ctx ? mcgpu_dev_reducer_kind(ctx->reducer) : "none"; }
int mcgpu_device_selftest(mcgpu_ctx* ctx, const char* name, unsigned long long* mismatches) {
  if (!ctx || !name || !mismatches) return MCGPU_E_ARG;
  mcgpu_devices_ready(ctx);
  if (ctx->num_devices < 1) return mcgpu_fail(ctx, MCGPU_E_CUDA, "selftest: no usable CUDA device");
  return mcgpu_dev_selftest(ctx->dev[0], name, mismatches, ctx->err, sizeof ctx->err) == 0 ? MCGPU_OK : MCGPU_E_CUDA;
}

int mcgpu_get_scan_stats(const mcgpu_ctx* ctx, double* out, int n) {
  int k;
  if (!ctx || !out) return MCGPU_E_ARG;
  for (k = 0; k < n && k < MCGPU_SCAN_STATS; k++) out[k] = ctx->scan_stats[k];
  return k;
}

/* ---- whole scan: projections dealt round-robin to the devices, one host thread per device ---- */

typedef struct scan_shared {
  mcgpu_ctx* ctx;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int* done;       /* per projection: 0 pending, 1 done, 2 skipped, <0 error */
  double* seconds; /* per projection */
  int hpt, blocks, write_raw;
  int* seeds;
} scan_shared;

typedef struct scan_worker {
  scan_shared* sh;
  int device, started;
  /* where this device's host thread spent the scan (mcgpu_get_scan_stats) */
  double kernel_ms, t_wait, t_report;
  int projections;
  char err[512];
} scan_worker;

static void scan_publish(scan_shared* sh, int p, int status, double dt) {
  pthread_mutex_lock(&sh->mu);
  sh->done[p] = status;
  sh->seconds[p] = dt;
  pthread_cond_broadcast(&sh->cv);
  pthread_mutex_unlock(&sh->mu);
}

/* One host thread per device.  The projection loop is software-pipelined three deep: while the GPU transports
 * projection p into one of its two images, the other image (projection p-1) is copied to a pinned host buffer on a second
 * stream and this thread formats and writes its ASCII file (63 MB of text, ~0.1 s), so neither the copy nor the report
 * costs GPU time (the reference serialises kernel, copy and fprintf, H:861-1040). */
static int scan_report(scan_worker* w, int p, const uint64_t* image, double dt) {
  scan_shared* sh = w->sh;
  mcgpu_ctx* ctx = sh->ctx;
  const double t0 = now_s();
  int rc;
  mcgpu_fail_into(w->err, sizeof w->err); /* the writers run on this worker's thread: their messages must not race on ctx->err */
  rc = mcgpu_write_projection_ascii(ctx, p, image, dt);
  if (rc == MCGPU_OK && sh->write_raw) rc = mcgpu_write_projection_raw(ctx, p, image);
  mcgpu_fail_into(NULL, 0);
  w->t_report += now_s() - t0;
  scan_publish(sh, p, rc == MCGPU_OK ? 1 : rc, dt);
  return rc;
}

static void* scan_thread(void* arg) {
  scan_worker* w = (scan_worker*)arg;
  scan_shared* sh = w->sh;
  mcgpu_ctx* ctx = sh->ctx;
  struct mcgpu_device* dev = ctx->dev[w->device];
  const int P = ctx->in.num_projections, n = ctx->num_devices;
  int p, slot = 0, prev = -1, prev_slot = 0, failed = 0;
  double prev_t0 = 0.0;
  if (mcgpu_dev_pipeline_begin(dev, w->err, sizeof w->err) != 0) {
    for (p = w->device; p < P; p += n) scan_publish(sh, p, MCGPU_E_CUDA, 0.0);
    return NULL;
  }
  for (p = w->device; p < P && !failed; p += n) {
    const double t0 = now_s();
    int launched = 0;
    if (projection_skipped(ctx, p)) {
      scan_publish(sh, p, 2, 0.0);
    } else {
      mcgpu_launch l;
      l.histories_per_thread = sh->hpt;
      l.seed_input = sh->seeds[p];
      l.threads_per_block = ctx->in.threads_per_block;
      l.stream_begin = 0;
      l.stream_end = (long long)sh->blocks * ctx->in.threads_per_block;
      l.zero_image = 1;
      l.image_slot = slot;
      if (mcgpu_dev_pipeline_launch(dev, &ctx->views[p], &l, w->err, sizeof w->err) != 0) {
        scan_publish(sh, p, MCGPU_E_CUDA, 0.0);
        failed = 1;
      } else
        launched = 1;
    }
    if (prev >= 0) { /* the previous projection: wait for its copy, report it while the kernel just launched runs */
      uint64_t* host = NULL;
      float ms = 0.f;
      const double tw = now_s();
      if (mcgpu_dev_pipeline_wait(dev, prev_slot, &ms, &host, w->err, sizeof w->err) != 0) {
        scan_publish(sh, prev, MCGPU_E_CUDA, 0.0);
        failed = 1;
      } else {
        w->t_wait += now_s() - tw;
        w->kernel_ms += ms;
        w->projections++;
        if (scan_report(w, prev, host, now_s() - prev_t0) != MCGPU_OK) failed = 1;
      }
      prev = -1;
    }
    if (launched) {
      prev = p, prev_slot = slot, prev_t0 = t0;
      slot ^= 1;
    }
  }
  if (prev >= 0) {
    uint64_t* host = NULL;
    float ms = 0.f;
    const double tw = now_s();
    if (mcgpu_dev_pipeline_wait(dev, prev_slot, &ms, &host, w->err, sizeof w->err) != 0)
      scan_publish(sh, prev, MCGPU_E_CUDA, 0.0);
    else {
      w->t_wait += now_s() - tw;
      w->kernel_ms += ms;
      w->projections++;
      scan_report(w, prev, host, now_s() - prev_t0);
    }
  }
  mcgpu_dev_pipeline_end(dev);
  return NULL;
}

int mcgpu_run_all(mcgpu_ctx* ctx, mcgpu_progress_cb cb, void* user) {
  int rc, P, p, n;
  if ((rc = ready_to_run(ctx, 0)) != MCGPU_OK) return rc;
  P = ctx->in.num_projections;
  n = ctx->num_devices;
  if ((rc = mcgpu_reset_dose(ctx)) != MCGPU_OK) return rc;

  if (P < n) { /* fewer projections than devices: split each projection's histories instead */
    const size_t words = (size_t)4 * ctx->views[0].total_num_pixels;
    uint64_t* image = (uint64_t*)malloc(words * sizeof(uint64_t));
    if (!image) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "run_all: out of memory for the host image");
    for (p = 0; p < P; p++) {
      double t0 = now_s(), dt;
      if (projection_skipped(ctx, p)) continue;
      if (cb) cb(p, P, -1.0, user); /* "Simulating Projection" marker before the work, like H:680 */
      if ((rc = mcgpu_run_projection(ctx, p, image)) != MCGPU_OK) break;
      dt = now_s() - t0;
      if ((rc = mcgpu_write_projection_ascii(ctx, p, image, dt)) != MCGPU_OK) break;
      if (getenv("MCGPU_WRITE_RAW") && atoi(getenv("MCGPU_WRITE_RAW")) != 0 && (rc = mcgpu_write_projection_raw(ctx, p, image)) != MCGPU_OK) break;
      if (cb) cb(p, P, dt, user);
    }
    free(image);
    if (rc == MCGPU_OK) rc = mcgpu_write_dose_reports(ctx, 0.0, P);
    return rc;
  }

  {
    scan_shared sh;
    scan_worker* workers = (scan_worker*)calloc((size_t)n, sizeof *workers);
    pthread_t* threads = (pthread_t*)calloc((size_t)n, sizeof *threads);
    int d, seed, hpt, blocks, verbose = ctx->verbose;
    unsigned long long launched;
    const double t_scan0 = now_s();
    memset(&sh, 0, sizeof sh);
    sh.ctx = ctx;
    sh.write_raw = getenv("MCGPU_WRITE_RAW") && atoi(getenv("MCGPU_WRITE_RAW")) != 0;
    sh.done = (int*)calloc((size_t)P, sizeof(int));
    sh.seconds = (double*)calloc((size_t)P, sizeof(double));
    sh.seeds = (int*)calloc((size_t)P, sizeof(int));
    if (!workers || !threads || !sh.done || !sh.seconds || !sh.seeds) {
      free(workers), free(threads), free(sh.done), free(sh.seconds), free(sh.seeds);
      return mcgpu_fail(ctx, MCGPU_E_NOMEM, "run_all: out of memory");
    }
    /* closed-form seed schedule (Q1), computed once for the whole scan */
    seed = ctx->in.seed_input;
    hpt = ctx->in.histories_per_thread;
    blocks = 1;
    launched = ctx->in.total_histories;
    for (p = 0; p < P; p++) {
      if (projection_skipped(ctx, p)) continue;
      mcgpu_grid_rule(launched, ctx->in.threads_per_block, &hpt, &blocks, &launched); /* sticky hpt (H:833) and history count (H:841) */
      sh.seeds[p] = seed;
      seed = mcgpu_ranecu_advance_projection_seed(seed, launched);
    }
    /* hpt and the history count only change at the first simulated projection (the rule is idempotent afterwards),
     * so every projection of the scan runs the same grid */
    sh.hpt = hpt;
    sh.blocks = blocks;
    ctx->hpt_current = hpt;
    ctx->hist_current = launched;
    pthread_mutex_init(&sh.mu, NULL);
    pthread_cond_init(&sh.cv, NULL);
    ctx->verbose = 0; /* per-projection report banners would interleave across workers */
    for (d = 0; d < n; d++) {
      workers[d].sh = &sh;
      workers[d].device = d;
      workers[d].started = pthread_create(&threads[d], NULL, scan_thread, &workers[d]) == 0;
      if (!workers[d].started) { /* e.g. the container's thread limit: fail this device's projections instead of waiting for them forever */
        snprintf(workers[d].err, sizeof workers[d].err, "run_all: cannot start the host thread of device %d", d);
        for (p = d; p < P; p += n) scan_publish(&sh, p, MCGPU_E_NOMEM, 0.0);
      }
    }
    rc = MCGPU_OK;
    for (p = 0; p < P && rc == MCGPU_OK; p++) {
      int st;
      pthread_mutex_lock(&sh.mu);
      while (sh.done[p] == 0) {
        /* a failed worker stops early: do not wait forever for its later projections */
        int failed = 0, q;
        for (q = 0; q < P; q++) failed |= sh.done[q] < 0;
        if (failed) break;
        pthread_cond_wait(&sh.cv, &sh.mu);
      }
      st = sh.done[p];
      pthread_mutex_unlock(&sh.mu);
      if (st == 1 && cb) {
        cb(p, P, -1.0, user);
        cb(p, P, sh.seconds[p], user);
      } else if (st <= 0) {
        int q;
        rc = MCGPU_E_CUDA;
        for (q = 0; q < P; q++)
          if (sh.done[q] < 0) rc = sh.done[q];
      }
    }
    for (d = 0; d < n; d++)
      if (workers[d].started) pthread_join(threads[d], NULL);
    ctx->verbose = verbose;
    memset(ctx->scan_stats, 0, sizeof ctx->scan_stats);
    ctx->scan_stats[0] = now_s() - t_scan0;
    for (d = 0; d < n; d++) {
      ctx->scan_stats[1] += 1e-3 * workers[d].kernel_ms;
      ctx->scan_stats[2] += workers[d].t_wait;
      ctx->scan_stats[3] += workers[d].t_report;
      ctx->scan_stats[4] += workers[d].projections;
    }
    ctx->scan_stats[5] = n;
    if (rc != MCGPU_OK)
      for (d = 0; d < n; d++)
        if (workers[d].err[0]) snprintf(ctx->err, sizeof ctx->err, "%s", workers[d].err);
    pthread_mutex_destroy(&sh.mu);
    pthread_cond_destroy(&sh.cv);
    free(workers), free(threads), free(sh.done), free(sh.seconds), free(sh.seeds);
    if (rc == MCGPU_OK) rc = mcgpu_write_dose_reports(ctx, 0.0, P);
    return rc;
  }
}

int mcgpu_get_info(const mcgpu_ctx* ctx, mcgpu_info* out) {
  int hpt, blocks = 0;
  unsigned long long launched = 0;
  if (!ctx || !out) return MCGPU_E_ARG;
  mcgpu_devices_ready((mcgpu_ctx*)ctx); /* num_devices counts the devices that really opened */
  memset(out, 0, sizeof *out);
  out->num_devices = ctx->num_devices;
  out->fast_math = ctx->fast_math;
  if (ctx->have_input) {
    const mcgpu_view* v = &ctx->views[0];
    mcgpu_current_grid(ctx, &hpt, &blocks, &launched);
    out->num_projections = ctx->in.num_projections;
    out->num_pixels_x = v->num_pixels_x;
    out->num_pixels_z = v->num_pixels_z;
    out->num_spectrum_bins = ctx->spc.num_bins;
    out->threads_per_block = ctx->in.threads_per_block;
    out->histories_per_thread = hpt;
    out->num_blocks = blocks;
    out->seed_input = ctx->in.seed_input;
    out->enable_specific_angles = ctx->in.enable_specific_angles;
    out->requested_histories = ctx->in.total_histories;
    out->launched_histories = launched;
    out->mean_energy_spectrum = ctx->spc.mean_energy;
  }
  if (ctx->have_voxels) {
    out->num_voxels_x = ctx->vol.nx;
    out->num_voxels_y = ctx->vol.ny;
    out->num_voxels_z = ctx->vol.nz;
    out->voxel_bits = ctx->vol.voxel_bits;
    out->palette_size = ctx->vol.palette_size;
  }
  if (ctx->have_tables) {
    out->num_materials_used = ctx->scene.num_slots;
    out->num_energy_values = ctx->tab.num_values;
    out->e0 = ctx->tab.e0;
    out->ide = ctx->tab.ide;
  }
  return MCGPU_OK;
}

long long mcgpu_copy_table(const mcgpu_ctx* ctx, const char* name, void* out, size_t cap) {
  const void* src = NULL;
  size_t bytes = 0;
  const mcgpu_tables* t;
  const size_t NR = (size_t)MCGPU_NP_RAYLEIGH * MCGPU_MAX_MATERIALS, NC = (size_t)MCGPU_MAX_MATERIALS * MCGPU_MAX_SHELLS;
  if (!ctx || !name) return MCGPU_E_ARG;
  t = &ctx->tab;
#define TAB(nm, ptr, nbytes, cond) \
  if (!strcmp(name, nm)) {         \
    if (!(cond)) return MCGPU_E_STATE; \
    src = (ptr);                   \
    bytes = (nbytes);              \
  }
  TAB("woodcock", t->woodcock, sizeof(mcgpu_f2) * t->num_values, ctx->have_tables)
  TAB("mfp_a", t->mfp_a, sizeof(mcgpu_f3) * t->num_values * MCGPU_MAX_MATERIALS, ctx->have_tables)
  TAB("mfp_b", t->mfp_b, sizeof(mcgpu_f3) * t->num_values * MCGPU_MAX_MATERIALS, ctx->have_tables)
  TAB("rayleigh_xco", t->ray_xco, 4 * NR, ctx->have_tables)
  TAB("rayleigh_pco", t->ray_pco, 4 * NR, ctx->have_tables)
  TAB("rayleigh_aco", t->ray_aco, 4 * NR, ctx->have_tables)
  TAB("rayleigh_bco", t->ray_bco, 4 * NR, ctx->have_tables)
  TAB("rayleigh_itlco", t->ray_itlco, NR, ctx->have_tables)
  TAB("rayleigh_ituco", t->ray_ituco, NR, ctx->have_tables)
  TAB("rayleigh_pmax", t->ray_pmax, (size_t)4 * t->num_values * MCGPU_MAX_MATERIALS, ctx->have_tables)
  TAB("compton_fco", t->cmp_fco, 4 * NC, ctx->have_tables)
  TAB("compton_uico", t->cmp_uico, 4 * NC, ctx->have_tables)
  TAB("compton_fj0", t->cmp_fj0, 4 * NC, ctx->have_tables)
  TAB("compton_noscco", t->cmp_noscco, sizeof(int) * MCGPU_MAX_MATERIALS, ctx->have_tables)
  TAB("density_nominal", t->density_nominal, sizeof(float) * MCGPU_MAX_MATERIALS, ctx->have_tables)
  TAB("espc", ctx->spc.espc, sizeof ctx->spc.espc, ctx->have_input)
  TAB("espc_cutoff", ctx->spc.cutoff, sizeof ctx->spc.cutoff, ctx->have_input)
  TAB("espc_alias", ctx->spc.alias, sizeof ctx->spc.alias, ctx->have_input)
  TAB("views", ctx->views, sizeof(mcgpu_view) * (size_t)ctx->in.num_projections, ctx->have_input)
  TAB("density_max", ctx->vol.density_max, sizeof ctx->vol.density_max, ctx->have_voxels)
  TAB("voxel_material", ctx->vol.material, (size_t)ctx->vol.nx * ctx->vol.ny * ctx->vol.nz, ctx->have_voxels)
  TAB("voxel_density", ctx->vol.density, (size_t)4 * ctx->vol.nx * ctx->vol.ny * ctx->vol.nz, ctx->have_voxels)
  TAB("voxel_packed", ctx->vol.packed, ctx->vol.packed_bytes, ctx->have_voxels)
#undef TAB
  if (!src) return MCGPU_E_ARG;
  if (!out) return (long long)bytes;
  if (bytes > cap) bytes = cap;
  memcpy(out, src, bytes);
  return (long long)bytes;
}

/* ---- dose tallies (optional; K:357-369, reports H:2976-3263) ------------------------------ */

int mcgpu_reset_dose(mcgpu_ctx* ctx) {
  int d;
  if (!ctx || !ctx->have_tables) return MCGPU_E_STATE;
  mcgpu_devices_ready(ctx);
  for (d = 0; d < ctx->num_devices; d++)
    if (mcgpu_dev_reset_dose(ctx->dev[d], ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
  return MCGPU_OK;
}

long long mcgpu_get_dose(mcgpu_ctx* ctx, const char* which, uint64_t* out, size_t cap_words) {
  const int materials = which && !strcmp(which, "materials");
  size_t words;
  int d;
  if (!ctx || !ctx->have_tables || !which || (!materials && strcmp(which, "voxels"))) return MCGPU_E_ARG;
  if (materials ? !ctx->scene.tally_material_dose : !ctx->scene.tally_voxel_dose) return 0;
  words = materials ? (size_t)2 * MCGPU_MAX_MATERIALS : (size_t)2 * (size_t)ctx->scene.dose_roi_voxels;
  if (!out) return (long long)words;
  if (cap_words < words) return mcgpu_fail(ctx, MCGPU_E_ARG, "get_dose: buffer of %zu words, %zu needed", cap_words, words);
  memset(out, 0, words * sizeof(uint64_t));
  mcgpu_devices_ready(ctx);
  for (d = 0; d < ctx->num_devices; d++)
    if (mcgpu_dev_add_dose(ctx->dev[d], materials ? out : NULL, materials ? NULL : out, ctx->err, sizeof ctx->err) != 0) return MCGPU_E_CUDA;
  return (long long)words;
}

int mcgpu_write_dose_reports(mcgpu_ctx* ctx, double seconds, int projections_simulated) {
  int rc = MCGPU_OK;
  if (!ctx || !ctx->have_tables) return MCGPU_E_STATE;
  if (projections_simulated < 1) projections_simulated = ctx->in.num_projections;
  if (ctx->scene.tally_voxel_dose) {
    const long long words = mcgpu_get_dose(ctx, "voxels", NULL, 0);
    uint64_t* buf = (uint64_t*)malloc((size_t)words * sizeof(uint64_t));
    if (!buf) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "dose report: out of memory");
    if (mcgpu_get_dose(ctx, "voxels", buf, (size_t)words) < 0)
      rc = MCGPU_E_CUDA;
    else
      rc = mcgpu_write_dose_files(ctx, buf, seconds, projections_simulated);
    free(buf);
  }
  if (rc == MCGPU_OK && ctx->scene.tally_material_dose) {
    uint64_t md[2 * MCGPU_MAX_MATERIALS];
    if (mcgpu_get_dose(ctx, "materials", md, 2 * MCGPU_MAX_MATERIALS) < 0)
      rc = MCGPU_E_CUDA;
    else
      rc = mcgpu_print_materials_dose(ctx, md, projections_simulated);
  }
  return rc;
}
