/* Energy spectrum (.spc) and material (.mcgpu[.gz]) tables.
 *
 * Table CONTENT restates, operation by operation and precision by precision, what the
 * reference builds on the host -- init_energy_spectrum + IRND0 (docker/mcgpu/MC-GPU_v1.3.cu:
 * 3498-3587, 3675-3734) and load_material (H:2177-2443) -- because these floats are the
 * kernel's constants.  Table LAYOUT on the device is our own (mcgpu_build_scene): compacted to
 * the materials present, one 32-byte record per (energy bin, material). */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include "mcgpu_host.h"

/* ------------------------------------------------------------------------------ spectrum */

/* Walker alias set-up as PENELOPE's IRND0 does it (H:3675-3734): repeatedly pair the lowest
 * under-full bin with the highest over-full one. */
static void walker_alias(const float* w, float* cutoff, short* alias, int n) {
  double ws = 0.0;
  int i, j;
  for (i = 0; i < n; i++) ws = ws + w[i];
  ws = ((double)n) / ws;
  for (i = 0; i < n; i++) {
    alias[i] = (short)i;
    cutoff[i] = w[i] * ws;
  }
  if (n == 1) return;
  for (i = 0; i < n - 1; i++) {
    float low = 1.0f, high = 1.0f;
    int ilow = -1, ihigh = -1;
    for (j = 0; j < n; j++) {
      if (alias[j] != j) continue;
      if (cutoff[j] < low) {
        low = cutoff[j];
        ilow = j;
      } else if (cutoff[j] > high) {
        high = cutoff[j];
        ihigh = j;
      }
    }
    if (ilow == -1 || ihigh == -1) return;
    alias[ilow] = (short)ihigh;
    cutoff[ihigh] = high + low - 1.0f;
  }
}

int mcgpu_read_spectrum(mcgpu_ctx* ctx, const char* path) {
  mcgpu_spectrum* s = &ctx->spc;
  float prob_bin[MCGPU_MAX_ENERGY_BINS];
  float e_low = 0.f, prob, all_energy = 0.0f, all_prob = 0.0f;
  char line[MCGPU_LINE];
  int bin = -1, i;
  FILE* f = fopen(path, "r");
  if (!f) return mcgpu_fail(ctx, MCGPU_E_ARG, "init_energy_spectrum: can not read the energy spectrum input file \"%s\"", path);
  memset(s, 0, sizeof *s);
  do {
    bin++;
    if (bin >= MCGPU_MAX_ENERGY_BINS) {
      fclose(f);
      return mcgpu_fail(ctx, MCGPU_E_ARG, "init_energy_spectrum: too many energy bins in the input spectrum (max %d); a negative probability marks the end", MCGPU_MAX_ENERGY_BINS);
    }
    if (!mcgpu_fgets_trimmed(line, MCGPU_LINE, f)) {
      fclose(f);
      return mcgpu_fail(ctx, MCGPU_E_ARG, "init_energy_spectrum: the spectrum file (%s) is incomplete (a negative probability marks the end)", path);
    }
    prob = -123456789.0f;
    sscanf(line, "%f %f", &e_low, &prob);
    prob_bin[bin] = prob;
    s->espc[bin] = e_low;
    if (prob == -123456789.0f) {
      fclose(f);
      return mcgpu_fail(ctx, MCGPU_E_ARG, "init_energy_spectrum: invalid energy bin number %d", bin);
    }
    if (e_low < s->espc[bin > 0 ? bin - 1 : 0]) {
      fclose(f);
      return mcgpu_fail(ctx, MCGPU_E_ARG, "init_energy_spectrum: input energy bins with decreasing energy at bin %d", bin);
    }
  } while (prob > -1.0e-11f);
  fclose(f);
  s->num_bins = bin;
  for (i = bin; i < MCGPU_MAX_ENERGY_BINS; i++) {
    s->espc[i] = e_low;
    prob_bin[i] = 0.0f;
  }
  for (i = 0; i < s->num_bins; i++) {
    all_energy += 0.5f * (s->espc[i] + s->espc[i + 1]) * prob_bin[i];
    all_prob += prob_bin[i];
  }
  s->mean_energy = all_energy / all_prob;
  for (i = 0; i < s->num_bins; i++)
    if (prob_bin[i] < 0.0f) return mcgpu_fail(ctx, MCGPU_E_ARG, "IRND0: negative point probability W(%d)=%f", i, prob_bin[i]);
  walker_alias(prob_bin, s->cutoff, s->alias, s->num_bins);
  return MCGPU_OK;
}

/* ------------------------------------------------------------------------------ materials */

void mcgpu_free_tables(mcgpu_tables* t) {
  free(t->woodcock), free(t->mfp_a), free(t->mfp_b);
  free(t->ray_xco), free(t->ray_pco), free(t->ray_aco), free(t->ray_bco);
  free(t->ray_itlco), free(t->ray_ituco), free(t->ray_pmax);
  memset(t, 0, sizeof *t);
}

static int seek_gz(gzFile f, const char* tag, char* line) {
  do {
    if (!gzgets(f, line, MCGPU_LINE)) return 0;
  } while (!strstr(line, tag));
  return 1;
}

/* One material file.  Everything it writes is private to the material (its column of the tables, its
 * Woodcock candidates in `wood`, its energy step), so the files are parsed concurrently; material 1 goes
 * first on its own because it allocates the tables and defines the energy grid (H:2240-2285). */
typedef struct {
  mcgpu_tables* t;
  float* density_max;
  const char* path;
  int mat;
  float* wood;     /* [num_values] total MFP * rho_nominal / rho_max of this material, or NULL when not loaded */
  double delta_e;  /* energy step read from this file */
  int loaded, rc;
  char err[MCGPU_LINE + 160];
} material_job;

#define MFAIL(code, ...)                      \
  do {                                        \
    if (f) gzclose(f);                        \
    snprintf(j->err, sizeof j->err, __VA_ARGS__); \
    j->rc = (code);                           \
    return;                                   \
  } while (0)

static void load_one_material(material_job* j) {
  mcgpu_tables* t = j->t;
  const int mat = j->mat;
  const size_t NR = (size_t)MCGPU_NP_RAYLEIGH * MCGPU_MAX_MATERIALS;
  char line[MCGPU_LINE];
  int n_values = 0, n_rayleigh = 0, n_shells = 0, i;
  double e_last = -1.0, delta_e = -99999.0;
  gzFile f = gzopen(j->path, "rb");
  j->rc = MCGPU_OK;
  j->loaded = 0;
  if (!f) MFAIL(MCGPU_E_PARSE, "load_material: file %d '%s' does not exist", mat, j->path);
  gzbuffer(f, 1 << 18);
  if (!seek_gz(f, "[NOMINAL DENSITY", line)) MFAIL(MCGPU_E_PARSE, "load_material: '%s' does not contain the string '[NOMINAL DENSITY'", j->path);
  gzgets(f, line, MCGPU_LINE);
  sscanf(line, "# %f", &t->density_nominal[mat]);

  if (!(j->density_max[mat] > 0)) { /* not in the voxels: only material 1 is read in full (H:2224-2233) */
    if (mat == 0)
      j->density_max[mat] = 0.01f * t->density_nominal[mat];
    else {
      gzclose(f);
      return;
    }
  }

  gzgets(f, line, MCGPU_LINE);
  if (strstr(line, "#[STUB")) /* header-only file (an asset staged without its rows) for a material the voxels DO use */
    MFAIL(MCGPU_E_PARSE, "load_material: '%s' is a header-only stub (no cross-section rows) but material %d is present in the voxels; stage the full .mcgpu file", j->path, mat + 1);
  gzgets(f, line, MCGPU_LINE);
  sscanf(line, "# %d", &n_values);
  if (mat == 0) {
    if (n_values < 2 || n_values > MCGPU_MAX_ENERGYBINS_RAYLEIGH)
      MFAIL(MCGPU_E_PARSE, "load_material: unsupported number of energy bins %d (max %d)", n_values, MCGPU_MAX_ENERGYBINS_RAYLEIGH);
    t->num_values = n_values;
    t->woodcock = (mcgpu_f2*)calloc(n_values, sizeof(mcgpu_f2));
    t->mfp_a = (mcgpu_f3*)calloc((size_t)n_values * MCGPU_MAX_MATERIALS, sizeof(mcgpu_f3));
    t->mfp_b = (mcgpu_f3*)calloc((size_t)n_values * MCGPU_MAX_MATERIALS, sizeof(mcgpu_f3));
    t->ray_pmax = (float*)calloc((size_t)(n_values + 1) * MCGPU_MAX_MATERIALS, sizeof(float)); /* zero row nE: Q3 */
    t->ray_xco = (float*)calloc(NR, sizeof(float));
    t->ray_pco = (float*)calloc(NR, sizeof(float));
    t->ray_aco = (float*)calloc(NR, sizeof(float));
    t->ray_bco = (float*)calloc(NR, sizeof(float));
    t->ray_itlco = (uint8_t*)calloc(NR, 1);
    t->ray_ituco = (uint8_t*)calloc(NR, 1);
    if (!t->woodcock || !t->mfp_a || !t->mfp_b || !t->ray_pmax || !t->ray_xco || !t->ray_pco || !t->ray_aco || !t->ray_bco || !t->ray_itlco || !t->ray_ituco)
      MFAIL(MCGPU_E_NOMEM, "load_material: not enough memory for the interpolation tables");
  } else if (n_values != t->num_values)
    MFAIL(MCGPU_E_PARSE, "load_material: incorrect number of energy values in material '%s': input=%d, expected=%d", j->path, n_values, t->num_values);
  j->wood = (float*)malloc(sizeof(float) * (size_t)n_values);
  if (!j->wood) MFAIL(MCGPU_E_NOMEM, "load_material: not enough memory for the interpolation tables");

  /* -- mean free paths -> inverse MFP per unit density at the bin edges (H:2287-2332) */
  gzgets(f, line, MCGPU_LINE);
  gzgets(f, line, MCGPU_LINE);
  for (i = 0; i < n_values; i++) {
    double e = 0, ray = 0, com = 0, pho = 0, tot = 0, pmax = 0;
    mcgpu_f3* a = &t->mfp_a[(size_t)i * MCGPU_MAX_MATERIALS + mat];
    if (!gzgets(f, line, MCGPU_LINE)) MFAIL(MCGPU_E_PARSE, "load_material: '%s' ends inside the mean free path table", j->path);
    sscanf(line, "  %le  %le  %le  %le  %le  %le", &e, &ray, &com, &pho, &tot, &pmax);
    j->wood[i] = tot * (t->density_nominal[mat]) / (j->density_max[mat]); /* the Woodcock candidate of this material (H:2294-2296) */
    a->x = 1.0 / (tot * t->density_nominal[mat]);
    a->y = 1.0 / (com * t->density_nominal[mat]);
    a->z = 1.0 / (ray * t->density_nominal[mat]);
    t->ray_pmax[(size_t)i * MCGPU_MAX_MATERIALS + mat] = pmax;
    if (i == 0 && mat == 0) t->e0 = e;
    if (i == 0) {
      if (fabs(e - t->e0) > 1.0e-9) MFAIL(MCGPU_E_PARSE, "load_material: incorrect first energy value in material '%s': input=%f, expected=%f", j->path, e, t->e0);
    } else if (i == 1)
      delta_e = e - e_last;
    else if (((fabs((e - e_last) - delta_e)) / delta_e) > 0.001)
      MFAIL(MCGPU_E_PARSE, "load_material: the energy step between mean free path values is not constant (material '%s', value %d)", j->path, i);
    e_last = e;
  }
  j->delta_e = delta_e;

  /* -- slopes, then re-base the intercepts to E=0 (H:2340-2358) */
  for (i = 0; i < n_values - 1; i++) {
    const size_t bin = (size_t)i * MCGPU_MAX_MATERIALS + mat;
    t->mfp_b[bin].x = (t->mfp_a[bin + MCGPU_MAX_MATERIALS].x - t->mfp_a[bin].x) / delta_e;
    t->mfp_b[bin].y = (t->mfp_a[bin + MCGPU_MAX_MATERIALS].y - t->mfp_a[bin].y) / delta_e;
    t->mfp_b[bin].z = (t->mfp_a[bin + MCGPU_MAX_MATERIALS].z - t->mfp_a[bin].z) / delta_e;
  }
  t->mfp_b[(size_t)(n_values - 1) * MCGPU_MAX_MATERIALS + mat] = t->mfp_b[(size_t)(n_values - 2) * MCGPU_MAX_MATERIALS + mat];
  for (i = 0; i < n_values; i++) {
    const size_t bin = (size_t)i * MCGPU_MAX_MATERIALS + mat;
    const double e = t->e0 + i * delta_e;
    t->mfp_a[bin].x = t->mfp_a[bin].x - e * t->mfp_b[bin].x;
    t->mfp_a[bin].y = t->mfp_a[bin].y - e * t->mfp_b[bin].y;
    t->mfp_a[bin].z = t->mfp_a[bin].z - e * t->mfp_b[bin].z;
  }

  /* -- Rayleigh RITA grid (H:2361-2394) */
  if (!seek_gz(f, "[DATA VALUES", line)) MFAIL(MCGPU_E_PARSE, "load_material: Rayleigh data not found in file '%s'", j->path);
  gzgets(f, line, MCGPU_LINE);
  sscanf(line, "# %d", &n_rayleigh);
  if (n_rayleigh != MCGPU_NP_RAYLEIGH) MFAIL(MCGPU_E_PARSE, "load_material: %d Rayleigh sampling values in '%s', expected %d", n_rayleigh, j->path, MCGPU_NP_RAYLEIGH);
  gzgets(f, line, MCGPU_LINE);
  for (i = 0; i < n_rayleigh; i++) {
    const int bin = MCGPU_NP_RAYLEIGH * mat + i;
    int itl = 0, itu = 0;
    gzgets(f, line, MCGPU_LINE);
    sscanf(line, "  %e  %e  %e  %e  %d  %d", &t->ray_xco[bin], &t->ray_pco[bin], &t->ray_aco[bin], &t->ray_bco[bin], &itl, &itu);
    t->ray_itlco[bin] = (uint8_t)itl;
    t->ray_ituco[bin] = (uint8_t)itu;
  }

  /* -- Compton shells (H:2398-2426) */
  if (!seek_gz(f, "[NUMBER OF SHELLS", line)) MFAIL(MCGPU_E_PARSE, "load_material: Compton data not found in file '%s'", j->path);
  gzgets(f, line, MCGPU_LINE);
  sscanf(line, "# %d", &n_shells);
  if (n_shells > MCGPU_MAX_SHELLS || n_shells < 0) MFAIL(MCGPU_E_PARSE, "load_material: too many Compton shells in '%s': %d (max %d)", j->path, n_shells, MCGPU_MAX_SHELLS);
  t->cmp_noscco[mat] = n_shells;
  gzgets(f, line, MCGPU_LINE);
  for (i = 0; i < n_shells; i++) {
    const int bin = mat + i * MCGPU_MAX_MATERIALS;
    int kz, ks;
    gzgets(f, line, MCGPU_LINE);
    sscanf(line, " %e  %e  %e  %d  %d", &t->cmp_fco[bin], &t->cmp_uico[bin], &t->cmp_fj0[bin], &kz, &ks);
  }
  gzclose(f);
  t->material_loaded[mat] = 1;
  j->loaded = 1;
}
#undef MFAIL

typedef struct {
  material_job* jobs;
  int n, next;
  pthread_mutex_t mu;
} material_pool;

static void* material_worker(void* arg) {
  material_pool* p = (material_pool*)arg;
  for (;;) {
    int k;
    pthread_mutex_lock(&p->mu);
    k = p->next < p->n ? p->next++ : -1;
    pthread_mutex_unlock(&p->mu);
    if (k < 0) return NULL;
    load_one_material(&p->jobs[k]);
  }
}

int mcgpu_read_materials(mcgpu_ctx* ctx, const char* const* paths, int n_paths) {
  mcgpu_tables* t = &ctx->tab;
  float* density_max = ctx->vol.density_max;
  material_job jobs[MCGPU_MAX_MATERIALS];
  double delta_e = -99999.0;
  int mat, i, n_jobs = 0, rc = MCGPU_OK;

  mcgpu_free_tables(t);
  for (mat = 0; mat < MCGPU_MAX_MATERIALS; mat++) t->density_nominal[mat] = -1.0f;
  if (n_paths > MCGPU_MAX_MATERIALS) n_paths = MCGPU_MAX_MATERIALS;

  /* a voxel material without a file would leave its rows uninitialised in the reference */
  for (mat = 0; mat < MCGPU_MAX_MATERIALS; mat++)
    if (density_max[mat] > 0 && (mat >= n_paths || !paths[mat] || paths[mat][0] == '\0' || paths[mat][0] == '\n'))
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_material: material %d is used by the voxels but has no material file", mat + 1);
  if (n_paths < 1 || !paths[0] || paths[0][0] == '\0' || paths[0][0] == '\n')
    return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_material: the first material file is required (it defines the energy grid)");

  for (mat = 0; mat < n_paths; mat++) {
    material_job* j;
    if (!paths[mat] || paths[mat][0] == '\0' || paths[mat][0] == '\n') continue;
    j = &jobs[n_jobs++];
    memset(j, 0, sizeof *j);
    j->t = t, j->density_max = density_max, j->path = paths[mat], j->mat = mat;
  }
  load_one_material(&jobs[0]); /* material 1: allocates the tables, fixes the energy grid */
  if (jobs[0].rc == MCGPU_OK && n_jobs > 1) {
    material_pool pool;
    pthread_t threads[8];
    int n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN), k;
    if (n_threads > 8) n_threads = 8;
    if (n_threads > n_jobs - 1) n_threads = n_jobs - 1;
    if (n_threads < 1) n_threads = 1;
    pool.jobs = jobs + 1, pool.n = n_jobs - 1, pool.next = 0;
    pthread_mutex_init(&pool.mu, NULL);
    for (k = 0; k < n_threads; k++)
      if (pthread_create(&threads[k], NULL, material_worker, &pool) != 0) break;
    if (k == 0) material_worker(&pool); /* no thread could be started: do the work here */
    while (k-- > 0) pthread_join(threads[k], NULL);
    pthread_mutex_destroy(&pool.mu);
  }

  /* results in material order: first error wins; Woodcock minimum over the loaded materials (H:2294-2296);
   * the energy step the tables keep is the one of the last material read, as in the reference's loop */
  for (i = 0; i < n_jobs && rc == MCGPU_OK; i++)
    if (jobs[i].rc != MCGPU_OK) rc = mcgpu_fail(ctx, jobs[i].rc, "%s", jobs[i].err);
  if (rc == MCGPU_OK) {
    for (i = 0; i < t->num_values; i++) t->woodcock[i].x = 99999999.99f;
    for (i = 0; i < n_jobs; i++) {
      int k;
      if (!jobs[i].loaded) continue;
      for (k = 0; k < t->num_values; k++)
        if (jobs[i].wood[k] < t->woodcock[k].x) t->woodcock[k].x = jobs[i].wood[k];
      delta_e = jobs[i].delta_e;
    }
    t->ide = 1.0f / delta_e;
    t->delta_e = delta_e;
  }
  for (i = 0; i < n_jobs; i++) free(jobs[i].wood);
  if (rc != MCGPU_OK) return rc;

  /* -- Woodcock majorant: slope and re-based intercept (H:2434-2441).  The reference leaves the
   *    last slope unassigned (Q3); it is defined here as the previous one, like mfp_b. */
  for (i = 0; i < t->num_values - 1; i++) t->woodcock[i].y = (t->woodcock[i + 1].x - t->woodcock[i].x) / delta_e;
  t->woodcock[t->num_values - 1].y = t->woodcock[t->num_values - 2].y;
  for (i = 0; i < t->num_values; i++) t->woodcock[i].x = t->woodcock[i].x - (t->e0 + i * delta_e) * t->woodcock[i].y;

  /* -- spectrum must lie inside the tabulated interval (H:565-577) */
  if (ctx->have_input) {
    const mcgpu_spectrum* s = &ctx->spc;
    if ((s->espc[0] < t->e0) || (s->espc[s->num_bins] > (t->e0 + (t->num_values - 1) / t->ide)))
      return mcgpu_fail(ctx, MCGPU_E_ARG, "the input x-ray spectrum [%.3f, %.3f] eV is outside the tabulated energy interval [%.3f, %.3f] eV of the material tables", s->espc[0],
                        s->espc[s->num_bins], t->e0, (t->e0 + (t->num_values - 1) / t->ide));
  }
  return MCGPU_OK;
}

/* ------------------------------------------------------------------------------ device layout */

void mcgpu_free_scene(mcgpu_scene* s) {
  free(s->mfp), free(s->woodcock), free(s->ray_xpab), free(s->ray_itl_itu), free(s->cmp_shells), free(s->palette);
  memset(s, 0, sizeof *s);
}

int mcgpu_build_scene(mcgpu_ctx* ctx) {
  const mcgpu_tables* t = &ctx->tab;
  const mcgpu_volume* v = &ctx->vol;
  mcgpu_scene* s = &ctx->scene;
  int present[MCGPU_MAX_MATERIALS] = {0};
  int m, k, i, ns = 0;
  mcgpu_free_scene(s);
  if (v->voxel_bits == 64) {
    for (m = 0; m < MCGPU_MAX_MATERIALS; m++) present[m] = v->density_max[m] > 0 && t->material_loaded[m];
  } else
    for (k = 0; k < v->palette_size; k++) present[v->palette_material[k] - 1] = 1;
  for (m = 0; m < MCGPU_MAX_MATERIALS; m++) {
    s->slot_of_material[m] = -1;
    if (present[m]) {
      if (!t->material_loaded[m]) return mcgpu_fail(ctx, MCGPU_E_PARSE, "material %d is used by the voxels but was not loaded", m + 1);
      s->slot_of_material[m] = ns;
      s->material_of_slot[ns] = m;
      ns++;
    }
  }
  s->num_slots = ns;
  s->num_values = t->num_values;
  s->e0 = t->e0;
  s->ide = t->ide;
  s->mfp = (mcgpu_mfp_record*)calloc((size_t)t->num_values * ns, sizeof(mcgpu_mfp_record));
  s->woodcock = (mcgpu_f2*)malloc(sizeof(mcgpu_f2) * t->num_values);
  s->ray_xpab = (float*)calloc((size_t)ns * MCGPU_NP_RAYLEIGH * 4, sizeof(float));
  s->ray_itl_itu = (uint8_t*)calloc((size_t)ns * MCGPU_NP_RAYLEIGH * 2, 1);
  s->cmp_shells = (float*)calloc((size_t)ns * MCGPU_MAX_SHELLS * 4, sizeof(float));
  if (!s->mfp || !s->woodcock || !s->ray_xpab || !s->ray_itl_itu || !s->cmp_shells) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "not enough memory for the device tables");
  memcpy(s->woodcock, t->woodcock, sizeof(mcgpu_f2) * t->num_values);
  for (k = 0; k < ns; k++) {
    m = s->material_of_slot[k];
    for (i = 0; i < t->num_values; i++) {
      mcgpu_mfp_record* r = &s->mfp[(size_t)i * ns + k];
      const mcgpu_f3 a = t->mfp_a[(size_t)i * MCGPU_MAX_MATERIALS + m], b = t->mfp_b[(size_t)i * MCGPU_MAX_MATERIALS + m];
      r->ax = a.x, r->ay = a.y, r->az = a.z;
      r->bx = b.x, r->by = b.y, r->bz = b.z;
      r->pmax_next = t->ray_pmax[(size_t)(i + 1) * MCGPU_MAX_MATERIALS + m]; /* K:336 reads row index+1 */
    }
    for (i = 0; i < MCGPU_NP_RAYLEIGH; i++) {
      const int src = m * MCGPU_NP_RAYLEIGH + i;
      float* d = &s->ray_xpab[((size_t)k * MCGPU_NP_RAYLEIGH + i) * 4];
      d[0] = t->ray_xco[src], d[1] = t->ray_pco[src], d[2] = t->ray_aco[src], d[3] = t->ray_bco[src];
      s->ray_itl_itu[((size_t)k * MCGPU_NP_RAYLEIGH + i) * 2] = t->ray_itlco[src];
      s->ray_itl_itu[((size_t)k * MCGPU_NP_RAYLEIGH + i) * 2 + 1] = t->ray_ituco[src];
    }
    s->cmp_noscco[k] = t->cmp_noscco[m];
    for (i = 0; i < t->cmp_noscco[m]; i++) {
      float* d = &s->cmp_shells[((size_t)k * MCGPU_MAX_SHELLS + i) * 4];
      d[0] = t->cmp_fco[m + i * MCGPU_MAX_MATERIALS];
      d[1] = t->cmp_uico[m + i * MCGPU_MAX_MATERIALS];
      d[2] = t->cmp_fj0[m + i * MCGPU_MAX_MATERIALS];
      d[3] = d[1] * 510998.918f; /* U * m_e c^2 [eV] of K:1329/1369, the float product the kernel would form per shell term */
    }
  }
  s->tally_material_dose = ctx->have_input && ctx->in.flag_material_dose == 1;
  s->tally_voxel_dose = 0;
  s->dose_roi_voxels = 0;
  if (ctx->have_input && ctx->in.flag_voxel_dose == 1) { /* clip the ROI to the volume like load_voxels (H:2057-2065) */
    const int nv[3] = {v->nx, v->ny, v->nz};
    for (i = 0; i < 3; i++) {
      s->dose_roi[2 * i] = ctx->in.dose_roi[2 * i];
      s->dose_roi[2 * i + 1] = ctx->in.dose_roi[2 * i + 1] < nv[i] - 1 ? ctx->in.dose_roi[2 * i + 1] : nv[i] - 1;
    }
    if (s->dose_roi[0] <= s->dose_roi[1] && s->dose_roi[2] <= s->dose_roi[3] && s->dose_roi[4] <= s->dose_roi[5]) {
      s->tally_voxel_dose = 1;
      s->dose_roi_voxels = (long long)(s->dose_roi[1] - s->dose_roi[0] + 1) * (s->dose_roi[3] - s->dose_roi[2] + 1) * (s->dose_roi[5] - s->dose_roi[4] + 1);
    }
  }
  s->voxel_bits = v->voxel_bits;
  s->palette_size = v->palette_size;
  if (v->palette_size > 0) {
    s->palette = (mcgpu_f2*)malloc(sizeof(mcgpu_f2) * v->palette_size);
    if (!s->palette) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "not enough memory for the voxel palette");
    for (k = 0; k < v->palette_size; k++) {
      int slot = s->slot_of_material[v->palette_material[k] - 1];
      s->palette[k].x = v->palette_density[k];
      memcpy(&s->palette[k].y, &slot, 4);
    }
  } else {
    /* direct volume: rewrite material0 -> slot in place */
    mcgpu_f2* p = (mcgpu_f2*)v->packed;
    const size_t n = (size_t)v->nx * v->ny * v->nz;
    size_t j;
    for (j = 0; j < n; j++) {
      int slot = s->slot_of_material[v->material[j] - 1];
      memcpy(&p[j].y, &slot, 4);
    }
  }
  return MCGPU_OK;
}
