/* .in parser: the drop-in grammar of MC-GPU's read_input (docker/mcgpu/MC-GPU_v1.3.cu:1240-1895,
 * "H").  Sections are located by substring in the order the reference expects; values are read
 * from comment-stripped lines.  Only parsing lives here; the poses derived from the values are
 * built in geometry.c. */
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "mcgpu_host.h"

/* Worker threads of a scan (api.c: scan_thread) report into their own buffer: the context's message buffer belongs to the
 * caller's thread.  mcgpu_fail_into(buf) redirects the failures raised on THIS thread until it is reset with NULL. */
static __thread char* tls_err_buf = NULL;
static __thread size_t tls_err_len = 0;

void mcgpu_fail_into(char* buf, size_t len) {
  tls_err_buf = buf;
  tls_err_len = buf ? len : 0;
}

int mcgpu_fail(mcgpu_ctx* ctx, int code, const char* fmt, ...) {
  char* out = tls_err_buf ? tls_err_buf : ctx->err;
  const size_t len = tls_err_buf ? tls_err_len : sizeof ctx->err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(out, len, fmt, ap);
  va_end(ap);
  if (ctx->verbose) { /* the reference reports on stdout, where cbctmc greps for "error" (Q9) */
    printf("\n\n   !!ERROR!! %s\n\n", out);
    fflush(stdout);
  }
  return code;
}

/* H:1907-1926: first blank-delimited token, '#' ends it too. */
void mcgpu_trim_name(const char* line, char* name) {
  int a = 0, b = 0;
  while (line[a] == ' ') a++;
  while (line[a] != ' ' && line[a] != '#' && line[a] != '\0' && line[a] != '\n' && line[a] != '\r' && b < MCGPU_LINE - 1)
    name[b++] = line[a++];
  name[b] = '\0';
}

/* H:1935-1965: next line that is not blank / comment-only, leading blanks and trailing comment removed. */
char* mcgpu_fgets_trimmed(char* out, int num, FILE* f) {
  char raw[MCGPU_LINE];
  char* got;
  int b;
  out[0] = '\0';
  do {
    int a = 0;
    b = 0;
    got = fgets(raw, num, f);
    if (got) {
      while (raw[a] == ' ') a++;
      while (raw[a] != '\n' && raw[a] != '#' && raw[a] != '\0') out[b++] = raw[a++];
    }
    out[b] = '\0';
  } while (got && out[0] == '\0');
  return got;
}

static int seek_section(mcgpu_ctx* ctx, FILE* f, const char* tag, char* line) {
  do {
    if (!fgets(line, MCGPU_LINE, f))
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "read_input: input file is not readable or does not contain the string '%s'", tag);
  } while (!strstr(line, tag));
  return MCGPU_OK;
}

static int yes_no(const char* s) {
  if (!strncmp("YE", s, 2) || !strncmp("Ye", s, 2) || !strncmp("ye", s, 2)) return 1;
  if (!strncmp("NO", s, 2) || !strncmp("No", s, 2) || !strncmp("no", s, 2)) return 0;
  return -1;
}

int mcgpu_parse_input(mcgpu_ctx* ctx, const char* in_path) {
  mcgpu_input* in = &ctx->in;
  char line[MCGPU_LINE];
  double d;
  int rc, i;
  FILE* f = fopen(in_path, "r");
  if (!f) return mcgpu_fail(ctx, MCGPU_E_ARG, "read_input: input file not found or not readable: '%s'", in_path);
  memset(in, 0, sizeof *in);

#define SECTION(tag)                                     \
  if ((rc = seek_section(ctx, f, tag, line)) != MCGPU_OK) { \
    fclose(f);                                           \
    return rc;                                           \
  }
#define BAIL(...)                                        \
  do {                                                   \
    fclose(f);                                           \
    return mcgpu_fail(ctx, MCGPU_E_PARSE, __VA_ARGS__);  \
  } while (0)

  /* -- simulation config (H:1280-1312) */
  SECTION("SECTION SIMULATION CONFIG v.2009-05-12");
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  d = 0.0;
  sscanf(line, "%lf", &d);
  in->total_histories = (unsigned long long)(d + 0.0001);
  /* H:654: values below 95 000 mean "simulate this many SECONDS" (speed test + extrapolation), a mode
   * that is irreproducible by construction and that cbctmc never uses (it always writes a history count). */
  if (in->total_histories < 95000ull)
    BAIL("read_input: %llu is below 95000 and would mean a simulation TIME in seconds in MC-GPU; time-limited runs are not supported, give a number of histories",
         in->total_histories);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%d", &in->seed_input);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%d", &in->gpu_id);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%d", &in->threads_per_block);
  if (in->threads_per_block <= 0 || in->threads_per_block % 32 != 0)
    BAIL("read_input: the number of GPU threads per CUDA block must be a multiple of 32 (input: %d)", in->threads_per_block);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%d", &in->histories_per_thread);
  if (in->histories_per_thread < 1) BAIL("read_input: histories per thread must be positive (input: %d)", in->histories_per_thread);

  /* -- source (H:1315-1395); the pose itself is derived in geometry.c from these raw values */
  SECTION("SECTION SOURCE v.2011-07-12");
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  mcgpu_trim_name(line, in->file_espc);
  ctx->views = (mcgpu_view*)calloc(MCGPU_MAX_PROJECTIONS, sizeof(mcgpu_view));
  if (!ctx->views) {
    fclose(f);
    return mcgpu_fail(ctx, MCGPU_E_NOMEM, "read_input: out of memory");
  }
  mcgpu_view* v0 = &ctx->views[0];
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%f %f %f", &v0->src_pos[0], &v0->src_pos[1], &v0->src_pos[2]);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%f %f %f", &v0->src_dir[0], &v0->src_dir[1], &v0->src_dir[2]);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  in->phi1_deg = in->phi2_deg = in->theta_deg = 0.0;
  sscanf(line, "%lf %lf %lf", &in->phi1_deg, &in->phi2_deg, &in->theta_deg);
  if (in->theta_deg > 180.0) BAIL("read_input: input polar aperture must be in [0,180] deg (theta=%f)", in->theta_deg);
  if (in->phi1_deg + in->phi2_deg > 360.0) BAIL("read_input: input azimuthal aperture must be in [0,360] deg (phi=%f)", in->phi1_deg + in->phi2_deg);

  /* -- detector (H:1398-1449) */
  SECTION("SECTION IMAGE DETECTOR v.2009-12-02");
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  mcgpu_trim_name(line, in->file_output);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  {
    float px = 0.f, pz = 0.f;
    sscanf(line, "%f %f", &px, &pz);
    v0->num_pixels_x = (int)(px + 0.001f);
    v0->num_pixels_z = (int)(pz + 0.001f);
    v0->total_num_pixels = v0->num_pixels_x * v0->num_pixels_z;
    if (v0->total_num_pixels < 1 || v0->total_num_pixels > 99999999)
      BAIL("read_input: the input number of pixels is incorrect: %d x %d", v0->num_pixels_x, v0->num_pixels_z);
  }
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%f %f", &v0->width_X, &v0->height_Z);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%f", &v0->sdd);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  sscanf(line, "%f", &v0->lateral_displacement);
  if (v0->sdd < 1.0e-6) BAIL("read_input: the source-to-detector distance must be positive (sdd=%f)", v0->sdd);

  /* -- explicit projection angles (fork section, H:1472-1533) */
  SECTION("SECTION ANGLES OF PROJ v.2023-09-06");
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  in->enable_specific_angles = yes_no(line);
  if (in->enable_specific_angles < 0) BAIL("read_input: answer YES or NO in the first line of 'SECTION ANGLES OF PROJ' (input: %s)", line);
  i = 0;
  while (!strstr(line, "SECTION CT SCAN TRAJECTORY v.2011-10-25")) {
    float angle = 99999.f;
    if (!fgets(line, MCGPU_LINE, f)) BAIL("read_input: input file does not contain the string 'SECTION CT SCAN TRAJECTORY v.2011-10-25'");
    if (sscanf(line, "%f", &angle) == 1 && angle != 99999.f) {
      if (i >= MCGPU_MAX_PROJECTIONS) BAIL("read_input: too many angles are specified (max %d)", MCGPU_MAX_PROJECTIONS);
      in->specific_angles[i++] = angle;
    }
  }
  in->num_specific_angles = i;
  if (in->enable_specific_angles == 1 && i == 0) BAIL("read_input: specific angles enabled but no angle was specified");

  /* -- CT trajectory (H:1535-1615) */
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  in->num_projections = 0;
  sscanf(line, "%d", &in->num_projections);
  if (in->num_projections == 0) in->num_projections = 1;
  if (in->enable_specific_angles == 1) in->num_projections = i;
  if (abs(in->num_projections) > MCGPU_MAX_PROJECTIONS)
    BAIL("read_input: the input number of projections is too large (max %d)", MCGPU_MAX_PROJECTIONS);
  in->D_angle = -1.0;
  in->angularROI_0 = 0.0;
  in->angularROI_1 = 360.0;
  in->initial_angle = 0.0;
  in->SRotAxisD = -1.0;
  in->vertical_translation = 0.0;
  if (in->num_projections != 1 || in->enable_specific_angles == 1) {
    mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
    sscanf(line, "%lf", &in->D_angle);
    mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
    sscanf(line, "%lf %lf", &in->angularROI_0, &in->angularROI_1);
    mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
    sscanf(line, "%lf", &in->SRotAxisD);
    if (in->SRotAxisD < 0.0 || in->SRotAxisD > v0->sdd)
      BAIL("read_input: invalid source-to-rotation axis distance %f (sdd=%f)", in->SRotAxisD, v0->sdd);
    mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
    sscanf(line, "%lf", &in->vertical_translation);
  }

  /* -- dose tallies (H:1619-1709) */
  do {
    if (!fgets(line, MCGPU_LINE, f)) BAIL("read_input: input file does not contain the string 'SECTION DOSE DEPOSITION v.2012-12-12'");
    if (strstr(line, "SECTION DOSE DEPOSITION v.2011-02-18")) BAIL("read_input: please update the input file to the MC-GPU v1.3 format (DOSE DEPOSITION v.2012-12-12)");
  } while (!strstr(line, "SECTION DOSE DEPOSITION v.2012-12-12"));
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  in->flag_material_dose = yes_no(line);
  if (in->flag_material_dose < 0) BAIL("read_input: answer YES or NO in 'SECTION DOSE DEPOSITION' (input: %s)", line);
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  in->flag_voxel_dose = yes_no(line);
  if (in->flag_voxel_dose < 0) BAIL("read_input: answer YES or NO in 'SECTION DOSE DEPOSITION' (input: %s)", line);
  if (in->flag_voxel_dose == 1) {
    short* r = in->dose_roi;
    mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
    mcgpu_trim_name(line, in->file_dose);
    for (i = 0; i < 3; i++) {
      mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
      sscanf(line, "%hd %hd", &r[2 * i], &r[2 * i + 1]);
      r[2 * i] -= 1;
      r[2 * i + 1] -= 1;
    }
    if (r[0] > r[1] || r[2] > r[3] || r[4] > r[5] || r[0] < 0 || r[2] < 0 || r[4] < 0)
      BAIL("read_input: the input region-of-interest in 'SECTION DOSE DEPOSITION' is not valid");
  } else {
    for (i = 0; i < 3; i++) {
      in->dose_roi[2 * i] = (short)32500;
      in->dose_roi[2 * i + 1] = (short)-32500;
    }
  }

  /* -- voxel and material files (H:1713-1745) */
  SECTION("SECTION VOXELIZED GEOMETRY FILE v.2009-11-30");
  mcgpu_fgets_trimmed(line, MCGPU_LINE, f);
  mcgpu_trim_name(line, in->file_voxels);
  SECTION("SECTION MATERIAL");
  for (i = 0; i < MCGPU_MAX_MATERIALS; i++) {
    if (!mcgpu_fgets_trimmed(line, MCGPU_LINE, f))
      in->file_materials[i][0] = '\0';
    else
      mcgpu_trim_name(line, in->file_materials[i]);
  }
  fclose(f);
#undef SECTION
#undef BAIL
  return MCGPU_OK;
}
