/* Projection writer: the ASCII format of report_image (docker/mcgpu/MC-GPU_v1.3.cu:2783-2953),
 * which cbctmc parses with np.loadtxt (cbctmc/mc/projection.py:36-51).  Same header lines, one
 * "%.8lf %.8lf %.8lf %.8lf" line per pixel, a blank line after each detector row, same trailer.
 *
 * The reference pays one fprintf per pixel (1.4 M per projection).  Here each value is
 * formatted by exact integer arithmetic: a double is m*2^e, so floor(v*10^8) and the rounding
 * remainder are computed exactly in 128 bits and rounded half-to-even like glibc's printf,
 * which makes the bytes identical to "%.8lf" at a fraction of the cost. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mcgpu_host.h"

#define PI 3.14159265358979323846

/* writes v with 8 decimals, returns number of chars; v >= 0 finite */
static int format_f8(char* out, double v) {
  uint64_t bits, mant;
  int exp2, n = 0;
  unsigned __int128 scaled, q;
  uint64_t ip, fp;
  char tmp[40];
  if (v == 0.0) {
    memcpy(out, "0.00000000", 10);
    return 10;
  }
  memcpy(&bits, &v, 8);
  exp2 = (int)((bits >> 52) & 0x7ff);
  mant = bits & 0xfffffffffffffull;
  if ((bits >> 63) || exp2 == 0x7ff) return sprintf(out, "%.8lf", v);
  if (exp2 == 0)
    exp2 = -1074;
  else {
    mant |= 1ull << 52;
    exp2 -= 1075;
  }
  /* v = mant * 2^exp2;  want round_half_even(mant * 10^8 * 2^exp2) */
  if (exp2 >= 0 || exp2 < -120) {
    if (exp2 < -120) { /* below 2^-67: prints as zero */
      memcpy(out, "0.00000000", 10);
      return 10;
    }
    return sprintf(out, "%.8lf", v);
  }
  scaled = (unsigned __int128)mant * 100000000ull;
  {
    const int sh = -exp2;
    const unsigned __int128 one = (unsigned __int128)1 << sh;
    const unsigned __int128 rem = scaled & (one - 1);
    const unsigned __int128 half = one >> 1;
    q = scaled >> sh;
    if (rem > half || (rem == half && (q & 1))) q += 1;
  }
  if (q >> 64) return sprintf(out, "%.8lf", v);
  ip = (uint64_t)q / 100000000ull;
  fp = (uint64_t)q % 100000000ull;
  do {
    tmp[n++] = (char)('0' + ip % 10);
    ip /= 10;
  } while (ip);
  {
    int i, k = 0;
    for (i = n - 1; i >= 0; i--) out[k++] = tmp[i];
    out[k++] = '.';
    for (i = 7; i >= 0; i--) {
      out[k + i] = (char)('0' + fp % 10);
      fp /= 10;
    }
    return k + 8;
  }
}

static void projection_angles(const mcgpu_ctx* ctx, int p, float* current, float* sequential) { /* H:2787-2800 */
  const mcgpu_input* in = &ctx->in;
  if (in->enable_specific_angles == 0) {
    *current = (in->initial_angle + p * in->D_angle) * 180.0 / PI;
    *sequential = *current;
    if (*current >= (360 - 0.0001)) *current -= 360;
  } else {
    *current = in->specific_angles[p];
    *sequential = *current;
  }
}

int mcgpu_projection_filename(const mcgpu_ctx* ctx, int p, char* out, size_t out_len) {
  float cur, seq;
  if (!ctx || !ctx->have_input || p < 0 || p >= ctx->in.num_projections) return MCGPU_E_ARG;
  projection_angles(ctx, p, &cur, &seq);
  return snprintf(out, out_len, "%s_%010.6fdeg", ctx->in.file_output, seq);
}

int mcgpu_write_projection_ascii(mcgpu_ctx* ctx, int p, const uint64_t* image, double seconds) {
  const mcgpu_view* v0;
  const mcgpu_view* vp;
  char name[MCGPU_LINE + 32];
  float cur, seq;
  unsigned long long total_histories, launched;
  int hpt, blocks, nx, nz, npix, i, j, pixel = 0;
  int max_x = 0, max_z = 0, max_pixel = 0;
  double norm, integral = 0.0, max_e = -100.0;
  const double scale = 1.0 / 100.0f;
  char* buf;
  size_t cap, len = 0;
  FILE* f;
  if (!ctx || !ctx->have_input || !image || p < 0 || p >= ctx->in.num_projections) return MCGPU_E_ARG;
  v0 = &ctx->views[0];
  vp = &ctx->views[p];
  mcgpu_current_grid(ctx, &hpt, &blocks, &launched);
  total_histories = launched;
  projection_angles(ctx, p, &cur, &seq);
  snprintf(name, sizeof name, "%s_%010.6fdeg", ctx->in.file_output, seq);
  nx = v0->num_pixels_x;
  nz = v0->num_pixels_z;
  npix = nx * nz;
  norm = scale * v0->inv_pixel_size_X * v0->inv_pixel_size_Z / ((double)total_histories);

  if (ctx->verbose) { /* H:2806-2816 */
    printf("\n\n          *** IMAGE TALLY PERFORMANCE REPORT ***\n");
    printf("              CT projection %d of %d: angle from X axis = %lf (initial angle=%lf)\n", p + 1, ctx->in.num_projections, cur, ctx->in.initial_angle);
    printf("              Simulated x rays:    %lld\n", total_histories);
    printf("              Simulation time [s]: %.2f\n", seconds);
    if (seconds > 0.000001) printf("              Speed [x-rays/s]:    %.2f\n\n", ((double)total_histories) / seconds);
    printf("              Specific angles enabled: %s\n", ctx->in.enable_specific_angles == 0 ? "NO" : "YES");
  }

  f = fopen(name, "w");
  if (!f) return mcgpu_fail(ctx, MCGPU_E_OUTPUT, "report_image: file %s can not be opened", name);
  cap = (size_t)1 << 22;
  buf = (char*)malloc(cap + 256);
  if (!buf) {
    fclose(f);
    return mcgpu_fail(ctx, MCGPU_E_NOMEM, "report_image: out of memory");
  }
  fprintf(f, "# \n");
  fprintf(f, "#     *****************************************************************************\n");
  fprintf(f, "#     ***         MC-GPU, version 1.3 (http://code.google.com/p/mcgpu/)         ***\n");
  fprintf(f, "#     ***                                                                       ***\n");
  fprintf(f, "#     ***                     Andreu Badal (Andreu.Badal-Soler@fda.hhs.gov)     ***\n");
  fprintf(f, "#     *****************************************************************************\n");
  fprintf(f, "# \n");
  fprintf(f, "#  *** SIMULATION IN THE GPU USING CUDA ***\n");
  fprintf(f, "#\n");
  fprintf(f, "#  Image created counting the energy arriving at each pixel: ideal energy integrating detector.\n");
  fprintf(f, "#  Pixel value units: eV/cm^2 per history (energy fluence).\n");
  fprintf(f, "#  CT projection %d of %d: angle from X axis = %lf (mod 360deg), %lf (no mod 360deg) \n", p + 1, ctx->in.num_projections, cur, seq);
  fprintf(f, "#  Focal spot position = (%.8f,%.8f,%.8f), cone beam direction = (%.8f,%.8f,%.8f)\n", vp->src_pos[0], vp->src_pos[1], vp->src_pos[2], vp->src_dir[0], vp->src_dir[1],
          vp->src_dir[2]);
  fprintf(f, "#  Specific angles enabled: %s\n", ctx->in.enable_specific_angles == 0 ? "NO" : "YES");
  fprintf(f, "#  Pixel size:  %lf x %lf = %lf cm^2\n", 1.0 / (double)(v0->inv_pixel_size_X), 1.0 / (double)(v0->inv_pixel_size_Z),
          1.0 / (double)(v0->inv_pixel_size_X * v0->inv_pixel_size_Z));
  fprintf(f, "#  Number of pixels in X and Z:  %d  %d\n", nx, nz);
  fprintf(f, "#  (X rows given first, a blank line separates the different Z values)\n");
  fprintf(f, "# \n");
  fprintf(f, "#  [NON-SCATTERED] [COMPTON] [RAYLEIGH] [MULTIPLE-SCATTING]\n");
  fprintf(f, "# ==========================================================\n");

  for (j = 0; j < nz; j++) {
    for (i = 0; i < nx; i++) {
      const double e0 = (double)image[pixel], e1 = (double)image[pixel + npix], e2 = (double)image[pixel + 2 * npix], e3 = (double)image[pixel + 3 * npix];
      const double tot = e0 + e1 + e2 + e3;
      len += format_f8(buf + len, norm * e0);
      buf[len++] = ' ';
      len += format_f8(buf + len, norm * e1);
      buf[len++] = ' ';
      len += format_f8(buf + len, norm * e2);
      buf[len++] = ' ';
      len += format_f8(buf + len, norm * e3);
      buf[len++] = '\n';
      if (tot > max_e) {
        max_e = tot;
        max_x = i;
        max_z = j;
        max_pixel = pixel;
      }
      integral += tot;
      pixel++;
      if (len > cap) {
        fwrite(buf, 1, len, f);
        len = 0;
      }
    }
    buf[len++] = '\n';
  }
  fwrite(buf, 1, len, f);
  free(buf);

  fprintf(f, "#   *** Simulation REPORT: ***\n");
  fprintf(f, "#       Fraction of energy detected (over the mean energy of the spectrum): %.3lf%%\n", 100.0 * scale * (integral / (double)(total_histories)) / (double)(ctx->spc.mean_energy));
  fprintf(f, "#       Maximum energy detected in pixel %i: (x,y)=(%i,%i) -> pixel value = %lf eV/cm^2\n", max_pixel, max_x, max_z, norm * max_e);
  fprintf(f, "#       Simulated x rays:    %lld\n", total_histories);
  fprintf(f, "#       Simulation time [s]: %.2f\n", seconds);
  if (seconds > 0.000001) fprintf(f, "#       Speed [x-rays/sec]:  %.2f\n\n", ((double)total_histories) / seconds);
  if (fclose(f) != 0) return mcgpu_fail(ctx, MCGPU_E_OUTPUT, "report_image: error writing %s", name);

  if (ctx->verbose) {
    printf("              Fraction of initial energy arriving at the detector (over the mean energy of the spectrum):  %.3lf%%\n",
           100.0 * scale * (integral / (double)(total_histories)) / (double)(ctx->spc.mean_energy));
    printf("              Maximum energy detected in pixel %i: (x,y)=(%i,%i). Maximum pixel value = %lf eV/cm^2\n\n", max_pixel, max_x, max_z, norm * max_e);
    fflush(stdout);
  }
  return MCGPU_OK;
}

/* Optional binary side-file: '<base>_%010.6fdeg.raw', little-endian float32 [4][Nz][Nx] of
 * the same NORM*count values as the ASCII columns (the reference's own .raw writer is commented out,
 * H:2911-2949).  4x smaller than the text and read with one np.fromfile instead of np.loadtxt. */
int mcgpu_write_projection_raw(mcgpu_ctx* ctx, int p, const uint64_t* image) {
  char name[MCGPU_LINE + 40];
  float cur, seq;
  unsigned long long launched;
  int hpt, blocks;
  size_t n, i;
  double norm;
  float* buf;
  FILE* f;
  if (!ctx || !ctx->have_input || !image || p < 0 || p >= ctx->in.num_projections) return MCGPU_E_ARG;
  mcgpu_current_grid(ctx, &hpt, &blocks, &launched);
  projection_angles(ctx, p, &cur, &seq);
  snprintf(name, sizeof name, "%s_%010.6fdeg.raw", ctx->in.file_output, seq);
  n = (size_t)4 * ctx->views[0].total_num_pixels;
  norm = (1.0 / 100.0f) * ctx->views[0].inv_pixel_size_X * ctx->views[0].inv_pixel_size_Z / ((double)launched);
  buf = (float*)malloc(n * sizeof(float));
  if (!buf) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "write_projection_raw: out of memory");
  for (i = 0; i < n; i++) buf[i] = (float)(norm * (double)image[i]);
  f = fopen(name, "wb");
  if (!f || fwrite(buf, sizeof(float), n, f) != n) {
    if (f) fclose(f);
    free(buf);
    return mcgpu_fail(ctx, MCGPU_E_OUTPUT, "write_projection_raw: file %s can not be written", name);
  }
  fclose(f);
  free(buf);
  return MCGPU_OK;
}
