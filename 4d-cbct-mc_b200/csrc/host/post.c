/* Projection post-processing entry points: the C-ABI face of csrc/cuda/postprocess.cu.
 * What cbctmc does after a simulation in NumPy/SciPy (cbctmc/mc/projection.py:36-51, 101-169,
 * cbctmc/mc/simulation.py:235-277) starting from the u64 tallies; see include/mcgpu_b200.h. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mcgpu_host.h"

int mcgpu_dev_post_intensity(struct mcgpu_device* d, const uint64_t* tally_host, double norm, int nx, int nz, int crop, float* total, float* unscattered, float* scattered,
                             float* min_positive, char* err, size_t errlen);
int mcgpu_dev_post_gaussian(struct mcgpu_device* d, const float* in, int n0, int n1, const double* w0, int r0, const double* w1, int r1, float* out, char* err, size_t errlen);
int mcgpu_dev_post_normalize(struct mcgpu_device* d, const float* air, float* stack, long long n_images, int n0, int n1, float min_nonzero, char* err, size_t errlen);

int mcgpu_post_intensity(mcgpu_ctx* ctx, const uint64_t* tally, unsigned long long launched_histories, int crop_x, float* total, float* unscattered, float* scattered,
                         float* min_positive) {
  const mcgpu_view* v0;
  double norm;
  const double scale = 1.0 / 100.0f; /* SCALE_eV, as in report_image (H:2860-2861) */
  if (!ctx || !ctx->have_input) return MCGPU_E_ARG;
  mcgpu_devices_ready(ctx);
  if (ctx->num_devices < 1) return mcgpu_fail(ctx, MCGPU_E_CUDA, "post_intensity: no CUDA device (this engine has no CPU path)");
  v0 = &ctx->views[0];
  if (crop_x <= 0 || crop_x > v0->num_pixels_x) crop_x = v0->num_pixels_x;
  if (launched_histories == 0) return mcgpu_fail(ctx, MCGPU_E_ARG, "post_intensity: launched_histories must be positive");
  norm = scale * v0->inv_pixel_size_X * v0->inv_pixel_size_Z / ((double)launched_histories);
  if (mcgpu_dev_post_intensity(ctx->dev[0], tally, norm, v0->num_pixels_x, v0->num_pixels_z, crop_x, total, unscattered, scattered, min_positive, ctx->err, sizeof ctx->err) != 0)
    return MCGPU_E_CUDA;
  return MCGPU_OK;
}

/* numpy's float64 add.reduce (pairwise_sum in loops_utils.h): plain loop below 8 elements, 8 accumulators up to
 * 128, halves (rounded down to a multiple of 8) above -- reproduced so that the kernel weights are bit-equal. */
static double numpy_pairwise_sum(const double* a, int n) {
  if (n < 8) {
    double res = 0.0;
    int i;
    for (i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8], res;
    int i, j;
    for (j = 0; j < 8; j++) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (j = 0; j < 8; j++) r[j] += a[i + j];
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return numpy_pairwise_sum(a, n2) + numpy_pairwise_sum(a + n2, n - n2);
  }
}

/* scipy.ndimage._filters._gaussian_kernel1d(sigma, order 0, radius): exp(-0.5 / sigma^2 * x^2) / sum in float64;
 * radius = int(truncate * sigma + 0.5) with truncate = 4.0.  w[k] = weight at distance k. */
static int gaussian_weights(double sigma, double** w_out) {
  const int radius = (int)(4.0 * sigma + 0.5);
  double* w = (double*)malloc((size_t)(radius + 1) * sizeof(double));
  double* full = (double*)malloc((size_t)(2 * radius + 1) * sizeof(double));
  double sum;
  int i;
  if (!w || !full) {
    free(w), free(full);
    return -1;
  }
  for (i = -radius; i <= radius; i++) full[i + radius] = exp(-0.5 / (sigma * sigma) * (double)(i * i));
  sum = numpy_pairwise_sum(full, 2 * radius + 1);
  for (i = 0; i <= radius; i++) w[i] = full[radius + i] / sum;
  free(full);
  *w_out = w;
  return radius;
}

int mcgpu_gaussian_weights(double sigma, double* w, int capacity) {
  double* tmp = NULL;
  int r, i;
  if (!(sigma > 1e-15)) return MCGPU_E_ARG;
  r = gaussian_weights(sigma, &tmp);
  if (r < 0) return MCGPU_E_NOMEM;
  if (w)
    for (i = 0; i <= r && i < capacity; i++) w[i] = tmp[i];
  free(tmp);
  return r;
}

int mcgpu_post_gaussian(mcgpu_ctx* ctx, const float* in, int n0, int n1, double sigma0, double sigma1, float* out) {
  double *w0 = NULL, *w1 = NULL;
  int r0 = 0, r1 = 0, rc;
  if (!ctx || !in || !out || n0 < 1 || n1 < 1) return MCGPU_E_ARG;
  mcgpu_devices_ready(ctx);
  if (ctx->num_devices < 1) return mcgpu_fail(ctx, MCGPU_E_CUDA, "post_gaussian: no CUDA device (this engine has no CPU path)");
  if (sigma0 > 1e-15 && (r0 = gaussian_weights(sigma0, &w0)) < 0) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "post_gaussian: out of memory");
  if (sigma1 > 1e-15 && (r1 = gaussian_weights(sigma1, &w1)) < 0) {
    free(w0);
    return mcgpu_fail(ctx, MCGPU_E_NOMEM, "post_gaussian: out of memory");
  }
  rc = mcgpu_dev_post_gaussian(ctx->dev[0], in, n0, n1, w0, r0, w1, r1, out, ctx->err, sizeof ctx->err);
  free(w0), free(w1);
  return rc == 0 ? MCGPU_OK : MCGPU_E_CUDA;
}

int mcgpu_post_normalize(mcgpu_ctx* ctx, const float* air, float* stack, long long n_images, int n0, int n1, float min_nonzero) {
  if (!ctx || !air || !stack || n_images < 0 || n0 < 1 || n1 < 1) return MCGPU_E_ARG;
  mcgpu_devices_ready(ctx);
  if (ctx->num_devices < 1) return mcgpu_fail(ctx, MCGPU_E_CUDA, "post_normalize: no CUDA device (this engine has no CPU path)");
  if (n_images == 0) return MCGPU_OK;
  return mcgpu_dev_post_normalize(ctx->dev[0], air, stack, n_images, n0, n1, min_nonzero, ctx->err, sizeof ctx->err) == 0 ? MCGPU_OK : MCGPU_E_CUDA;
}
