// Projection post-processing on the device: what cbctmc does in NumPy/SciPy after the
// simulation (cbctmc/mc/projection.py:36-51, 101-169; cbctmc/mc/simulation.py:235-277), starting from
// the u64 tallies instead of 64 MB of text per projection:
//   intensity : tally -> the float32 values np.loadtxt(...).astype(float32) would read back from the
//               "%.8lf" text, detector rows flipped, x cropped to the half-fan width, planes summed
//               (total / unscattered / scattered), smallest positive value of each image;
//   gaussian  : scipy.ndimage.gaussian_filter(float32 image, sigma=(sz, sx)) -- mode 'reflect',
//               truncate 4.0, float64 accumulation in SciPy's order, float32 between the two passes;
//   normalize : np.where(p == 0, min_non_zero, p) then np.log(air / p) in float32 (Beer-Lambert).
// Built like the rest of the device layer: sm_100a, -fmad=false, no fast-math.
#include "device_internal.h"

#define POST_MAX_RADIUS 512

// The decimal the reference's report writes with "%.8lf" and np.loadtxt parses back: round-half-even of
// v*1e8 on the exact product (hi + lo by FMA), then the correctly rounded quotient q/1e8 (= strtod of
// that decimal string), then float32 like .astype(np.float32).
__device__ __forceinline__ float text_round_trip(double v) {
  const double hi = __dmul_rn(v, 1.0e8);
  const double lo = __fma_rn(v, 1.0e8, -hi);  // hi + lo = v * 1e8 exactly
  const double q = floor(hi);
  const double g = __dsub_rn(__dsub_rn(hi, q), 0.5);  // exact; a non-zero g is at least one ulp(hi), i.e. larger than |lo|
  const bool up = g > 0.0 || (g == 0.0 && (lo > 0.0 || (lo == 0.0 && fmod(q, 2.0) != 0.0)));  // round half to even, like printf
  return (float)__ddiv_rn(up ? q + 1.0 : q, 1.0e8);  // correctly rounded quotient = strtod of the decimal string
}

__global__ void post_intensity(const unsigned long long* __restrict__ tally, double norm, int nx, int nz, int crop, float* __restrict__ total,
                               float* __restrict__ unscattered, float* __restrict__ scattered, unsigned* __restrict__ min_bits) {
  const size_t npix = (size_t)nx * nz;
  unsigned m0 = 0x7f800000u, m1 = 0x7f800000u, m2 = 0x7f800000u;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)crop * nz; i += (size_t)gridDim.x * blockDim.x) {
    const int zo = (int)(i / crop), x = (int)(i % crop);
    const size_t src = (size_t)(nz - 1 - zo) * nx + x;  // np.flip(data, axis=0), then data[:, :crop]
    const float v0 = text_round_trip(norm * (double)tally[src]);
    const float v1 = text_round_trip(norm * (double)tally[src + npix]);
    const float v2 = text_round_trip(norm * (double)tally[src + 2 * npix]);
    const float v3 = text_round_trip(norm * (double)tally[src + 3 * npix]);
    const float t = __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), v3);  // float32 sum over the last axis, in order
    const float s = __fadd_rn(__fadd_rn(v1, v2), v3);
    if (total) total[i] = t;
    if (unscattered) unscattered[i] = v0;
    if (scattered) scattered[i] = s;
    if (t > 0.0f) m0 = min(m0, __float_as_uint(t));  // positive floats order like their bit patterns
    if (v0 > 0.0f) m1 = min(m1, __float_as_uint(v0));
    if (s > 0.0f) m2 = min(m2, __float_as_uint(s));
  }
  for (int o = 16; o > 0; o >>= 1) {
    m0 = min(m0, __shfl_xor_sync(0xffffffffu, m0, o));
    m1 = min(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(min_bits + 0, m0);
    atomicMin(min_bits + 1, m1);
    atomicMin(min_bits + 2, m2);
  }
}

// one pass of scipy.ndimage.correlate1d with a symmetric kernel (ni_filters.c, NI_Correlate1D): double
// accumulation, centre tap first, then the pairs from the outermost inwards; 'reflect' boundary
// (d c b a | a b c d | d c b a); result stored as float32.  w[0..radius]: w[k] = weight at distance k.
__global__ void post_correlate1d(const float* __restrict__ in, float* __restrict__ out, int n0, int n1, int axis, const double* __restrict__ w, int radius) {
  const size_t n = (size_t)n0 * n1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / n1), c = (int)(i % n1);
    const int len = axis == 0 ? n0 : n1, pos = axis == 0 ? r : c;
    const size_t stride = axis == 0 ? (size_t)n1 : 1, base = axis == 0 ? (size_t)c : (size_t)r * n1;
    double acc = __dmul_rn((double)in[base + (size_t)pos * stride], w[0]);
    for (int k = radius; k >= 1; k--) {
      int a = pos - k, b = pos + k;
      const int period = 2 * len;
      a = ((a % period) + period) % period;
      if (a >= len) a = period - 1 - a;
      b = b % period;
      if (b >= len) b = period - 1 - b;
      const double pair = __dadd_rn((double)in[base + (size_t)a * stride], (double)in[base + (size_t)b * stride]);
      acc = __dadd_rn(acc, __dmul_rn(pair, w[k]));
    }
    out[i] = (float)acc;
  }
}

__global__ void post_normalize(const float* __restrict__ air, float* __restrict__ stack, size_t per_image, size_t n, float min_nonzero) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float p = stack[i];
    if (p == 0.0f) p = min_nonzero;
    stack[i] = logf(__fdiv_rn(air[i % per_image], p));
  }
}

// ------------------------------------------------------------------------------------------ extern "C"
// one device workspace per device, grown on demand and released with the device: no cudaMalloc/cudaFree per call
static int post_workspace(mcgpu_device* d, size_t bytes, void** out, char* err, size_t errlen) {
  if (d->post_ws_bytes < bytes) {
    CK(cudaStreamSynchronize(d->stream));
    cudaFree(d->post_ws);
    d->post_ws = NULL, d->post_ws_bytes = 0;
    CK(cudaMalloc(&d->post_ws, bytes));
    d->post_ws_bytes = bytes;
  }
  *out = d->post_ws;
  return 0;
}

static int post_grid(const mcgpu_device* d, size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)d->sm_count * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

extern "C" int mcgpu_dev_post_intensity(struct mcgpu_device* d, const uint64_t* tally_host, double norm, int nx, int nz, int crop, float* total, float* unscattered,
                                        float* scattered, float* min_positive, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  const size_t words = (size_t)4 * nx * nz, nout = (size_t)crop * nz;
  const size_t off_out = (words * sizeof(unsigned long long) + 255) & ~size_t(255), off_min = off_out + ((3 * nout * sizeof(float) + 255) & ~size_t(255));
  unsigned char* ws = NULL;
  if (post_workspace(d, off_min + 256, (void**)&ws, err, errlen)) return -1;
  const unsigned long long* src = d->d_image;
  if (tally_host) {
    CK(cudaMemcpyAsync(ws, tally_host, words * sizeof(unsigned long long), cudaMemcpyHostToDevice, d->stream));
    src = reinterpret_cast<const unsigned long long*>(ws);
  } else if (!src || d->image_words != words) {
    snprintf(err, errlen, "post_intensity: no tally on device %d", d->ordinal);
    return -1;
  }
  float* buf = reinterpret_cast<float*>(ws + off_out);
  unsigned* mins = reinterpret_cast<unsigned*>(ws + off_min);
  CK(cudaMemsetAsync(mins, 0x7f, 3 * sizeof(unsigned), d->stream));  // 0x7f7f7f7f: a large finite float
  post_intensity<<<post_grid(d, nout), 256, 0, d->stream>>>(src, norm, nx, nz, crop, buf, buf + nout, buf + 2 * nout, mins);
  CK(cudaGetLastError());
  if (total) CK(cudaMemcpyAsync(total, buf, nout * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  if (unscattered) CK(cudaMemcpyAsync(unscattered, buf + nout, nout * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  if (scattered) CK(cudaMemcpyAsync(scattered, buf + 2 * nout, nout * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  unsigned h[3];
  CK(cudaMemcpyAsync(h, mins, sizeof h, cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  if (min_positive)
    for (int k = 0; k < 3; k++) memcpy(&min_positive[k], &h[k], sizeof(float));
  return 0;
}

extern "C" int mcgpu_dev_post_gaussian(struct mcgpu_device* d, const float* in, int n0, int n1, const double* w0, int r0, const double* w1, int r1, float* out, char* err,
                                       size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  const size_t n = (size_t)n0 * n1, img = (n * sizeof(float) + 255) & ~size_t(255);
  if (r0 > POST_MAX_RADIUS || r1 > POST_MAX_RADIUS) {
    snprintf(err, errlen, "post_gaussian: radius %d/%d above %d", r0, r1, POST_MAX_RADIUS);
    return -1;
  }
  unsigned char* ws = NULL;
  if (post_workspace(d, 2 * img + (size_t)(POST_MAX_RADIUS + 1) * 2 * sizeof(double), (void**)&ws, err, errlen)) return -1;
  float *cur = reinterpret_cast<float*>(ws), *nxt = reinterpret_cast<float*>(ws + img);
  double* w = reinterpret_cast<double*>(ws + 2 * img);
  CK(cudaMemcpyAsync(cur, in, n * sizeof(float), cudaMemcpyHostToDevice, d->stream));
  if (w0) {  // axis 0 first, like gaussian_filter's loop over the axes
    CK(cudaMemcpyAsync(w, w0, (size_t)(r0 + 1) * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    post_correlate1d<<<post_grid(d, n), 256, 0, d->stream>>>(cur, nxt, n0, n1, 0, w, r0);
    float* t = cur;
    cur = nxt, nxt = t;
  }
  if (w1) {
    CK(cudaMemcpyAsync(w + POST_MAX_RADIUS + 1, w1, (size_t)(r1 + 1) * sizeof(double), cudaMemcpyHostToDevice, d->stream));
    post_correlate1d<<<post_grid(d, n), 256, 0, d->stream>>>(cur, nxt, n0, n1, 1, w + POST_MAX_RADIUS + 1, r1);
    float* t = cur;
    cur = nxt, nxt = t;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, cur, n * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}

extern "C" int mcgpu_dev_post_normalize(struct mcgpu_device* d, const float* air, float* stack, long long n_images, int n0, int n1, float min_nonzero, char* err, size_t errlen) {
  CK(cudaSetDevice(d->ordinal));
  const size_t per = (size_t)n0 * n1, img = (per * sizeof(float) + 255) & ~size_t(255);
  const long long chunk_images = 16;  // 16 x 3 MB at the reference size: bounded device footprint for any stack length
  unsigned char* ws = NULL;
  if (post_workspace(d, img + per * sizeof(float) * (size_t)chunk_images, (void**)&ws, err, errlen)) return -1;
  float *dair = reinterpret_cast<float*>(ws), *dbuf = reinterpret_cast<float*>(ws + img);
  CK(cudaMemcpyAsync(dair, air, per * sizeof(float), cudaMemcpyHostToDevice, d->stream));
  for (long long first = 0; first < n_images; first += chunk_images) {
    const long long m = n_images - first < chunk_images ? n_images - first : chunk_images;
    const size_t n = per * (size_t)m;
    CK(cudaMemcpyAsync(dbuf, stack + (size_t)first * per, n * sizeof(float), cudaMemcpyHostToDevice, d->stream));
    post_normalize<<<post_grid(d, n), 256, 0, d->stream>>>(dair, dbuf, per, n, min_nonzero);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(stack + (size_t)first * per, dbuf, n * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  }
  CK(cudaStreamSynchronize(d->stream));
  return 0;
}
