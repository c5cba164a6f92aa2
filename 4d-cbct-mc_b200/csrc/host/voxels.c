/* Voxel geometry: penEasy-2008 text (.vox / .vox.gz) -> packed material+density volume.
 *
 * File semantics follow load_voxels (docker/mcgpu/MC-GPU_v1.3.cu:1996-2145): header located by
 * "[SECTION VOXELS", then "Nx Ny Nz", then "dx dy dz" [cm], body after "[END OF VXH SECTION",
 * one "<material> <density>" line per voxel with x running fastest, blank and '#' lines skipped,
 * 1 <= material <= 25 and density >= 1e-9 enforced, per-material maximum density recorded
 * (it sets the Woodcock majorant, H:2294).
 *
 * B200 layout (not the reference's float2 per voxel): cbctmc geometries are piecewise constant
 * (one density per material, cbctmc/mc/geometry.py:72-74), so the distinct (material, density)
 * pairs form a small palette.  The volume is stored as 4-, 8- or 16-bit palette indices
 * (Catphan604 500^3: 62.5 MB instead of 1.0 GB, i.e. L2-resident on B200), falling back to
 * 8 bytes per voxel only for geometries with more than 65536 distinct pairs.  The palette holds
 * the exact float density, so the arithmetic of the transport kernel is unchanged. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "mcgpu_host.h"

void mcgpu_free_volume(mcgpu_volume* v) {
  free(v->material);
  free(v->density);
  free(v->palette_density);
  free(v->palette_material);
  free(v->packed);
  memset(v, 0, sizeof *v);
}

static int alloc_volume(mcgpu_ctx* ctx, int nx, int ny, int nz, const float* size) {
  mcgpu_volume* v = &ctx->vol;
  size_t n;
  int k;
  mcgpu_free_volume(v);
  if (nx < 1 || ny < 1 || nz < 1 || (double)nx * ny * nz > 2147483647.0)
    return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: invalid number of voxels %d x %d x %d", nx, ny, nz);
  for (k = 0; k < 3; k++)
    if (!(size[k] > 0.0f)) return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: invalid voxel size %g", size[k]);
  v->nx = nx;
  v->ny = ny;
  v->nz = nz;
  for (k = 0; k < 3; k++) {
    v->voxel_size[k] = size[k];
    v->inv_voxel_size[k] = 1.0f / size[k];
  }
  v->size_bbox[0] = nx * size[0];
  v->size_bbox[1] = ny * size[1];
  v->size_bbox[2] = nz * size[2];
  n = (size_t)nx * ny * nz;
  v->material = (uint8_t*)malloc(n);
  v->density = (float*)malloc(n * sizeof(float));
  if (!v->material || !v->density) return mcgpu_fail(ctx, MCGPU_E_NOMEM, "load_voxels: not enough memory for %zu voxels", n);
  return MCGPU_OK;
}

int mcgpu_read_voxels(mcgpu_ctx* ctx, const char* path) {
  char line[MCGPU_LINE];
  int nx = 0, ny = 0, nz = 0, rc;
  float size[3] = {0.f, 0.f, 0.f};
  size_t n, i;
  gzFile f = gzopen(path, "rb");
  if (!f) return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: file '%s' does not exist", path);
  gzbuffer(f, 1 << 20);
  do {
    if (!gzgets(f, line, MCGPU_LINE)) {
      gzclose(f);
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: file does not contain the string '[SECTION VOXELS HEADER'");
    }
  } while (!strstr(line, "[SECTION VOXELS"));
  if (gzgets(f, line, MCGPU_LINE)) sscanf(line, "%d %d %d", &nx, &ny, &nz);
  if (gzgets(f, line, MCGPU_LINE)) sscanf(line, "%f %f %f", &size[0], &size[1], &size[2]);
  do {
    if (!gzgets(f, line, MCGPU_LINE)) {
      gzclose(f);
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: file does not contain the string '[END OF VXH SECTION]'");
    }
  } while (!strstr(line, "[END OF VXH SECTION"));
  if ((rc = alloc_volume(ctx, nx, ny, nz, size)) != MCGPU_OK) {
    gzclose(f);
    return rc;
  }
  n = (size_t)nx * ny * nz;
  for (i = 0; i < n; i++) {
    char* end;
    long m;
    float rho;
    do {
      if (!gzgets(f, line, MCGPU_LINE)) {
        gzclose(f);
        return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: premature end of file after %zu of %zu voxels", i, n);
      }
    } while (line[0] == '\n' || line[1] == '\n' || line[0] == '#' || line[1] == '#');
    m = strtol(line, &end, 10);
    if (end == line) {
      gzclose(f);
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: expecting material and density at voxel number %zu", i + 1);
    }
    {
      char* end2;
      rho = strtof(end, &end2);
      if (end2 == end) {
        gzclose(f);
        return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: expecting material and density at voxel number %zu", i + 1);
      }
    }
    if (m > MCGPU_MAX_MATERIALS || m < 1) {
      gzclose(f);
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: voxel material number %ld out of range [1,%d] at voxel number %zu", m, MCGPU_MAX_MATERIALS, i + 1);
    }
    if (rho < 1.0e-9f) {
      gzclose(f);
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "load_voxels: voxel density can not be 0 or negative: material %ld, density %f, voxel number %zu", m, rho, i + 1);
    }
    ctx->vol.material[i] = (uint8_t)m;
    ctx->vol.density[i] = rho;
  }
  gzclose(f);
  return mcgpu_finish_volume(ctx);
}

int mcgpu_set_voxels(mcgpu_ctx* ctx, int nx, int ny, int nz, float dx, float dy, float dz, const uint8_t* material, const float* density) {
  float size[3];
  size_t n, i;
  int rc;
  if (!ctx || !material || !density) return MCGPU_E_ARG;
  size[0] = dx;
  size[1] = dy;
  size[2] = dz;
  if ((rc = alloc_volume(ctx, nx, ny, nz, size)) != MCGPU_OK) return rc;
  n = (size_t)nx * ny * nz;
  for (i = 0; i < n; i++) {
    if (material[i] > MCGPU_MAX_MATERIALS || material[i] < 1)
      return mcgpu_fail(ctx, MCGPU_E_PARSE, "set_voxels: voxel material number %d out of range [1,%d] at voxel number %zu", material[i], MCGPU_MAX_MATERIALS, i + 1);
    if (density[i] < 1.0e-9f) return mcgpu_fail(ctx, MCGPU_E_PARSE, "set_voxels: voxel density can not be 0 or negative at voxel number %zu", i + 1);
  }
  memcpy(ctx->vol.material, material, n);
  memcpy(ctx->vol.density, density, n * sizeof(float));
  return mcgpu_finish_volume(ctx);
}

/* density_max per material, palette of distinct (material, density) pairs, packed indices */
int mcgpu_finish_volume(mcgpu_ctx* ctx) {
  mcgpu_volume* v = &ctx->vol;
  const size_t n = (size_t)v->nx * v->ny * v->nz;
  enum { HBITS = 18, HSIZE = 1 << HBITS, MAXPAL = 65536 };
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * HSIZE);
  int32_t* vals = (int32_t*)malloc(sizeof(int32_t) * HSIZE);
  uint16_t* idx = (uint16_t*)malloc(sizeof(uint16_t) * n);
  uint64_t* pal = (uint64_t*)malloc(sizeof(uint64_t) * MAXPAL);
  uint64_t last_key = ~0ull;
  int last_val = -1, npal = 0, overflow = 0, k, min_bits = 0;
  size_t i;
  if (!keys || !vals || !idx || !pal) {
    free(keys), free(vals), free(idx), free(pal);
    return mcgpu_fail(ctx, MCGPU_E_NOMEM, "load_voxels: not enough memory to pack %zu voxels", n);
  }
  for (k = 0; k < MCGPU_MAX_MATERIALS; k++) v->density_max[k] = -999.0f;
  memset(vals, 0xff, sizeof(int32_t) * HSIZE);
  for (i = 0; i < n; i++) {
    uint32_t bits;
    uint64_t key;
    const int m = v->material[i];
    const float rho = v->density[i];
    if (rho > v->density_max[m - 1]) v->density_max[m - 1] = rho;
    if (overflow) continue;
    memcpy(&bits, &rho, 4);
    key = ((uint64_t)m << 32) | bits;
    if (key != last_key) {
      uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> (64 - HBITS));
      while (vals[h] >= 0 && keys[h] != key) h = (h + 1) & (HSIZE - 1);
      if (vals[h] < 0) {
        if (npal == MAXPAL) {
          overflow = 1;
          continue;
        }
        keys[h] = key;
        vals[h] = npal;
        pal[npal++] = key;
      }
      last_key = key;
      last_val = vals[h];
    }
    idx[i] = (uint16_t)last_val;
  }
  free(v->palette_density), free(v->palette_material), free(v->packed);
  v->palette_density = NULL, v->palette_material = NULL, v->packed = NULL;
  { /* MCGPU_VOXEL_BITS=8|16|64 forces a wider packing than needed (parity tests of every kernel variant) */
    const char* force = getenv("MCGPU_VOXEL_BITS");
    min_bits = force ? atoi(force) : 0;
    if (min_bits >= 64) overflow = 1;
  }
  if (overflow) {
    mcgpu_f2* p = (mcgpu_f2*)malloc(sizeof(mcgpu_f2) * n);
    if (!p) {
      free(keys), free(vals), free(idx), free(pal);
      return mcgpu_fail(ctx, MCGPU_E_NOMEM, "load_voxels: not enough memory to pack %zu voxels", n);
    }
    /* (density, material0) pairs; material is remapped to a slot when the scene is built */
    for (i = 0; i < n; i++) {
      int m0 = v->material[i] - 1;
      p[i].x = v->density[i];
      memcpy(&p[i].y, &m0, 4);
    }
    v->packed = p;
    v->packed_bytes = sizeof(mcgpu_f2) * n;
    v->voxel_bits = 64;
    v->palette_size = 0;
  } else {
    v->palette_size = npal;
    v->palette_density = (float*)malloc(sizeof(float) * npal);
    v->palette_material = (uint8_t*)malloc(npal);
    for (k = 0; k < npal; k++) {
      uint32_t bits = (uint32_t)(pal[k] & 0xffffffffu);
      memcpy(&v->palette_density[k], &bits, 4);
      v->palette_material[k] = (uint8_t)(pal[k] >> 32);
    }
    if (npal <= 16 && min_bits <= 4) {
      uint8_t* p = (uint8_t*)calloc((n + 1) / 2, 1);
      for (i = 0; i < n; i++) p[i >> 1] |= (uint8_t)(idx[i] << ((i & 1) * 4));
      v->packed = p;
      v->packed_bytes = (n + 1) / 2;
      v->voxel_bits = 4;
    } else if (npal <= 256 && min_bits <= 8) {
      uint8_t* p = (uint8_t*)malloc(n);
      for (i = 0; i < n; i++) p[i] = (uint8_t)idx[i];
      v->packed = p;
      v->packed_bytes = n;
      v->voxel_bits = 8;
    } else {
      v->packed = idx;
      idx = NULL;
      v->packed_bytes = n * 2;
      v->voxel_bits = 16;
    }
  }
  free(keys), free(vals), free(idx), free(pal);
  ctx->have_voxels = 1;
  ctx->have_tables = 0;
  return MCGPU_OK;
}
