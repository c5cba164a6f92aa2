/* Internal types shared by the C host (csrc/host) and the CUDA layer (csrc/cuda).
 * Plain C; everything the kernel consumes is laid out here once and uploaded as-is. */
#ifndef MCGPU_HOST_H_
#define MCGPU_HOST_H_

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "mcgpu_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* limits of the reference (MC-GPU_v1.3.h:59-70) */
#define MCGPU_MAX_PROJECTIONS 1024
#define MCGPU_MAX_MATERIALS 25
#define MCGPU_MAX_SHELLS 40
#define MCGPU_NP_RAYLEIGH 128
#define MCGPU_MAX_ENERGYBINS_RAYLEIGH 25005
#define MCGPU_MAX_ENERGY_BINS 256
#define MCGPU_LINE 250
#define MCGPU_SCAN_STATS 8

typedef struct { float x, y; } mcgpu_f2;
typedef struct { float x, y, z; } mcgpu_f3;

/* One projection's source + detector pose: the union of the reference's source_struct and
 * detector_struct (MC-GPU_v1.3.h:155-208) minus fields the kernel never reads.  Passed to
 * the kernel by value (constant bank), 45 words. */
typedef struct mcgpu_view {
  float src_pos[3];
  float src_dir[3];
  float rot_fan[9];
  float cos_theta_low, phi_low, D_cos_theta, D_phi, max_height_at_y1cm;
  float det_center[3];
  float det_corner[3]; /* corner_min_rotated_to_Y */
  float rot_inv[9];
  float inv_pixel_size_X, inv_pixel_size_Z;
  float sdd, lateral_displacement, width_X, height_Z;
  int num_pixels_x, num_pixels_z, total_num_pixels, rotation_flag;
} mcgpu_view;

/* Energy spectrum sampled with Walker's alias method (MC-GPU_v1.3.h:173-183). */
typedef struct mcgpu_spectrum {
  int num_bins;
  float espc[MCGPU_MAX_ENERGY_BINS];
  float cutoff[MCGPU_MAX_ENERGY_BINS];
  short alias[MCGPU_MAX_ENERGY_BINS];
  float mean_energy;
} mcgpu_spectrum;

/* Everything parsed from the .in file (read_input, H:1240-1895). */
typedef struct mcgpu_input {
  unsigned long long total_histories;
  int seed_input, gpu_id, threads_per_block, histories_per_thread;
  char file_espc[MCGPU_LINE], file_output[MCGPU_LINE], file_voxels[MCGPU_LINE], file_dose[MCGPU_LINE];
  char file_materials[MCGPU_MAX_MATERIALS][MCGPU_LINE];
  int num_projections, enable_specific_angles, num_specific_angles;
  float specific_angles[MCGPU_MAX_PROJECTIONS];
  double D_angle, angularROI_0, angularROI_1, initial_angle, SRotAxisD, vertical_translation;
  double phi1_deg, phi2_deg, theta_deg; /* final apertures, for the banner (H:1468) */
  int flag_material_dose, flag_voxel_dose;
  short dose_roi[6]; /* x_min,x_max,y_min,y_max,z_min,z_max (0-based) */
} mcgpu_input;

/* Voxel volume in host memory, reference semantics (H:1996-2145) but stored split. */
typedef struct mcgpu_volume {
  int nx, ny, nz;
  float voxel_size[3];
  float inv_voxel_size[3]; /* 1.0f/size (H:2073-2075) */
  float size_bbox[3];      /* n*size    (H:2047-2049) */
  uint8_t* material;       /* 1-based */
  float* density;
  float density_max[MCGPU_MAX_MATERIALS]; /* -999 when the material is absent (H:2094-2095) */
  /* palette built at load time: distinct (material, density) pairs */
  int palette_size;        /* 0 when more than 65536 distinct pairs */
  int voxel_bits;          /* 4, 8, 16 or 64 */
  float* palette_density;
  uint8_t* palette_material;
  void* packed;            /* palette indices (4/8/16 bit) or float2 pairs */
  size_t packed_bytes;
} mcgpu_volume;

/* Tables in the REFERENCE layout (for parity tests and for deriving the device layout). */
typedef struct mcgpu_tables {
  int num_values;
  float e0, ide;
  double delta_e;
  float density_nominal[MCGPU_MAX_MATERIALS];
  int material_loaded[MCGPU_MAX_MATERIALS];
  mcgpu_f2* woodcock;                 /* [nE] */
  mcgpu_f3* mfp_a;                    /* [nE*25] */
  mcgpu_f3* mfp_b;                    /* [nE*25] */
  float* ray_xco; float* ray_pco; float* ray_aco; float* ray_bco; /* [128*25] */
  uint8_t* ray_itlco; uint8_t* ray_ituco;                          /* [128*25] */
  float* ray_pmax;                    /* [(nE+1)*25], zero-initialised (Q3) */
  float cmp_fco[MCGPU_MAX_MATERIALS * MCGPU_MAX_SHELLS];
  float cmp_uico[MCGPU_MAX_MATERIALS * MCGPU_MAX_SHELLS];
  float cmp_fj0[MCGPU_MAX_MATERIALS * MCGPU_MAX_SHELLS];
  int cmp_noscco[MCGPU_MAX_MATERIALS];
} mcgpu_tables;

/* Device-side layout, compacted to the materials present ("slots"). */
typedef struct mcgpu_mfp_record { /* 32 B = one L2 sector: a (total, Compton, Rayleigh) then b */
  float ax, ay, az, bx, by, bz, pmax_next, pad;
} mcgpu_mfp_record;

typedef struct mcgpu_scene {
  int num_slots;
  int slot_of_material[MCGPU_MAX_MATERIALS]; /* material0 (0-based) -> slot, -1 if absent */
  int material_of_slot[MCGPU_MAX_MATERIALS];
  int num_values;
  float e0, ide;
  mcgpu_mfp_record* mfp;  /* [nE][num_slots] */
  mcgpu_f2* woodcock;     /* [nE] */
  float* ray_xpab;        /* [num_slots][128][4] = xco, pco, aco, bco */
  uint8_t* ray_itl_itu;   /* [num_slots][128][2] */
  float* cmp_shells;      /* [num_slots][40][4] = fco, uico, fj0, uico*510998.918f */
  int cmp_noscco[MCGPU_MAX_MATERIALS]; /* per slot */
  /* palette remapped to slots */
  int palette_size, voxel_bits;
  mcgpu_f2* palette;      /* [palette_size] = (density, slot as int bits) */
  /* optional dose tallies (SECTION DOSE DEPOSITION, H:1619-1709), ROI clipped to the volume (H:2057-2065) */
  int tally_material_dose, tally_voxel_dose;
  int dose_roi[6];        /* x_min,x_max,y_min,y_max,z_min,z_max, 0-based inclusive */
  long long dose_roi_voxels;
} mcgpu_scene;

struct mcgpu_device; /* opaque, csrc/cuda/device.cu */

struct mcgpu_ctx {
  char err[512];
  int verbose;
  int have_input, have_voxels, have_tables;
  mcgpu_input in;
  mcgpu_spectrum spc;
  mcgpu_view* views; /* [num_projections] */
  mcgpu_volume vol;
  mcgpu_tables tab;
  mcgpu_scene scene;
  /* launch state */
  int hpt_current;   /* sticky histories_per_thread (H:833) */
  unsigned long long hist_current; /* sticky history count: the reference overwrites total_histories with the launched count (H:841),
                                      which the NEXT projection's grid rule starts from; 0 = the .in value */
  int num_devices;
  struct mcgpu_device** dev;
  void* opening;     /* device opens still running in the background (mcgpu_create); joined by mcgpu_devices_ready */
  int n_opening;
  double last_kernel_ms;
  double last_reduce_ms;            /* history-split runs: device time of the reduction of the partial images */
  struct mcgpu_reducer* reducer;    /* created on the first multi-device projection, for reducer_devices devices */
  int reducer_devices;
  int fast_math;
  double scan_stats[MCGPU_SCAN_STATS]; /* mcgpu_run_all: summed over the devices' host threads, see mcgpu_get_scan_stats */
};

/* ---- host stages (each returns MCGPU_OK or an error code, message in ctx->err) ---------- */
int mcgpu_parse_input(mcgpu_ctx* ctx, const char* in_path);           /* input.c */
int mcgpu_read_spectrum(mcgpu_ctx* ctx, const char* path);            /* tables.c */
int mcgpu_build_views(mcgpu_ctx* ctx);                                /* geometry.c */
int mcgpu_read_voxels(mcgpu_ctx* ctx, const char* path);              /* voxels.c */
int mcgpu_finish_volume(mcgpu_ctx* ctx);                              /* voxels.c: density_max, palette, packing */
int mcgpu_read_materials(mcgpu_ctx* ctx, const char* const* paths, int n); /* tables.c */
int mcgpu_build_scene(mcgpu_ctx* ctx);                                /* tables.c */
void mcgpu_free_volume(mcgpu_volume* v);
void mcgpu_free_tables(mcgpu_tables* t);
void mcgpu_free_scene(mcgpu_scene* s);
int mcgpu_fail(mcgpu_ctx* ctx, int code, const char* fmt, ...);
void mcgpu_fail_into(char* buf, size_t len); /* failures raised on the calling thread go to buf instead of ctx->err (NULL: back to ctx->err) */
void mcgpu_devices_ready(mcgpu_ctx* ctx); /* api.c: join the background device opens; call before touching ctx->dev / num_devices */
/* grid of the projection(s) simulated last (or of the first one before any run): sticky hpt AND sticky history count */
void mcgpu_current_grid(const mcgpu_ctx* ctx, int* hpt, int* blocks, unsigned long long* launched);
char* mcgpu_fgets_trimmed(char* out, int num, FILE* f);
void mcgpu_trim_name(const char* line, char* name);

/* ---- CUDA layer (csrc/cuda/device.cu) ---------------------------------------------------- */
typedef struct mcgpu_launch {
  int histories_per_thread;
  int seed_input;
  int threads_per_block;
  long long stream_begin, stream_end; /* reference global thread ids */
  int zero_image;
  int image_slot; /* 0 = the device's image, 1 = the second image of the pipelined scan */
} mcgpu_launch;

int mcgpu_dev_count(void);
struct mcgpu_device* mcgpu_dev_open(int ordinal, char* err, size_t errlen);
void mcgpu_dev_close(struct mcgpu_device* d);
int mcgpu_dev_ordinal(const struct mcgpu_device* d);
int mcgpu_dev_upload(struct mcgpu_device* d, const mcgpu_scene* s, const mcgpu_volume* v, const mcgpu_spectrum* spc,
                     int npix_total, char* err, size_t errlen);
int mcgpu_dev_launch(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen);
int mcgpu_dev_sync(struct mcgpu_device* d, float* kernel_ms, char* err, size_t errlen);
int mcgpu_dev_fetch(struct mcgpu_device* d, uint64_t* host, char* err, size_t errlen);
/* history-split reduction: sum of the devices' images on the first one -- ncclReduce over NVLink, or one kernel reading all peers */
struct mcgpu_reducer;
struct mcgpu_reducer* mcgpu_dev_reducer_create(struct mcgpu_device** devs, int n, char* err, size_t errlen);
void mcgpu_dev_reducer_free(struct mcgpu_reducer* r);
const char* mcgpu_dev_reducer_kind(const struct mcgpu_reducer* r);
int mcgpu_dev_reduce(struct mcgpu_reducer* r, float* reduce_ms, char* err, size_t errlen);
void* mcgpu_dev_image_ptr(struct mcgpu_device* d);
/* exhaustive device checks of the kernel's arithmetic shortcuts against the CUDA functions they replace (launch.cu) */
int mcgpu_dev_selftest(struct mcgpu_device* d, const char* name, unsigned long long* mismatches, char* err, size_t errlen);
/* pipelined scan: launch projection work into image slot 0/1 (asynchronous; the copy to a pinned host buffer is queued behind
 * it on a second stream), wait for a slot's copy (kernel_ms = device time of its kernel, *host = the pinned buffer) */
int mcgpu_dev_pipeline_begin(struct mcgpu_device* d, char* err, size_t errlen);
int mcgpu_dev_pipeline_launch(struct mcgpu_device* d, const mcgpu_view* view, const mcgpu_launch* l, char* err, size_t errlen);
int mcgpu_dev_pipeline_wait(struct mcgpu_device* d, int slot, float* kernel_ms, uint64_t** host, char* err, size_t errlen);
void mcgpu_dev_pipeline_end(struct mcgpu_device* d);
void mcgpu_dev_set_fast_math(struct mcgpu_device* d, int on);
/* dose tallies accumulate over launches until reset; fetch ADDS the device's counters to the host arrays */
int mcgpu_dev_reset_dose(struct mcgpu_device* d, char* err, size_t errlen);
int mcgpu_dev_add_dose(struct mcgpu_device* d, uint64_t* materials_2x25, uint64_t* voxels_2xroi, char* err, size_t errlen);
int mcgpu_write_dose_files(mcgpu_ctx* ctx, const uint64_t* voxels_edep, double seconds, int projections); /* dose.c */
int mcgpu_print_materials_dose(mcgpu_ctx* ctx, const uint64_t* materials_dose, int projections);         /* dose.c */

#ifdef __cplusplus
}
#endif
#endif
