// Device-side description of a loaded scene (global namespace: shared by the exact and the
// fast-math builds of the kernels and by the device management code).
#pragma once
#include <cuda_runtime.h>

#include "../host/mcgpu_host.h"

// ------------------------------------------------------------------------------------------
// Device-side description of a loaded scene; passed to kernels by value (constant bank).
struct McgpuSceneDev {
  const void* volume;             // packed voxels (4/8/16-bit palette indices or float2 pairs)
  const float2* palette;          // (density, slot-as-int-bits), global copy
  const mcgpu_mfp_record* mfp;    // [nE][num_slots]
  const float2* woodcock;         // [nE]
  const float4* ray_xpab;         // [num_slots][128] (xco, pco, aco, bco)
  const uchar2* ray_itl_itu;      // [num_slots][128]
  const float4* cmp_shells;       // [num_slots][40] (fco, uico, fj0, uico*510998.918f)
  const mcgpu_spectrum* spectrum; // global copy
  unsigned long long* image;      // [4][Npix]
  int cmp_noscco[MCGPU_MAX_MATERIALS];
  int num_slots, palette_size, num_values;
  int max_shells;  // largest cmp_noscco over the slots in use (sizes the per-warp shell scratch)
  // optional dose tallies (K:357-369), both NULL in every cbctmc run (mcgpu_input.jinja2:37-38)
  unsigned long long* materials_dose;  // [25][2]: sum of round(Edep*100), sum of round(Edep^2), indexed by material0
  unsigned long long* voxels_edep;     // [ROI voxels][2], same two sums
  int dose_roi[6];                     // x_min, x_max, y_min, y_max, z_min, z_max (0-based, inclusive)
  int material_of_slot[MCGPU_MAX_MATERIALS];
  int nvx, nvy, nvz;
  float inv_voxel[3];
  float bbox[3];
  unsigned box_hi[3];  // bits(bbox[k] - EPS_SOURCE) - bits(EPS_SOURCE): right-hand sides of locate_voxel_fast's unsigned comparisons
  float e0, ide;
};

