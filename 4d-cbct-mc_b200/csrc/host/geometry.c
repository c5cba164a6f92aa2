/* Source / detector pose of every projection.
 *
 * Restates the geometric part of read_input (docker/mcgpu/MC-GPU_v1.3.cu:1337-1395, 1437-1465,
 * 1543-1583, 1750-1841) and set_CT_trajectory (H:3280-3434) with the operation order and the
 * precision of every intermediate kept, because these floats feed the transport kernel and a
 * single differing rounding breaks bit-exact tallies.
 *
 * Precision note: the production reference is compiled by nvcc, i.e. as C++, where
 * acos/sqrt/atan2 applied to float arguments resolve to the float overloads.  This file is C,
 * so those calls are spelled acosf/sqrtf/atan2f explicitly wherever the reference's argument
 * is a float expression (H:1753, 1759-1761, 1823-1824, 3372-3374).  The reference's CPU build
 * (plain C) uses the double versions there; that variant is what oracle/ restates. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mcgpu_host.h"

#define PI 3.14159265358979323846
/* the reference's RAD2DEG / DEG2RAD macros expand without parentheses (MC-GPU_v1.3.h:73-75):
 * x*RAD2DEG == (x*180.0)/PI and x*DEG2RAD == (x*PI)/180.0 */
#define TO_DEG(x) ((x) * 180.0 / PI)
#define TO_RAD(x) ((x) * PI / 180.0)

static void fill_rot_inv(float* r, double rotX, double rotZ) { /* H:1767-1781, 3381-3393 */
  double cX = cos(rotX), cZ = cos(rotZ), sX = sin(rotX), sZ = sin(rotZ);
  r[0] = cZ;
  r[1] = -sZ;
  r[2] = 0.0f;
  r[3] = cX * sZ;
  r[4] = cX * cZ;
  r[5] = -sX;
  r[6] = sX * sZ;
  r[7] = sX * cZ;
  r[8] = cX;
}

static void fill_rot_fan(float* r, double rotX, double rotZ) { /* H:1825-1838, 3408-3421 */
  double cX = cos(rotX), cZ = cos(rotZ), sX = sin(rotX), sZ = sin(rotZ);
  r[0] = cZ;
  r[1] = -cX * sZ;
  r[2] = sX * sZ;
  r[3] = sZ;
  r[4] = cX * cZ;
  r[5] = -sX * cZ;
  r[6] = 0.0f;
  r[7] = sX;
  r[8] = cX;
}

/* rotation about Z that brings the beam direction (u,v) to +Y (H:1755-1764, 3370-3376) */
static double rot_z_to_plus_y(float u, float v) {
  if ((u * u + v * v) > 1.0e-8) {
    float c = acosf(u / sqrtf(u * u + v * v));
    if (v >= 0.0f) return 0.5 * PI - c;
    return 0.5 * PI - (-c);
  }
  return 0.0;
}

static void corner_from_center(mcgpu_view* v, const float* c) { /* H:1800-1810, 3396-3403 */
  const float* r = v->rot_inv;
  v->det_corner[0] = c[0] * r[0] + c[1] * r[1] + c[2] * r[2];
  v->det_corner[1] = c[0] * r[3] + c[1] * r[4] + c[2] * r[5];
  v->det_corner[2] = c[0] * r[6] + c[1] * r[7] + c[2] * r[8];
}

int mcgpu_build_views(mcgpu_ctx* ctx) {
  mcgpu_input* in = &ctx->in;
  mcgpu_view* v = &ctx->views[0];
  double n, theta = in->theta_deg, phi1 = in->phi1_deg, phi2 = in->phi2_deg, phi = phi1 + phi2;
  float center[3];
  int i;

  /* -- normalise the beam direction (H:1338-1341) */
  n = 1.0 / sqrt((double)(v->src_dir[0] * v->src_dir[0] + v->src_dir[1] * v->src_dir[1] + v->src_dir[2] * v->src_dir[2]));
  for (i = 0; i < 3; i++) v->src_dir[i] = (float)(((double)v->src_dir[i]) * n);

  /* -- rectangular beam centred on (0,1,0) (H:1374-1395) */
  v->cos_theta_low = (float)(cos(TO_RAD(90.0 - 0.5 * theta)));
  v->D_cos_theta = (float)(-2.0 * v->cos_theta_low);
  v->phi_low = (float)(TO_RAD(90.0 - phi1));
  v->D_phi = (float)(TO_RAD(phi));
  v->max_height_at_y1cm = (float)(tan(TO_RAD(0.5 * theta)));
  if (fabs(theta) < 1.0e-7) {
    theta = +1.00e-7;
    v->cos_theta_low = 0.0f;
    v->D_cos_theta = 0.0f;
    v->max_height_at_y1cm = 0.0f;
  }
  if (fabs(phi) < 1.0e-7) {
    phi = +1.00e-7;
    v->phi_low = (float)(TO_RAD(90.0));
    v->D_phi = 0.0f;
  }

  /* -- pixel pitch and detector centre (H:1429-1440) */
  v->inv_pixel_size_X = v->num_pixels_x / v->width_X;
  v->inv_pixel_size_Z = v->num_pixels_z / v->height_Z;
  for (i = 0; i < 3; i++) center[i] = v->src_pos[i] + v->src_dir[i] * v->sdd;

  /* -- negative apertures: cover exactly the detector (H:1451-1465; the phi branch keeps D_phi, Q8) */
  if (phi < -1.0e-7) {
    phi1 = TO_DEG(atan((v->width_X / 2.0) / v->sdd));
    phi2 = TO_DEG(atan((v->width_X / 2.0) / v->sdd));
    v->phi_low = (float)(TO_RAD(90.0 - phi1));
    v->D_phi = (float)(TO_RAD(phi));
  }
  if (theta < -1.0e-7) {
    theta = TO_DEG(2.0 * atan(0.5 * v->height_Z / (v->sdd)));
    v->cos_theta_low = (float)(cos(TO_RAD(90.0 - 0.5 * theta)));
    v->D_cos_theta = (float)(-2.0 * v->cos_theta_low);
    v->max_height_at_y1cm = (float)(tan(TO_RAD(0.5 * theta)));
  }
  in->phi1_deg = phi1;
  in->phi2_deg = phi2;
  in->theta_deg = theta;

  /* -- CT scans need a beam perpendicular to Z (H:1543-1550) */
  if (abs(in->num_projections) > 1 && fabsf(v->src_dir[2]) > 0.00001f)
    return mcgpu_fail(ctx, MCGPU_E_PARSE, "read_input: CT scans can only be simulated when the source direction is perpendicular to the Z axis (w=0)");

  /* -- angle of projection 0 (H:1561-1583); D_angle and the ROI are stored in radians (H:1567, 1597-1598) */
  if (in->num_projections != 1 || in->enable_specific_angles == 1) {
    double a;
    in->D_angle = TO_RAD(in->D_angle);
    a = acos((double)(v->src_dir[0]));
    if (v->src_dir[1] < 0) a = -a;
    if (a < 0.0) a = a + 2.0 * PI;
    a = a - PI;
    if (a < 0.0) a = a + 2.0 * PI;
    if (in->enable_specific_angles == 1) {
      a = TO_RAD(in->specific_angles[0]);
      if (a >= (2.0 * PI - 0.0001)) a -= 2.0 * PI;
    }
    in->initial_angle = a;
    in->angularROI_0 = TO_RAD(in->angularROI_0 - 0.00001);
    in->angularROI_1 = TO_RAD(in->angularROI_1 + 0.00001);
  }

  /* -- detector rotation to +Y for projection 0 (H:1750-1814) */
  {
    double rotX = acosf(v->src_dir[2]) - 0.5 * PI;
    double rotZ = rot_z_to_plus_y(v->src_dir[0], v->src_dir[1]);
    fill_rot_inv(v->rot_inv, rotX, rotZ);
  }
  if (v->src_dir[1] > 0.99999f && in->num_projections == 1) {
    v->rotation_flag = 0;
    for (i = 0; i < 3; i++) v->det_corner[i] = center[i];
  } else {
    v->rotation_flag = 1;
    corner_from_center(v, center);
  }
  v->det_corner[0] = v->det_corner[0] - 0.5 * v->width_X;
  v->det_corner[2] = v->det_corner[2] - 0.5 * v->height_Z;
  for (i = 0; i < 3; i++) v->det_center[i] = center[i];

  /* -- fan-beam rotation for projection 0 (H:1820-1841); identity storage is never read when flag==0 */
  if (v->rotation_flag == 1) {
    double rotX = 0.5 * PI - acosf(v->src_dir[2]);
    double rotZ = atan2f(v->src_dir[1], v->src_dir[0]) - 0.5 * PI;
    fill_rot_fan(v->rot_fan, rotX, rotZ);
  }

  /* -- remaining projections (set_CT_trajectory, H:3280-3434); only when num_projections != 1 (H:548) */
  if (in->num_projections != 1) {
    const double R = in->SRotAxisD;
    float rot_center[3];
    double angle;
    rot_center[0] = v->src_pos[0] + v->src_dir[0] * R;
    rot_center[1] = v->src_pos[1] + v->src_dir[1] * R;
    rot_center[2] = v->src_pos[2];
    if (in->enable_specific_angles == 0) {
      angle = acos((double)v->src_dir[0]);
      if (v->src_dir[1] < 0) angle = -angle;
      if (angle < 0.0) angle += 2.0 * PI;
      angle = angle - PI;
      if (angle < 0.0) angle += 2.0 * PI;
    } else {
      angle = TO_RAD(in->specific_angles[0]);
      if (angle >= (2.0 * PI - 0.0001)) angle -= 2.0 * PI;
    }
    for (i = 1; i < in->num_projections; i++) {
      mcgpu_view* w = &ctx->views[i];
      double norm, rotZ;
      *w = *v; /* constant members (H:3316-3329); pose members are overwritten below */
      if (in->enable_specific_angles) {
        angle = TO_RAD(in->specific_angles[i]);
        if (angle >= (2.0 * PI - 0.0001)) angle -= 2.0 * PI;
      } else {
        angle += in->D_angle;
        if (angle >= (2.0 * PI - 0.0001)) angle -= 2.0 * PI;
      }
      w->src_pos[0] = rot_center[0] + R * cos(angle);
      w->src_pos[1] = rot_center[1] + R * sin(angle);
      w->src_pos[2] = ctx->views[i - 1].src_pos[2] + in->vertical_translation;
      w->src_dir[0] = rot_center[0] - w->src_pos[0];
      w->src_dir[1] = rot_center[1] - w->src_pos[1];
      w->src_dir[2] = 0.0f;
      norm = 1.0 / sqrt((double)w->src_dir[0] * (double)w->src_dir[0] + (double)w->src_dir[1] * (double)w->src_dir[1]);
      w->src_dir[0] = (float)(((double)w->src_dir[0]) * norm);
      w->src_dir[1] = (float)(((double)w->src_dir[1]) * norm);
      w->det_center[0] = w->src_pos[0] + w->src_dir[0] * w->sdd;
      w->det_center[1] = w->src_pos[1] + w->src_dir[1] * w->sdd;
      w->det_center[2] = w->src_pos[2];
      rotZ = rot_z_to_plus_y(w->src_dir[0], w->src_dir[1]);
      fill_rot_inv(w->rot_inv, 0.0, rotZ);
      corner_from_center(w, w->det_center);
      w->det_corner[0] = w->det_corner[0] - 0.5 * w->width_X;
      w->det_corner[2] = w->det_corner[2] - 0.5 * w->height_Z;
      fill_rot_fan(w->rot_fan, 0.0, -rotZ);
    }
  }
  return MCGPU_OK;
}
