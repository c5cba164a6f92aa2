/* Dose tally reports: the voxel-dose files of report_voxels_dose (docker/mcgpu/MC-GPU_v1.3.cu:
 * 2976-3200: '<name>' ASCII plane + '<name>.raw' / '<name>_2sigma.raw' float32 volumes) and the
 * per-material table of report_materials_dose (H:3214-3263).  Both tallies are off in every cbctmc
 * run (mcgpu_input.jinja2:37-38); they exist so that a .in that enables them keeps working. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mcgpu_host.h"

#define SCALE_eV 100.0f

static unsigned long long launched_per_projection(mcgpu_ctx* ctx) {
  int hpt, blocks;
  unsigned long long launched;
  mcgpu_current_grid(ctx, &hpt, &blocks, &launched);
  return launched;
}

int mcgpu_write_dose_files(mcgpu_ctx* ctx, const uint64_t* edep, double seconds, int projections) {
  const mcgpu_scene* sc = &ctx->scene;
  const mcgpu_volume* v = &ctx->vol;
  const int* r = sc->dose_roi;
  const int DX = r[1] - r[0] + 1, DY = r[3] - r[2] + 1, DZ = r[5] - r[4] + 1;
  const unsigned long long total_histories = launched_per_projection(ctx);
  char name_mean[MCGPU_LINE + 16], name_sigma[MCGPU_LINE + 16];
  FILE *f, *fm, *fs;
  int z_plane, z_plane_roi, i, j, k, voxel = 0;
  double max_dose = -1.0, max_dose_sd = -1.0;
  int max_geom = 0, max_x = -1, max_y = -1, max_z = -1;
  unsigned long long total_edep = 0;
  const double inv_scale = 1.0 / SCALE_eV, inv_n = 1.0 / (double)(total_histories * ((unsigned long long)projections));
  const double voxel_volume = 1.0 / (((double)v->inv_voxel_size[0]) * ((double)v->inv_voxel_size[1]) * ((double)v->inv_voxel_size[2]));
  double mat_edep[MCGPU_MAX_MATERIALS] = {0}, mat_edep2[MCGPU_MAX_MATERIALS] = {0}, mat_mass[MCGPU_MAX_MATERIALS] = {0};
  unsigned int mat_voxels[MCGPU_MAX_MATERIALS] = {0};

  if (!sc->tally_voxel_dose) return MCGPU_OK;
  snprintf(name_mean, sizeof name_mean, "%s.raw", ctx->in.file_dose);
  snprintf(name_sigma, sizeof name_sigma, "%s_2sigma.raw", ctx->in.file_dose);
  f = fopen(ctx->in.file_dose, "w");
  fm = fopen(name_mean, "w");
  fs = fopen(name_sigma, "w");
  if (!f || !fm || !fs) {
    if (f) fclose(f);
    if (fm) fclose(fm);
    if (fs) fclose(fs);
    return mcgpu_fail(ctx, MCGPU_E_OUTPUT, "report_voxels_dose: file %s can not be opened", ctx->in.file_dose);
  }
  z_plane = (int)(ctx->views[0].src_pos[2] * v->inv_voxel_size[2] + 0.00001f);
  if ((z_plane < r[4]) || (z_plane > r[5])) z_plane = (r[5] + r[4]) / 2;
  z_plane_roi = z_plane - r[4];

  fprintf(f, "# \n");
  fprintf(f, "#     *****************************************************************************\n");
  fprintf(f, "#     ***         MC-GPU, version 1.3 (http://code.google.com/p/mcgpu/)         ***\n");
  fprintf(f, "#     ***                                                                       ***\n");
  fprintf(f, "#     ***                     Andreu Badal (Andreu.Badal-Soler@fda.hhs.gov)     ***\n");
  fprintf(f, "#     *****************************************************************************\n");
  fprintf(f, "# \n");
  fprintf(f, "#  *** SIMULATION IN THE GPU USING CUDA ***\n");
  fprintf(f, "#\n");
  fprintf(f, "#\n");
  fprintf(f, "#  3D dose deposition map (and dose uncertainty) created tallying the energy deposited by photons inside each voxel of the input geometry.\n");
  fprintf(f, "#  Electrons were not transported and therefore we are approximating that the dose is equal to the KERMA (energy released by the photons alone).\n");
  fprintf(f, "#  This approximation is acceptable when there is electronic equilibrium and when the range of the secondary electrons is shorter than the voxel size.\n");
  fprintf(f, "#  Usually the doses will be acceptable for photon energies below 1 MeV. The dose estimates may not be accurate at the interface of low density volumes.\n");
  fprintf(f, "#\n");
  fprintf(f, "#  The 3D dose deposition is reported in binary form in the .raw files (data given as 32-bit floats). \n");
  fprintf(f, "#  To reduce the memory use and the reporting time this text output reports only the 2D dose at the Z plane at the level\n");
  fprintf(f, "#  of the source focal spot: z_coord = %d (z_coord in ROI = %d)\n", z_plane, z_plane_roi);
  fprintf(f, "#\n");
  fprintf(f, "#  The total dose deposited in each different material is reported to the standard output.\n");
  fprintf(f, "#  The dose is calculated adding the energy deposited in the individual voxels within the dose ROI and dividing by the total mass of the material in the ROI.\n");
  fprintf(f, "#\n");
  fprintf(f, "#\n");
  fprintf(f, "#  Voxel size:  %lf x %lf x %lf = %lf cm^3\n", 1.0 / (double)(v->inv_voxel_size[0]), 1.0 / (double)(v->inv_voxel_size[1]), 1.0 / (double)(v->inv_voxel_size[2]),
          1.0 / (double)(v->inv_voxel_size[0] * v->inv_voxel_size[1] * v->inv_voxel_size[2]));
  fprintf(f, "#  Number of voxels in the reported region of interest (ROI) X, Y and Z:\n");
  fprintf(f, "#      %d  %d  %d\n", DX, DY, DZ);
  fprintf(f, "#  Coordinates of the ROI inside the voxel volume = X[%d,%d], Y[%d,%d], Z[%d,%d]\n", r[0] + 1, r[1] + 1, r[2] + 1, r[3] + 1, r[4] + 1, r[5] + 1);
  fprintf(f, "#\n");
  fprintf(f, "#  Voxel dose units: eV/g per history\n");
  fprintf(f, "#  X rows given first, then Y, then Z. One blank line separates the different Y, and two blanks the Z values (GNUPLOT format).\n");
  fprintf(f, "#  The dose distribution is also reported with binary FLOAT values (.raw file) for easy visualization in ImageJ.\n");
  fprintf(f, "# \n");
  fprintf(f, "#    [DOSE]   [2*standard_deviation]\n");
  fprintf(f, "# =====================================\n");

  for (k = 0; k < DZ; k++) {
    for (j = 0; j < DY; j++) {
      for (i = 0; i < DX; i++) {
        const size_t geom = (size_t)(i + r[0]) + (size_t)(j + r[2]) * v->nx + (size_t)(k + r[4]) * v->nx * v->ny;
        const float rho = v->density[geom];
        const int mat = v->material[geom] - 1;
        const double inv_mass = 1.0 / (rho * voxel_volume);
        const uint64_t e1 = edep[2 * (size_t)voxel], e2 = edep[2 * (size_t)voxel + 1];
        double dose, sd;
        float dose_f, sigma_f;
        mat_mass[mat] += rho * voxel_volume;
        mat_edep[mat] += (double)e1;
        mat_edep2[mat] += (double)e2;
        mat_voxels[mat]++;
        dose = ((double)e1) * inv_n * inv_mass * inv_scale;
        total_edep += e1;
        sd = (((double)e2) * inv_n * inv_scale * inv_mass - dose * dose) * inv_n;
        if (sd > 0.0) sd = sqrt(sd);
        if (dose > max_dose) {
          max_dose = dose;
          max_dose_sd = sd;
          max_x = i + r[0], max_y = j + r[2], max_z = k + r[4];
          max_geom = (int)geom;
        }
        if (k == z_plane_roi) fprintf(f, "%.6lf %.6lf\n", dose, 2.0 * sd);
        dose_f = (float)dose;
        sigma_f = 2.0f * (float)(sd);
        fwrite(&dose_f, sizeof(float), 1, fm);
        fwrite(&sigma_f, sizeof(float), 1, fs);
        voxel++;
      }
      if (k == z_plane_roi) fprintf(f, "\n");
    }
    if (k == z_plane_roi) fprintf(f, "\n");
  }
  fprintf(f, "#   ****** DOSE REPORT: TOTAL SIMULATION PERFORMANCE FOR ALL PROJECTIONS ******\n");
  fprintf(f, "#       Total number of simulated x rays: %lld\n", total_histories * ((unsigned long long)projections));
  fprintf(f, "#       Simulated x rays per projection:  %lld\n", total_histories);
  fprintf(f, "#       Total simulation time [s]:  %.2f\n", seconds);
  if (seconds > 0.000001) fprintf(f, "#       Total speed [x-rays/s]:  %.2f\n", (double)(total_histories * ((unsigned long long)projections)) / seconds);
  fprintf(f, "\n#       Total energy absorved inside the dose ROI: %.5lf keV/hist\n\n", 0.001 * ((double)total_edep) * inv_n * inv_scale);
  fclose(f);
  fclose(fm);
  fclose(fs);

  if (ctx->verbose) {
    const double mass_max = voxel_volume * v->density[max_geom];
    printf("\n\n          *** VOXEL ROI DOSE TALLY REPORT ***\n\n");
    printf("              Total energy absorved inside the dose deposition ROI: %.5lf keV/hist\n", 0.001 * ((double)total_edep) * inv_n * inv_scale);
    printf("              Maximum voxel dose (+-2 sigma): %lf +- %lf eV/g per history (E_dep_voxel=%lf eV/hist)\n", max_dose, max_dose_sd, (max_dose * mass_max));
    printf("              for the voxel: material=%d, density=%.8f g/cm^3, voxel_mass=%.8lf g, voxel coord in geometry=(%d,%d,%d)\n\n", (int)v->material[max_geom],
           v->density[max_geom], mass_max, max_x, max_y, max_z);
    printf("    [MATERIAL]  [DOSE_ROI, eV/g/hist]  [2*std_dev]  [Rel error 2*std_dev, %%]  [E_dep [eV/hist]  [MASS_ROI, g]  [NUM_VOXELS_ROI]\n");
    for (i = 0; i < MCGPU_MAX_MATERIALS; i++)
      if (mat_voxels[i] > 0) {
        const double e = mat_edep[i] * inv_n * inv_scale;
        double sd = (mat_edep2[i] * inv_n - e * e) * inv_n, dose, rel = 0.0;
        if (sd > 0.0) sd = sqrt(sd);
        dose = e / mat_mass[i];
        sd = sd / mat_mass[i];
        if (dose > 0.0) rel = sd / dose;
        printf("\t%d\t%.5lf\t\t%.5lf\t\t%.2lf\t\t%.2lf\t\t%.5lf\t%u\n", (i + 1), dose, 2.0 * sd, (2.0 * 100.0 * rel), e, mat_mass[i], mat_voxels[i]);
      }
    printf("\n");
    fflush(stdout);
  }
  return MCGPU_OK;
}

int mcgpu_print_materials_dose(mcgpu_ctx* ctx, const uint64_t* md, int projections) {
  const mcgpu_volume* v = &ctx->vol;
  const unsigned long long total_histories = launched_per_projection(ctx);
  const double inv_n = 1.0 / (double)(total_histories * ((unsigned long long)projections));
  const double voxel_volume = 1.0 / (((double)v->inv_voxel_size[0]) * ((double)v->inv_voxel_size[1]) * ((double)v->inv_voxel_size[2]));
  double mass[MCGPU_MAX_MATERIALS] = {0};
  const size_t n = (size_t)v->nx * v->ny * v->nz;
  size_t kk;
  int i;
  if (!ctx->scene.tally_material_dose || !ctx->verbose) return MCGPU_OK;
  for (kk = 0; kk < n; kk++) mass[v->material[kk] - 1] += ((double)v->density[kk]) * voxel_volume; /* H:579-585 */
  printf("\n\n          *** MATERIALS TOTAL DOSE TALLY REPORT ***\n\n");
  printf("              Dose deposited in each material defined in the input file (tallied directly per material, not per voxel):\n");
  printf("    [MAT]  [DOSE, eV/g/hist]  [2*std_dev]  [Rel_error 2*std_dev, %%]  [E_dep [eV/hist]  [MASS_TOTAL, g]\n");
  printf("   ====================================================================================================\n");
  for (i = 0; i < MCGPU_MAX_MATERIALS; i++) {
    double edep, sd, rel, dose;
    if (ctx->tab.density_nominal[i] < 0.0f) break;
    edep = ((double)md[2 * i]) / SCALE_eV * inv_n;
    sd = sqrt((((double)md[2 * i + 1]) * inv_n - edep * edep) * inv_n);
    rel = edep > 0.0 ? sd / edep : 0.0;
    dose = edep / mass[i];
    sd = sd / mass[i];
    printf("\t%d\t%.5lf\t\t%.5lf\t\t%.2lf\t\t%.2lf\t\t%.5lf\n", (i + 1), dose, 2.0 * sd, 2.0 * 100.0 * rel, edep, mass[i]);
  }
  fflush(stdout);
  return MCGPU_OK;
}
