// Test harness around the UNMODIFIED reference host code (compiled from where it lies; nothing of
// it is copied here): includes docker/mcgpu/MC-GPU_v1.3.cu with its main() renamed, calls the
// reference's own read_input / init_energy_spectrum / set_CT_trajectory / load_voxels /
// load_material exactly as its main does (MC-GPU_v1.3.cu:490-562) and dumps the resulting structs
// and tables as raw bytes.  Compiled by nvcc (C++ host semantics, the production flavour); no GPU
// is needed because no CUDA call is made.  This pins the host side of oracle/ (cxx flavour) and of
// the product (csrc/host) against the reference's real host arithmetic.
//   usage: ref_host_dump <input.in> <out.bin>
#define main mcgpu_reference_main
#include "MC-GPU_v1.3.cu"
#undef main

static void put(FILE* f, const char* name, const void* p, size_t n) {
  unsigned long long len = n;
  char tag[32] = {0};
  strncpy(tag, name, 31);
  fwrite(tag, 1, 32, f);
  fwrite(&len, 8, 1, f);
  fwrite(p, 1, n, f);
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  static struct detector_struct detector_data[MAX_NUM_PROJECTIONS];
  static struct source_struct source_data[MAX_NUM_PROJECTIONS];
  static struct source_energy_struct source_energy_data;
  static struct rayleigh_struct rayleigh_table;
  static struct compton_struct compton_table;
  struct voxel_struct voxel_data;
  struct linear_interp mfp_table_data;
  float2* voxel_mat_dens = NULL;
  unsigned int voxel_mat_dens_bytes = 0;
  float density_max[MAX_MATERIALS], density_nominal[MAX_MATERIALS];
  unsigned long long int* image = NULL;
  int image_bytes = -1, mfp_table_bytes = -1, mfp_Woodcock_table_bytes = -1;
  float2* mfp_Woodcock_table = NULL;
  float3 *mfp_table_a = NULL, *mfp_table_b = NULL;
  short int rx0, rx1, ry0, ry1, rz0, rz1;
  ulonglong2* voxels_Edep = NULL;
  int voxels_Edep_bytes = 0;
  unsigned long long int total_histories;
  int histories_per_thread, seed_input, num_threads_per_block, gpu_id, num_projections;
  int flag_material_dose = -2, enable_specific_angles = -2;
  double D_angle = -1.0, angularROI_0 = 0.0, angularROI_1 = 360.0, initial_angle = 0.0, SRotAxisD = -1.0, vtrans = 0.0;
  static char file_name_voxels[250], file_name_materials[MAX_MATERIALS][250], file_name_output[250], file_dose_output[250], file_name_espc[250];
  static float specific_angles[MAX_NUM_ANGLES];
  memset(&rayleigh_table, 0, sizeof rayleigh_table);
  for (int k = 0; k < MAX_MATERIALS; k++) density_nominal[k] = -1.0f;
  char* fake_argv[2] = {argv[0], argv[1]};
  FILE* devnull = freopen("/dev/null", "w", stdout);
  (void)devnull;
  read_input(2, fake_argv, 0, &total_histories, &seed_input, &gpu_id, &num_threads_per_block, &histories_per_thread, detector_data, &image, &image_bytes, source_data,
             &source_energy_data, file_name_voxels, file_name_materials, file_name_output, file_name_espc, &num_projections, &D_angle, &angularROI_0, &angularROI_1,
             &initial_angle, &voxels_Edep, &voxels_Edep_bytes, file_dose_output, &rx0, &rx1, &ry0, &ry1, &rz0, &rz1, &SRotAxisD, &vtrans, &flag_material_dose,
             &enable_specific_angles, specific_angles);
  float mean_energy_spectrum = 0.0f;
  init_energy_spectrum(file_name_espc, &source_energy_data, &mean_energy_spectrum);
  if (num_projections != 1)
    set_CT_trajectory(0, num_projections, D_angle, angularROI_0, angularROI_1, SRotAxisD, source_data, detector_data, vtrans, &enable_specific_angles, specific_angles);
  load_voxels(0, file_name_voxels, density_max, &voxel_data, &voxel_mat_dens, &voxel_mat_dens_bytes, &rx1, &ry1, &rz1);
  load_material(0, file_name_materials, density_max, density_nominal, &mfp_table_data, &mfp_Woodcock_table, &mfp_Woodcock_table_bytes, &mfp_table_a, &mfp_table_b,
                &mfp_table_bytes, &rayleigh_table, &compton_table);
  FILE* f = fopen(argv[2], "wb");
  if (!f) return 3;
  double angles[6] = {D_angle, angularROI_0, angularROI_1, initial_angle, SRotAxisD, vtrans};
  long long ints[8] = {num_projections, (long long)total_histories, seed_input, num_threads_per_block, histories_per_thread, enable_specific_angles, mfp_table_data.num_values, 0};
  float scal[3] = {mfp_table_data.e0, mfp_table_data.ide, mean_energy_spectrum};
  put(f, "ints", ints, sizeof ints);
  put(f, "angles", angles, sizeof angles);
  put(f, "scalars", scal, sizeof scal);
  put(f, "source", source_data, sizeof(struct source_struct) * num_projections);
  put(f, "detector", detector_data, sizeof(struct detector_struct) * num_projections);
  put(f, "spectrum", &source_energy_data, sizeof source_energy_data);
  put(f, "voxel_struct", &voxel_data, sizeof voxel_data);
  put(f, "density_max", density_max, sizeof density_max);
  put(f, "density_nominal", density_nominal, sizeof density_nominal);
  put(f, "woodcock", mfp_Woodcock_table, mfp_Woodcock_table_bytes);
  put(f, "mfp_a", mfp_table_a, mfp_table_bytes);
  put(f, "mfp_b", mfp_table_b, mfp_table_bytes);
  put(f, "rayleigh", &rayleigh_table, sizeof rayleigh_table);
  put(f, "compton", &compton_table, sizeof compton_table);
  fclose(f);
  return 0;
}
