/* Host side of the RANECU stream partition and the launch-size rule.
 *
 * RANECU = two multiplicative LCGs (L'Ecuyer 1988).  The reference jumps ahead with a
 * square-and-multiply over a 32-bit Russian-peasant product (abMODm, docker/mcgpu/
 * MC-GPU_kernel_v1.3.cu:919-950); that product is exactly (a*s) mod m, so a 64-bit
 * multiply-and-reduce gives the same seeds (checked against the literal restatement in
 * oracle/ by tests/test_ranecu.py). */
#include "mcgpu_host.h"

#define M1 2147483563LL
#define M2 2147483399LL
#define A1 40014LL
#define A2 40692LL
#define LEAP_DISTANCE 256

static long long pow_mod(long long a, unsigned long long n, long long m) {
  long long y = 1, z = a % m;
  while (n) {
    if (n & 1ull) y = (y * z) % m;
    z = (z * z) % m;
    n >>= 1;
  }
  return y;
}

/* init_PRNG (K:841-894): both generators start from seed_input advanced by
 * (stream+1)*histories_per_thread*256 draws. */
void mcgpu_ranecu_init_stream(long long stream, int histories_per_thread, int seed_input, int* s1, int* s2) {
  const unsigned long long leap = ((unsigned long long)(stream + 1)) * (unsigned long long)(histories_per_thread * LEAP_DISTANCE);
  *s1 = (int)(((long long)seed_input * pow_mod(A1, leap, M1)) % M1);
  *s2 = (int)(((long long)seed_input * pow_mod(A2, leap, M2)) % M2);
}

/* ranecu (K:965-986) */
float mcgpu_ranecu_next(int* s1, int* s2) {
  int i1 = *s1 / 53668, i2;
  *s1 = 40014 * (*s1 - i1 * 53668) - i1 * 12211;
  i2 = *s2 / 52774;
  *s2 = 40692 * (*s2 - i2 * 52774) - i2 * 3791;
  if (*s1 < 0) *s1 += 2147483563;
  if (*s2 < 0) *s2 += 2147483399;
  i2 = *s1 - *s2;
  if (i2 < 1) i2 += 2147483562;
  return ((float)i2) * 4.65661305739e-10f;
}

/* update_seed_PRNG(1, total_histories, &seed) (H:3456-3485, called at H:869): only the first
 * generator's multiplier is used and the result seeds both generators of the next projection. */
int mcgpu_ranecu_advance_projection_seed(int seed, unsigned long long total_histories) {
  const unsigned long long leap = total_histories * (unsigned long long)(1 * LEAP_DISTANCE);
  return (int)(((long long)seed * pow_mod(A1, leap, M1)) % M1);
}

/* H:823-841.  The doubles and the +0.9990 are the reference's; histories_per_thread sticks. */
void mcgpu_grid_rule(unsigned long long requested, int threads_per_block, int* histories_per_thread, int* num_blocks, unsigned long long* launched) {
  int total_threads = (int)(((double)requested) / ((double)*histories_per_thread) + 0.9990);
  int blocks = (int)(((double)total_threads) / ((double)threads_per_block) + 0.9990);
  if (blocks > 65535) {
    blocks = 65000;
    *histories_per_thread = (int)(((double)requested) / ((double)(blocks * threads_per_block)) + 0.9990);
  } else if (blocks < 1)
    blocks = 1;
  *num_blocks = blocks;
  *launched = ((unsigned long long)(blocks * threads_per_block)) * (unsigned long long)*histories_per_thread;
}
