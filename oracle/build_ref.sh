#!/bin/sh
# Compile the UNMODIFIED reference sources, where they lie, into oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the snapshot).  No reference source is copied into this repo.
#   MC-GPU_v1.3_CPU.x          gcc, plain C, the documented CPU build (MC-GPU_v1.3.cu:287) -> pins oracle/
#   MC-GPU_v1.3_sm100_exact.x  nvcc sm_100, -fmad=false, no fast-math -> the bit-exactness oracle for the CUDA kernel
#   MC-GPU_v1.3_sm100_fast.x   nvcc sm_100 with the shipped flags (docker/compile.sh:36, -use_fast_math) -> the kernel to beat
set -e
REF=${1:-/root/reference}
SRC=$REF/docker/mcgpu
OUT=$(dirname "$0")/_ref
mkdir -p "$OUT"
up_to_date() { [ -x "$1" ] && [ "$1" -nt "$SRC/MC-GPU_v1.3.cu" ] && [ "$1" -nt "$SRC/MC-GPU_kernel_v1.3.cu" ]; }
up_to_date "$OUT/MC-GPU_v1.3_CPU.x" || gcc -x c -O3 -fgnu89-inline "$SRC/MC-GPU_v1.3.cu" -o "$OUT/MC-GPU_v1.3_CPU.x" -I"$SRC" -lm -lz 2>/dev/null
if command -v nvcc >/dev/null 2>&1; then
  COMMON="-m64 -O3 -DUSING_CUDA -I$SRC -I$REF/docker/cuda-samples/Common -lz -gencode=arch=compute_100,code=sm_100 -Wno-deprecated-gpu-targets -w"
  up_to_date "$OUT/MC-GPU_v1.3_sm100_exact.x" || nvcc "$SRC/MC-GPU_v1.3.cu" -o "$OUT/MC-GPU_v1.3_sm100_exact.x" $COMMON -fmad=false
  # the reference's own host stages (read_input ... load_material), nvcc/C++ host semantics, dumped as raw bytes
  HERE=$(cd "$(dirname "$0")" && pwd)
  up_to_date "$OUT/ref_host_dump.x" || nvcc "$HERE/ref_host_dump.cu" -o "$OUT/ref_host_dump.x" $COMMON
  up_to_date "$OUT/MC-GPU_v1.3_sm100_fast.x" || nvcc "$SRC/MC-GPU_v1.3.cu" -o "$OUT/MC-GPU_v1.3_sm100_fast.x" $COMMON -use_fast_math
fi
ls -la "$OUT"
