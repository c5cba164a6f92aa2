# Builds libmcgpu_b200.so (C host + sm_100a CUDA), the MC-GPU_v1.3.x drop-in executable, the
# CPU oracle (test infrastructure) and -- when /root/reference is present -- the reference's own
# binaries under oracle/_ref/ (checker only; see oracle/README.md).
PKG      := 4d-cbct-mc_b200
HOSTDIR  := $(PKG)/csrc/host
CUDADIR  := $(PKG)/csrc/cuda
BUILD    := build
LIBDIR   := $(PKG)/lib
BINDIR   := $(PKG)/bin

CC       ?= gcc
NVCC     ?= nvcc
# -ffp-contract=off: the table builders must round exactly like the reference's host code
CFLAGS   := -std=gnu11 -O2 -g -fPIC -Wall -Wextra -Wno-unused-result -ffp-contract=off -Iinclude -I$(HOSTDIR)
# -fmad=false and no fast-math: bit-exact tallies against the reference CUDA source
NVFLAGS  := -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
            -Xcompiler -fPIC -Iinclude -I$(HOSTDIR)

HOST_SRC := input.c geometry.c voxels.c tables.c ranecu_host.c report.c dose.c api.c post.c
HOST_OBJ := $(HOST_SRC:%.c=$(BUILD)/%.o)
CUDA_OBJ := $(BUILD)/device.o $(BUILD)/postprocess.o $(BUILD)/launch_exact.o $(BUILD)/launch_fast.o
CUDA_HDR := $(CUDADIR)/transport.cuh $(CUDADIR)/regroup.cuh $(CUDADIR)/streams.cuh $(CUDADIR)/wavefront.cuh $(CUDADIR)/scene_dev.h $(CUDADIR)/device_internal.h $(HOSTDIR)/mcgpu_host.h
# AB=1 also compiles the two earlier kernel generations (MCGPU_KERNEL=1|2) for A/B measurements; the shipped library carries only the product kernel
ifeq ($(AB),1)
NVFLAGS += -DMCGPU_AB_KERNELS
endif

all: lib exe oracle ubench

lib: $(LIBDIR)/libmcgpu_b200.so
exe: $(BINDIR)/MC-GPU_v1.3.x $(BINDIR)/MC-GPU_v1.3_batch.x
oracle:
	$(MAKE) -C oracle

$(BUILD)/%.o: $(HOSTDIR)/%.c $(HOSTDIR)/mcgpu_host.h include/mcgpu_b200.h
	@mkdir -p $(BUILD)
	$(CC) $(CFLAGS) -c $< -o $@

$(BUILD)/device.o: $(CUDADIR)/device.cu $(CUDA_HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false $(XFLAGS) -c $< -o $@

$(BUILD)/postprocess.o: $(CUDADIR)/postprocess.cu $(CUDA_HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false -c $< -o $@

# the kernels, twice: bit-exact arithmetic (default path) and the reference's shipped fast-math flags (opt-in)
$(BUILD)/launch_exact.o: $(CUDADIR)/launch.cu $(CUDA_HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -fmad=false $(XFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/ptxas_exact.log || (cat $(BUILD)/ptxas_exact.log; false)
	@grep -E "registers|spill" $(BUILD)/ptxas_exact.log | sort | uniq -c | head -20

$(BUILD)/launch_fast.o: $(CUDADIR)/launch.cu $(CUDA_HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -use_fast_math -DMCGPU_FAST_MATH -DMCGPU_NS=mcgpu_fast $(XFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/ptxas_fast.log || (cat $(BUILD)/ptxas_fast.log; false)
	@grep -E "registers|spill" $(BUILD)/ptxas_fast.log | sort | uniq -c | head -20

$(LIBDIR)/libmcgpu_b200.so: $(HOST_OBJ) $(CUDA_OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared -gencode arch=compute_100a,code=sm_100a -o $@ $^ -lz -lpthread -lm -Xlinker -soname=libmcgpu_b200.so

$(BINDIR)/MC-GPU_v1.3.x: $(HOSTDIR)/main.c $(LIBDIR)/libmcgpu_b200.so
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) -fPIE $< -o $@ -L$(LIBDIR) -lmcgpu_b200 -Wl,-rpath,'$$ORIGIN/../lib'

# several input files (the respiratory phases of a 4D scan) in one process: one CUDA context
$(BINDIR)/MC-GPU_v1.3_batch.x: $(HOSTDIR)/main.c $(LIBDIR)/libmcgpu_b200.so
	@mkdir -p $(BINDIR)
	$(CC) $(CFLAGS) -DMCGPU_BATCH_MAIN -fPIE $< -o $@ -L$(LIBDIR) -lmcgpu_b200 -lpthread -Wl,-rpath,'$$ORIGIN/../lib'

# microbenchmarks that pin the hardware denominators of the rooflines (instruction cache, u64 atomics, L2 random gather)
# diagnostics build of the library: batch sizes per queue, tracking steps per batch, idle polls printed to stderr after every launch
#   make stats  ->  4d-cbct-mc_b200/lib_stats/libmcgpu_b200.so   (use with MCGPU_B200_LIB=...)
stats:
	$(MAKE) BUILD=build_stats LIBDIR=$(PKG)/lib_stats XFLAGS=-DMCGPU_WF_STATS lib

ubench: tools/ubench/bin/icache tools/ubench/bin/atomics tools/ubench/bin/l2gather
tools/ubench/bin/%: tools/ubench/%.cu
	@mkdir -p tools/ubench/bin
	$(NVCC) -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -o $@ $<

clean:
	rm -rf $(BUILD) $(LIBDIR) $(BINDIR) tools/ubench/bin
	$(MAKE) -C oracle clean

.PHONY: all lib exe oracle ubench stats clean
