#!/bin/bash
# quick A/B (one B200): MFP record prefetched at batch start (product) vs fetched on demand (lib_x2)
# (lib_x2 was the shipped source with -DMCGPU_NO_REC_PREFETCH while the prefetch was in the source; both the prefetch and the switch have been removed, see wavefront.cuh)
set -u
O=gpurun_out/r02f
mkdir -p $O
echo "== product"; timeout 300 python tools/sweep.py catphan thorax patient --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee $O/product.txt
echo "== no prefetch"; MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_x2/libmcgpu_b200.so timeout 300 python tools/sweep.py catphan thorax patient --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee $O/no_prefetch.txt
echo "== product again"; timeout 300 python tools/sweep.py catphan thorax linepairs --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/product.txt
