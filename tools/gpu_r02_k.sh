#!/bin/bash
# one B200: phase chaining (source -> tracking, tracking -> tally without a queue exchange) -- parity, then A/B against the same kernel without it (lib_x5)
# (lib_x5 = make BUILD=build_x5 LIBDIR=4d-cbct-mc_b200/lib_x5 XFLAGS=-DMCGPU_WF_CHAIN_MIN=33 lib: the shipped kernel with phase chaining switched off)
set -u
O=gpurun_out/r02k
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
for v in product x5; do
  echo "== $v"
  if [ $v = product ]; then unset MCGPU_B200_LIB; else export MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_$v/libmcgpu_b200.so; fi
  timeout 400 python tools/sweep.py catphan thorax patient linepairs --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/ab_$v.txt
  timeout 300 python tools/sweep.py air --hist=5000000000 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/ab_$v.txt
done
unset MCGPU_B200_LIB
echo "== fast"; timeout 300 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=16 --fast=1 2>&1 | grep "hist/s" | tee $O/fast.txt
echo "== bench (sampler check)"; timeout 300 python bench.py --legs none > $O/bench_catphan_nolegs.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench_catphan_nolegs.json')); print('value %.4g e2e %.4g ms/step %.2f kernel %.2f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks']))"
timeout 300 python bench.py --workload thorax --legs none > $O/bench_thorax_nolegs.json 2>> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench_thorax_nolegs.json')); print('value %.4g e2e %.4g ms/step %.2f kernel %.2f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks']))"
