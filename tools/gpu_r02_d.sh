#!/bin/bash
# Round-2 GPU session D (one B200): wavefront diagnostics (batch sizes per queue), experiment builds, statistical protocol,
# final ncu captures of the shipped kernel, gpu_check.
# (needs the diagnostics and experiment libraries of that session: `make stats`; lib_x1 = make BUILD=build_x1 LIBDIR=4d-cbct-mc_b200/lib_x1 XFLAGS=-DMCGPU_EXP_UNROLL2 lib, a switch since removed)
set -u
O=gpurun_out/r02d
mkdir -p $O
echo "== wavefront stats"; for t in 12 20; do MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_stats/libmcgpu_b200.so timeout 300 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=$t 2>&1 | grep -E "wf_stats|hist/s" | awk '!seen[$0]++' | tail -16; done | tee $O/wf_stats.txt
echo "== experiment: shell-term loop unrolled by 2"; MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_x1/libmcgpu_b200.so timeout 300 python tools/sweep.py catphan thorax patient --hist=595166015 --kernels=3 --t3=12 2>&1 | grep "hist/s" | tee $O/exp_unroll2.txt
echo "== product for comparison"; timeout 300 python tools/sweep.py catphan thorax patient --hist=595166015 --kernels=3 --t3=12 2>&1 | grep "hist/s" | tee $O/product.txt
echo "== gpu_check"; timeout 600 python tests/gpu_check.py --big > $O/gpu_check.log 2>&1; tail -4 $O/gpu_check.log; cp gpurun_out/gpu_check.json $O/gpu_check.json
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum
for wl in catphan thorax air; do
  echo "== ncu full $wl"
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:transport_ --launch-skip 1 -c 1 -o $O/prof_$wl -f python bench.py --workload $wl --steps 1 --warmup 1 --legs none > $O/ncu_$wl.log 2>&1
  timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 --csv --log-file $O/extra_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --legs none > /dev/null 2>&1
done
echo "== stat protocol"; timeout 900 python tools/stat_protocol.py --out $O/stat_protocol.json > $O/stat.log 2>&1; tail -30 $O/stat.log
ls -la $O
