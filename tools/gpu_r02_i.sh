#!/bin/bash
# one B200: queue-protocol A/B (product: head CAS + tail add | x3: published counter, 4 atomics | x4: product + sleep in the entry wait),
# the bench's clock sampler fix, ncu of the line-pair workload
# (lib_x3 / lib_x4 were built from git fa86281..1c0ec66 with -DMCGPU_WF_PUBLISHED / -DMCGPU_WF_SPIN_SLEEP=40; the shipped source is the x3 protocol)
set -u
O=gpurun_out/r02i
mkdir -p $O
for v in product x3 x4; do
  echo "== $v"
  if [ $v = product ]; then unset MCGPU_B200_LIB; else export MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_$v/libmcgpu_b200.so; fi
  timeout 300 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/ab_$v.txt
  timeout 300 python tools/sweep.py air --hist=5000000000 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/ab_$v.txt
done
unset MCGPU_B200_LIB
echo "== bench (sampler)"; timeout 300 python bench.py --legs none > $O/bench_catphan_nolegs.json 2> $O/bench.err; cut -c1-330 $O/bench_catphan_nolegs.json; python -c "
import json; d=json.load(open('$O/bench_catphan_nolegs.json')); print('value %.4g e2e %.4g ms/step %.2f kernel %.2f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks']))"
timeout 300 python bench.py --workload thorax --legs none > $O/bench_thorax_nolegs.json 2>> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench_thorax_nolegs.json')); print('value %.4g e2e %.4g ms/step %.2f kernel %.2f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['clocks']))"
echo "== ncu linepairs"
M=lts__t_sectors.sum,lts__t_sectors_op_atom.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum
timeout 500 ncu --set full --import-source on --clock-control none -k regex:transport_ --launch-skip 1 -c 1 -o $O/prof_linepairs -f python bench.py --workload linepairs --steps 1 --warmup 1 --legs none > $O/ncu_linepairs.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 --csv --log-file $O/extra_linepairs.csv python bench.py --workload linepairs --steps 1 --warmup 1 --legs none > /dev/null 2>&1
ls -la $O
