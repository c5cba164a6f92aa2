#!/bin/bash
# Round-2 final single-GPU session for the shipped kernel (with phase chaining): bench lines, launch list, ncu captures for the committed constants.
set -u
O=gpurun_out/r02l
mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
echo "== bench default"; timeout 600 python bench.py > $O/bench_catphan.json 2> $O/bench_catphan.err; echo "rc=$?"; cut -c1-300 $O/bench_catphan.json; tail -3 $O/bench_catphan.err
echo "== bench thorax"; timeout 400 python bench.py --workload thorax --legs refcuda,cpu,split > $O/bench_thorax.json 2> $O/bench_thorax.err; echo "rc=$?"; cut -c1-200 $O/bench_thorax.json
for wl in air patient linepairs; do echo "== bench $wl"; timeout 300 python bench.py --workload $wl --legs cpu > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "rc=$?"; cut -c1-200 $O/bench_$wl.json; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --legs none > $O/bench_under_ncu.log 2>&1
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum
for wl in catphan thorax air; do
  echo "== ncu full $wl"
  timeout 500 ncu --set full --metrics $M --import-source on --clock-control none -k regex:transport_ --launch-skip 1 -c 1 -o $O/prof_$wl -f python bench.py --workload $wl --steps 1 --warmup 1 --legs none > $O/ncu_$wl.log 2>&1
done
ls -la $O
