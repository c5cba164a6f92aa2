#!/bin/bash
# Round-2 GPU session A (one B200): parity at scale, the new bench legs, microbenchmarks, ncu captures.
# Usage (from the repo root on the GPU box): bash tools/gpu_r02_a.sh
set -u
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -rs --durations=15 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -30 $O/pytest_gpu.log
echo "== ubench"; tools/ubench/bin/atomics > $O/ubench_atomics.txt 2>&1; tools/ubench/bin/l2gather > $O/ubench_l2gather.txt 2>&1; cat $O/ubench_atomics.txt $O/ubench_l2gather.txt
echo "== bench default"; timeout 600 python bench.py > $O/bench_catphan.json 2> $O/bench_catphan.err; echo "rc=$?"; cut -c1-1500 $O/bench_catphan.json; tail -5 $O/bench_catphan.err
echo "== bench thorax"; timeout 400 python bench.py --workload thorax --legs refcuda,cpu,split > $O/bench_thorax.json 2> $O/bench_thorax.err; echo "rc=$?"; cut -c1-600 $O/bench_thorax.json
echo "== bench patient"; timeout 300 python bench.py --workload patient --legs cpu > $O/bench_patient.json 2> $O/bench_patient.err; echo "rc=$?"; cut -c1-400 $O/bench_patient.json
echo "== bench air"; timeout 300 python bench.py --workload air --legs cpu > $O/bench_air.json 2> $O/bench_air.err; echo "rc=$?"; cut -c1-400 $O/bench_air.json
echo "== bench linepairs"; timeout 300 python bench.py --workload linepairs --legs cpu > $O/bench_linepairs.json 2> $O/bench_linepairs.err; echo "rc=$?"; cut -c1-400 $O/bench_linepairs.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --legs none > $O/bench_under_ncu.log 2>&1
M=lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_srcunit_tex.sum,lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_sectors.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed_op_global_red.sum,smsp__inst_executed_op_global_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum
for wl in catphan thorax air; do
  echo "== ncu full $wl"
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:transport_ --launch-skip 1 -c 1 -o $O/prof_$wl -f python bench.py --workload $wl --steps 1 --warmup 1 --legs none > $O/ncu_$wl.log 2>&1
  timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 --csv --log-file $O/extra_$wl.csv python bench.py --workload $wl --steps 1 --warmup 1 --legs none > /dev/null 2>&1
done
ls -la $O
