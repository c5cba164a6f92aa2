#!/bin/bash
# Round-2 GPU session B (one B200): parity after the kernel changes, tuning sweep, statistical protocol.
set -u
O=gpurun_out/r02b
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -rs --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -22 $O/pytest_gpu.log
echo "== sweep"; timeout 600 python tools/sweep.py catphan thorax patient --hist=595166015 --kernels=3 --t3=12,16,20 > $O/sweep.log 2>&1; cat $O/sweep.log; cp gpurun_out/sweep.json $O/sweep.json
echo "== sweep bits"; timeout 300 python tools/sweep.py thorax linepairs --hist=595166015 --kernels=3 --t3=12 --bits=0,8 > $O/sweep_bits.log 2>&1; cat $O/sweep_bits.log
echo "== sweep fast"; timeout 300 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=12 --fast=1 > $O/sweep_fast.log 2>&1; cat $O/sweep_fast.log
echo "== icc"; M=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum
for wl in catphan thorax; do timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 python tools/sweep.py $wl --hist=595166015 --kernels=3 --t3=12 2>&1 | grep -E "icc|gcc|inst_executed|issue_active|duration|hist/s" | tee -a $O/icc.txt; done
echo "== stat protocol"; timeout 900 python tools/stat_protocol.py --out $O/stat_protocol.json > $O/stat.log 2>&1; tail -25 $O/stat.log
ls -la $O
