#!/bin/bash
# Round-2 GPU session E (one B200): parity + speed after the queue protocol change (two atomics per exchange, Woodcock MFP in the context).
# (needs `make stats` for the wf_stats step)
set -u
O=gpurun_out/r02e
mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -8 $O/pytest_gpu.log
echo "== sweep"; timeout 600 python tools/sweep.py catphan thorax patient air --hist=595166015 --kernels=3 --t3=12,16 > $O/sweep.log 2>&1; cat $O/sweep.log
echo "== sweep fast"; timeout 300 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=12 --fast=1 > $O/sweep_fast.log 2>&1; cat $O/sweep_fast.log
echo "== wavefront stats"; MCGPU_B200_LIB=$PWD/4d-cbct-mc_b200/lib_stats/libmcgpu_b200.so timeout 300 python tools/sweep.py catphan --hist=595166015 --kernels=3 --t3=12,16 2>&1 | grep -E "wf_stats|hist/s" | awk '!seen[$0]++' | tee $O/wf_stats.txt
echo "== icc"; M=sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__warps_eligible.avg.per_cycle_active
for wl in catphan thorax; do timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 python tools/sweep.py $wl --hist=595166015 --kernels=3 --t3=12 2>&1 | grep -E "icc|gcc|inst_executed|issue_active|duration|eligible|hist/s" | tee -a $O/icc.txt; done
ls -la $O
