#!/bin/bash
# Instruction-cache probe (GPU box): ICC hit rate / GCC instruction-request load of the transport kernel
# for a list of "generation:w_threshold[:cta]" settings.  Usage: [WL=catphan] [H=200000000] [FAST=0] tools/icc_probe.sh 2:8 3:12:512 ...
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum
for cfg in "$@"; do
  IFS=: read -r gen thr blk <<< "$cfg"
  if [ "$gen" = 2 ]; then A="--kernels=2 --thresholds=$thr"; else A="--kernels=3 --t3=$thr:${blk:-512}"; fi
  echo "=== generation $gen threshold $thr block ${blk:-} fast ${FAST:-0}"
  timeout 300 ncu --metrics $M --clock-control none -k regex:transport_ --launch-skip 1 -c 1 python tools/sweep.py ${WL:-catphan} --hist=${H:-200000000} $A --fast=${FAST:-0} 2>&1 | grep -E "icc|gcc|inst_executed|issue_active|duration|hist/s"
done
