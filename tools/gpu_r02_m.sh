#!/bin/bash
# one B200: stream-initialisation queue (Q_I) -- parity, then speed
set -u
O=gpurun_out/r02m
mkdir -p $O
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -x -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
echo "== sweep"; timeout 200 python tools/sweep.py catphan thorax --hist=595166015 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee $O/sweep.txt
timeout 100 python tools/sweep.py air --hist=5000000000 --kernels=3 --t3=16 2>&1 | grep "hist/s" | tee -a $O/sweep.txt
