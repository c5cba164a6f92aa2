#!/bin/bash
# Round-2 multi-GPU session (N GPUs of one box): the torchrun bench with its history_split and scan_e2e legs, with FULL=1 the real
# 894-projection scan and the 4D scan through the drop-in executables, then the in-process history split (ncclReduce / peer kernel).
# Usage: N=2 bash tools/gpu_r02_c.sh      (run under gpurun --gpus N)
set -u
N=${N:-2}
O=gpurun_out/r02c_n$N
mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nvidia-smi topo -m > $O/topo.txt 2>&1; nproc > $O/nproc.txt
echo "== bench under torchrun, N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"; cut -c1-3500 $O/bench_n$N.json; tail -5 $O/bench_n$N.err
if [ "${FULL:-0}" = "1" ]; then
  echo "== full 894-projection scan, $N GPUs"; timeout 900 python tools/scan_e2e.py --gpus $N --projections 894 --out $O/scan_894_n$N.json 2>&1 | tail -30
  echo "== 4D scan (config 4): 10 phases, 894+10 projections, $N GPUs"; timeout 900 python tools/scan4d_e2e.py --gpus $N --phases 10 --projections 894 --out $O/scan4d_n$N.json 2>&1 | tail -30
fi
echo "== 2-GPU parity test"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "history_split" -rs 2>&1 | tail -5 | tee $O/pytest_split.log
echo "== in-process split: air"; timeout 300 python tools/split_inprocess.py --gpus $N --workload air --out $O/split_inprocess_air.json 2>&1 | tail -45
echo "== in-process split: thorax"; timeout 300 python tools/split_inprocess.py --gpus $N --workload thorax --out $O/split_inprocess_thorax.json 2>&1 | tail -45
echo "== air scan through the executable (history split inside MC-GPU_v1.3.x)"
python - <<PY 2>&1 | tail -12 | tee $O/exe_air.log
import subprocess, sys, tempfile, time, re
from pathlib import Path
sys.path.insert(0, ".")
from __graft_entry__ import import_package
pkg = import_package()
ph = pkg.phantoms.air_scan()
tmp = Path(tempfile.mkdtemp())
vox = pkg.mcio.write_vox(tmp / "geometry.vox", ph.materials, ph.densities, ph.spacing_cm)
cfg = pkg.mcio.ScanConfig(n_histories=50_000_000_000, n_projections=1, source_position=pkg.mcio.default_source_position(ph.size_mm))
inp = pkg.mcio.write_input(cfg, vox, tmp, tmp / "input.in")
t0 = time.time()
res = subprocess.run(["4d-cbct-mc_b200/bin/MC-GPU_v1.3.x", str(inp)], capture_output=True, text=True)
print("rc", res.returncode, "wall %.2f s" % (time.time() - t0), [f.name for f in tmp.glob("projection_*")])
print(res.stdout[-900:])
PY
ls -la $O
