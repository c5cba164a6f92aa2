#!/bin/bash
# Round-2 final multi-GPU session (shipped kernel): torchrun bench with its legs, the whole 894-projection scan through the executable,
# the 2-GPU parity test with both reducers, the air scan through the executable (in-process history split).
# Usage: N=8 bash tools/gpu_r02_h.sh   (under gpurun --gpus N)
set -u
N=${N:-8}
O=gpurun_out/r02h_n$N
mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc > $O/nproc.txt
echo "== bench under torchrun, N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"; cut -c1-600 $O/bench_n$N.json; tail -3 $O/bench_n$N.err
echo "== full 894-projection scan, $N GPUs"; timeout 600 python tools/scan_e2e.py --gpus $N --projections 894 --out $O/scan_894_n$N.json 2>&1 | tail -28
echo "== 2-GPU parity test (both reducers)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "history_split" -rs 2>&1 | tail -4 | tee $O/pytest_split.log
echo "== in-process split: air"; timeout 300 python tools/split_inprocess.py --gpus $N --workload air --out $O/split_inprocess_air.json 2>&1 | grep -E "kernel_ms_max|reduce_ms|\"reduce\"|first_call|bit_identical|speedup"
echo "== air scan through the executable"
python - <<PY 2>&1 | tail -14 | tee $O/exe_air.log
import subprocess, sys, tempfile, time
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
print("\n".join(l for l in res.stdout.splitlines() if "driver initialised" in l or "input + spectrum" in l or ">>>" in l))
PY
ls -la $O
