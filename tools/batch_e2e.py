#!/usr/bin/env python3
"""Fixed cost per invocation (GPU box): K respiratory-phase inputs of the thorax phantom (2 projections each
at a small history count) run as K separate `MC-GPU_v1.3.x` processes and as one `MC-GPU_v1.3_batch.x`
process.  Usage: python tools/batch_e2e.py [K]"""
import json
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    tmp = Path(tempfile.mkdtemp(prefix="mcgpu_batch_"))
    ph = pkg.phantoms.thorax()
    inputs = []
    for k in range(K):
        d = tmp / f"phase_{k:02d}"
        d.mkdir()
        vox = pkg.mcio.write_vox(d / "geometry.vox.gz", ph.materials, ph.densities, ph.spacing_cm)
        cfg = pkg.mcio.ScanConfig(n_histories=2_000_000, n_projections=2, angle_between_projections=7.0 + k, random_seed=42 + k,
                                  source_position=pkg.mcio.default_source_position(ph.size_mm))
        inputs.append(str(pkg.mcio.write_input(cfg, vox, d, d / "input.in")))
    bindir = ROOT / "4d-cbct-mc_b200" / "bin"
    subprocess.run([str(bindir / "MC-GPU_v1.3.x"), inputs[0]], capture_output=True)  # page the binaries and libraries in
    t0 = time.time()
    for i in inputs:
        assert subprocess.run([str(bindir / "MC-GPU_v1.3.x"), i], capture_output=True).returncode == 0
    t_sep = time.time() - t0
    t0 = time.time()
    assert subprocess.run([str(bindir / "MC-GPU_v1.3_batch.x")] + inputs, capture_output=True).returncode == 0
    t_batch = time.time() - t0
    out = {"inputs": K, "separate_processes_s": t_sep, "one_process_s": t_batch, "saved_per_input_s": (t_sep - t_batch) / K}
    print(json.dumps(out, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "batch_e2e.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
