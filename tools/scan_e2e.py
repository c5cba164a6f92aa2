#!/usr/bin/env python3
"""End-to-end timing of the drop-in executable on a short scan (GPU box): thorax phantom written as
a real .vox.gz, P projections at the speed-up-20 history count, `MC-GPU_v1.3.x input.in` as cbctmc
would start it.  Reports wall time per projection next to the transport-only time, i.e. what the
parsers, the pipelined ASCII writer and the D2H copies cost.  Usage: python tools/scan_e2e.py [P]"""
import json
import re
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
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    tmp = Path(tempfile.mkdtemp(prefix="mcgpu_e2e_"))
    ph = pkg.phantoms.thorax()
    t0 = time.time()
    vox = pkg.mcio.write_vox(tmp / "geometry.vox.gz", ph.materials, ph.densities, ph.spacing_cm)
    t_vox = time.time() - t0
    cfg = pkg.mcio.ScanConfig(n_histories=11_903_320_312 // 20, n_projections=P, angle_between_projections=360.0 / P,
                              source_position=pkg.mcio.default_source_position(ph.size_mm))
    inp = pkg.mcio.write_input(cfg, vox, tmp, tmp / "input.in")
    exe = ROOT / "4d-cbct-mc_b200" / "bin" / "MC-GPU_v1.3.x"
    t0 = time.time()
    res = subprocess.run([str(exe), str(inp)], capture_output=True, text=True)
    wall = time.time() - t0
    assert res.returncode == 0, res.stdout[-3000:]
    init = float(re.search(r"INITIALIZATION finished: elapsed time = ([0-9.]+)", res.stdout).group(1))
    files = sorted(tmp.glob("projection_*deg"))
    launched = pkg.mcio.launched_histories(cfg.n_histories, 128, 150)[2]
    out = {"projections": P, "files": len(files), "bytes_per_file": files[0].stat().st_size, "wall_s": wall, "init_s": init,
           "loop_s_per_projection": (wall - init) / P, "histories_per_projection": launched,
           "hist_per_s_end_to_end": P * launched / (wall - init), "vox_write_s_python": t_vox,
           "markers": len(re.findall(r"Simulating Projection", res.stdout))}
    print(json.dumps(out, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "scan_e2e.json").write_text(json.dumps(out, indent=1))
    print(res.stdout[-1500:])


if __name__ == "__main__":
    main()
