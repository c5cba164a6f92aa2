#!/usr/bin/env python3
"""BASELINE config 4 measured end to end on the GPU box: a 4D CBCT = one MC-GPU input per respiratory phase, each with its own
geometry and its own list of projection angles (cbctmc/mc/simulation.py:596-692: the angles of a 894-projection scan are dealt
to the phases by the respiratory signal; every phase repeats its first angle once, the reference's work-around for projection
0), all phases run by ONE `MC-GPU_v1.3_batch.x` process on N GPUs (next input parsed and uploaded while the current one is
simulated).  Reports wall time, seconds per projection and the fixed cost per phase, next to the same inputs run as separate
`MC-GPU_v1.3.x` processes when --separate is given.
Usage: python tools/scan4d_e2e.py [--gpus 8] [--phases 10] [--projections 894] [--separate]"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

pkg = bench.pkg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--phases", type=int, default=10)
    ap.add_argument("--projections", type=int, default=894)
    ap.add_argument("--separate", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    work = bench.scratch_dir("mcgpu_4d_")
    n_hist = 11_903_320_312 // 20
    step = 360.0 / 894
    inputs, total = [], 0
    t0 = time.perf_counter()
    for k in range(args.phases):
        ph = pkg.phantoms.thorax(diaphragm_shift_mm=4.0 * k)  # a different geometry per phase
        d = work / f"phase_{k:02d}"
        d.mkdir(parents=True)
        vox = pkg.mcio.write_voxb(d / "geometry.voxb", ph.materials, ph.densities, ph.spacing_cm)
        angles = [270.0 + i * step for i in range(args.projections) if i % args.phases == k]
        angles = angles[0:1] + angles  # simulation.py:655-657
        total += len(angles)
        cfg = pkg.mcio.ScanConfig(n_histories=n_hist, n_projections=len(angles), projection_angles=angles, angle_between_projections=step,
                                  source_position=pkg.mcio.default_source_position(ph.size_mm))
        inputs.append(str(pkg.mcio.write_input(cfg, vox, d, d / "input.in")))
    t_inputs = time.perf_counter() - t0
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=bench.visible_gpus(args.gpus))
    stop = threading.Event()
    written = {"files": 0}

    def reaper():
        while True:
            done = stop.is_set()
            files = sorted(work.glob("phase_*/projection_*deg"), key=lambda f: f.stat().st_mtime)
            for f in (files if done else files[: max(0, len(files) - 3 * args.gpus)]):
                try:
                    f.unlink()
                    written["files"] += 1
                except OSError:
                    pass
            if done:
                return
            stop.wait(0.25)

    bindir = ROOT / "4d-cbct-mc_b200" / "bin"
    out = {"phases": args.phases, "projections_total": total, "devices": args.gpus, "histories_per_projection": pkg.mcio.launched_histories(n_hist, 128, 150)[2],
           "inputs_write_s_python": t_inputs, "workload": "thorax 256x256x100 @2mm, one geometry per phase (diaphragm shifted), specific angle lists"}
    th = threading.Thread(target=reaper, daemon=True)
    th.start()
    t0 = time.perf_counter()
    res = subprocess.run([str(bindir / "MC-GPU_v1.3_batch.x")] + inputs, capture_output=True, text=True, env=env)
    out["one_process_wall_s"] = time.perf_counter() - t0
    assert res.returncode == 0, res.stdout[-2000:]
    loops = [float(x) for x in re.findall(r"Projection loop: ([0-9.]+) s wall", res.stdout)]
    out["projection_loops_s"] = loops
    out["s_per_projection_in_loops"] = sum(loops) / total if loops else None
    out["fixed_cost_per_phase_s"] = (out["one_process_wall_s"] - sum(loops)) / args.phases if loops else None
    out["markers"] = len(re.findall(r"Simulating Projection", res.stdout))
    if args.separate:
        t0 = time.perf_counter()
        for i in inputs:
            assert subprocess.run([str(bindir / "MC-GPU_v1.3.x"), i], capture_output=True, env=env).returncode == 0
        out["separate_processes_wall_s"] = time.perf_counter() - t0
    stop.set()
    th.join()
    out["report_files_written"] = written["files"]
    shutil.rmtree(work, ignore_errors=True)
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text)


if __name__ == "__main__":
    main()
