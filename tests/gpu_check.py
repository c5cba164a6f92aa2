#!/usr/bin/env python3
"""GPU-box sanity script: bit-exact parity of the CUDA engine against the reference's own CUDA
source (oracle/_ref/MC-GPU_v1.3_sm100_exact.x) and a first timing next to the reference's
shipped-flags build.  Writes a JSON summary to gpurun_out/gpu_check.json.

Usage: python tests/gpu_check.py [--big]"""
import json
import re
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
import oracle_py  # noqa: E402

OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
WORK = Path("/tmp/mcgpu_check")


def make_case(name, phantom, **cfg_kw):
    d = WORK / name
    d.mkdir(parents=True, exist_ok=True)
    vox = d / "geometry.vox.gz"
    pkg.mcio.write_vox(vox, phantom.materials, phantom.densities, phantom.spacing_cm)
    cfg = pkg.mcio.ScanConfig(source_position=pkg.mcio.default_source_position(phantom.size_mm), **cfg_kw)
    for sub in ("ref", "ours"):
        (d / sub).mkdir(exist_ok=True)
        pkg.mcio.write_input(cfg, vox, d / sub, d / f"input_{sub}.in")
    return d, cfg


def run_ref(binary, in_path):
    t = time.time()
    res = subprocess.run([str(binary), str(in_path)], capture_output=True, text=True)
    dt = time.time() - t
    speeds = [float(x) for x in re.findall(r"Speed \[x-rays/s\]:\s+([0-9.eE+]+)", res.stdout)]
    if res.returncode != 0:
        print(res.stdout[-3000:], res.stderr[-2000:])
        raise SystemExit(f"{binary} failed with {res.returncode}")
    return dt, speeds, res.stdout


def compare_case(name, phantom, report, **cfg_kw):
    d, cfg = make_case(name, phantom, **cfg_kw)
    dt_ref, speeds, log = run_ref(oracle_py.REF_CUDA_EXACT, d / "input_ref.in")
    (OUT / f"{name}_ref_exact.log").write_text(log)
    eng = pkg.engine.Engine([0])
    eng.load_input(d / "input_ours.in").load_voxels().load_materials()
    info = eng.info
    npx = (info.num_pixels_x, info.num_pixels_z)
    det_cm = (cfg.detector_size[0] / 10, cfg.detector_size[1] / 10)
    ok_all = True
    per_proj = []
    last_writer = {}
    for p in range(info.num_projections):  # with specific angles later projections overwrite earlier files of the same name (Q7)
        last_writer[Path(eng.projection_filename(p)).name] = p
    for p in sorted(last_writer.values()):
        img = eng.run_projection(p)
        ms = eng.last_kernel_ms
        fname = Path(eng.projection_filename(p)).name
        ref_vals = pkg.mcio.read_projection(d / "ref" / fname, npx)
        ref_cnt = pkg.mcio.projection_counts(ref_vals, npx, det_cm, info.launched_histories)
        ndiff = int((ref_cnt != img).sum())
        per_proj.append({"p": p, "file": fname, "ndiff": ndiff, "sum_ref": int(ref_cnt.sum()), "sum_ours": int(img.sum()), "kernel_ms": ms})
        ok_all &= ndiff == 0
        eng.write_projection(p, img, ms / 1e3)
        # our ASCII file must parse to the same numbers as the reference's
        ours_vals = pkg.mcio.read_projection(d / "ours" / fname, npx)
        per_proj[-1]["ascii_max_abs_diff"] = float(np.abs(ours_vals - ref_vals).max())
    report[name] = {"bit_exact": bool(ok_all), "launched": int(info.launched_histories), "voxel_bits": info.voxel_bits,
                    "palette": info.palette_size, "ref_exact_speed": speeds, "projections": per_proj}
    print(name, "bit_exact =", ok_all, per_proj)
    eng.close()
    return ok_all


def main():
    big = "--big" in sys.argv
    report = {}
    ph = pkg.phantoms
    ok = True
    # 1: water cylinder, single projection (rotation_flag = 0 fast path), 90 kVp
    WORK.mkdir(parents=True, exist_ok=True)
    spc = pkg.mcio.write_truncated_spectrum(WORK / "90kVp.spc", 90)
    ok &= compare_case("water_p1", ph.water_cylinder(n=125, spacing_mm=4.0), report, n_histories=2_000_000, spectrum=spc, n_detector_pixels=(462, 192))
    # 2: thorax (6 materials, density gradient), 5 projections with rotation
    ok &= compare_case("thorax_p5", ph.thorax(shape=(128, 128, 50), spacing_mm=4.0), report, n_histories=1_000_000, n_detector_pixels=(462, 192),
                       n_projections=5, angle_between_projections=72.0)
    # 3: catphan-like (10 materials), specific angles incl. the duplicated first one (Q7)
    ok &= compare_case("catphan_angles", ph.catphan604(n=125, spacing_mm=4.0), report, n_histories=1_000_000, n_detector_pixels=(462, 192),
                       projection_angles=[30.0, 30.0, 200.5], n_projections=3)
    # 4: air scan (one 200 cm voxel), source inside the box
    air = ph.air_scan()
    # cbctmc puts the source at (size/2, size/2 - SAD, size/2) = (100, 0, 100) cm
    ok &= compare_case("air", air, report, n_histories=3_000_000, n_detector_pixels=(462, 192))
    report["all_bit_exact"] = bool(ok)

    # timing: thorax 256x256x100 @ 2 mm, full detector, 1 projection of 5 -> rotation path
    if big:
        phantom = ph.thorax()
        d, cfg = make_case("thorax_big", phantom, n_histories=200_000_000, n_projections=2, angle_between_projections=90.0)
        dt, speeds, log = run_ref(oracle_py.REF_CUDA_FAST, d / "input_ref.in")
        (OUT / "thorax_big_ref_fast.log").write_text(log)
        eng = pkg.engine.Engine([0])
        eng.load_input(d / "input_ours.in").load_voxels().load_materials()
        ours = []
        for p in range(2):
            eng.run_streams(p, 0, eng.info.num_blocks * 128, fetch=False)
            ours.append(eng.info.launched_histories / (eng.last_kernel_ms / 1e3))
        report["thorax_big"] = {"ref_fast_hist_per_s": speeds, "ours_hist_per_s": ours, "launched": int(eng.info.launched_histories)}
        print("thorax_big", report["thorax_big"])
    (OUT / "gpu_check.json").write_text(json.dumps(report, indent=1))
    print("ALL BIT EXACT" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
