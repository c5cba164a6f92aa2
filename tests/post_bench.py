#!/usr/bin/env python3
"""Post-processing at the reference size (GPU box): 1848x768 tallies of a short thorax scan -> cropped float32
stacks + air normalisation on the device (mcgpu_post_*), next to the reference's route (ASCII file ->
np.loadtxt -> NumPy/SciPy, restated in oracle/post_oracle.py) on the same data.  Usage: python tests/post_bench.py [P]"""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
import post_oracle  # noqa: E402  (checker / CPU baseline only)


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    tmp = Path(tempfile.mkdtemp(prefix="mcgpu_post_"))
    ph = pkg.phantoms.thorax()
    cfg = pkg.mcio.ScanConfig(n_histories=20_000_000, n_projections=P, angle_between_projections=360.0 / P, source_position=pkg.mcio.default_source_position(ph.size_mm))
    inp = pkg.mcio.write_input(cfg, tmp / "x.vox", tmp, tmp / "input.in")
    air_ph = pkg.phantoms.air_scan()
    (tmp / "air").mkdir()
    air_cfg = pkg.mcio.ScanConfig(n_histories=200_000_000, n_projections=1, source_position=pkg.mcio.default_source_position(air_ph.size_mm))
    air_inp = pkg.mcio.write_input(air_cfg, tmp / "air" / "x.vox", tmp / "air", tmp / "air" / "input.in")

    with pkg.engine.Engine([0]) as eng:
        eng.load_input(air_inp).set_voxels(air_ph.materials, air_ph.densities, air_ph.spacing_cm).load_materials()
        air_tally = eng.run_projection(0)
        air_total = eng.post_intensity(air_tally, None, 1024)[0]
        eng.write_projection(0, air_tally, 0.0)
        air_file, air_npix = eng.projection_filename(0), air_cfg.n_detector_pixels
    with pkg.engine.Engine([0]) as eng:
        eng.load_input(inp).set_voxels(ph.materials, ph.densities, ph.spacing_cm).load_materials()
        tallies = [eng.run_projection(p) for p in range(P)]
        eng.post_intensity(tallies[0], None, 1024)  # warm-up
        t0 = time.time()
        ours = pkg.postprocess.postprocess_scan(eng, tallies, P, air_total, crop_x=1024, sigma=(10, 10))
        t_gpu = time.time() - t0
        t0 = time.time()
        for t in tallies:
            eng.post_intensity(t, None, 1024)
        t_int_host = (time.time() - t0) / P
        n_streams = eng.info.num_blocks * eng.info.threads_per_block
        eng.run_streams(0, 0, n_streams, fetch=False)
        t0 = time.time()
        for _ in range(P):
            eng.post_intensity(None, None, 1024)  # tally still on the device: no 45 MB round trip
        t_int_dev = (time.time() - t0) / P
        t0 = time.time()
        for p in range(P):
            eng.write_projection(p, tallies[p], 0.0)
        t_write = time.time() - t0
        files = [eng.projection_filename(p) for p in range(P)]
    t0 = time.time()
    stack4 = np.stack([post_oracle.read_raw(f, cfg.n_detector_pixels, (1024, 768)) for f in files])
    t_load = time.time() - t0
    t0 = time.time()
    air4 = post_oracle.read_raw(air_file, air_npix, (1024, 768))[None]
    ref_air = post_oracle.projections_stack(air4, "total")[0]
    ref = {m: post_oracle.projections_stack(stack4, m) for m in ("total", "unscattered", "scattered")}
    ref["total_normalized"] = post_oracle.projections_stack(stack4, "total", air=ref_air, sigma=(10, 10))
    t_numpy = time.time() - t0
    same = {m: bool(np.array_equal(ours[m], ref[m])) for m in ("total", "unscattered", "scattered")}
    d = np.abs(ours["total_normalized"] - ref["total_normalized"])
    ulp = np.spacing(np.abs(ref["total_normalized"]).astype(np.float32))
    out = {"projections": P, "detector": "1848x768 -> 1024x768", "device_s_per_projection": t_gpu / P, "post_intensity_s": {"tally_on_host": t_int_host, "tally_on_device": t_int_dev},
           "reference_route_s_per_projection": {"ascii_write": t_write / P, "np.loadtxt": t_load / P, "numpy_scipy": t_numpy / P},
           "intensity_stacks_bit_equal": same, "air_image_bit_equal": bool(np.array_equal(air_total, ref_air)),
           "normalized_max_abs_diff": float(d.max()), "normalized_max_diff_in_ulp": float((d / ulp).max())}
    print(json.dumps(out, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "post_bench.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
