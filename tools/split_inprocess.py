#!/usr/bin/env python3
"""History split INSIDE one process (the C library's own multi-device path): mcgpu_run_projection on a context with N
devices cuts the reference grid into N block ranges and sums the u64 tallies on device 0 -- with ncclReduce (libnccl opened
at run time) or with one kernel that reads every peer's image over NVLink (MCGPU_REDUCE=peer).  Reports kernel and reduce
times per variant and checks the image against one device, bit for bit.
Usage: python tools/split_inprocess.py [--gpus 2] [--workload air|catphan|thorax] [--histories N]"""
import argparse
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--workload", default="air")
    ap.add_argument("--histories", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    factories = {"air": (pkg.phantoms.air_scan, 50_000_000_000), "catphan": (pkg.phantoms.catphan604, 11_903_320_312), "thorax": (pkg.phantoms.thorax, 11_903_320_312)}
    factory, n_hist = factories[args.workload]
    n_hist = args.histories or n_hist
    ph = factory()
    tmp = Path(tempfile.mkdtemp(prefix="mcgpu_split_"))
    cfg = pkg.mcio.ScanConfig(n_histories=n_hist, n_projections=1 if ph.shape == (1, 1, 1) else 894, source_position=pkg.mcio.default_source_position(ph.size_mm))
    inp = pkg.mcio.write_input(cfg, tmp / "x.vox", tmp, tmp / "input.in")
    p = 0
    report = {"workload": args.workload, "requested_histories": n_hist, "devices": args.gpus, "variants": {}}

    def run(devices, reduce_env):
        if reduce_env:
            os.environ["MCGPU_REDUCE"] = reduce_env
        else:
            os.environ.pop("MCGPU_REDUCE", None)
        with pkg.engine.Engine(devices) as eng:
            eng.load_input(inp).set_voxels(ph.materials, ph.densities, ph.spacing_cm).load_materials()
            assert eng.info.num_devices == len(devices), (eng.info.num_devices, devices)
            img = eng.new_image()
            t0 = time.perf_counter()
            eng.run_projection(p, out=img)  # first call: creates the reducer (NCCL communicators / peer mappings), kernel attributes
            first = time.perf_counter() - t0
            t0 = time.perf_counter()
            eng.run_projection(p, out=img)
            wall = time.perf_counter() - t0
            return img.copy(), {"kernel_ms_max": eng.last_kernel_ms, "reduce_ms": eng.last_reduce_ms, "reduce": eng.reduce_kind, "wall_ms_with_d2h": 1e3 * wall, "first_call_wall_ms": 1e3 * first,
                                "launched": int(eng.info.launched_histories)}

    one, report["one_device"] = run([0], None)
    for name, env in (("nccl", "nccl"), ("peer_kernel", "peer")):
        img, r = run(list(range(args.gpus)), env)
        r["bit_identical_to_one_device"] = bool(np.array_equal(img, one))
        r["speedup_kernel_plus_reduce"] = report["one_device"]["kernel_ms_max"] / (r["kernel_ms_max"] + r["reduce_ms"])
        r["reduce_GB_per_s_into_device0"] = (args.gpus - 1) * one.nbytes / (r["reduce_ms"] * 1e-3) / 1e9 if r["reduce_ms"] > 0 else None
        report["variants"][name] = r
    text = json.dumps(report, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text)
    ok = all(v["bit_identical_to_one_device"] for v in report["variants"].values())
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
