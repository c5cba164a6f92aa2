#!/usr/bin/env python3
"""Kernel A/B and tuning sweep on the GPU box: histories/s of one projection for each workload and
each (MCGPU_KERNEL, MCGPU_W_THRESHOLD) setting (--kernels=1,2,3 --thresholds=8 --t3=12:512,12:1024 = W-batch threshold : CTA size); every variant is also checked for bit-identical
tallies against the first one.  Usage: python tools/sweep.py [workloads...] [--hist N] [--thresholds a,b,c]"""
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
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    opts = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--") and "=" in a)
    hist = int(opts.get("hist", 100_000_000))
    thresholds = opts.get("thresholds", "8").split(",")  # w_threshold of the regrouping kernel
    kernels = [int(x) for x in opts.get("kernels", "1,2").split(",")]
    workloads = args or ["thorax", "catphan"]
    factories = {"thorax": pkg.phantoms.thorax, "catphan": pkg.phantoms.catphan604, "water": pkg.phantoms.water_cylinder,
                 "air": pkg.phantoms.air_scan, "linepairs": pkg.phantoms.line_pairs, "patient": pkg.phantoms.patient}
    bits_list = opts.get("bits", "0").split(",")  # voxel packing: 0 = the engine's choice, 8 / 16 / 64 force a wider one (MCGPU_VOXEL_BITS)
    out = {}
    for wl in workloads:
        ph = factories[wl]()
        tmp = Path(tempfile.mkdtemp())
        cfg = pkg.mcio.ScanConfig(n_histories=hist, n_projections=1 if wl == "air" else 894, source_position=pkg.mcio.default_source_position(ph.size_mm))
        inp = pkg.mcio.write_input(cfg, tmp / "x.vox", tmp, tmp / "input.in")
        base = None
        t3 = opts.get("t3", "16").split(",")  # W-batch exit threshold of the wavefront kernel
        configs = []
        for k in kernels:
            if k == 1:
                configs.append((k, "0"))
            elif k == 2:
                configs += [(k, t) for t in thresholds]
            else:
                configs += [(k, t) for t in t3]
        configs = [(k, t, b) for (k, t) in configs for b in bits_list]
        for k, t, bits in configs:
            os.environ["MCGPU_KERNEL"] = str(k)
            if bits != "0":
                os.environ["MCGPU_VOXEL_BITS"] = bits
            else:
                os.environ.pop("MCGPU_VOXEL_BITS", None)
            if k != 1:
                os.environ["MCGPU_W_THRESHOLD"] = t.split(":")[0]
                os.environ["MCGPU_WF_BLOCK"] = t.split(":")[1] if ":" in t else "512"
                os.environ["MCGPU_WF_ROWS"] = t.split(":")[2] if t.count(":") > 1 else "0"
            eng = pkg.engine.Engine([0])
            eng.load_input(inp).set_voxels(ph.materials, ph.densities, ph.spacing_cm).load_materials()
            eng.set_fast_math(opts.get("fast", "0") != "0")
            info = eng.info
            n = info.num_blocks * info.threads_per_block
            p = 0 if wl == "air" else 100
            eng.run_streams(p, 0, n, fetch=False)  # warm-up
            ms = []
            for _ in range(2):
                eng.run_streams(p, 0, n, fetch=False)
                ms.append(eng.last_kernel_ms)
            img = eng.run_projection(p)
            same = True if base is None else bool(np.array_equal(img, base))
            if base is None:
                base = img
            rate = info.launched_histories / (min(ms) / 1e3)
            out[f"{wl}/k{k}/t{t}/bits{info.voxel_bits}"] = {"hist_per_s": rate, "ms": min(ms), "identical_to_first": same}
            print(f"{wl:10s} kernel v{k} cfg {t:>14s} voxel bits {info.voxel_bits:2d}: {rate:.4g} hist/s ({min(ms):.1f} ms) identical={same}", flush=True)
            eng.close()
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "sweep.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
