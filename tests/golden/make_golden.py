#!/usr/bin/env python3
"""Generate the golden vectors from the REFERENCE ITSELF (its unmodified CPU build,
oracle/_ref/MC-GPU_v1.3_CPU.x, compiled from /root/reference by oracle/build_ref.sh).

The reference ships no golden vectors for this path, so these are what pins the oracle
on boxes where /root/reference -- and therefore oracle/_ref -- may be absent.  For every case of
tests/conftest.py:CASES the reference binary is run on the generated inputs and the integer
tallies recovered from its ASCII projections are stored per output file name
(tests/golden/<case>.npz); one raw ASCII projection is kept for the writer test.

Run in the build container:  python tests/golden/make_golden.py"""
import gzip
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import import_package  # noqa: E402

pkg = import_package()
import oracle_py  # noqa: E402
from conftest import CASES, build_case  # noqa: E402


def main():
    assert oracle_py.REF_CPU.exists(), "run `make -C oracle ref` first"
    only = sys.argv[1:]  # optional: regenerate just these cases
    for name in CASES:
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            inp, cfg, _ = build_case(pkg, name, tmp)
            oracle_py.run_reference_binary(oracle_py.REF_CPU, inp, cwd=tmp)
            # histories the CPU build simulates: ceil(N/hpt)*hpt (MC-GPU_v1.3.cu:920-922)
            launched = int(float(cfg.n_histories) / float(cfg.histories_per_thread) + 0.9990) * cfg.histories_per_thread
            det_cm = (round(cfg.detector_size[0] / 10, 6), round(cfg.detector_size[1] / 10, 6))
            out = {"launched": np.array(launched, dtype=np.uint64)}
            for f in sorted(tmp.glob("projection_*deg")):
                vals = pkg.mcio.read_projection(f, cfg.n_detector_pixels)
                out[f.name] = pkg.mcio.projection_counts(vals, cfg.n_detector_pixels, det_cm, launched)
                if name == "water_p1":
                    with open(f, "rb") as src, gzip.GzipFile(HERE / "water_p1_projection_ascii.gz", "wb", mtime=0) as dst:
                        shutil.copyfileobj(src, dst)
            np.savez_compressed(HERE / f"{name}.npz", **out)
            print(name, launched, [k for k in out if k != "launched"], {k: int(v.sum()) for k, v in out.items() if k != "launched"})


if __name__ == "__main__":
    main()
