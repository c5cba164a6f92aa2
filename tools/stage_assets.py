#!/usr/bin/env python3
"""Stage the reference's MC-GPU input DATA (not code) under assets/.

The engine is a drop-in for MC-GPU_v1.3.x, so it has to be exercised with the
very material / spectrum files cbctmc hands to MC-GPU
(reference: cbctmc/assets/material_files/*.mcgpu, cbctmc/assets/spectra/*.spc;
material order = density-sorted list, cbctmc/mc/materials.py:112-119).

/root/reference does not exist on the GPU box, so the files the synthetic
workloads need are gzip'ed (byte-exact payload, mtime=0 so the output is
reproducible) into assets/materials/.  MC-GPU reads .gz material files through
zlib (MC-GPU_v1.3.cu:2199), and it only reads the *header* of materials that
do not appear in the voxel file (MC-GPU_v1.3.cu:2220-2233).  ALL 22 files are staged in full:
cbctmc's patient geometries use blood (40 shells = MAX_SHELLS), red_marrow (36), muscle, liver,
stomach/intestines, glands and cartilage (cbctmc/mc/geometry.py:161, 214-229).

Run once in the build container:  python tools/stage_assets.py
"""
import gzip
import os
import re
import shutil
import sys
from pathlib import Path

REF = Path(os.environ.get("MCGPU_REFERENCE", "/root/reference")) / "cbctmc" / "assets"
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "assets"

SPECTRA = ["125kVp_0.89mmTi_varian_norm.spc"]


def nominal_density(path: Path) -> float:
    with open(path, "rt") as f:
        for line in f:
            if "NOMINAL DENSITY" in line:
                return float(next(f).strip("# \n"))
    raise RuntimeError(f"no nominal density in {path}")


def main() -> int:
    if not REF.is_dir():
        print(f"reference assets not found at {REF}; nothing staged", file=sys.stderr)
        return 0
    (OUT / "materials").mkdir(parents=True, exist_ok=True)
    (OUT / "spectra").mkdir(parents=True, exist_ok=True)
    files = sorted((REF / "material_files").glob("*__5_125kev.mcgpu"))
    order = sorted(files, key=nominal_density)  # stable: ties keep filename order
    with open(OUT / "materials" / "ORDER.txt", "wt") as order_f:
        for number, path in enumerate(order, start=1):
            ident = path.name.split("__")[0]
            rho = nominal_density(path)
            full = True
            order_f.write(f"{number} {ident} {rho:g} {'full' if full else 'stub'}\n")
            dst = OUT / "materials" / (path.name + ".gz")
            with open(path, "rb") as src:
                payload = src.read()
            if not full:
                # keep everything up to (and including) the nominal-density value line
                m = re.search(rb"\[NOMINAL DENSITY[^\n]*\n[^\n]*\n", payload)
                payload = payload[: m.end()] + b"#[STUB: rows omitted; material unused by the synthetic workloads]\n"
            with open(dst, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", compresslevel=9, mtime=0, filename="") as gz:
                gz.write(payload)
    for name in SPECTRA:
        shutil.copyfile(REF / "spectra" / name, OUT / "spectra" / name)
    print("staged", len(order), "materials ->", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
