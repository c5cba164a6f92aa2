import sys, subprocess
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import import_package
pkg = import_package()
import oracle_py
import conftest

def run(name, scan_over, ph_kw=dict(shape=(32, 32, 12), spacing_mm=16.0)):
    conftest.CASES["diag"] = (("thorax", ph_kw), scan_over)
    tmp = Path("/tmp/diag") / name
    inp, cfg, _ = conftest.build_case(pkg, "diag", tmp)
    oracle_py.run_reference_binary(oracle_py.REF_CUDA_EXACT, inp, cwd=tmp)
    eng = pkg.engine.Engine([0]); eng.load_input(inp).load_voxels().load_materials()
    info = eng.info
    det = (round(cfg.detector_size[0] / 10, 6), round(cfg.detector_size[1] / 10, 6))
    for p in range(info.num_projections):
        f = Path(eng.projection_filename(p))
        ours = eng.run_projection(p)
        ref = pkg.mcio.projection_counts(pkg.mcio.read_projection(f, cfg.n_detector_pixels), cfg.n_detector_pixels, det, info.launched_histories)
        d = ours.astype(np.int64) - ref.astype(np.int64)
        print(name, "p", p, f.name, "ndiff", (d != 0).sum(), "per plane", [(d[k] != 0).sum() for k in range(4)], "sums", [int(ours[k].sum()) for k in range(4)], [int(ref[k].sum()) for k in range(4)])
    eng.close()

base = dict(n_histories=50_000, n_detector_pixels=(66, 28), n_projections=2, angle_between_projections=45.0)
run("oblique_full", dict(base, polar_aperture=(10.0, 5.0), azimuthal_aperture=8.0, source_direction=(1.0, 1.0, 0.0), sad=300.0))
run("oblique_sad1000", dict(base, polar_aperture=(10.0, 5.0), azimuthal_aperture=8.0, source_direction=(1.0, 1.0, 0.0)))
run("axis_apertures", dict(base, polar_aperture=(10.0, 5.0), azimuthal_aperture=8.0))
run("oblique_default_ap", dict(base, source_direction=(1.0, 1.0, 0.0)))
run("oblique_x", dict(base, source_direction=(1.0, 0.0, 0.0)))
