#!/usr/bin/env python3
"""Golden vectors of the projection post-processing: runs the reference's OWN functions
(cbctmc/mc/projection.py from /root/reference; SimpleITK and ipmi are not installed here and are replaced by
stubs -- the functions exercised are pure NumPy/SciPy) on ASCII projection files written by this repo's
writer from the golden tallies, and stores the results in tests/golden/post_reference.npz.
Run in the build container only (needs /root/reference): python tests/make_golden_post.py"""
import importlib.util
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import import_package  # noqa: E402
from conftest import build_case  # noqa: E402

pkg = import_package()


def load_reference():
    class Img:  # what the two SimpleITK calls of projections_to_itk need
        def __init__(self, a):
            self.arr, self.sp = np.asarray(a), (1, 1, 1)

        def SetSpacing(self, s):
            self.sp = tuple(s)

        def SetOrigin(self, o):
            self.org = tuple(o)

        def GetSize(self):
            return self.arr.shape[::-1]

        def GetSpacing(self):
            return self.sp

    sitk = types.ModuleType("SimpleITK")
    sitk.GetImageFromArray = lambda a: Img(a)
    sitk.GetArrayFromImage = lambda i: i.arr
    ipmi, common, logger = types.ModuleType("ipmi"), types.ModuleType("ipmi.common"), types.ModuleType("ipmi.common.logger")
    logger.init_fancy_logging = lambda *a, **k: None
    sys.modules.update({"SimpleITK": sitk, "ipmi": ipmi, "ipmi.common": common, "ipmi.common.logger": logger})
    sys.path.insert(0, "/root/reference")
    mc = types.ModuleType("cbctmc.mc")  # skip cbctmc/mc/__init__.py (pyximport of the voxel writer)
    mc.__path__ = ["/root/reference/cbctmc/mc"]
    sys.modules["cbctmc.mc"] = mc
    spec = importlib.util.spec_from_file_location("cbctmc.mc.projection", "/root/reference/cbctmc/mc/projection.py")
    m = importlib.util.module_from_spec(spec)
    sys.modules["cbctmc.mc.projection"] = m
    spec.loader.exec_module(m)
    return m


def main():
    ref = load_reference()
    tmp = Path(tempfile.mkdtemp())
    golden = ROOT / "tests" / "golden"
    npix, crop, sigma = (66, 28), (48, 28), (2, 3)
    files = {}
    for case in ("thorax_p4", "air"):
        inp, cfg, _ = build_case(pkg, case, tmp / case)
        g = np.load(golden / f"{case}.npz")
        with pkg.engine.Engine() as eng:
            eng.load_input(inp)
            eng.set_histories(cfg.n_histories)
            names = sorted(k for k in g.files if k.startswith("projection_"))
            for p in range(eng.info.num_projections):
                name = Path(eng.projection_filename(p)).name
                eng.write_projection(p, g[name], 0.0)
            files[case] = [tmp / case / n for n in names]
    projs = [ref.MCProjection.from_file(f, n_detector_pixels=npix, n_detector_pixels_half_fan=crop) for f in files["thorax_p4"]]
    air = ref.MCProjection.from_file(files["air"][0], n_detector_pixels=npix, n_detector_pixels_half_fan=crop)
    out = {"read_raw_0": np.asarray(projs[0]), "crop": np.array(crop), "sigma": np.array(sigma)}
    for mode in ("total", "unscattered", "scattered"):
        out[f"stack_{mode}"] = ref.projections_to_itk(projs, mode=mode).arr
    air_total = ref.projections_to_itk([air], mode="total").arr.squeeze(axis=0)  # what _read_itk returns for projections_total.mha
    out["air_total"] = air_total
    air_obj = ref.MCProjection(air_total, detector_pixel_size=(0.776, 0.776))
    out["stack_total_normalized"] = ref.projections_to_itk(projs, air_projection=air_obj, air_projection_denoise_kernel_size=sigma, mode="total").arr
    rng = np.random.default_rng(5)
    img = (1000.0 * rng.random((96, 128)) + np.linspace(0, 500, 128)[None, :]).astype(np.float32)
    out["gauss_in"] = img
    out["gauss_out_10_10"] = ref.ndi.gaussian_filter(img, sigma=(10, 10))  # the call of normalize_projections (projection.py:110)
    out["gauss_out_0_17"] = ref.ndi.gaussian_filter(img, sigma=(0, 17))
    np.savez_compressed(golden / "post_reference.npz", **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
