"""Shared fixtures.  `-m "not gpu"` covers the oracle (pinned to the reference), the host logic
and the C-ABI surface; `-m gpu` holds the parity tests proper, all through the C ABI."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    lib = ROOT / "4d-cbct-mc_b200" / "lib" / "libmcgpu_b200.so"
    if not lib.exists() or not (ROOT / "oracle" / "liboracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT), "lib", "exe"], check=True, capture_output=True)
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "liboracle.so"], check=True, capture_output=True)
    from __graft_entry__ import import_package

    p = import_package()
    p.engine  # noqa: B018
    return p


@pytest.fixture(scope="session")
def oracle_py(pkg):
    import oracle_py as o

    o.lib()
    return o


CASES = {
    # name: (phantom factory name, kwargs), scan kwargs.  History counts stay >= 95 000: below that the
    # reference's GPU build reads the number as SECONDS (MC-GPU_v1.3.cu:654), a mode cbctmc never uses.
    "water_p1": (("water_cylinder", dict(n=50, spacing_mm=10.0)), dict(n_histories=200_000, n_detector_pixels=(66, 28), kvp=90)),
    "thorax_p4": (("thorax", dict(shape=(64, 64, 25), spacing_mm=8.0)),
                  dict(n_histories=100_000, n_detector_pixels=(66, 28), n_projections=4, angle_between_projections=90.0)),
    "catphan_angles": (("catphan604", dict(n=50, spacing_mm=10.0)),
                       dict(n_histories=100_000, n_detector_pixels=(66, 28), n_projections=3, projection_angles=[30.0, 30.0, 200.5])),
    "air": (("air_scan", dict()), dict(n_histories=200_000, n_detector_pixels=(66, 28))),
    # oblique initial beam (atan2 / acos paths of the pose builder), explicit theta aperture
    "thorax_oblique": (("thorax", dict(shape=(32, 32, 12), spacing_mm=16.0)),
                       dict(n_histories=100_000, n_detector_pixels=(66, 28), polar_aperture=(10.0, 5.0), azimuthal_aperture=8.0,
                            n_projections=2, angle_between_projections=45.0, source_direction=(1.0, 1.0, 0.0), sad=300.0)),
    # cbctmc's patient material set: blood (40 Compton shells = MAX_SHELLS), red marrow (36), muscle, liver, stomach, glands, cartilage
    "patient_p2": (("patient", dict(shape=(64, 64, 25), spacing_mm=8.0)),
                   dict(n_histories=120_000, n_detector_pixels=(66, 28), n_projections=2, angle_between_projections=135.0)),
}


def build_case(pkg, name: str, folder: Path, compressed: bool = True):
    (ph_name, ph_kw), scan = CASES[name]
    scan = dict(scan)
    phantom = getattr(pkg.phantoms, ph_name)(**ph_kw)
    folder.mkdir(parents=True, exist_ok=True)
    vox = folder / ("geometry.vox.gz" if compressed else "geometry.vox")
    pkg.mcio.write_vox(vox, phantom.materials, phantom.densities, phantom.spacing_cm)
    kvp = scan.pop("kvp", None)
    if kvp:
        scan["spectrum"] = pkg.mcio.write_truncated_spectrum(folder / f"{kvp}kVp.spc", kvp)
    cfg = pkg.mcio.ScanConfig(source_position=pkg.mcio.default_source_position(phantom.size_mm), **scan)
    inp = pkg.mcio.write_input(cfg, vox, folder, folder / "input.in")
    return inp, cfg, phantom


@pytest.fixture(scope="session")
def case_dir(tmp_path_factory):
    return tmp_path_factory.mktemp("cases")


@pytest.fixture(scope="session")
def cases(pkg, case_dir):
    """name -> (input path, ScanConfig, Phantom); built once per session."""
    return {name: build_case(pkg, name, case_dir / name) for name in CASES}


def gpu_available() -> bool:
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_engine_factory(pkg):
    if not gpu_available():
        pytest.fail("gpu-marked test started without a visible GPU; the engine has no CPU fallback")

    def make(in_path, device_ids=(0,)):
        eng = pkg.engine.Engine(list(device_ids))
        eng.load_input(in_path).load_voxels().load_materials()
        assert eng.info.num_devices >= 1
        return eng

    return make
