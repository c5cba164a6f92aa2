#!/usr/bin/env python3
"""Fixtures written by the REFERENCE'S OWN writers (run in the build container, where /root/reference exists):

  * MCSimulation.create_mcgpu_input  (cbctmc/mc/simulation.py:288-357)  -> tests/golden/refwriter_*.in
    rendering cbctmc/assets/templates/mcgpu_input.jinja2 with the values of cbctmc/defaults.py;
  * MCGeometry.create_mcgpu_geometry (cbctmc/mc/geometry.py:579-623)    -> tests/golden/refwriter_geometry.vox.gz
    rendering mcgpu_geometry.jinja2 around the text produced by the Cython module cbctmc/mc/voxel_data.pyx
    (compiled here from the reference's source file), including the rot90 / spacing swap it applies.

The reference package imports SimpleITK, matplotlib, docker, torch-free helpers ... that this image does not have; none of
them is touched by the two functions, so missing top-level packages are replaced by inert stub modules for the import.
The voxel arrays as handed to create_mcgpu_geometry are stored next to the files (refwriter_expected.npz) so that the
tests can check the parsed volume voxel by voxel.  Absolute paths of this container in the rendered .in files
(/root/reference/cbctmc/assets/...) are kept as the reference wrote them; the tests map them to the staged copies.

Run:  python tests/golden/make_reference_writer_fixtures.py"""
import gzip
import importlib
import importlib.abc
import importlib.machinery
import subprocess
import sys
import tempfile
import types
import warnings
from pathlib import Path
from unittest import mock

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Third-party packages the reference imports but this image lacks become inert stubs (attributes are MagicMocks).
    Only names collected in `missing` are stubbed -- never the standard library's optional modules."""

    missing: set = set()

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] not in self.missing:
            return None
        return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda attr: mock.MagicMock(name=f"{spec.name}.{attr}")
        return m

    def exec_module(self, module):
        pass


def import_with_stubs(name):
    """Import a reference module, stubbing one missing third-party package per attempt until it loads."""
    for _ in range(40):
        try:
            return importlib.import_module(name)
        except ModuleNotFoundError as e:
            top = (e.name or "").split(".")[0]
            if not top or top == "cbctmc" or top in _StubFinder.missing:
                raise
            _StubFinder.missing.add(top)
            for k in [k for k in sys.modules if k.startswith("cbctmc")]:
                if k != "cbctmc.mc.voxel_data":
                    del sys.modules[k]
    raise RuntimeError(f"cannot import {name}")


def import_reference_writers():
    warnings.simplefilter("ignore")
    build = Path(tempfile.mkdtemp(prefix="voxel_data_"))
    (build / "voxel_data.pyx").write_bytes((REF / "cbctmc/mc/voxel_data.pyx").read_bytes())
    subprocess.run([sys.executable, "-m", "cython", "-3", "voxel_data.pyx"], cwd=build, check=True)
    import sysconfig

    inc = sysconfig.get_paths()["include"]
    so = build / ("voxel_data" + sysconfig.get_config_var("EXT_SUFFIX"))
    subprocess.run(["gcc", "-shared", "-fPIC", "-O2", f"-I{inc}", f"-I{np.get_include()}", "voxel_data.c", "-o", str(so)], cwd=build, check=True)
    sys.path.insert(0, str(REF))
    sys.meta_path.append(_StubFinder())
    spec = importlib.util.spec_from_file_location("cbctmc.mc.voxel_data", so)
    voxel_data = importlib.util.module_from_spec(spec)
    sys.modules["cbctmc.mc.voxel_data"] = voxel_data
    spec.loader.exec_module(voxel_data)
    MCDefaults = import_with_stubs("cbctmc.defaults").DefaultMCSimulationParameters
    MCGeometry = import_with_stubs("cbctmc.mc.geometry").MCGeometry
    MATERIALS_125KEV = import_with_stubs("cbctmc.mc.materials").MATERIALS_125KEV
    MCSimulation = import_with_stubs("cbctmc.mc.simulation").MCSimulation
    print("stubbed third-party packages:", sorted(_StubFinder.missing))
    return MCSimulation, MCGeometry, MCDefaults, MATERIALS_125KEV, voxel_data


def small_patient(materials):
    """A 20 x 24 x 12 volume in cbctmc's (pre-rot90) image frame with an anisotropic spacing."""
    num = {name: m.number for name, m in materials.items()}
    rho = {name: m.density for name, m in materials.items()}
    shape = (20, 24, 12)
    mat = np.full(shape, num["air"], dtype=np.uint8)
    den = np.full(shape, rho["air"], dtype=np.float32)
    x, y, z = np.meshgrid(*(np.arange(n) for n in shape), indexing="ij", sparse=True)
    body = ((x - 10) / 8.5) ** 2 + ((y - 12) / 10.5) ** 2 <= 1
    for name, mask in [("adipose", body), ("muscle_tissue", ((x - 10) / 7.5) ** 2 + ((y - 12) / 9.5) ** 2 <= 1),
                       ("soft_tissue", ((x - 10) / 6.5) ** 2 + ((y - 12) / 8.5) ** 2 <= 1),
                       ("lung", (((x - 6) / 2.5) ** 2 + ((y - 12) / 5.0) ** 2 <= 1) & (z >= 5)),
                       ("blood", (x == 6) & (y == 12) & (z >= 5)), ("liver", (((x - 13) / 3.0) ** 2 + ((y - 10) / 4.0) ** 2 <= 1) & (z < 5)),
                       ("bone_100", (x - 10) ** 2 + (y - 18) ** 2 <= 4), ("red_marrow", (x - 10) ** 2 + (y - 18) ** 2 <= 1)]:
        full = np.broadcast_to(mask, shape)
        mat[full] = num[name]
        den[full] = rho[name]
    # a density gradient inside the lung: several distinct "%.6f" strings per material
    lung = mat == num["lung"]
    den[lung] = (0.2 + 0.013 * np.broadcast_to(z, shape)[lung]).astype(np.float32)
    return mat, den, (3.0, 2.0, 5.0)  # mm


def main():
    MCSimulation, MCGeometry, MCDefaults, MATERIALS, voxel_data = import_reference_writers()
    mat, den, spacing = small_patient(MATERIALS)
    geometry_text = MCGeometry.create_mcgpu_geometry(materials=mat, densities=den, image_spacing=spacing)
    with gzip.GzipFile(HERE / "refwriter_geometry.vox.gz", "wb", mtime=0) as f:  # MCGeometry.save_mcgpu_geometry gzips when asked to
        f.write(geometry_text.encode())
    np.savez_compressed(HERE / "refwriter_expected.npz", materials=mat, densities=den, spacing_mm=np.array(spacing))

    common = dict(voxel_geometry_filepath="@GEOMETRY@", material_filepaths=MCDefaults.material_filepaths, xray_spectrum_filepath=MCDefaults.spectrum_filepath,
                  output_folder="@OUTPUT@")
    size_mm = (mat.shape[1] * spacing[1], mat.shape[0] * spacing[0], mat.shape[2] * spacing[2])  # after the rot90 of create_mcgpu_geometry
    source = (size_mm[0] / 2, size_mm[1] / 2 - MCDefaults.source_to_isocenter_distance, size_mm[2] / 2)  # simulation.py:132-136
    # (1) the defaults of a 3D scan, shortened to 3 projections and a history count that finishes in a second
    text = MCSimulation.create_mcgpu_input(source_position=source, n_histories=240_000, n_projections=3, angle_between_projections=MCDefaults.angle_between_projections,
                                           gpu_ids=(0,), **common)
    (HERE / "refwriter_default.in").write_text(text)
    # (2) a 4D phase: explicit projection angles, two GPUs (gpu id -1), another seed
    text = MCSimulation.create_mcgpu_input(source_position=source, n_histories=150_000, projection_angles=[10.5, 131.25, 250.0, 359.9], n_projections=4,
                                           angle_between_projections=0.0, random_seed=4711, gpu_ids=(0, 1), **common)
    (HERE / "refwriter_angles.in").write_text(text)
    print("wrote", sorted(p.name for p in HERE.glob("refwriter_*")))
    print(geometry_text[:600])


if __name__ == "__main__":
    main()
