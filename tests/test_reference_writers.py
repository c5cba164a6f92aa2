"""The boundary against the REFERENCE'S OWN writers: input files rendered by cbctmc's MCSimulation.create_mcgpu_input
(mcgpu_input.jinja2 + cbctmc/defaults.py) and a geometry written by MCGeometry.create_mcgpu_geometry
(mcgpu_geometry.jinja2 + the Cython voxel_data.pyx), committed as fixtures by tests/golden/make_reference_writer_fixtures.py
(which imports and calls those functions of /root/reference).  CPU tests: the engine parses them to the values the reference
put in; GPU tests: `MC-GPU_v1.3.x` runs them as cbctmc would (simulation.py:187-226: one copy per GPU under mpirun,
stdout scraped for progress and for the word "error", files named projection_<angle>deg), bit-exact against the reference
CUDA build on the same files."""
import os
import re
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, gpu_available

GOLDEN = ROOT / "tests" / "golden"
EXE = ROOT / "4d-cbct-mc_b200" / "bin" / "MC-GPU_v1.3.x"
FILE_PATTERN = re.compile(r"^projection_\d{3}\.\d{6}deg$")  # simulation.py:283 (_clean_simulation_folder)


def materialise(name: str, folder: Path) -> Path:
    """Give the fixture the paths of this machine: the reference wrote /root/reference/cbctmc/assets/... (plain .mcgpu
    files); the staged copies are assets/materials/*.mcgpu.gz and assets/spectra/*."""
    folder.mkdir(parents=True, exist_ok=True)
    shutil.copyfile(GOLDEN / "refwriter_geometry.vox.gz", folder / "geometry.vox.gz")
    text = (GOLDEN / name).read_text()
    text = re.sub(r"/root/reference/cbctmc/assets/material_files/(\S+?)\.mcgpu\b", lambda m: str(ROOT / "assets" / "materials" / (m.group(1) + ".mcgpu.gz")), text)
    text = text.replace("/root/reference/cbctmc/assets/spectra/", str(ROOT / "assets" / "spectra") + "/")
    text = text.replace("@GEOMETRY@", str(folder / "geometry.vox.gz")).replace("@OUTPUT@", str(folder))
    assert "/root/reference" not in text
    out = folder / "input.in"
    out.write_text(text)
    return out


def expected_volume():
    z = np.load(GOLDEN / "refwriter_expected.npz")
    mat = np.rot90(z["materials"], k=3, axes=(0, 1))  # geometry.py:589-590
    den = np.rot90(z["densities"], k=3, axes=(0, 1))
    sp = z["spacing_mm"]
    spacing_cm = (sp[1] / 10.0, sp[0] / 10.0, sp[2] / 10.0)  # geometry.py:596-598: "switch 0, 1 due to rot"
    den_text = np.array([np.float32(f"{v:.6f}") for v in den.reshape(-1)], dtype=np.float32).reshape(den.shape)  # voxel_data.pyx:27
    return mat, den_text, spacing_cm


def test_reference_rendered_geometry_parses_to_the_voxels_cbctmc_wrote(pkg, tmp_path):
    inp = materialise("refwriter_default.in", tmp_path)
    mat, den, spacing_cm = expected_volume()
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels().load_materials()
        info = eng.info
        assert (info.num_voxels_x, info.num_voxels_y, info.num_voxels_z) == mat.shape == (24, 20, 12)
        got_mat = eng.table("voxel_material").reshape(mat.shape[::-1]).transpose(2, 1, 0)  # x fastest in the file
        got_den = eng.table("voxel_density").reshape(mat.shape[::-1]).transpose(2, 1, 0)
        assert np.array_equal(got_mat, mat)
        assert np.array_equal(got_den.view(np.uint32), den.view(np.uint32))
        assert info.num_materials_used == len(np.unique(mat)) == 9
        nosc = eng.table("compton_noscco")
        assert nosc.max() == 40  # blood is in this volume


@pytest.mark.parametrize("name,expect", [
    ("refwriter_default.in", dict(P=3, hist=240_000, seed=42, angles=False)),
    ("refwriter_angles.in", dict(P=4, hist=150_000, seed=4711, angles=True)),
])
def test_reference_rendered_input_parses_to_the_values_of_defaults_py(pkg, tmp_path, name, expect):
    inp = materialise(name, tmp_path)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        info = eng.info
        assert (info.num_projections, info.requested_histories, info.seed_input) == (expect["P"], expect["hist"], expect["seed"])
        assert (info.num_pixels_x, info.num_pixels_z, info.threads_per_block, info.histories_per_thread) == (1848, 768, 128, 150)
        assert bool(info.enable_specific_angles) == expect["angles"]
        names = [Path(eng.projection_filename(p)).name for p in range(info.num_projections)]
        assert all(FILE_PATTERN.match(n) for n in names), names
        # the same scan written by this repo's own writer gives the very same poses
        sp_cm = expected_volume()[2]
        size_mm = tuple(10.0 * n * s for n, s in zip((24, 20, 12), sp_cm))
        kw = dict(projection_angles=[10.5, 131.25, 250.0, 359.9], angle_between_projections=0.0) if expect["angles"] else {}
        cfg = pkg.mcio.ScanConfig(n_histories=expect["hist"], n_projections=expect["P"], random_seed=expect["seed"],
                                  source_position=pkg.mcio.default_source_position(size_mm), **kw)
        mine = pkg.mcio.write_input(cfg, tmp_path / "geometry.vox.gz", tmp_path / "mine", tmp_path / "mine.in")
        views_ref = eng.views().copy()
    with pkg.engine.Engine() as eng:
        eng.load_input(mine)
        assert np.array_equal(eng.views().view(np.uint32), views_ref.view(np.uint32))


# ---------------------------------------------------------------------------------------------------- GPU
def read_counts(pkg, path, launched):
    n_pix, det_cm = (1848, 768), (71.7024, 29.7984)
    return pkg.mcio.projection_counts(pkg.mcio.read_projection(path, n_pix), n_pix, det_cm, launched)


@pytest.mark.gpu
def test_executable_runs_the_reference_rendered_scan_like_cbctmc_starts_it(pkg, oracle_py, tmp_path):
    assert gpu_available()
    inp = materialise("refwriter_default.in", tmp_path / "ours")
    res = subprocess.run([str(EXE), str(inp)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
    marks = re.findall(r"Simulating Projection (?P<i>\d{1,4}) of (?P<n>\d{1,4})", res.stdout)  # simulation.py:213-216
    assert marks == [("1", "3"), ("2", "3"), ("3", "3")]
    assert not re.search(r"(?i)error", res.stdout)  # simulation.py:217
    files = sorted(f for f in (tmp_path / "ours").iterdir() if FILE_PATTERN.match(f.name))
    assert len(files) == 3
    # bit-exact against the reference's CUDA source on the very same reference-written files
    assert oracle_py.REF_CUDA_EXACT.exists(), "oracle/_ref/MC-GPU_v1.3_sm100_exact.x is missing"
    ref_inp = materialise("refwriter_default.in", tmp_path / "ref")
    oracle_py.run_reference_binary(oracle_py.REF_CUDA_EXACT, ref_inp, cwd=tmp_path / "ref")
    launched = pkg.mcio.launched_histories(240_000, 128, 150)[2]
    for f in files:
        ours, ref = read_counts(pkg, f, launched), read_counts(pkg, tmp_path / "ref" / f.name, launched)
        assert ours.sum() > 0 and np.array_equal(ours, ref), f.name


@pytest.mark.gpu
def test_four_mpirun_style_copies_leave_one_clean_file_set(pkg, tmp_path):
    """cbctmc starts `mpirun -n <gpus> MC-GPU_v1.3.x input.in` (simulation.py:187-198): N concurrent copies of the same
    command line, told apart only by the launcher's rank variable.  Rank 0 simulates on every GPU, the others exit."""
    inp = materialise("refwriter_angles.in", tmp_path / "mpi")
    procs = [subprocess.Popen([str(EXE), str(inp)], stdout=subprocess.PIPE, text=True, env=dict(os.environ, OMPI_COMM_WORLD_RANK=str(r), OMPI_COMM_WORLD_SIZE="4"))
             for r in range(4)]
    outs = [p.communicate()[0] for p in procs]
    assert [p.returncode for p in procs] == [0, 0, 0, 0]
    assert len(re.findall(r"Simulating Projection", outs[0])) == 4
    for o in outs[1:]:
        assert "Simulating Projection" not in o and not re.search(r"(?i)error", o) and len(o.strip().splitlines()) == 1
    files = sorted(f.name for f in (tmp_path / "mpi").iterdir() if FILE_PATTERN.match(f.name))
    assert files == ["projection_010.500000deg", "projection_131.250000deg", "projection_250.000000deg", "projection_359.899994deg"]
    # the same scan run once, alone: identical numbers
    alone = materialise("refwriter_angles.in", tmp_path / "alone")
    assert subprocess.run([str(EXE), str(alone)], capture_output=True, text=True).returncode == 0
    launched = pkg.mcio.launched_histories(150_000, 128, 150)[2]
    for name in files:
        a, b = read_counts(pkg, tmp_path / "mpi" / name, launched), read_counts(pkg, tmp_path / "alone" / name, launched)
        assert a.sum() > 0 and np.array_equal(a, b), name
