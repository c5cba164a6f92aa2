"""The C-ABI library loads, exports every symbol include/mcgpu_b200.h declares, and fails loudly
(no CPU fallback) when asked to simulate without a GPU."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, gpu_available


def declared_symbols():
    text = (ROOT / "include" / "mcgpu_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcgpu_[a-z0-9_]+)\s*\(", text)) - {"mcgpu_progress_cb"})


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(str(ROOT / "4d-cbct-mc_b200" / "lib" / "libmcgpu_b200.so"))
    names = declared_symbols()
    assert len(names) >= 24
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcgpu_b200.h but not exported"
    assert set(pkg.engine.EXPORTED_SYMBOLS) == set(names)


def test_no_torch_or_cxx_types_in_signatures():
    text = re.sub(r"/\*.*?\*/", "", (ROOT / "include" / "mcgpu_b200.h").read_text(), flags=re.S)  # comments stripped
    assert "torch" not in text and "std::" not in text and "at::" not in text


def test_loaded_library_is_the_in_tree_one(pkg):
    maps = Path("/proc/self/maps").read_text()
    assert str(ROOT / "4d-cbct-mc_b200" / "lib" / "libmcgpu_b200.so") in maps


def test_library_contains_sm100a_code_only():
    so = ROOT / "4d-cbct-mc_b200" / "lib" / "libmcgpu_b200.so"
    out = subprocess.run(["cuobjdump", "--list-elf", str(so)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_error_codes_mirror_reference_exit_codes(pkg, tmp_path):
    E = pkg.engine
    with E.Engine() as eng:
        with pytest.raises(E.McgpuError) as e:
            eng.load_input(tmp_path / "missing.in")
        assert e.value.code == -1  # reference exit(-1): input file not found (H:1255)
        bad = tmp_path / "bad.in"
        bad.write_text("# nothing useful\n")
        with pytest.raises(E.McgpuError) as e:
            eng.load_input(bad)
        assert e.value.code == -2  # reference exit(-2): section not found (H:1287)
        with pytest.raises(E.McgpuError) as e:
            eng.load_materials()
        assert e.value.code == -6  # stage order


def test_threads_per_block_must_be_multiple_of_32(pkg, cases, tmp_path):
    text = Path(cases["water_p1"][0]).read_text().replace("128  # GPU THREADS", "100  # GPU THREADS")
    f = tmp_path / "tpb.in"
    f.write_text(text)
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_input(f)
        assert e.value.code == -2 and "multiple of 32" in str(e.value)


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_run_without_gpu_fails_loudly(pkg, cases):
    inp, _, _ = cases["water_p1"]
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels().load_materials()
        assert eng.info.num_devices == 0
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.run_projection(0)
        assert e.value.code == -5 and "no CPU fallback" in str(e.value)
        with pytest.raises(pkg.engine.McgpuError):
            eng.run_all()


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure mode")
def test_executable_fails_without_gpu_and_with_bad_args(pkg, cases):
    exe = ROOT / "4d-cbct-mc_b200" / "bin" / "MC-GPU_v1.3.x"
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode != 0 and "not given" in res.stdout
    res = subprocess.run([str(exe), str(cases["water_p1"][0])], capture_output=True, text=True)
    assert res.returncode != 0


def test_executable_tolerates_mpirun_style_launch(pkg, cases):
    """`mpirun -n N MC-GPU_v1.3.x input.in` starts N copies; ranks > 0 must exit 0 with a one-line notice (nothing cbctmc's
    stdout scraping reacts to: no "error", no "Simulating Projection").  SLURM_PROCID alone is NOT an MPI launch
    (`srun -n4 MC-GPU_v1.3.x phase_$SLURM_PROCID.in` = independent runs), so it must not silence a process."""
    import os

    exe = ROOT / "4d-cbct-mc_b200" / "bin" / "MC-GPU_v1.3.x"
    for var in ("OMPI_COMM_WORLD_RANK", "PMI_RANK", "PMIX_RANK", "MV2_COMM_WORLD_RANK"):
        env = dict(os.environ, **{var: "1"})
        res = subprocess.run([str(exe), str(cases["water_p1"][0])], capture_output=True, text=True, env=env)
        assert res.returncode == 0 and len(res.stdout.strip().splitlines()) == 1 and "rank 1 has nothing to do" in res.stdout
        assert "error" not in res.stdout.lower() and "Simulating Projection" not in res.stdout
    env = dict(os.environ, SLURM_PROCID="3")
    res = subprocess.run([str(exe), str(cases["water_p1"][0])], capture_output=True, text=True, env=env)
    assert "Reading the input file" in res.stdout  # it runs (and, without a GPU, fails loudly later)


def test_time_limited_mode_is_rejected(pkg, cases, tmp_path):
    """MC-GPU reads a history count below 95 000 as seconds (H:654); cbctmc never uses that mode."""
    text = Path(cases["water_p1"][0]).read_text().replace("200000  # TOTAL NUMBER", "600  # TOTAL NUMBER")
    f = tmp_path / "time.in"
    f.write_text(text)
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_input(f)
        assert e.value.code == -2 and "seconds" in str(e.value)


def test_a_new_input_or_a_failed_voxel_load_invalidates_what_the_devices_hold(pkg, cases, tmp_path):
    """The devices keep the spectrum and an image sized for the detector of the input they were uploaded with: after
    mcgpu_load_input the context must ask for load_materials again instead of running with stale device state, and a
    failed load_voxels must not leave a context that still looks runnable."""
    with pkg.engine.Engine() as eng:
        eng.load_input(cases["water_p1"][0]).load_voxels().load_materials()
        eng.load_input(cases["thorax_p4"][0])
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.run_projection(0)
        assert e.value.code == -6  # MCGPU_E_STATE, before the "no device" check
        eng.load_voxels().load_materials()
        bad = tmp_path / "truncated.vox"  # valid header, then the data ends: fails AFTER the old volume was released
        bad.write_text("[SECTION VOXELS HEADER v.2008-04-13]\n4 4 4\n1.0 1.0 1.0\n1\n2\n1\n[END OF VXH SECTION]\n1 1.0\n1 1.0\n")
        with pytest.raises(pkg.engine.McgpuError):
            eng.load_voxels(bad)
        assert eng.info.num_voxels_x == 0
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.run_projection(0)
        assert e.value.code == -6
        with pytest.raises(pkg.engine.McgpuError):  # an error, not a NULL dereference, in the dose path either
            eng.dose("voxels")
