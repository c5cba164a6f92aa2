"""Bit-exact parity at BASELINE.json's scale (needs a B200): the default (wavefront) kernel, through the C ABI,
against the reference's own CUDA source compiled for sm_100 without fast-math
(oracle/_ref/MC-GPU_v1.3_sm100_exact.x, built by oracle/build_ref.sh) on the same input files.

tests/test_gpu_parity.py keeps its cases small so that they finish in seconds; there a projection is < 1000 RANECU
streams, i.e. ONE CTA of the product kernel.  The cases here cover what that leaves out:

  (a) BASELINE config 1 as stated: water cylinder, 1 projection, 1e7 histories, 90 kVp, rotation_flag 0, 1848x768
      detector: 66 800 streams over ~66 CTAs, the full pixel index range;
  (b) patient-like thorax 256x256x100 @ 2 mm, 1848x768, 3 projections at 5e6 histories (rotated poses);
  (c) cbctmc's PATIENT material set at full size (blood: 40 Compton shells = MAX_SHELLS, red marrow 36, muscle,
      liver, stomach, glands, cartilage): the 41-float scratch stride and the 16-row scratch decision of launch.cu;
  (d) the > 65 535-block rule (H:823-841) on the device: 32 threads/block and 1 history/thread make the grid
      65 000 x 32 with 2 histories per thread, sticky for the second projection;
  (e) the one MC-GPU input the reference repository commits: scripts/run_35000000000_run_00/air/geometry.vox.gz
      (copied byte for byte to tests/golden/ref_air_geometry.vox.gz).

A missing oracle/_ref is a FAILURE here, not a skip: these tests are the parity claim.  Every reference run is < 60 s."""
import shutil
import time
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def det_cm(cfg):
    return (round(cfg.detector_size[0] / 10, 6), round(cfg.detector_size[1] / 10, 6))


def compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, folder: Path, phantom, vox_path=None, **scan):
    """Write the inputs once, run the reference binary and the engine on the SAME files, compare every projection
    file's u64 tallies.  Returns (info of the engine, per-projection kernel ms, seconds the reference took)."""
    assert oracle_py.REF_CUDA_EXACT.exists(), "oracle/_ref/MC-GPU_v1.3_sm100_exact.x is missing: run oracle/build_ref.sh where /root/reference exists"
    folder.mkdir(parents=True, exist_ok=True)
    if vox_path is None:
        vox_path = pkg.mcio.write_vox(folder / "geometry.vox.gz", phantom.materials, phantom.densities, phantom.spacing_cm)
    kvp = scan.pop("kvp", None)
    if kvp:
        scan["spectrum"] = pkg.mcio.write_truncated_spectrum(folder / f"{kvp}kVp.spc", kvp)
    size_mm = scan.pop("size_mm", None) or phantom.size_mm
    cfg = pkg.mcio.ScanConfig(source_position=pkg.mcio.default_source_position(size_mm), **scan)
    inp = pkg.mcio.write_input(cfg, vox_path, folder, folder / "input.in")
    t0 = time.time()
    log = oracle_py.run_reference_binary(oracle_py.REF_CUDA_EXACT, inp, cwd=folder, timeout=600)
    t_ref = time.time() - t0
    assert "CUDA SIMULATION IN THE GPU" in log
    ref_dir = folder / "ref_out"
    ref_dir.mkdir()
    for f in folder.glob("projection_*deg"):
        shutil.move(str(f), ref_dir / f.name)
    eng = gpu_engine_factory(inp)
    info = eng.info
    last_writer = {}
    for p in range(info.num_projections):
        last_writer[Path(eng.projection_filename(p)).name] = p
    assert {f.name for f in ref_dir.iterdir()} == set(last_writer)
    ms = {}
    for fname, p in sorted(last_writer.items(), key=lambda kv: kv[1]):
        ours = eng.run_projection(p)
        ms[p] = eng.last_kernel_ms
        info = eng.info  # histories_per_thread / launched after the (sticky) grid rule
        ref = pkg.mcio.projection_counts(pkg.mcio.read_projection(ref_dir / fname, cfg.n_detector_pixels), cfg.n_detector_pixels, det_cm(cfg), info.launched_histories)
        assert ours.sum() > 0
        ndiff = int((ours != ref).sum())
        assert ndiff == 0, f"{fname}: {ndiff} of {ours.size} tallies differ (sum ours {int(ours.sum())}, reference {int(ref.sum())})"
    eng.close()
    return info, ms, t_ref, log


def test_config1_water_cylinder_1e7_histories_90kvp(pkg, oracle_py, gpu_engine_factory, tmp_path):
    ph = pkg.phantoms.water_cylinder(n=200, spacing_mm=2.5)
    info, ms, t_ref, _ = compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, tmp_path, ph, n_histories=10_000_000, kvp=90)
    assert (info.num_pixels_x, info.num_pixels_z) == (1848, 768)
    assert info.launched_histories == 10_003_200 and info.num_blocks == 521  # 521 blocks x 128 x 150
    assert t_ref < 60


def test_thorax_full_size_full_detector_three_projections(pkg, oracle_py, gpu_engine_factory, tmp_path):
    ph = pkg.phantoms.thorax()
    assert ph.shape == (256, 256, 100)
    info, ms, t_ref, _ = compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, tmp_path, ph, n_histories=5_000_000, n_projections=3,
                                                     angle_between_projections=117.0)
    assert info.num_blocks * info.threads_per_block >= 33_000  # >= 33 photon pools of 1024 contexts: many CTAs share the stream counter
    assert t_ref < 60


def test_patient_material_set_40_and_36_shells(pkg, oracle_py, gpu_engine_factory, tmp_path):
    ph = pkg.phantoms.patient()
    info, ms, t_ref, _ = compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, tmp_path, ph, n_histories=4_000_000, n_projections=2,
                                                     angle_between_projections=75.0)
    assert info.num_materials_used == 14
    with pkg.engine.Engine([0]) as eng:  # host-side view of the same files: the shell counts that size the scratch
        eng.load_input(tmp_path / "input.in").load_voxels().load_materials()
        nosc = eng.table("compton_noscco")
    assert nosc.max() == 40 and 36 in nosc
    assert t_ref < 60


def test_more_than_65535_blocks_rule_is_sticky_on_the_device(pkg, oracle_py, gpu_engine_factory, tmp_path):
    ph = pkg.phantoms.thorax(shape=(64, 64, 25), spacing_mm=8.0)
    info, ms, t_ref, log = compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, tmp_path, ph, n_histories=3_000_000, threads_per_block=32,
                                                       histories_per_thread=1, n_projections=2, angle_between_projections=90.0, n_detector_pixels=(924, 384))
    assert info.num_blocks == 65_000 and info.histories_per_thread == 2 and info.launched_histories == 65_000 * 32 * 2
    assert "65000" in log
    assert t_ref < 60


def test_reference_repository_air_fixture(pkg, oracle_py, gpu_engine_factory, tmp_path):
    """The air-scan geometry file committed in the reference repository, read by both programs as it is."""
    vox = tmp_path / "geometry.vox.gz"
    shutil.copyfile(ROOT / "tests" / "golden" / "ref_air_geometry.vox.gz", vox)
    info, ms, t_ref, _ = compare_with_reference_cuda(pkg, oracle_py, gpu_engine_factory, tmp_path, None, vox_path=vox, size_mm=(2000.0, 2000.0, 2000.0),
                                                     n_histories=20_000_000)
    assert (info.num_voxels_x, info.num_voxels_y, info.num_voxels_z) == (1, 1, 1)
    assert t_ref < 60
