"""Limits of the input formats (MC-GPU_v1.3.h:59-70): maximum sizes are accepted, one more is rejected with
the reference's error class, and the edge values behave like the reference (CPU only, host parsers)."""
import dataclasses

import numpy as np
import pytest


def make(pkg, cases, tmp_path, **changes):
    inp, cfg, ph = cases["water_p1"]
    cfg = dataclasses.replace(cfg, **changes)
    vox = inp.parent / "geometry.vox.gz"
    return pkg.mcio.write_input(cfg, vox, tmp_path, tmp_path / "input.in"), cfg


def test_1024_projections_are_accepted_and_1025_rejected(pkg, cases, tmp_path):
    inp, _ = make(pkg, cases, tmp_path, n_projections=1024, angle_between_projections=360.0 / 1024)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        assert eng.info.num_projections == 1024 and eng.views().shape[0] == 1024
        seeds = [eng.projection_seed(p) for p in (0, 1, 1023)]
        assert len(set(seeds)) == 3 and all(0 < s < 2147483563 for s in seeds)
    inp, _ = make(pkg, cases, tmp_path, n_projections=1025)
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError, match="too large"):
            eng.load_input(inp)


def test_specific_angle_list_limits(pkg, cases, tmp_path):
    angles = [float(i % 360) for i in range(1024)]
    inp, _ = make(pkg, cases, tmp_path, n_projections=1024, projection_angles=angles)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        assert eng.info.num_projections == 1024
        assert eng.projection_filename(1023).endswith("_%010.6fdeg" % angles[1023])
    inp, _ = make(pkg, cases, tmp_path, n_projections=1025, projection_angles=angles + [1.0])
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError):
            eng.load_input(inp)


def write_spectrum(path, n_bins):
    e = np.linspace(20e3, 80e3, n_bins + 1)
    rows = [f"{e[i]:.3f} {1.0 + (i % 7):.3f}" for i in range(n_bins)] + [f"{e[-1]:.3f} -1"]
    path.write_text("\n".join(rows) + "\n")
    return path


def test_spectrum_bin_limit(pkg, cases, tmp_path):
    inp, _ = make(pkg, cases, tmp_path, spectrum=write_spectrum(tmp_path / "max.spc", 255))
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        assert eng.info.num_spectrum_bins == 255
        assert 20.0 < eng.info.mean_energy_spectrum * 1e-3 < 80.0
    inp, _ = make(pkg, cases, tmp_path, spectrum=write_spectrum(tmp_path / "over.spc", 300))
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError, match="too many energy bins"):
            eng.load_input(inp)


def test_history_counts_at_the_seconds_boundary(pkg, cases, tmp_path):
    """< 95 000 means SECONDS in the reference's GPU build (MC-GPU_v1.3.cu:654): refused, never misread as a count"""
    inp, _ = make(pkg, cases, tmp_path, n_histories=94_999)
    with pkg.engine.Engine() as eng:
        with pytest.raises(pkg.engine.McgpuError, match="seconds"):
            eng.load_input(inp)
    inp, _ = make(pkg, cases, tmp_path, n_histories=95_000)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        assert eng.info.launched_histories == pkg.mcio.launched_histories(95_000, 128, 150)[2]


def test_grid_rule_above_65535_blocks_is_sticky(pkg, cases, tmp_path):
    """H:823-841: more than 65535 blocks -> 65000 blocks and a larger histories-per-thread, kept for later projections"""
    n = 11_903_320_312
    inp, _ = make(pkg, cases, tmp_path, n_histories=n, n_projections=3, angle_between_projections=1.0)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        info = eng.info
        blocks, hpt, launched = pkg.mcio.launched_histories(n, 128, 150)
        assert (info.num_blocks, info.histories_per_thread, info.launched_histories) == (blocks, hpt, launched) == (65000, 1431, 11_905_920_000)
        assert len({eng.projection_seed(p) for p in range(3)}) == 3


def test_history_count_is_sticky_too_after_the_65535_block_correction(pkg, oracle_py, cases, tmp_path):
    """H:841 overwrites total_histories with the launched count, and the NEXT projection's grid rule starts from it: at the
    reference count, projection 2 runs 65000 blocks of 1431 histories per thread again (11 905 920 000 histories) -- not the
    64 986 blocks the .in value (11 903 320 312) would give with the sticky 1431.  Visible on the host through the seed
    schedule (each projection advances the seed by the histories launched before it) and through the oracle."""
    n = 11_903_320_312
    inp, _ = make(pkg, cases, tmp_path, n_histories=n, n_projections=3, angle_between_projections=1.0)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        s0 = eng.projection_seed(0)
        s1 = pkg.engine.advance_projection_seed(s0, 11_905_920_000)
        s2 = pkg.engine.advance_projection_seed(s1, 11_905_920_000)
        assert (eng.projection_seed(1), eng.projection_seed(2)) == (s1, s2)
        assert s2 != pkg.engine.advance_projection_seed(s1, 64_986 * 128 * 1431)
    # small case, through the oracle's GPU-rule schedule: 32 threads/block, 1 history/thread, 3e6 histories
    inp, _ = make(pkg, cases, tmp_path, n_histories=3_000_000, n_projections=2, angle_between_projections=90.0, threads_per_block=32, histories_per_thread=1)
    ora = oracle_py.Oracle(inp, cxx_host_math=True)
    _, launched = ora.run_gpu_rule(1, threads=8)
    assert launched == 65_000 * 32 * 2
