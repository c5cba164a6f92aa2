"""Pin the oracle before trusting it: oracle/mcgpu_oracle.c (plain-C host flavour, CPU-build
stream partition) must reproduce the REFERENCE's own CPU binary count for count -- against the
committed fixtures generated from that binary (tests/golden/make_golden.py), and live against
oracle/_ref/MC-GPU_v1.3_CPU.x when it is present."""
from pathlib import Path

import numpy as np
import pytest

from conftest import CASES, ROOT

GOLDEN = ROOT / "tests" / "golden"


def oracle_files(pkg, ora, cfg):
    """Run every projection the way the reference CPU build does and key the images by the file
    name report_image would give them (later projections overwrite earlier ones, Q7)."""
    out = {}
    launched = None
    for p in range(ora.num_projections):
        img, launched = ora.run_cpu_rule(p, threads=4)
        if len(cfg.projection_angles):
            seq = cfg.projection_angles[p]
        else:
            seq = (ora.angle("initial_angle") + p * ora.angle("D_angle")) * 180.0 / np.pi
        out[Path(pkg.mcio.projection_filename("projection", seq)).name] = img
    return out, launched


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reproduces_reference_cpu_golden(pkg, oracle_py, cases, name):
    inp, cfg, _ = cases[name]
    gold = np.load(GOLDEN / f"{name}.npz")
    ora = oracle_py.Oracle(inp, cxx_host_math=False)
    files, launched = oracle_files(pkg, ora, cfg)
    assert launched == int(gold["launched"])
    assert set(files) == set(gold.files) - {"launched"}
    for fname, img in files.items():
        assert np.array_equal(img, gold[fname]), f"{name}/{fname}: {(img != gold[fname]).sum()} pixels differ"
        assert img.sum() > 0


@pytest.mark.parametrize("name", ["water_p1", "thorax_p4"])
def test_oracle_reproduces_reference_cpu_binary_live(pkg, oracle_py, cases, name, tmp_path):
    if not oracle_py.REF_CPU.exists():
        pytest.skip("oracle/_ref not built here (no /root/reference); the committed golden vectors cover this")
    from conftest import build_case

    inp, cfg, _ = build_case(pkg, name, tmp_path)
    # a different seed and history count than the committed fixtures
    text = inp.read_text().replace("42  # RANDOM SEED", "20231  # RANDOM SEED").replace(f"{cfg.n_histories}  # TOTAL", "60000  # TOTAL")
    inp.write_text(text)
    oracle_py.run_reference_binary(oracle_py.REF_CPU, inp, cwd=tmp_path)
    ora = oracle_py.Oracle(inp, cxx_host_math=False)
    assert ora.seed == 20231
    files, launched = oracle_files(pkg, ora, cfg)
    det_cm = (round(cfg.detector_size[0] / 10, 6), round(cfg.detector_size[1] / 10, 6))
    for fname, img in files.items():
        ref = pkg.mcio.projection_counts(pkg.mcio.read_projection(tmp_path / fname, cfg.n_detector_pixels), cfg.n_detector_pixels, det_cm, launched)
        assert np.array_equal(img, ref), fname


def test_threaded_oracle_is_deterministic(oracle_py, cases):
    inp, _, _ = cases["water_p1"]
    ora = oracle_py.Oracle(inp)
    a, _ = ora.run_cpu_rule(0, threads=1)
    b, _ = ora.run_cpu_rule(0, threads=7)
    assert np.array_equal(a, b)


def test_gpu_partition_of_the_oracle_conserves_energy_and_differs_from_cpu_partition(oracle_py, cases):
    inp, cfg, _ = cases["water_p1"]
    ora = oracle_py.Oracle(inp, cxx_host_math=True)
    g, n_gpu = ora.run_gpu_rule(0, threads=8)
    c, n_cpu = ora.run_cpu_rule(0, threads=8)
    assert n_gpu == 211_200 and n_cpu == 200_100  # 11 blocks x 128 x 150 vs ceil(N/150) x 150
    # same physics, different stream partition: total detected energy per history agrees statistically
    assert abs(g.sum() / n_gpu - c.sum() / n_cpu) / (c.sum() / n_cpu) < 0.02
    # nothing can arrive with more than 90 keV (scaled by 100) per history
    assert g.sum() <= n_gpu * 90_000 * 100


def test_event_counters_match_survey_regime(oracle_py, cases):
    """SURVEY §8d B_hist inputs: voxel fetches, MFP fetches, Compton, ..., detector hits per history."""
    inp, _, _ = cases["water_p1"]
    ora = oracle_py.Oracle(inp)
    _, ev = ora.run_batches(0, 42, 150, 0, 200, threads=4, count_events=True)
    hist = float(ev[0])
    assert hist == 200 * 150
    v, f, c, r, pe, hits = [float(x) / hist for x in ev[1:7]]
    assert 3.0 < v < 30.0 and f <= v and 0.2 < c < 1.5 and 0.0 < r < 0.2 and 0.0 < pe < 0.3 and 0.2 < hits <= 1.0
