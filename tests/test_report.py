"""Projection writer: byte-compatible with report_image (MC-GPU_v1.3.cu:2783-2953).
Structure is checked against a projection file written by the reference CPU binary
(tests/golden/water_p1_projection_ascii.gz); the number formatting (our exact integer "%.8lf") is
checked byte for byte against an independent correctly-rounded formatter."""
import gzip
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = ROOT / "tests" / "golden"


@pytest.fixture()
def written(pkg, cases, tmp_path):
    inp, cfg, _ = cases["water_p1"]
    gold = np.load(GOLDEN / "water_p1.npz")
    counts = gold["projection_000.000000deg"]
    text = open(inp).read().replace(str(Path(inp).parent) + "/projection", str(tmp_path / "projection"))
    f = tmp_path / "w.in"
    f.write_text(text)
    with pkg.engine.Engine() as eng:
        eng.load_input(f)
        eng.write_projection(0, counts, 1.25)
        launched = eng.info.launched_histories
        name = eng.projection_filename(0)
    return Path(name).read_text().splitlines(), counts, launched, cfg


def test_structure_matches_reference_file(written):
    ours, counts, launched, cfg = written
    ref = gzip.open(GOLDEN / "water_p1_projection_ascii.gz", "rt").read().splitlines()
    assert len(ours) == len(ref)
    volatile = ("SIMULATION IN THE", "Simulated x rays", "Simulation time", "Speed [x-rays/sec]", "Fraction of energy", "Maximum energy detected")
    for a, b in zip(ours, ref):
        if a.startswith("#") or b.startswith("#"):
            if any(v in b for v in volatile):
                assert a.startswith("#") and any(v in a for v in volatile)
            else:
                assert a == b  # banner, angle, focal spot, pixel size, column legend: identical text
        else:
            assert (a == "") == (b == "")  # blank line after every detector row
            if a:
                assert len(a.split()) == 4


def test_numbers_are_formatted_like_printf_8f(written, pkg):
    ours, counts, launched, cfg = written
    nx, nz = cfg.n_detector_pixels
    inv_x = float(np.float32(nx) / np.float32(round(cfg.detector_size[0] / 10, 6)))
    inv_z = float(np.float32(nz) / np.float32(round(cfg.detector_size[1] / 10, 6)))
    norm = (1.0 / 100.0) * inv_x * inv_z / float(launched)
    data = [line for line in ours if line and not line.startswith("#")]
    assert len(data) == nx * nz
    flat = counts.reshape(4, -1).astype(np.float64)
    for k, line in enumerate(data):
        expect = " ".join("%.8f" % (norm * flat[j, k]) for j in range(4))
        assert line == expect, (k, line, expect)


def test_f8_formatter_on_adversarial_values(pkg, cases, tmp_path):
    """Ties, tiny and huge tallies: the writer must agree with correctly-rounded %.8f for any u64 count."""
    inp, cfg, _ = cases["water_p1"]
    text = open(inp).read().replace(str(Path(inp).parent) + "/projection", str(tmp_path / "projection"))
    f = tmp_path / "w.in"
    f.write_text(text)
    nx, nz = cfg.n_detector_pixels
    rng = np.random.default_rng(5)
    counts = np.zeros((4, nz, nx), dtype=np.uint64)
    flat = counts.reshape(-1)
    flat[:] = rng.integers(0, 2**40, size=flat.size, dtype=np.uint64)
    flat[::7] = rng.integers(0, 50, size=flat[::7].size, dtype=np.uint64)
    flat[1::13] = rng.integers(2**52, 2**62, size=flat[1::13].size, dtype=np.uint64)
    with pkg.engine.Engine() as eng:
        eng.load_input(f)
        for hist in (200_000, 11_903_320_312):
            eng.set_histories(hist)
            eng.write_projection(0, counts, 0.0)
            launched = eng.info.launched_histories
            lines = [l for l in Path(eng.projection_filename(0)).read_text().splitlines() if l and not l.startswith("#")]
            inv_x = float(np.float32(nx) / np.float32(round(cfg.detector_size[0] / 10, 6)))
            inv_z = float(np.float32(nz) / np.float32(round(cfg.detector_size[1] / 10, 6)))
            norm = (1.0 / 100.0) * inv_x * inv_z / float(launched)
            c = counts.reshape(4, -1).astype(np.float64)
            for k in range(0, len(lines), 3):
                assert lines[k] == " ".join("%.8f" % (norm * c[j, k]) for j in range(4))


def test_reader_round_trip(written, pkg, tmp_path):
    ours, counts, launched, cfg = written
    f = tmp_path / "again"
    f.write_text("\n".join(ours) + "\n")
    vals = pkg.mcio.read_projection(f, cfg.n_detector_pixels)
    det_cm = (round(cfg.detector_size[0] / 10, 6), round(cfg.detector_size[1] / 10, 6))
    assert np.array_equal(pkg.mcio.projection_counts(vals, cfg.n_detector_pixels, det_cm, launched), counts)


def test_binary_side_file_matches_ascii(pkg, cases, tmp_path):
    inp, cfg, _ = cases["water_p1"]
    counts = np.load(GOLDEN / "water_p1.npz")["projection_000.000000deg"]
    text = open(inp).read().replace(str(Path(inp).parent) + "/projection", str(tmp_path / "projection"))
    f = tmp_path / "w.in"
    f.write_text(text)
    with pkg.engine.Engine() as eng:
        eng.load_input(f)
        eng.write_projection(0, counts, 0.0)
        eng.write_projection_raw(0, counts)
        name = eng.projection_filename(0)
    raw = pkg.mcio.read_projection_raw(name + ".raw", cfg.n_detector_pixels)
    txt = pkg.mcio.read_projection(name, cfg.n_detector_pixels)
    assert raw.shape == txt.shape == (4, cfg.n_detector_pixels[1], cfg.n_detector_pixels[0])
    assert np.allclose(raw, txt, rtol=1e-6, atol=1e-8) and raw.sum() > 0
