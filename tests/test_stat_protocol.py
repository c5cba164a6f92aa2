"""The statistics of tools/stat_protocol.py, checked on synthetic tallies (CPU): under the null hypothesis (two programs sampling
the same distribution with independent seeds) the pooled-variance criterion passes with ~99.7 % of pixels inside |z| < 3, the
raw 16-seed variances give the Student-t expectation (~99.46 %), and a 1 % bias of one program or a localised defect fails."""
import importlib.util
import sys

import numpy as np

from conftest import ROOT


def load_compare():
    sys.argv = ["stat_protocol.py"]
    spec = importlib.util.spec_from_file_location("stat_protocol", ROOT / "tools" / "stat_protocol.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.compare


def synthetic(rng, K, bias=1.0, defect=None):
    """K seeds x 4 planes x 96 x 231 'tallies': compound-Poisson counts (number of hits x energy), smooth mean image"""
    z, x = np.mgrid[0:96, 0:231]
    mean_hits = 4000.0 * np.exp(-(((x - 115) / 90.0) ** 2 + ((z - 48) / 40.0) ** 2))
    planes = []
    for scale in (1.0, 0.2, 0.03, 0.1):
        lam = mean_hits * scale * bias
        if defect is not None:
            lam = lam.copy()
            lam[defect] *= 1.05
        hits = rng.poisson(lam, size=(K,) + lam.shape).astype(np.float64)
        energy = 5.0e6 * (1.0 + 0.25 * rng.standard_normal(size=hits.shape) / np.sqrt(np.maximum(hits, 1.0)))
        planes.append(hits * energy)
    return np.stack(planes, axis=1)


def test_null_hypothesis_passes_and_matches_the_expected_fractions(pkg):
    compare = load_compare()
    rng = np.random.default_rng(11)
    out = compare(synthetic(rng, 16), synthetic(rng, 16), "null")
    assert out["pass_z_pooled_997"] and out["pass_mean_rel_diff_0p5pct"]
    t = out["total"]
    assert 0.9955 < t["frac_abs_z_pooled_lt_3"] < 0.9990
    assert 0.9915 < t["frac_abs_z_lt_3"] < 0.9970  # Student-t with ~30 degrees of freedom: 99.46 % expected
    assert 0.95 < t["z_pooled_std"] < 1.05 and abs(t["z_mean"]) < 0.05


def test_a_one_percent_bias_or_a_local_defect_fails(pkg):
    compare = load_compare()
    rng = np.random.default_rng(12)
    biased = compare(synthetic(rng, 16, bias=1.01), synthetic(rng, 16), "bias")
    assert not biased["pass_mean_rel_diff_0p5pct"]
    assert abs(biased["total"]["rel_diff_of_plane_sum"] - 0.01) < 2e-3
    defect = (slice(30, 60), slice(80, 150))  # 5 % more counts in a 30 x 70 pixel patch, < 0.5 % of the plane sums
    local = compare(synthetic(rng, 16, defect=defect), synthetic(rng, 16), "defect")
    assert not local["pass_z_pooled_997"]
