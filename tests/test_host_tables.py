"""Product host code (csrc/host) vs the oracle's independent restatement, bit for bit: poses of
every projection, spectrum alias tables, MFP / Rayleigh / Compton tables.  The oracle is used in
its nvcc-host flavour here (float overloads), the flavour of the production reference build; its
plain-C flavour is pinned to the reference CPU binary in test_oracle_pinned.py."""
import numpy as np
import pytest

from conftest import CASES

TABLES = ["woodcock", "rayleigh_xco", "rayleigh_pco", "rayleigh_aco", "rayleigh_bco", "rayleigh_itlco", "rayleigh_ituco",
          "compton_fco", "compton_uico", "compton_fj0", "compton_noscco", "density_nominal", "density_max", "espc",
          "espc_cutoff", "espc_alias"]


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("name", list(CASES))
def test_tables_and_views_match_oracle(pkg, oracle_py, cases, name):
    inp, cfg, _ = cases[name]
    ora = oracle_py.Oracle(inp, cxx_host_math=True)
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels().load_materials()
        info = eng.info
        assert info.num_projections == ora.num_projections
        assert info.num_energy_values == ora.num_values == 24001
        assert info.e0 == ora.scalar("e0") and info.ide == ora.scalar("ide")
        assert info.mean_energy_spectrum == ora.scalar("mean_energy")
        for t in TABLES:
            assert np.array_equal(bits(eng.table(t)), bits(ora.table(t))), t
        used = np.nonzero(ora.table("compton_noscco"))[0]
        assert info.num_materials_used <= len(used)
        for t, w in (("mfp_a", 3), ("mfp_b", 3), ("rayleigh_pmax", 1)):
            a = eng.table(t).reshape(-1, 25, w)[:, used]
            b = ora.table(t).reshape(-1, 25, w)[:, used]
            assert np.array_equal(bits(a), bits(b)), t
        v = eng.views()
        P = info.num_projections
        src = ora.table("source").reshape(P, 20)
        det = ora.table("detector").view(np.float32).reshape(P, 25)
        assert np.array_equal(bits(v[:, 0:6]), bits(src[:, 0:6])), "source position/direction"
        flag = v[:, 44].view(np.int32)
        assert np.array_equal(flag, det[:, 24].view(np.int32))
        if flag[0] == 1:
            assert np.array_equal(bits(v[:, 6:15]), bits(src[:, 6:15])), "rot_fan"
        assert np.array_equal(bits(v[:, 15:20]), bits(src[:, 15:20])), "apertures"
        assert np.array_equal(bits(v[:, 20:23]), bits(det[:, 5:8])), "detector centre"
        assert np.array_equal(bits(v[:, 23:26]), bits(det[:, 2:5])), "detector corner"
        assert np.array_equal(bits(v[:, 26:35]), bits(det[:, 8:17])), "rot_inv"
        assert np.array_equal(bits(v[:, 35:37]), bits(det[:, 19:21])), "inverse pixel size"
    ora.close()


def test_cxx_and_c_host_flavours_differ_where_predicted(oracle_py, cases):
    """acos(float) is acosf under nvcc (C++) and acos under gcc -x c: rotX = acosf(0) - pi/2 = 4.37e-8."""
    inp, _, _ = cases["thorax_p4"]
    a = oracle_py.Oracle(inp, cxx_host_math=True).table("detector").view(np.float32).reshape(-1, 25)
    b = oracle_py.Oracle(inp, cxx_host_math=False).table("detector").view(np.float32).reshape(-1, 25)
    assert abs(a[0, 8 + 5]) == pytest.approx(4.3711388e-08, rel=1e-6) and b[0, 8 + 5] == 0.0


def test_views_of_a_full_scan_are_a_circle(pkg, cases, tmp_path):
    """894 projections: sources on a circle of radius SAD about the isocentre, detector centre SDD away."""
    from conftest import build_case

    inp, cfg, ph = cases["thorax_p4"]
    text = open(inp).read().replace("4  # NUMBER OF PROJECTIONS", "894  # NUMBER OF PROJECTIONS").replace(
        "90.0  # ANGLE BETWEEN", f"{360.0 / 894}  # ANGLE BETWEEN")
    f = tmp_path / "full.in"
    f.write_text(text)
    with pkg.engine.Engine() as eng:
        eng.load_input(f)
        v = eng.views().astype(np.float64)
        assert v.shape == (894, 45)
        iso = np.array(ph.size_mm) / 20.0
        r = np.hypot(v[:, 0] - iso[0], v[:, 1] - iso[1])
        assert np.allclose(r, 100.0, atol=1e-3)
        d = np.linalg.norm(v[:, 20:23] - v[:, 0:3], axis=1)
        assert np.allclose(d, 150.0, atol=1e-3)
        assert np.allclose(np.linalg.norm(v[:, 3:6], axis=1), 1.0, atol=1e-6)
        names = [eng.projection_filename(p).rsplit("_", 1)[1] for p in (0, 1, 893)]
        assert names[0] == "270.000000deg" and names[2] == "629.597290deg"  # sequential angle, Q5
