"""RANECU stream partition and launch-size rule: product host code vs the oracle's literal
restatement of abMODm / init_PRNG / update_seed_PRNG, plus the worked examples of SURVEY §8a."""
import numpy as np
import pytest


def test_mulmod_matches_russian_peasant(oracle_py):
    L = oracle_py.lib()
    rng = np.random.default_rng(1)
    for m in (2147483563, 2147483399):
        for a, s in rng.integers(1, m, size=(2000, 2)):
            assert L.oracle_abmodm(m, int(a), int(s)) == (int(a) * int(s)) % m


@pytest.mark.parametrize("hpt", [1, 150, 1431, 6010])
@pytest.mark.parametrize("seed", [1, 42, 657632199])
def test_stream_seeds_match_init_prng(pkg, oracle_py, hpt, seed):
    import ctypes as C

    L = oracle_py.lib()
    for stream in [0, 1, 2, 31, 127, 128, 4095, 65000 * 128 - 1]:
        a, b = C.c_int(), C.c_int()
        L.oracle_init_prng(stream, hpt, seed, C.byref(a), C.byref(b))
        assert pkg.engine.ranecu_init_stream(stream, hpt, seed) == (a.value, b.value)


def test_ranecu_sequence_matches_oracle(pkg, oracle_py):
    import ctypes as C

    L = oracle_py.lib()
    s1, s2 = pkg.engine.ranecu_init_stream(7, 150, 42)
    ours = pkg.engine.ranecu_sequence(s1, s2, 1000)
    a, b = C.c_int(s1), C.c_int(s2)
    ref = np.array([L.oracle_ranecu(C.byref(a), C.byref(b)) for _ in range(1000)], dtype=np.float32)
    assert np.array_equal(ours, ref)
    assert ours.min() > 0.0 and ours.max() < 1.0


def test_projection_seed_advance_matches_update_seed(pkg, oracle_py):
    L = oracle_py.lib()
    seed = 42
    for launched in (10_003_200, 595_180_800, 11_905_920_000, 50_003_200_000):
        assert pkg.engine.advance_projection_seed(seed, launched) == L.oracle_update_seed(1, launched, seed)
    # the reference log of the first GPU run (tests/gpu_check.py, 200 006 400 histories, seed 42)
    assert pkg.engine.advance_projection_seed(42, 200_006_400) == 657632199


@pytest.mark.parametrize("requested,blocks,hpt,launched", [
    (10_000_000, 521, 150, 10_003_200),              # config 1
    (595_166_015, 30_999, 150, 595_180_800),         # speed-up 20
    (1_190_332_031, 61_997, 150, 1_190_342_400),     # speed-up 10
    (238_066_406, 12_400, 150, 238_080_000),         # speed-up 50
    (11_903_320_312, 65_000, 1431, 11_905_920_000),  # reference quality: > 65535 blocks -> sticky hpt
    (50_000_000_000, 65_000, 6010, 50_003_200_000),  # air scan
    (1, 1, 150, 19_200),
])
def test_grid_rule_worked_examples(pkg, requested, blocks, hpt, launched):
    assert pkg.engine.grid_rule(requested, 128, 150) == (blocks, hpt, launched)
    assert pkg.mcio.launched_histories(requested, 128, 150) == (blocks, hpt, launched)


def test_projection_seed_schedule_is_closed_form(pkg, cases):
    inp, cfg, _ = cases["thorax_p4"]
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        launched = eng.info.launched_histories
        seed = cfg.random_seed
        for p in range(4):
            assert eng.projection_seed(p) == seed
            seed = pkg.engine.advance_projection_seed(seed, launched)
