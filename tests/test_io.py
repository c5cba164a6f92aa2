"""Voxel-file ingest (text, gzip, in-memory), palette packing, and their error behaviour
(reference: load_voxels, MC-GPU_v1.3.cu:1996-2145)."""
import gzip
from pathlib import Path

import numpy as np  # noqa
import pytest


def unpack(packed, bits, n):
    if bits == 4:
        b = packed[: (n + 1) // 2]
        out = np.empty(2 * len(b), dtype=np.uint8)
        out[0::2] = b & 15
        out[1::2] = b >> 4
        return out[:n].astype(np.int64)
    if bits == 8:
        return packed[:n].astype(np.int64)
    if bits == 16:
        return packed.view(np.uint16)[:n].astype(np.int64)
    raise AssertionError(bits)


def load(pkg, inp, **kw):
    eng = pkg.engine.Engine()
    eng.load_input(inp)
    return eng


def test_vox_gz_plain_and_memory_agree(pkg, cases, tmp_path):
    from conftest import build_case

    inp_gz, cfg, ph = cases["thorax_p4"]
    inp_txt, _, _ = build_case(pkg, "thorax_p4", tmp_path / "plain", compressed=False)
    with load(pkg, inp_gz) as a, load(pkg, inp_txt) as b, load(pkg, inp_gz) as c:
        a.load_voxels()
        b.load_voxels()
        c.set_voxels(ph.materials, ph.densities, ph.spacing_cm)
        for t in ("voxel_material", "voxel_density", "voxel_packed", "density_max"):
            assert np.array_equal(a.table(t), b.table(t)), t
        # the text file carries %.6f densities; in-memory ones are the float32 originals
        assert np.array_equal(a.table("voxel_material"), c.table("voxel_material"))
        assert np.allclose(a.table("voxel_density"), c.table("voxel_density"), atol=5e-7)
        mat = a.table("voxel_material").reshape(ph.shape[::-1])  # z, y, x
        assert np.array_equal(mat, ph.materials.transpose(2, 1, 0))
        i = a.info
        assert (i.num_voxels_x, i.num_voxels_y, i.num_voxels_z) == ph.shape


@pytest.mark.parametrize("n_pairs,bits", [(3, 4), (16, 4), (17, 8), (256, 8), (257, 16), (5000, 16), (70000, 64)])
def test_palette_packing_round_trips(pkg, cases, n_pairs, bits):
    inp, _, _ = cases["water_p1"]
    rng = np.random.default_rng(n_pairs)
    shape = (41, 37, 53)  # odd sizes: exercises the nibble tail
    n = int(np.prod(shape))
    pair = rng.integers(0, n_pairs, size=n)
    pair[:n_pairs] = np.arange(n_pairs)  # every pair occurs
    mats = (pair % 22 + 1).astype(np.uint8).reshape(shape[::-1]).transpose(2, 1, 0)
    dens = (0.001 + (pair // 22 + 1) * 1e-4 + (pair % 22) * 0.05).astype(np.float32).reshape(shape[::-1]).transpose(2, 1, 0)
    with load(pkg, inp) as eng:
        eng.set_voxels(mats, dens, (0.1, 0.2, 0.3))
        info = eng.info
        assert info.voxel_bits == bits
        m = eng.table("voxel_material").astype(np.int64)
        d = eng.table("voxel_density")
        assert np.array_equal(m, mats.transpose(2, 1, 0).reshape(-1))
        assert np.array_equal(d, dens.transpose(2, 1, 0).reshape(-1))
        if bits != 64:
            assert info.palette_size == n_pairs
            idx = unpack(eng.table("voxel_packed"), bits, n)
            assert idx.max() == n_pairs - 1
            # same index <=> same (material, density) pair
            key = m * (1 << 32) + d.view(np.uint32).astype(np.int64)
            first = {}
            for i, k in zip(idx[:20000], key[:20000]):
                assert first.setdefault(int(i), int(k)) == int(k)
        dmax = eng.table("density_max")
        for mat in range(1, 23):
            sel = m == mat
            assert dmax[mat - 1] == (d[sel].max() if sel.any() else np.float32(-999.0))


def write_tiny_vox(path, body, n=(2, 1, 1)):
    text = f"[SECTION VOXELS HEADER v.2008-04-13]\n{n[0]} {n[1]} {n[2]}\n1.0 1.0 1.0\n[END OF VXH SECTION]\n" + body
    if str(path).endswith(".gz"):
        with gzip.open(path, "wt") as f:
            f.write(text)
    else:
        path.write_text(text)


@pytest.mark.parametrize("body,msg", [
    ("0 1.0\n1 1.0\n", "out of range"),        # material 0 (H:2120)
    ("26 1.0\n1 1.0\n", "out of range"),       # material > MAX_MATERIALS (H:2115)
    ("1 0.0\n1 1.0\n", "density"),             # density below 1e-9 (H:2126)
    ("1 1.0\n", "premature end"),              # ragged: fewer voxels than the header promises
    ("1 1.0\nxx yy\n", "expecting material"),  # garbage line
])
def test_bad_voxel_files_are_rejected(pkg, cases, tmp_path, body, msg):
    inp, _, _ = cases["water_p1"]
    vox = tmp_path / "bad.vox"
    write_tiny_vox(vox, body)
    with load(pkg, inp) as eng:
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_voxels(vox)
        assert e.value.code == -2 and msg in str(e.value)


def test_blank_and_comment_lines_are_skipped_like_the_reference(pkg, cases, tmp_path):
    inp, _, _ = cases["water_p1"]
    vox = tmp_path / "ok.vox.gz"
    write_tiny_vox(vox, "\n# comment\n1 0.001300\n\n\n 6 1.000000\n\n", n=(2, 1, 1))
    with load(pkg, inp) as eng:
        eng.load_voxels(vox)
        assert list(eng.table("voxel_material")) == [1, 6]
        assert list(eng.table("voxel_density")) == [np.float32(0.0013), np.float32(1.0)]
        assert eng.info.voxel_bits == 4 and eng.info.palette_size == 2


def test_missing_files(pkg, cases, tmp_path):
    inp, _, _ = cases["water_p1"]
    with load(pkg, inp) as eng:
        with pytest.raises(pkg.engine.McgpuError):
            eng.load_voxels(tmp_path / "nope.vox")
        eng.load_voxels()
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_materials([tmp_path / "nope.mcgpu"])
        assert e.value.code == -2


def test_material_used_by_voxels_but_not_listed_is_an_error(pkg, cases):
    inp, _, _ = cases["thorax_p4"]
    with load(pkg, inp) as eng:
        eng.load_voxels()
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_materials(pkg.mcio.material_paths()[:3])
        assert "has no material file" in str(e.value)


def test_spectrum_outside_table_range_is_rejected(pkg, cases, tmp_path):
    inp, _, _ = cases["water_p1"]
    spc = tmp_path / "hot.spc"
    spc.write_text("100.0e3 1.0\n150.0e3 1.0\n200.0e3 -1\n")
    text = open(inp).read()
    old = [line for line in text.splitlines() if "X-RAY ENERGY SPECTRUM FILE" in line][0]
    f = tmp_path / "hot.in"
    f.write_text(text.replace(old, f"{spc}  # X-RAY ENERGY SPECTRUM FILE"))
    with pkg.engine.Engine() as eng:
        eng.load_input(f).load_voxels()
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_materials()
        assert e.value.code == -1 and "outside the tabulated energy interval" in str(e.value)


def test_fast_density_scanner_equals_strtof(pkg, cases, tmp_path):
    """The streaming ingest scans densities with integer arithmetic; every value must equal glibc's
    strtof (what the reference's sscanf("%f") gives), including exponent forms and long mantissas."""
    import ctypes

    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    rng = np.random.default_rng(11)
    n = 60000
    vals = np.concatenate([rng.random(n // 2) * 3.0, 10.0 ** rng.uniform(-8.5, 1.5, n // 2)])
    fmts = ["%.6f", "%.3f", "%.9f", "%.12f", "%.17g", "%.7e", "%g", "%.1f0000"]
    tokens = []
    for i, v in enumerate(vals):
        tok = fmts[i % len(fmts)] % v
        if float(tok) < 1e-9:
            tok = "0.001300"
        tokens.append(tok)
    tokens[:6] = ["1", "2.", ".5", "1e0", "+0.25", "0.0013000000000000000001"]
    body = "".join(f"{1 + i % 22} {t}\n" for i, t in enumerate(tokens))
    vox = tmp_path / "dens.vox"
    vox.write_text(f"[SECTION VOXELS HEADER v.2008-04-13]\n{n} 1 1\n1.0 1.0 1.0\n[END OF VXH SECTION]\n" + body)
    inp, _, _ = cases["water_p1"]
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels(vox)
        got = eng.table("voxel_density")
    want = np.array([libc.strtof(t.encode(), None) for t in tokens], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_large_file_spans_many_blocks_and_ignores_trailing_lines(pkg, cases, tmp_path):
    """More than one 4 MB block, rows separated by the blank lines cbctmc writes, extra lines after the last voxel."""
    ph = pkg.phantoms.thorax(shape=(96, 96, 60), spacing_mm=4.0)
    vox = pkg.mcio.write_vox(tmp_path / "big.vox", ph.materials, ph.densities, ph.spacing_cm)
    assert vox.stat().st_size > 5 * (4 << 20) // 4
    with open(vox, "a") as f:
        f.write("\n# trailing comment\n7 1.000000\n")
    inp, _, _ = cases["water_p1"]
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels(vox)
        assert np.array_equal(eng.table("voxel_material").reshape(60, 96, 96), ph.materials.transpose(2, 1, 0))
        want = np.array([float("%.6f" % d) for d in np.unique(ph.densities)], dtype=np.float32)
        assert set(np.unique(eng.table("voxel_density")).tolist()) == set(want.tolist())


def test_binary_geometry_gives_the_same_volume_as_the_text_file(pkg, cases, tmp_path):
    """.voxb: same packed voxels, palette and density maxima as the .vox.gz of the same arrays"""
    inp, cfg, ph = cases["thorax_p4"]
    with pkg.engine.Engine() as eng:
        eng.load_input(inp).load_voxels()
        ref = {t: eng.table(t).copy() for t in ("voxel_packed", "voxel_material", "voxel_density")}
        ref_info = eng.info
    for compress in (False, True):
        f = pkg.mcio.write_voxb(tmp_path / f"g{int(compress)}.voxb", ph.materials, ph.densities, ph.spacing_cm, compress=compress)
        with pkg.engine.Engine() as eng:
            eng.load_input(inp).load_voxels(f)
            for t, a in ref.items():
                assert np.array_equal(eng.table(t), a), t
            i = eng.info
            assert (i.num_voxels_x, i.num_voxels_y, i.num_voxels_z, i.voxel_bits, i.palette_size) == (
                ref_info.num_voxels_x, ref_info.num_voxels_y, ref_info.num_voxels_z, ref_info.voxel_bits, ref_info.palette_size)
    raw = (tmp_path / "g0.voxb").read_bytes()
    (tmp_path / "short.voxb").write_bytes(raw[: len(raw) // 2])
    bad = bytearray(raw)
    bad[36] = 77  # first material byte
    (tmp_path / "bad.voxb").write_bytes(bytes(bad))
    with pkg.engine.Engine() as eng:
        eng.load_input(inp)
        with pytest.raises(pkg.engine.McgpuError, match="ends inside"):
            eng.load_voxels(tmp_path / "short.voxb")
        with pytest.raises(pkg.engine.McgpuError, match="out of range"):
            eng.load_voxels(tmp_path / "bad.voxb")


def test_header_only_material_stub_is_named_as_such(pkg, cases, tmp_path):
    """A material file cut after its nominal density (an asset staged without rows) is fine for a material the voxels do
    not use (H:2220-2233 reads only the header) but must be reported as a stub -- not as 'incorrect number of energy
    values: input=0' -- when the voxels use it."""
    import gzip
    import re

    inp, _, _ = cases["thorax_p4"]
    paths = list(pkg.mcio.material_paths())
    numbers = pkg.mcio.material_numbers()

    def stub_of(ident):
        src = paths[numbers[ident] - 1]
        payload = gzip.open(src, "rb").read()
        m = re.search(rb"\[NOMINAL DENSITY[^\n]*\n[^\n]*\n", payload)
        dst = tmp_path / f"{ident}_stub.mcgpu.gz"
        with gzip.open(dst, "wb") as f:
            f.write(payload[: m.end()] + b"#[STUB: rows omitted]\n")
        return dst

    with load(pkg, inp) as eng:
        eng.load_voxels()
        unused = list(paths)
        unused[numbers["teflon"] - 1] = stub_of("teflon")  # not in the thorax phantom
        eng.load_materials(unused)
        used = list(paths)
        used[numbers["lung"] - 1] = stub_of("lung")
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.load_materials(used)
        assert e.value.code == -2 and "header-only stub" in str(e.value) and "lung" in str(e.value)


def test_every_material_of_the_patient_set_ships_in_full(pkg):
    """cbctmc's patient geometries use blood (40 shells = MAX_SHELLS), red_marrow (36), liver, muscle, stomach, glands,
    cartilage (cbctmc/mc/geometry.py:161, 214-229): all 22 files carry their tables."""
    import gzip

    rows = pkg.mcio.material_table()
    assert len(rows) == 22
    for _, ident, _, path in rows:
        text = gzip.open(path, "rt").read()
        assert "STUB" not in text and text.count("\n") > 20000, ident


def test_content_addressed_geometry_cache(pkg, cases, tmp_path, monkeypatch):
    """MCGPU_CACHE_DIR: a text geometry is stored once as a binary file named after the hash of its bytes;
    the same bytes load from it (same volume, same packing), other bytes get another entry, a corrupt entry is ignored,
    and without the variable nothing is written."""
    inp, cfg, ph = cases["thorax_p4"]
    vox = Path(inp).parent / "geometry.vox.gz"
    cache = tmp_path / "cache"
    cache.mkdir()

    def volume():
        with load(pkg, inp) as eng:
            eng.load_voxels(vox)
            return eng.table("voxel_material").copy(), eng.table("voxel_density").copy(), eng.table("voxel_packed").copy(), eng.info.voxel_bits

    plain = volume()
    assert list(cache.iterdir()) == []
    monkeypatch.setenv("MCGPU_CACHE_DIR", str(cache))
    first = volume()
    entries = list(cache.iterdir())
    assert len(entries) == 1 and entries[0].name.startswith("vox_") and entries[0].suffix == ".voxb"
    mtime = entries[0].stat().st_mtime_ns
    second = volume()  # served from the cache: the entry is not rewritten
    assert list(cache.iterdir()) == entries and entries[0].stat().st_mtime_ns == mtime
    for a, b, c in zip(plain, first, second):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    # the cache file is the documented binary geometry: loading it directly gives the same volume
    with load(pkg, inp) as eng:
        eng.load_voxels(entries[0])
        assert np.array_equal(eng.table("voxel_packed"), plain[2])
    # other bytes -> another entry
    other = tmp_path / "other.vox.gz"
    pkg.mcio.write_vox(other, ph.materials[::-1].copy(), ph.densities[::-1].copy(), ph.spacing_cm)
    with load(pkg, inp) as eng:
        eng.load_voxels(other)
    assert len(list(cache.iterdir())) == 2
    # a damaged entry is not trusted: the text is parsed again and the entry replaced
    entries[0].write_bytes(b"MCGPUVXB" + b"\0" * 10)
    third = volume()
    assert np.array_equal(third[2], plain[2]) and entries[0].stat().st_size > 1000


@pytest.mark.parametrize("shape,n_pairs,bits", [((101, 103, 101), 11, 4), ((128, 128, 70), 200, 8), ((128, 128, 70), 5000, 16)])
def test_chunk_parallel_packing_equals_first_occurrence_order(pkg, cases, shape, n_pairs, bits):
    """Volumes above 2^20 voxels are packed by several threads (voxels.c: finish_volume): the palette must still be in order of
    first occurrence in the whole volume and the packed indices what one sequential pass would write -- also for an odd
    number of voxels at 4 bits -- and a bad voxel must be reported at its own (first) position."""
    rng = np.random.default_rng(7)
    n = int(np.prod(shape))
    assert n > (1 << 20)
    mats = rng.integers(1, 8, size=n_pairs).astype(np.uint8)
    dens = (0.05 + np.arange(n_pairs) * 1e-3).astype(np.float32)
    # long runs of one pair (like anatomy), every pair present, pair 0 only in the last quarter (late first occurrence)
    runs = rng.integers(0, n_pairs, size=n // 97 + 1)
    runs[: 3 * len(runs) // 4][runs[: 3 * len(runs) // 4] == 0] = 1
    which = np.repeat(runs, 97)[:n]
    which[-n_pairs:] = np.arange(n_pairs)
    m = mats[which].reshape(shape[::-1])  # [z][y][x], x fastest
    r = dens[which].reshape(shape[::-1])
    inp, _, _ = cases["water_p1"]
    with load(pkg, inp) as eng:
        eng.set_voxels(m.transpose(2, 1, 0), r.transpose(2, 1, 0), (0.1, 0.1, 0.1))
        info = eng.info
        assert info.voxel_bits == bits
        key = (m.reshape(-1).astype(np.uint64) << np.uint64(32)) | r.reshape(-1).view(np.uint32).astype(np.uint64)
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        order = np.argsort(first)
        rank = np.empty_like(order)
        rank[order] = np.arange(len(order))
        gi = rank[inv]
        assert info.palette_size == len(uniq)
        packed = eng.table("voxel_packed")
        if bits == 4:
            pad = np.concatenate([gi, [0]]) if n % 2 else gi
            want = (pad[0::2] | (pad[1::2] << 4)).astype(np.uint8)
        elif bits == 8:
            want = gi.astype(np.uint8)
        else:
            want = gi.astype(np.uint16).view(np.uint8)
        assert np.array_equal(packed, want)
        dmax = eng.table("density_max")
        for mat in range(1, 8):
            sel = m.reshape(-1) == mat
            assert dmax[mat - 1] == (r.reshape(-1)[sel].max() if sel.any() else np.float32(-999.0))
        bad = m.copy().reshape(-1)
        bad[n - 5] = 0
        bad[n // 2 + 3] = 26
        with pytest.raises(pkg.engine.McgpuError) as e:
            eng.set_voxels(bad.reshape(shape[::-1]).transpose(2, 1, 0), r.transpose(2, 1, 0), (0.1, 0.1, 0.1))
        assert f"voxel number {n // 2 + 4}" in str(e.value) and "26" in str(e.value)
