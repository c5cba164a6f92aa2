"""Host-side pieces of bench.py that the GPU legs rely on (no GPU needed): the stdout line of the executable that the
scan_e2e leg parses, the device list handed to the single-process scan, the committed ncu constants."""
import json
import re
import subprocess
import sys

from conftest import ROOT


def test_projection_loop_line_of_the_executable_matches_the_parser_of_the_scan_leg():
    src = (ROOT / "4d-cbct-mc_b200" / "csrc" / "host" / "main.c").read_text()
    m = re.search(r'printf\("(\s+>>> Projection loop:.*?)\\n",\s*st\[0\]', src, re.S)
    assert m, "main.c no longer prints the projection-loop summary"
    fmt = re.sub(r'"\s*\n\s*"', "", m.group(1))  # adjacent C string literals
    line = fmt.replace("%.3f", "{:.3f}").replace("%.0f", "{:.0f}").replace("%.4f", "{:.4f}").format(18.475, 894, 8, 0.0207, 17.879, 8.129, 9.831)
    bench_src = (ROOT / "bench.py").read_text()
    pat = re.search(r'm = re\.search\(r"(Projection loop:.*?)", res\.stdout\)', bench_src, re.S)
    assert pat
    pattern = re.sub(r'"\s*\n\s*r"', "", pat.group(1))
    got = re.search(pattern, line)
    assert got and [float(x) for x in got.groups()] == [18.475, 894, 8, 0.0207, 17.879, 8.129, 9.831]


def test_visible_gpus_respects_the_environment():
    code = "import bench; print(bench.visible_gpus(2)); print(bench.visible_gpus(8))"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, env={"PATH": "/usr/bin:/bin", "CUDA_VISIBLE_DEVICES": "4,5,6"})
    assert out.stdout.split() == ["4,5", "4,5,6"], out.stderr[-500:]
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, env={"PATH": "/usr/bin:/bin"})
    assert out.stdout.split() == ["0,1", "0,1,2,3,4,5,6,7"], out.stderr[-500:]


def test_committed_ncu_constants_are_complete():
    c = json.loads((ROOT / "profiles" / "ncu_constants.json").read_text())
    for wl in ("catphan", "thorax", "air"):
        for key in ("warp_inst_per_history", "lanes_per_instruction", "lts_sectors_per_history", "l2_gather_peak_gbs", "atomics_per_history", "atomic_peak_gps", "source"):
            assert key in c[wl], (wl, key)
        assert (ROOT / c[wl]["source"].split(" ")[0]).exists(), c[wl]["source"]
