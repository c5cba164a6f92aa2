#!/usr/bin/env python3
"""Attribute an ncu SASS-level profile to CUDA source lines (ncu's CSV export of the CUDA view has
no metrics): nvdisasm -g gives the line of every SASS instruction of the kernel in the in-tree
library, ncu --page source --csv gives per-instruction counters in the same order.
Usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-mangled-substring> [top_n]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def line_map(kernel_sub):
    tmp = tempfile.mkdtemp()
    # the library holds two cubins of the same name (exact / fast-math build of launch.cu): take the object file of the build asked for
    obj = ROOT / "build" / ("launch_fast.o" if "mcgpu_fast" in kernel_sub else "launch_exact.o")
    subprocess.run(["cuobjdump", "-xelf", "all", str(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = next(Path(tmp).glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and kernel_sub in l)
    out, cur = [], ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l):
            out.append(cur)
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    lm = line_map(ksub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    body = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
    assert len(body) == len(lm), (len(body), len(lm))
    agg = {}
    for (f, ln), d in zip(lm, body):
        a = agg.setdefault((f, ln), [0, 0, 0])
        a[0] += int(d["Instructions Executed"] or 0)
        a[1] += int(d["Thread Instructions Executed"] or 0)
        a[2] += int(d["# Samples"] or 0)
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[2] for a in agg.values())
    tht = sum(a[1] for a in agg.values())
    print(f"total warp-inst {tot:.4g}, thread-inst {tht:.4g}, lanes/inst {tht / tot:.2f}, samples {tots}")
    src = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            p = ROOT / "4d-cbct-mc_b200/csrc/cuda" / f
            src[f] = p.read_text().splitlines() if p.exists() else []
        text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
        print(f"{f:14s} L{ln:<4d} inst {100 * a[0] / tot:5.2f}%  lanes {a[1] / max(a[0], 1):5.1f}  stall-samples {100 * a[2] / max(tots, 1):5.2f}% | {text}")


if __name__ == "__main__":
    main()


def regions():
    """function name per line of transport.cuh / regroup.cuh (top-level __device__ functions and kernel phases)"""
    out = {}
    for fname in ("transport.cuh", "regroup.cuh"):
        text = (ROOT / "4d-cbct-mc_b200/csrc/cuda" / fname).read_text().splitlines()
        cur = "-"
        for i, l in enumerate(text, start=1):
            m = re.match(r"\s*(?:template <[^>]*>\s*)?__device__ __forceinline__ [\w:<> &*]+?\s+(\w+)\(", l) or re.match(r"\s*__device__ __forceinline__ explicit (\w+)\(", l)
            if m:
                cur = m.group(1)
            m = re.match(r"\s*// -{20,} (\w+):", l)
            if m and fname == "regroup.cuh":
                cur = "phase_" + m.group(1)
            out[(fname, i)] = cur
    return out
