#!/usr/bin/env python3
"""Attribute an ncu SASS-level profile of transport_wavefront to FUNCTIONS: every SASS instruction carries the source line it
was inlined from (nvdisasm -g); transport.cuh lines are mapped to the enclosing __device__ function, wavefront.cuh lines to
the section of the kernel loop (acquire / load / tracking / N / C-R / store+push).  RANECU and libm lines are shared by
all phases and listed as such.  Usage: python tools/ncu_functions.py <report.ncu-rep> <kernel-mangled-substring>"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
import ncu_lines  # noqa: E402


def function_of_line():
    out = {}
    text = (ROOT / "4d-cbct-mc_b200/csrc/cuda/transport.cuh").read_text().splitlines()
    cur = "-"
    for i, l in enumerate(text, start=1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?__device__ __forceinline__ [\w:<> &*]+?\s+(\w+)\(", l) or re.match(r"\s*__device__ __forceinline__ explicit (\w+)\(", l)
        if m:
            cur = m.group(1)
        if re.match(r"\s*struct Ranecu", l):
            cur = "Ranecu"
        out[("transport.cuh", i)] = cur
    w = (ROOT / "4d-cbct-mc_b200/csrc/cuda/wavefront.cuh").read_text().splitlines()
    marks = [(r"acquire a batch", "wavefront: acquire"), (r"^\s*Photon p;", "wavefront: load context"), (r"if \(q == Q_W\)", "wavefront: tracking loop"),
             (r"else if \(q == Q_N\)", "wavefront: tally / stream / source section"), (r"C / CT / R: one scattering", "wavefront: Compton / Rayleigh section"),
             (r"store what every kind", "wavefront: store + push")]
    cur = "wavefront: prologue"
    for i, l in enumerate(w, start=1):
        for pat, name in marks:
            if re.search(pat, l):
                cur = name
        out[("wavefront.cuh", i)] = cur
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    lm = ncu_lines.line_map(ksub)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    body = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
    assert len(body) == len(lm), (len(body), len(lm))
    fol = function_of_line()
    agg = defaultdict(lambda: [0, 0, 0])
    for (f, ln), d in zip(lm, body):
        key = fol.get((f, ln)) or f"[{f}]"
        a = agg[key]
        a[0] += int(d["Instructions Executed"] or 0)
        a[1] += int(d["Thread Instructions Executed"] or 0)
        a[2] += int(d["# Samples"] or 0)
    tot = sum(a[0] for a in agg.values())
    ts = sum(a[2] for a in agg.values())
    th = sum(a[1] for a in agg.values())
    print(f"total warp-instructions {tot:.4g}, lanes per instruction {th / tot:.2f}")
    print(f"{'function / section':48s} {'% instr':>8s} {'lanes':>6s} {'% stall samples':>16s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] / tot < 0.002:
            continue
        print(f"{k:48s} {100 * a[0] / tot:8.2f} {a[1] / max(a[0], 1):6.1f} {100 * a[2] / max(ts, 1):16.2f}")


if __name__ == "__main__":
    main()
