#!/usr/bin/env python3
"""The scan as cbctmc runs it, measured end to end on the GPU box: `MC-GPU_v1.3.x input.in` (ONE process driving N GPUs,
mcgpu_run_all) on P projections of a workload at the speed-up-20 history count, one ASCII report per projection written to
files.  Prints the same object as bench.py's `scan_e2e` leg.
Usage: python tools/scan_e2e.py [--workload catphan] [--gpus 8] [--projections 894]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="catphan", choices=list(bench.WORKLOADS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--projections", type=int, default=894)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    factory, n_hist, desc = bench.WORKLOADS[args.workload]
    out = bench.scan_e2e_leg(args, factory(), n_hist, args.gpus, args.projections)
    out["workload"] = desc
    text = json.dumps(out, indent=1)
    print(text)
    if args.out:
        Path(args.out).write_text(text)


if __name__ == "__main__":
    main()
