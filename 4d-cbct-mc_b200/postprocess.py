"""Host-side mirror of cbctmc's projection post-processing (cbctmc/mc/projection.py:125-176,
cbctmc/mc/simulation.py:235-277) on top of the device entry points of include/mcgpu_b200.h
(`mcgpu_post_intensity / _gaussian / _normalize`): stacks of total / unscattered / scattered intensity and the
air-normalised line integrals, from the u64 tallies -- no text files, no np.loadtxt.  All arithmetic runs on
the GPU; this module only orders the calls and writes MetaImage (.mha) files like `sitk.WriteImage` would."""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

MODES = ("total", "unscattered", "scattered")


class ProjectionStacks:
    """Accumulates the three intensity stacks projection by projection (one `post_intensity` call each)."""

    def __init__(self, engine, n_projections: int, crop_x: int = 1024):
        info = engine.info
        self.engine = engine
        crop = crop_x if 0 < crop_x <= info.num_pixels_x else info.num_pixels_x
        self.stacks = {m: np.empty((n_projections, info.num_pixels_z, crop), dtype=np.float32) for m in MODES}
        self.min_positive = {m: np.float32(np.inf) for m in MODES}
        self.crop, self.count = crop, 0

    def add(self, tally: np.ndarray | None = None, launched: int | None = None):
        """tally None: the projection the engine simulated last (still on the device)."""
        t, u, s, mins = self.engine.post_intensity(tally, launched, self.crop)
        for k, (m, img) in enumerate(zip(MODES, (t, u, s))):
            self.stacks[m][self.count] = img
            self.min_positive[m] = min(self.min_positive[m], mins[k])
        self.count += 1

    def stack(self, mode: str = "total", air: np.ndarray | None = None, sigma: Sequence[float] | None = (10, 10)) -> np.ndarray:
        """projections_to_itk: zeros -> the stack's smallest positive value; with `air` (mode total) Beer-Lambert line integrals."""
        s = self.stacks[mode][: self.count]
        mn = self.min_positive[mode]
        if air is not None and mode == "total":
            a = np.asarray(air, dtype=np.float32)
            if sigma:
                a = self.engine.post_gaussian(a, sigma)
            return self.engine.post_normalize(a, s.copy(), mn)
        return np.where(s == 0, mn, s)


def write_mha(path, stack: np.ndarray, pixel_size: Sequence[float]):
    """MetaImage file as projections_to_itk + sitk.WriteImage produce it (projection.py:167-176): spacing (dx, dy, 1), centred origin."""
    a = np.ascontiguousarray(stack, dtype="<f4")
    nz, ny, nx = a.shape
    header = (
        "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\nCompressedData = False\n"
        "TransformMatrix = 1 0 0 0 1 0 0 0 1\n"
        f"Offset = {-nx * pixel_size[0] / 2:.17g} {-ny * pixel_size[1] / 2:.17g} 0\n"
        "CenterOfRotation = 0 0 0\nAnatomicalOrientation = RAI\n"
        f"ElementSpacing = {pixel_size[0]:.17g} {pixel_size[1]:.17g} 1\n"
        f"DimSize = {nx} {ny} {nz}\nElementType = MET_FLOAT\nElementDataFile = LOCAL\n"
    )
    with open(path, "wb") as f:
        f.write(header.encode())
        f.write(a.tobytes())
    return Path(path)


def read_mha(path) -> np.ndarray:
    raw = Path(path).read_bytes()
    marker = b"ElementDataFile = LOCAL\n"
    head = raw[: raw.index(marker)].decode()
    dims = [int(x) for x in next(l for l in head.splitlines() if l.startswith("DimSize")).split("=")[1].split()]
    return np.frombuffer(raw[raw.index(marker) + len(marker):], dtype="<f4").reshape(dims[::-1])


def postprocess_scan(engine, tallies: Iterable[np.ndarray | None], n_projections: int, air_total: np.ndarray | None, folder=None, crop_x: int = 1024,
                     sigma: Sequence[float] | None = (10, 10), pixel_size: Sequence[float] = (0.388, 0.388)):
    """postprocess_simulation (simulation.py:235-277): the three stacks and, with an air image, the normalised stack;
    written as projections_<mode>.mha / projections_total_normalized.mha when `folder` is given."""
    acc = ProjectionStacks(engine, n_projections, crop_x)
    for t in tallies:
        acc.add(t)
    out = {m: acc.stack(m) for m in MODES}
    if air_total is not None:
        out["total_normalized"] = acc.stack("total", air=air_total, sigma=sigma)
    if folder is not None:
        for name, a in out.items():
            write_mha(Path(folder) / f"projections_{name}.mha", a, pixel_size)
    return out
