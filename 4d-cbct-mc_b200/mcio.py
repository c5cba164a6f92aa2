"""File-format boundary of the MC-GPU path: writers for the inputs cbctmc hands
to MC-GPU and a reader for the projections it gets back.

These mirror (behaviour, not code) the reference's L3 layer:
  * `.in`   : cbctmc/assets/templates/mcgpu_input.jinja2 rendered by
              cbctmc/mc/simulation.py:288-357 (mm -> cm, rounding to 6 digits,
              gpu id -1 when more than one GPU)
  * `.vox`  : cbctmc/mc/voxel_data.pyx:12-29 ("<mat> <rho:.6f>" per voxel, x
              fastest, one blank line per x-row, one more per z-slab) inside
              cbctmc/assets/templates/mcgpu_geometry.jinja2
  * reader  : cbctmc/mc/projection.py:36-51 (np.loadtxt of 4 columns)
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass, field
from pathlib import Path
from typing import Sequence

import numpy as np

REPO_ROOT = Path(__file__).resolve().parents[1]
ASSETS = REPO_ROOT / "assets"

# --------------------------------------------------------------------------- materials


def material_table() -> list[tuple[int, str, float, Path]]:
    """(number, identifier, nominal density, path) in MC-GPU material order
    (density-sorted, reference cbctmc/mc/materials.py:112-119)."""
    rows = []
    for line in (ASSETS / "materials" / "ORDER.txt").read_text().splitlines():
        number, ident, rho, _kind = line.split()
        rows.append((int(number), ident, float(rho), ASSETS / "materials" / f"{ident}__5_125kev.mcgpu.gz"))
    return rows


def material_numbers() -> dict[str, int]:
    return {ident: number for number, ident, _, _ in material_table()}


def material_densities() -> dict[str, float]:
    return {ident: rho for _, ident, rho, _ in material_table()}


def material_paths() -> list[Path]:
    return [p for _, _, _, p in material_table()]


DEFAULT_SPECTRUM = ASSETS / "spectra" / "125kVp_0.89mmTi_varian_norm.spc"


def write_truncated_spectrum(dst: Path, kvp: float, src: Path = DEFAULT_SPECTRUM) -> Path:
    """No 90 kVp spectrum ships with the reference: cut the
    125 kVp Varian spectrum at `kvp` keV and terminate it with a negative
    probability row, which is how MC-GPU detects the end (MC-GPU_v1.3.cu:3551)."""
    out = []
    for line in Path(src).read_text().splitlines():
        s = line.strip()
        if not s or s.startswith("#"):
            out.append(line)
            continue
        e, p = s.split()[:2]
        if float(e) >= kvp * 1e3 or float(p) < 0:
            break
        out.append(line)
    out.append(f"{kvp:.1f}e3 -1")
    Path(dst).write_text("\n".join(out) + "\n")
    return Path(dst)


# --------------------------------------------------------------------------- .in file


@dataclass
class ScanConfig:
    """Parameters of one MC-GPU invocation; lengths in mm like cbctmc
    (cbctmc/defaults.py:41-110), converted to cm when written."""

    n_histories: int = 10_000_000
    random_seed: int = 42
    gpu_id: int = 0
    threads_per_block: int = 128
    histories_per_thread: int = 150
    spectrum: Path = DEFAULT_SPECTRUM
    source_position: tuple[float, float, float] = (0.0, 0.0, 0.0)  # mm
    source_direction: tuple[float, float, float] = (0.0, 1.0, 0.0)
    polar_aperture: tuple[float, float] = (1.481720423651376, 13.441979314886868)
    azimuthal_aperture: float = -1
    n_detector_pixels: tuple[int, int] = (1848, 768)
    detector_size: tuple[float, float] = (717.024, 297.984)  # mm
    sdd: float = 1500.0  # mm
    sad: float = 1000.0  # mm
    lateral_displacement: float = -159.856  # mm (parsed but unused by the kernel, Q4)
    n_projections: int = 1
    angle_between_projections: float = 360.0 / 894
    projection_angles: Sequence[float] = field(default_factory=list)
    angular_roi: tuple[float, float] = (0.0, 5000.0)
    tally_material_dose: bool = False
    tally_voxel_dose: bool = False
    dose_roi: tuple[tuple[int, int], tuple[int, int], tuple[int, int]] = ((1, 1), (1, 1), (1, 1))
    materials: Sequence[Path] = field(default_factory=material_paths)


def default_source_position(volume_size_mm: Sequence[float], sad: float = 1000.0) -> tuple[float, float, float]:
    """cbctmc/mc/simulation.py:132-136: source on the -Y side of the volume centre."""
    return (volume_size_mm[0] / 2, volume_size_mm[1] / 2 - sad, volume_size_mm[2] / 2)


def render_input(cfg: ScanConfig, vox_path: Path, output_folder: Path) -> str:
    r6 = lambda v: round(v / 10.0, 6)  # noqa: E731  (mm -> cm as simulation.py:318-346)
    lines = [
        "# >>>> INPUT FILE FOR MC-GPU v1.3 >>>>",
        "",
        "#[SECTION SIMULATION CONFIG v.2009-05-12]",
        f"{cfg.n_histories}  # TOTAL NUMBER OF HISTORIES, OR SIMULATION TIME IN SECONDS IF VALUE < 100000",
        f"{cfg.random_seed}  # RANDOM SEED (ranecu PRNG)",
        f"{cfg.gpu_id}  # GPU NUMBER TO USE WHEN MPI IS NOT USED, OR TO BE AVOIDED IN MPI RUNS",
        f"{cfg.threads_per_block}  # GPU THREADS PER CUDA BLOCK (multiple of 32)",
        f"{cfg.histories_per_thread}  # SIMULATED HISTORIES PER GPU THREAD",
        "",
        "#[SECTION SOURCE v.2011-07-12]",
        f"{cfg.spectrum}  # X-RAY ENERGY SPECTRUM FILE",
        f"{r6(cfg.source_position[0])} {r6(cfg.source_position[1])} {r6(cfg.source_position[2])}  # SOURCE POSITION: X Y Z [cm]",
        f"{cfg.source_direction[0]} {cfg.source_direction[1]} {cfg.source_direction[2]}  # SOURCE DIRECTION COSINES: U V W",
        f"{cfg.polar_aperture[0]} {cfg.polar_aperture[1]} {cfg.azimuthal_aperture}  # POLAR (PHI 1, PHI 2) AND AZIMUTHAL (THETA) APERTURES [degrees]",
        "",
        "#[SECTION IMAGE DETECTOR v.2009-12-02]",
        f"{output_folder}/projection  # OUTPUT IMAGE FILE NAME",
        f"{cfg.n_detector_pixels[0]} {cfg.n_detector_pixels[1]}  # NUMBER OF PIXELS IN THE IMAGE: Nx Nz",
        f"{r6(cfg.detector_size[0])} {r6(cfg.detector_size[1])}  # IMAGE SIZE (width, height): Dx Dz [cm]",
        f"{r6(cfg.sdd)}  # SOURCE-TO-DETECTOR DISTANCE",
        f"{r6(cfg.lateral_displacement)}  # LATERAL DETECTOR DISPLACEMENT (along x axis [cm])",
        "",
        "#[SECTION ANGLES OF PROJ v.2023-09-06]",
        f"{'YES' if len(cfg.projection_angles) else 'NO'}  # DEFINE ANGLES SPECIFICALLY? [YES/NO]",
    ]
    for i, a in enumerate(cfg.projection_angles, start=1):
        lines.append(f"{a}  # PROJECTION ANGLE {i}")
    lines += [
        "",
        "#[SECTION CT SCAN TRAJECTORY v.2011-10-25]",
        f"{cfg.n_projections}  # NUMBER OF PROJECTIONS",
        f"{cfg.angle_between_projections}  # ANGLE BETWEEN PROJECTIONS [degrees]",
        f"{cfg.angular_roi[0]} {cfg.angular_roi[1]}  # ANGLES OF INTEREST",
        f"{r6(cfg.sad)}  # SOURCE-TO-ROTATION AXIS DISTANCE",
        "0.0  # VERTICAL TRANSLATION BETWEEN PROJECTIONS (HELICAL SCAN)",
        "",
        "#[SECTION DOSE DEPOSITION v.2012-12-12]",
        f"{'YES' if cfg.tally_material_dose else 'NO'}  # TALLY MATERIAL DOSE? [YES/NO]",
        f"{'YES' if cfg.tally_voxel_dose else 'NO'}  # TALLY 3D VOXEL DOSE? [YES/NO]",
        f"{output_folder}/dose.dat  # OUTPUT VOXEL DOSE FILE NAME",
        f"{cfg.dose_roi[0][0]} {cfg.dose_roi[0][1]}  # Dose ROI X",
        f"{cfg.dose_roi[1][0]} {cfg.dose_roi[1][1]}  # Dose ROI Y",
        f"{cfg.dose_roi[2][0]} {cfg.dose_roi[2][1]}  # Dose ROI Z",
        "",
        "#[SECTION VOXELIZED GEOMETRY FILE v.2009-11-30]",
        f"{vox_path}  # VOXELIZED GEOMETRY FILE",
        "",
        "#[SECTION MATERIAL FILE LIST v.2009-11-30]",
    ]
    for i, m in enumerate(cfg.materials, start=1):
        lines.append(f"{m}  # MATERIAL FILE {i}")
    lines += ["", "# >>>> END INPUT FILE >>>>", ""]
    return "\n".join(lines)


def write_input(cfg: ScanConfig, vox_path: Path, output_folder: Path, in_path: Path) -> Path:
    Path(output_folder).mkdir(parents=True, exist_ok=True)
    Path(in_path).write_text(render_input(cfg, vox_path, output_folder))
    return Path(in_path)


# --------------------------------------------------------------------------- .vox file

_VOX_HEADER = """# voxel geometry written by 4d-cbct-mc_b200 (penEasy 2008 format)
[SECTION VOXELS HEADER v.2008-04-13]
{nx} {ny} {nz}  # SIZE IN X, Y, Z
{dx} {dy} {dz}  # VOXEL SPACING IN X, Y, Z
1  # COLUMN NUMBER WHERE MATERIAL ID IS LOCATED
2  # COLUMN NUMBER WHERE MASS DENSITY IS LOCATED
1  # BLANK LINES AT END OF X,Y-CYCLES (1=YES, 0=NO)
[END OF VXH SECTION]
# >>>> DATA BEGINS >>>>
"""


def write_vox(path: Path, materials: np.ndarray, densities: np.ndarray, spacing_cm: Sequence[float]) -> Path:
    """`materials`/`densities` are indexed [x, y, z] in MC-GPU's frame (i.e. already
    rotated the way cbctmc/mc/geometry.py:589-599 does); x runs fastest in the file."""
    assert materials.shape == densities.shape and materials.ndim == 3
    nx, ny, nz = materials.shape
    mat = np.ascontiguousarray(materials.transpose(2, 1, 0)).reshape(-1)  # z, y, x -> x fastest
    rho = np.ascontiguousarray(densities.transpose(2, 1, 0)).reshape(-1).astype(np.float32)
    # one text line per distinct (material, density) pair, gathered through a palette
    key = mat.astype(np.uint64) << np.uint64(32) | rho.view(np.uint32).astype(np.uint64)
    uniq, inv = np.unique(key, return_inverse=True)
    palette = []
    for k in uniq:
        m = int(k >> np.uint64(32))
        d = np.array([int(k & np.uint64(0xFFFFFFFF))], dtype=np.uint32).view(np.float32)[0]
        palette.append(f"{m} {float(d):.6f}\n".encode())
    width = max(len(p) for p in palette)
    table = np.zeros((len(palette), width), dtype=np.uint8)
    lens = np.zeros(len(palette), dtype=np.int64)
    for i, p in enumerate(palette):
        table[i, : len(p)] = np.frombuffer(p, dtype=np.uint8)
        lens[i] = len(p)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(_VOX_HEADER.format(nx=nx, ny=ny, nz=nz, dx=spacing_cm[0], dy=spacing_cm[1], dz=spacing_cm[2]).encode())
        inv = inv.reshape(nz, ny, nx)
        for k in range(nz):
            rows = table[inv[k]]  # (ny, nx, width)
            keep = np.arange(width)[None, None, :] < lens[inv[k]][..., None]
            for j in range(ny):
                f.write(rows[j][keep[j]].tobytes())
                f.write(b"\n")
            f.write(b"\n")
    return Path(path)


# --------------------------------------------------------------------------- projections


def projection_filename(base: str, angle_deg: float) -> str:
    """MC-GPU_v1.3.cu:2803 -- '<base>_%010.6fdeg' with the (float) sequential angle."""
    return "%s_%010.6fdeg" % (base, float(np.float32(angle_deg)))


def read_projection(path: Path, n_pixels: tuple[int, int]) -> np.ndarray:
    """ASCII projection -> float64 array [4, Nz, Nx] (scatter plane, detector row, column)."""
    data = np.loadtxt(path, dtype=np.float64)
    nx, nz = n_pixels
    return data.reshape(nz, nx, 4).transpose(2, 0, 1).copy()


def write_voxb(path: Path, materials: np.ndarray, densities: np.ndarray, spacing_cm, compress: bool = False) -> Path:
    """Binary geometry the engine reads without tokenising text (csrc/host/voxels.c: read_voxels_binary).  Same
    arguments as write_vox: arrays indexed [x, y, z], materials 1-based; stored x fastest like the text file."""
    import struct

    assert materials.shape == densities.shape and materials.ndim == 3
    nx, ny, nz = materials.shape
    m = np.ascontiguousarray(materials.transpose(2, 1, 0), dtype=np.uint8)
    d = np.ascontiguousarray(densities.transpose(2, 1, 0), dtype="<f4")
    head = b"MCGPUVXB" + struct.pack("<4I3f", 1, nx, ny, nz, *[float(s) for s in spacing_cm])
    opener = gzip.open if compress else open
    with opener(path, "wb") as f:
        f.write(head)
        f.write(m.tobytes())
        f.write(d.tobytes())
    return Path(path)


def read_projection_raw(path: Path, n_pixels: tuple[int, int]) -> np.ndarray:
    """Binary side-file (mcgpu_write_projection_raw): float32 [4, Nz, Nx], same values as the ASCII columns."""
    nx, nz = n_pixels
    return np.fromfile(path, dtype="<f4").reshape(4, nz, nx)


def projection_counts(values: np.ndarray, n_pixels: tuple[int, int], detector_size_cm: tuple[float, float],
                      total_histories: int) -> np.ndarray:
    """Invert report_image's normalisation (MC-GPU_v1.3.cu:2860-2879) to recover the
    u64 tallies.  Only exact while one count (= NORM) is well above the 1e-8 print
    resolution, i.e. for modest history counts."""
    inv_x = np.float32(n_pixels[0]) / np.float32(detector_size_cm[0])
    inv_z = np.float32(n_pixels[1]) / np.float32(detector_size_cm[1])
    norm = (1.0 / 100.0) * float(inv_x) * float(inv_z) / float(total_histories)  # python floats: NEP-50 would keep float32
    if norm < 4e-8:
        raise ValueError(f"NORM={norm:g} too close to the 1e-8 print resolution to recover integer tallies")
    return np.rint(values / norm).astype(np.uint64)


def launched_histories(n_histories: int, threads_per_block: int, histories_per_thread: int) -> tuple[int, int, int]:
    """Grid rule of MC-GPU_v1.3.cu:823-841 -> (blocks, histories_per_thread, launched)."""
    total_threads = int(float(n_histories) / float(histories_per_thread) + 0.9990)
    blocks = int(float(total_threads) / float(threads_per_block) + 0.9990)
    if blocks > 65535:
        blocks = 65000
        histories_per_thread = int(float(n_histories) / float(blocks * threads_per_block) + 0.9990)
    elif blocks < 1:
        blocks = 1
    return blocks, histories_per_thread, blocks * threads_per_block * histories_per_thread
