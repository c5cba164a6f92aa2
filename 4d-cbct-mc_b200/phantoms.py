"""Synthetic CT-like phantoms of the shapes BASELINE.json names.

Arrays are indexed [x, y, z] in MC-GPU's frame (x fastest in the .vox file) and
hold the MC-GPU material number (1-based, density-sorted order) and the mass
density in g/cm^3.  cbctmc assigns one density per material
(cbctmc/mc/geometry.py:72-74), which is what these generators do too, except
for the lungs of the thorax phantom which carry a density gradient.

Phantom definitions restate the *parameters* of the reference geometries:
  water cylinder : cbctmc/mc/geometry.py:1106-1165 (MCWaterPhantomGeometry)
  Catphan604     : cbctmc/mc/geometry.py:902-1068 (body, sensitometry rods, air rods)
  line pairs     : cbctmc/mc/geometry.py:797-862 (x upsampled x4, Al bars in a CIRS-like body)
  air scan       : cbctmc/mc/geometry.py:626-639 (one 200 cm voxel of air)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mcio import material_densities, material_numbers


@dataclass
class Phantom:
    name: str
    materials: np.ndarray  # uint8 [x, y, z]
    densities: np.ndarray  # float32 [x, y, z]
    spacing_cm: tuple[float, float, float]

    @property
    def shape(self):
        return self.materials.shape

    @property
    def size_mm(self):
        return tuple(10.0 * n * s for n, s in zip(self.shape, self.spacing_cm))


def _blank(shape, fill="air"):
    num, rho = material_numbers(), material_densities()
    mats = np.full(shape, num[fill], dtype=np.uint8)
    dens = np.full(shape, rho[fill], dtype=np.float32)
    return mats, dens


def _paint(mats, dens, mask, ident, density=None):
    num, rho = material_numbers(), material_densities()
    mats[mask] = num[ident]
    dens[mask] = rho[ident] if density is None else density


def _grid(shape):
    return np.meshgrid(*(np.arange(n, dtype=np.float32) for n in shape), indexing="ij", sparse=True)


def _cyl(shape, center, radius, height):
    x, y, z = _grid(shape)
    return ((x - center[0]) ** 2 + (y - center[1]) ** 2 <= radius**2) & (z >= center[2] - height / 2) & (z < center[2] + height / 2)


def air_scan() -> Phantom:
    mats, dens = _blank((1, 1, 1))
    return Phantom("air_1x1x1_200cm", mats, dens, (200.0, 200.0, 200.0))


def water_cylinder(n: int = 500, spacing_mm: float = 1.0, radius_mm: float = 100.0, height_mm: float = 150.0) -> Phantom:
    shape = (n, n, n)
    mats, dens = _blank(shape)
    c = np.array(shape) / 2
    _paint(mats, dens, _cyl(shape, c, radius_mm / spacing_mm, height_mm / spacing_mm), "h2o")
    return Phantom(f"water_cylinder_{n}", mats, dens, (spacing_mm / 10,) * 3)


_CATPHAN_SENSITOMETRY = [  # (material, angle deg, distance mm, radius mm, length mm)
    ("air", 90, 58.7, 6.5, 24.0), ("teflon", 60, 58.7, 6.5, 24.0), ("delrin", 0, 58.7, 6.5, 24.0),
    ("bone_020", 330, 58.7, 6.5, 24.0), ("acrylic", 300, 58.7, 6.5, 24.0), ("air", 270, 58.7, 6.5, 24.0),
    ("polystyrene", 240, 58.7, 6.5, 24.0), ("ldpe", 180, 58.7, 6.5, 24.0), ("bone_050", 150, 58.7, 6.5, 24.0),
    ("pmp", 120, 58.7, 6.5, 24.0), ("h2o", 0, 0.0, 30.0, 40.0),
]
_CATPHAN_AIR_RODS = [("air", a, 35.355, 1.5, 24.0) for a in (135, 45, 315, 225)]


def catphan604(n: int = 500, spacing_mm: float = 1.0) -> Phantom:
    shape = (n, n, n)
    mats, dens = _blank(shape)
    c = np.array(shape, dtype=np.float64) / 2
    rois = [("h2o", 0.0, 0.0, 100.0, 100.0)] + _CATPHAN_SENSITOMETRY + _CATPHAN_AIR_RODS
    for ident, angle, dist, radius, length in rois:
        phi = np.deg2rad(angle)
        centre = c + np.array([np.cos(phi), -np.sin(phi), 0.0]) * dist / spacing_mm
        _paint(mats, dens, _cyl(shape, centre, radius / spacing_mm, length / spacing_mm), ident)
    return Phantom(f"catphan604_{n}", mats, dens, (spacing_mm / 10,) * 3)


def thorax(shape=(256, 256, 100), spacing_mm: float = 2.0, diaphragm_shift_mm: float = 0.0) -> Phantom:
    """Patient-like thorax: elliptical soft-tissue body 34x24 cm with an adipose rim,
    two lungs with a density gradient 0.10-0.26, bone_050 spine with a bone_100 shell,
    ribs of bone_100.  `diaphragm_shift_mm` moves the lung base (the 4D variants)."""
    mats, dens = _blank(shape)
    x, y, z = _grid(shape)
    cx, cy = shape[0] / 2, shape[1] / 2
    ax, ay = 170.0 / spacing_mm, 120.0 / spacing_mm
    body = ((x - cx) / ax) ** 2 + ((y - cy) / ay) ** 2 <= 1.0
    inner = ((x - cx) / (ax - 10 / spacing_mm)) ** 2 + ((y - cy) / (ay - 10 / spacing_mm)) ** 2 <= 1.0
    allz = z >= 0
    _paint(mats, dens, body & allz, "adipose")
    _paint(mats, dens, inner & allz, "soft_tissue")
    # ribs: an elliptical shell of bone, present in alternating z bands
    rib_out = ((x - cx) / (ax - 14 / spacing_mm)) ** 2 + ((y - cy) / (ay - 14 / spacing_mm)) ** 2 <= 1.0
    rib_in = ((x - cx) / (ax - 22 / spacing_mm)) ** 2 + ((y - cy) / (ay - 22 / spacing_mm)) ** 2 <= 1.0
    band = (np.floor(z * spacing_mm / 12.0) % 2) == 0
    _paint(mats, dens, rib_out & ~rib_in & band, "bone_100")
    # lungs
    z_base = (30.0 + diaphragm_shift_mm) / spacing_mm
    for sgn in (-1, 1):
        lx = cx + sgn * 75.0 / spacing_mm
        lung = (((x - lx) / (55.0 / spacing_mm)) ** 2 + ((y - cy) / (80.0 / spacing_mm)) ** 2 <= 1.0) & (z >= z_base)
        grad = (0.10 + 0.16 * np.clip((y - (cy - 80.0 / spacing_mm)) / (160.0 / spacing_mm), 0, 1)).astype(np.float32)
        grad = np.round(grad * 50) / 50  # 0.02 g/cm3 steps keep the (material, density) palette small
        full = np.broadcast_to(lung, shape)
        mats[full] = material_numbers()["lung"]
        dens[full] = np.broadcast_to(grad, shape)[full]
    # spine
    sy = cy + 70.0 / spacing_mm
    _paint(mats, dens, _cyl(shape, (cx, sy, shape[2] / 2), 18.0 / spacing_mm, shape[2] * 2), "bone_100")
    _paint(mats, dens, _cyl(shape, (cx, sy, shape[2] / 2), 14.0 / spacing_mm, shape[2] * 2), "bone_050")
    return Phantom(f"thorax_{shape[0]}x{shape[1]}x{shape[2]}_shift{diaphragm_shift_mm:g}", mats, dens, (spacing_mm / 10,) * 3)


def line_pairs(shape=(305, 300, 152), spacing_mm=(1.0, 1.0, 1.0), upsample_x: int = 4) -> Phantom:
    """CIRS-like body with an aluminium line-pair insert, x upsampled x4 -> 0.25 mm
    (1220x300x152 voxels): the geometry that exceeds L2 in the reference layout."""
    fine = (shape[0] * upsample_x, shape[1], shape[2])
    mats, dens = _blank(fine)
    x, y, z = _grid(fine)
    sx = spacing_mm[0] / upsample_x
    cx, cy = fine[0] / 2, fine[1] / 2
    body = ((x - cx) * sx / 150.0) ** 2 + ((y - cy) * spacing_mm[1] / 100.0) ** 2 <= 1.0
    allz = z >= 0
    _paint(mats, dens, body & allz, "soft_tissue")
    for sgn in (-1, 1):
        lung = (((x - cx - sgn * 70.0 / sx) * sx / 45.0) ** 2 + ((y - cy) * spacing_mm[1] / 65.0) ** 2 <= 1.0)
        _paint(mats, dens, lung & allz, "h2o", density=0.207)
    _paint(mats, dens, (((x - cx) * sx) ** 2 + ((y - cy - 70.0) * spacing_mm[1]) ** 2 <= 15.0**2) & allz, "bone_050")
    # line-pair bars along x inside a water insert at the centre
    insert = (np.abs((x - cx) * sx) <= 30.0) & (np.abs((y - cy) * spacing_mm[1]) <= 15.0) & (np.abs(z - fine[2] / 2) * spacing_mm[2] <= 20.0)
    _paint(mats, dens, insert, "h2o")
    x_mm = (np.arange(fine[0]) - cx) * sx
    bars = np.zeros(fine[0], dtype=bool)
    pos = -28.0
    for gap in (4.0, 3.0, 2.0, 1.5, 1.0, 0.75, 0.5):
        for _ in range(3):
            bars |= (x_mm >= pos) & (x_mm < pos + gap)
            pos += 2 * gap
        pos += 2.0
    _paint(mats, dens, insert & bars[:, None, None], "aluminium")
    return Phantom(f"line_pairs_{fine[0]}x{fine[1]}x{fine[2]}", mats, dens, (sx / 10, spacing_mm[1] / 10, spacing_mm[2] / 10))


def patient(shape=(256, 256, 100), spacing_mm: float = 2.0) -> Phantom:
    """Patient-like trunk in cbctmc's PATIENT material set (cbctmc/mc/geometry.py:130-229: bone mapper with red
    marrow, lung vessels = blood, liver, stomach/intestines, muscle, fat) -- the tissues whose Compton profiles have the
    most shells (blood 40 = MAX_SHELLS, red marrow 36): adipose rim, muscle layer, soft-tissue interior, lungs with
    blood vessels, liver and stomach below the diaphragm, vertebra and ribs of bone_020/050/100 around red marrow,
    cartilage discs and a gland."""
    mats, dens = _blank(shape)
    x, y, z = _grid(shape)
    s = spacing_mm
    cx, cy, nz = shape[0] / 2, shape[1] / 2, shape[2]
    allz = z >= 0

    def ell(ax_mm, ay_mm, ox_mm=0.0, oy_mm=0.0):
        return ((x - cx - ox_mm / s) / (ax_mm / s)) ** 2 + ((y - cy - oy_mm / s) / (ay_mm / s)) ** 2 <= 1.0

    _paint(mats, dens, ell(170, 120) & allz, "adipose")
    _paint(mats, dens, ell(158, 108) & allz, "muscle_tissue")
    _paint(mats, dens, ell(146, 96) & allz, "soft_tissue")
    z_dia = 0.45 * nz  # diaphragm: abdomen organs below, lungs above
    upper, lower = z >= z_dia, z < z_dia
    for sgn in (-1, 1):
        lung = ell(52, 74, sgn * 74.0, -6.0) & upper
        _paint(mats, dens, lung, "lung", density=0.26)
        # vessel tree: blood cylinders along z and a few oblique branches
        for k, (ox, oy, r) in enumerate([(60, -20, 5.0), (85, 10, 3.5), (70, 25, 3.0), (95, -30, 2.5)]):
            tilt = 0.15 * (k - 1.5)
            vessel = ((x - cx - sgn * ox / s - tilt * (z - z_dia)) ** 2 + (y - cy - oy / s) ** 2 <= (r / s) ** 2) & upper
            _paint(mats, dens, vessel & lung, "blood")
    _paint(mats, dens, ell(28, 32, 10.0, -10.0) & upper, "blood")  # heart / great vessels
    _paint(mats, dens, ell(70, 60, -55.0, -10.0) & lower, "liver")
    _paint(mats, dens, ell(45, 40, 65.0, -15.0) & lower, "stomach_intestines")
    _paint(mats, dens, ell(12, 10, 60.0, 40.0) & lower, "glands_others")
    # ribs: shell of cortical bone around red marrow, alternating z bands
    band = (np.floor(z * s / 12.0) % 2) == 0
    _paint(mats, dens, ell(144, 94) & ~ell(132, 82) & band, "bone_100")
    _paint(mats, dens, ell(140, 90) & ~ell(136, 86) & band, "red_marrow")
    # vertebral column: bone_100 shell, bone_050 / bone_020 spongiosa, red marrow core, cartilage discs
    disc = (np.floor(z * s / 30.0) % 2) == 1
    sp = lambda r: ((x - cx) ** 2 + (y - cy - 62.0 / s) ** 2 <= (r / s) ** 2) & allz  # noqa: E731
    _paint(mats, dens, sp(20), "bone_100")
    _paint(mats, dens, sp(17), "bone_050")
    _paint(mats, dens, sp(13), "bone_020")
    _paint(mats, dens, sp(8), "red_marrow")
    _paint(mats, dens, sp(20) & disc & (np.floor(z * s / 6.0) % 5 == 0), "cartilage")
    return Phantom(f"patient_{shape[0]}x{shape[1]}x{shape[2]}", mats, dens, (spacing_mm / 10,) * 3)
