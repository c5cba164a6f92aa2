"""CPU oracle of the projection post-processing -- TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy restatement of what cbctmc does after a simulation; each function cites the reference lines
it follows.  Pinned: tests/golden/post_reference.npz holds the outputs of the reference's OWN functions
(cbctmc/mc/projection.py imported from /root/reference with stub SimpleITK/ipmi modules by
tests/make_golden_post.py) on ASCII files written from the golden tallies; tests/test_post.py checks this
restatement against them bit for bit.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
leg may import this module; the product path (csrc/cuda/postprocess.cu behind include/mcgpu_b200.h) never does.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as ndi


def read_raw(path, n_detector_pixels, n_detector_pixels_half_fan=None) -> np.ndarray:
    """MCProjection._read_raw (cbctmc/mc/projection.py:36-51): text -> float64 -> float32 [Nz][Nx][4], rows flipped, x cropped."""
    data = np.loadtxt(path, dtype=np.float64).astype(np.float32)
    data = data.reshape(*n_detector_pixels[::-1], 4)
    data = np.flip(data, axis=0)
    if n_detector_pixels_half_fan:
        data = data[:, : n_detector_pixels_half_fan[0]]
    return data


def values_from_tally(tally: np.ndarray, norm: float) -> np.ndarray:
    """The same float32 array without the text file: "%.8lf" of NORM*count parsed back (report_image, MC-GPU_v1.3.cu:2860-2895)."""
    t = np.asarray(tally, dtype=np.uint64)
    text = np.array([float("%.8f" % (norm * float(c))) for c in t.reshape(-1)], dtype=np.float64).astype(np.float32)
    return np.moveaxis(text.reshape(t.shape), 0, -1)  # [4][Nz][Nx] -> [Nz][Nx][4]


def select_mode(stack: np.ndarray, mode: str) -> np.ndarray:
    """projections_to_itk (projection.py:143-149): sum over the 4 planes in float32 / plane 0 / planes 1..3."""
    if mode == "total":
        return stack.sum(axis=-1)
    if mode == "unscattered":
        return stack[..., 0]
    if mode == "scattered":
        return stack[..., 1:].sum(axis=-1)
    raise ValueError(mode)


def fill_zeros(stack: np.ndarray) -> np.ndarray:
    """projection.py:151-153: zeros become the smallest positive value of the whole stack."""
    min_non_zero = stack[stack > 0.0].min()
    return np.where(stack == 0, min_non_zero, stack)


def normalize(stack: np.ndarray, air: np.ndarray, sigma=None) -> np.ndarray:
    """normalize_projections (projection.py:101-122): Gaussian-filtered air image, log(air / p) in float32."""
    if sigma:
        air = ndi.gaussian_filter(air, sigma=sigma)
    return np.log(air / stack)


def projections_stack(stack4: np.ndarray, mode: str = "total", air: np.ndarray | None = None, sigma=None) -> np.ndarray:
    """projections_to_itk (projection.py:125-176) without the ITK wrapper: stack4 = [P][Nz][Nx'][4] float32."""
    s = fill_zeros(select_mode(stack4, mode))
    if air is not None and mode == "total":
        a = np.asarray(air)
        if a.ndim == 3:
            a = a.sum(-1)
        s = normalize(s, a, sigma)
    return s
