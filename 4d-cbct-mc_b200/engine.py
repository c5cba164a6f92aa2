"""ctypes binding of libmcgpu_b200.so (include/mcgpu_b200.h) and the host-side mirror of the
reference's call into MC-GPU.

The reference drives this path by shelling out to `MC-GPU_v1.3.x input.in` inside Docker
(cbctmc/mc/simulation.py:176-233, cbctmc/docker.py:31-65) and parsing the ASCII projections
back.  `run_mcgpu()` below is that same operation in-process; `Engine` exposes the stages of the
C ABI for callers that want the integer tallies without the text round trip.

There is no fallback of any kind: a missing library raises at import, and every run call raises
`McgpuError` when no B200 is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

_LIB_PATH = Path(os.environ.get("MCGPU_B200_LIB") or Path(__file__).resolve().parent / "lib" / "libmcgpu_b200.so")  # env: developer A/B builds only
if not _LIB_PATH.exists():
    raise ImportError(f"{_LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); there is no CPU fallback")
_lib = C.CDLL(str(_LIB_PATH))


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "num_projections", "num_pixels_x", "num_pixels_z", "num_voxels_x", "num_voxels_y", "num_voxels_z",
        "num_materials_used", "num_energy_values", "num_spectrum_bins", "threads_per_block", "histories_per_thread",
        "num_blocks", "seed_input", "enable_specific_angles", "num_devices", "voxel_bits", "palette_size")] + [
        ("requested_histories", C.c_ulonglong), ("launched_histories", C.c_ulonglong),
        ("mean_energy_spectrum", C.c_float), ("e0", C.c_float), ("ide", C.c_float), ("fast_math", C.c_int)]


PROGRESS_CB = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_double, C.c_void_p)

_SIGS = {
    "mcgpu_create": (C.c_void_p, [C.POINTER(C.c_int), C.c_int]),
    "mcgpu_destroy": (None, [C.c_void_p]),
    "mcgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "mcgpu_set_verbose": (None, [C.c_void_p, C.c_int]),
    "mcgpu_load_input": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mcgpu_load_voxels": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mcgpu_set_voxels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "mcgpu_load_materials": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_int]),
    "mcgpu_set_histories": (C.c_int, [C.c_void_p, C.c_ulonglong]),
    "mcgpu_set_seed": (C.c_int, [C.c_void_p, C.c_int]),
    "mcgpu_set_fast_math": (C.c_int, [C.c_void_p, C.c_int]),
    "mcgpu_run_projection": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mcgpu_run_streams": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p]),
    "mcgpu_device_image": (C.c_void_p, [C.c_void_p]),
    "mcgpu_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "mcgpu_last_reduce_ms": (C.c_double, [C.c_void_p]),
    "mcgpu_reduce_kind": (C.c_char_p, [C.c_void_p]),
    "mcgpu_get_scan_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "mcgpu_device_selftest": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_ulonglong)]),
    "mcgpu_run_all": (C.c_int, [C.c_void_p, PROGRESS_CB, C.c_void_p]),
    "mcgpu_write_projection_ascii": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_double]),
    "mcgpu_write_projection_raw": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "mcgpu_post_intensity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcgpu_post_gaussian": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]),
    "mcgpu_gaussian_weights": (C.c_int, [C.c_double, C.c_void_p, C.c_int]),
    "mcgpu_post_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_float]),
    "mcgpu_projection_filename": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_size_t]),
    "mcgpu_reset_dose": (C.c_int, [C.c_void_p]),
    "mcgpu_get_dose": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "mcgpu_write_dose_reports": (C.c_int, [C.c_void_p, C.c_double, C.c_int]),
    "mcgpu_get_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "mcgpu_projection_seed": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "mcgpu_copy_table": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "mcgpu_ranecu_init_stream": (None, [C.c_longlong, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mcgpu_ranecu_next": (C.c_float, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mcgpu_ranecu_advance_projection_seed": (C.c_int, [C.c_int, C.c_ulonglong]),
    "mcgpu_grid_rule": (None, [C.c_ulonglong, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(_lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype, _fn.argtypes = _res, _args

EXPORTED_SYMBOLS = tuple(_SIGS)

_TABLE_DTYPES = {
    "woodcock": np.float32, "mfp_a": np.float32, "mfp_b": np.float32, "rayleigh_xco": np.float32, "rayleigh_pco": np.float32,
    "rayleigh_aco": np.float32, "rayleigh_bco": np.float32, "rayleigh_itlco": np.uint8, "rayleigh_ituco": np.uint8,
    "rayleigh_pmax": np.float32, "compton_fco": np.float32, "compton_uico": np.float32, "compton_fj0": np.float32,
    "compton_noscco": np.int32, "density_nominal": np.float32, "density_max": np.float32, "espc": np.float32,
    "espc_cutoff": np.float32, "espc_alias": np.int16, "views": np.float32, "voxel_material": np.uint8,
    "voxel_density": np.float32, "voxel_packed": np.uint8,
}

VIEW_WORDS = 45  # sizeof(mcgpu_view)/4, see csrc/host/mcgpu_host.h


class McgpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"mcgpu error {code}: {message}")
        self.code = code


class Engine:
    """One MC-GPU simulation context (mcgpu_ctx).  Stage order: load_input -> load_voxels /
    set_voxels -> load_materials -> run_*."""

    def __init__(self, device_ids: Sequence[int] | None = None, verbose: bool = False):
        if device_ids is None:
            self._h = _lib.mcgpu_create(None, 0)
        else:
            arr = (C.c_int * len(device_ids))(*device_ids)
            self._h = _lib.mcgpu_create(arr, len(device_ids))
        if not self._h:
            raise MemoryError("mcgpu_create failed")
        _lib.mcgpu_set_verbose(self._h, int(verbose))

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            _lib.mcgpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc < 0:
            raise McgpuError(rc, (_lib.mcgpu_last_error(self._h) or b"").decode(errors="replace"))
        return rc

    # -- stages
    def load_input(self, in_path):
        self._check(_lib.mcgpu_load_input(self._h, str(in_path).encode()))
        return self

    def load_voxels(self, vox_path=None):
        self._check(_lib.mcgpu_load_voxels(self._h, None if vox_path is None else str(vox_path).encode()))
        return self

    def set_voxels(self, materials: np.ndarray, densities: np.ndarray, spacing_cm: Sequence[float]):
        """materials/densities indexed [x, y, z] (MC-GPU frame)."""
        nx, ny, nz = materials.shape
        m = np.ascontiguousarray(materials.transpose(2, 1, 0), dtype=np.uint8)
        r = np.ascontiguousarray(densities.transpose(2, 1, 0), dtype=np.float32)
        self._check(_lib.mcgpu_set_voxels(self._h, nx, ny, nz, *[np.float32(s) for s in spacing_cm], m.ctypes.data, r.ctypes.data))
        return self

    def load_materials(self, paths: Sequence | None = None):
        if paths is None:
            self._check(_lib.mcgpu_load_materials(self._h, None, 0))
        else:
            arr = (C.c_char_p * len(paths))(*[str(p).encode() for p in paths])
            self._check(_lib.mcgpu_load_materials(self._h, arr, len(paths)))
        return self

    def set_histories(self, n: int):
        self._check(_lib.mcgpu_set_histories(self._h, n))

    def set_seed(self, seed: int):
        self._check(_lib.mcgpu_set_seed(self._h, seed))

    def set_fast_math(self, on: bool):
        """False (default): bit-exact arithmetic.  True: the reference's shipped -use_fast_math flags."""
        self._check(_lib.mcgpu_set_fast_math(self._h, int(on)))

    # -- info
    @property
    def info(self) -> Info:
        out = Info()
        self._check(_lib.mcgpu_get_info(self._h, C.byref(out)))
        return out

    def table(self, name: str) -> np.ndarray:
        n = self._check(_lib.mcgpu_copy_table(self._h, name.encode(), None, 0))
        buf = np.empty(n, dtype=np.uint8)
        self._check(_lib.mcgpu_copy_table(self._h, name.encode(), buf.ctypes.data, n))
        return buf.view(_TABLE_DTYPES[name])

    def views(self) -> np.ndarray:
        return self.table("views").reshape(-1, VIEW_WORDS)

    def projection_seed(self, p: int) -> int:
        s = C.c_int()
        self._check(_lib.mcgpu_projection_seed(self._h, p, C.byref(s)))
        return s.value

    def projection_filename(self, p: int) -> str:
        buf = C.create_string_buffer(512)
        self._check(_lib.mcgpu_projection_filename(self._h, p, buf, 512))
        return buf.value.decode()

    # -- simulation
    def new_image(self, pinned: bool = False) -> np.ndarray:
        i = self.info
        return np.zeros((4, i.num_pixels_z, i.num_pixels_x), dtype=np.uint64)

    def run_projection(self, p: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        img = self.new_image() if out is None else out
        self._check(_lib.mcgpu_run_projection(self._h, p, img.ctypes.data))
        return img

    def run_streams(self, p: int, begin: int, end: int, out: np.ndarray | None = None, fetch: bool = True):
        img = (self.new_image() if out is None else out) if fetch else None
        self._check(_lib.mcgpu_run_streams(self._h, p, begin, end, img.ctypes.data if fetch else None))
        return img

    @property
    def last_kernel_ms(self) -> float:
        return float(_lib.mcgpu_last_kernel_ms(self._h))

    @property
    def last_reduce_ms(self) -> float:
        """History-split runs: device time of the reduction of the partial images (ncclReduce / peer kernel)."""
        return float(_lib.mcgpu_last_reduce_ms(self._h))

    @property
    def reduce_kind(self) -> str:
        return (_lib.mcgpu_reduce_kind(self._h) or b"none").decode()

    def selftest(self, name: str) -> int:
        """Mismatches of a device self test (include/mcgpu_b200.h: mcgpu_device_selftest); 0 = the shortcut is exact."""
        n = C.c_ulonglong(1)
        self._check(_lib.mcgpu_device_selftest(self._h, name.encode(), C.byref(n)))
        return int(n.value)

    def scan_stats(self) -> dict:
        out = (C.c_double * 6)()
        _lib.mcgpu_get_scan_stats(self._h, out, 6)
        return dict(zip(("wall_s", "kernel_s", "wait_s", "report_s", "projections", "devices"), [float(v) for v in out]))

    @property
    def device_image_ptr(self) -> int:
        return int(_lib.mcgpu_device_image(self._h) or 0)

    def run_all(self, progress: Callable[[int, int, float], None] | None = None):
        cb = PROGRESS_CB((lambda p, n, s, u: progress(p, n, s)) if progress else (lambda p, n, s, u: None))
        self._check(_lib.mcgpu_run_all(self._h, cb, None))

    def write_projection_raw(self, p: int, image: np.ndarray):
        img = np.ascontiguousarray(image, dtype=np.uint64)
        self._check(_lib.mcgpu_write_projection_raw(self._h, p, img.ctypes.data))

    # ---- projection post-processing on the device (include/mcgpu_b200.h)
    def post_intensity(self, tally: np.ndarray | None = None, launched: int | None = None, crop_x: int = 0):
        """(total, unscattered, scattered, min_positive[3]) float32 [Nz][crop_x]; tally None = the device's last projection."""
        info = self.info
        crop = crop_x if 0 < crop_x <= info.num_pixels_x else info.num_pixels_x
        outs = [np.empty((info.num_pixels_z, crop), dtype=np.float32) for _ in range(3)]
        mins = np.empty(3, dtype=np.float32)
        t = None if tally is None else np.ascontiguousarray(tally, dtype=np.uint64)
        self._check(_lib.mcgpu_post_intensity(self._h, None if t is None else t.ctypes.data, int(launched or info.launched_histories), crop,
                                              outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, mins.ctypes.data))
        return outs[0], outs[1], outs[2], mins

    def post_gaussian(self, image: np.ndarray, sigma) -> np.ndarray:
        a = np.ascontiguousarray(image, dtype=np.float32)
        out = np.empty_like(a)
        self._check(_lib.mcgpu_post_gaussian(self._h, a.ctypes.data, a.shape[0], a.shape[1], float(sigma[0]), float(sigma[1]), out.ctypes.data))
        return out

    @staticmethod
    def gaussian_weights(sigma: float) -> np.ndarray:
        r = _lib.mcgpu_gaussian_weights(float(sigma), None, 0)
        if r < 0:
            raise McgpuError(r, "gaussian_weights: sigma must be positive")
        w = np.empty(r + 1, dtype=np.float64)
        _lib.mcgpu_gaussian_weights(float(sigma), w.ctypes.data, r + 1)
        return w

    def post_normalize(self, air: np.ndarray, stack: np.ndarray, min_nonzero: float) -> np.ndarray:
        """In place on a C-contiguous float32 stack [P][Nz][Nx]."""
        a = np.ascontiguousarray(air, dtype=np.float32)
        assert stack.dtype == np.float32 and stack.flags.c_contiguous and stack.shape[1:] == a.shape
        self._check(_lib.mcgpu_post_normalize(self._h, a.ctypes.data, stack.ctypes.data, stack.shape[0], a.shape[0], a.shape[1], float(min_nonzero)))
        return stack

    def reset_dose(self):
        self._check(_lib.mcgpu_reset_dose(self._h))

    def dose(self, which: str) -> np.ndarray:
        """uint64 [n, 2] dose counters ("materials": n = 25, "voxels": n = ROI voxels); empty when off."""
        n = self._check(_lib.mcgpu_get_dose(self._h, which.encode(), None, 0))
        out = np.zeros(n, dtype=np.uint64)
        if n:
            self._check(_lib.mcgpu_get_dose(self._h, which.encode(), out.ctypes.data, n))
        return out.reshape(-1, 2)

    def write_dose_reports(self, seconds: float = 0.0, projections: int = 0):
        self._check(_lib.mcgpu_write_dose_reports(self._h, seconds, projections))

    def write_projection(self, p: int, image: np.ndarray, seconds: float = 0.0):
        img = np.ascontiguousarray(image, dtype=np.uint64)
        self._check(_lib.mcgpu_write_projection_ascii(self._h, p, img.ctypes.data, seconds))


def run_mcgpu(input_filepath, device_ids: Sequence[int] | None = None, progress=None) -> Info:
    """In-process equivalent of `MC-GPU_v1.3.x <input.in>` (what cbctmc's
    MCSimulation._run_simulation executes in Docker): parse, load, simulate every projection,
    write the `<base>_<angle>deg` ASCII files.  Raises McgpuError on any failure."""
    with Engine(device_ids) as eng:
        eng.load_input(input_filepath).load_voxels().load_materials()
        eng.run_all(progress)
        return eng.info


# -- RANECU helpers (known-answer tests)
def ranecu_init_stream(stream: int, hpt: int, seed: int) -> tuple[int, int]:
    a, b = C.c_int(), C.c_int()
    _lib.mcgpu_ranecu_init_stream(stream, hpt, seed, C.byref(a), C.byref(b))
    return a.value, b.value


def ranecu_sequence(s1: int, s2: int, n: int) -> np.ndarray:
    a, b = C.c_int(s1), C.c_int(s2)
    return np.array([_lib.mcgpu_ranecu_next(C.byref(a), C.byref(b)) for _ in range(n)], dtype=np.float32)


def advance_projection_seed(seed: int, total_histories: int) -> int:
    return int(_lib.mcgpu_ranecu_advance_projection_seed(seed, total_histories))


def grid_rule(requested: int, tpb: int, hpt: int) -> tuple[int, int, int]:
    h, b, n = C.c_int(hpt), C.c_int(), C.c_ulonglong()
    _lib.mcgpu_grid_rule(requested, tpb, C.byref(h), C.byref(b), C.byref(n))
    return b.value, h.value, n.value
