"""ctypes binding of oracle/liboracle.so.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
cpu_baseline / reference arm); the product package never imports this module."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
REF_CPU = REF_DIR / "MC-GPU_v1.3_CPU.x"
REF_CUDA_EXACT = REF_DIR / "MC-GPU_v1.3_sm100_exact.x"
REF_CUDA_FAST = REF_DIR / "MC-GPU_v1.3_sm100_fast.x"
REF_HOST_DUMP = REF_DIR / "ref_host_dump.x"

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = HERE / "liboracle.so"
        if not so.exists():
            subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)
        L = C.CDLL(str(so))
        L.oracle_load.restype = C.c_void_p
        L.oracle_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                  C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_info.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_histories.argtypes = [C.c_void_p, C.c_ulonglong]
        L.oracle_run_batches.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_run_projection_cpu_rule.restype = C.c_ulonglong
        L.oracle_run_projection_cpu_rule.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_run_projection_gpu_rule.restype = C.c_ulonglong
        L.oracle_run_projection_gpu_rule.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_table.restype = C.c_longlong
        L.oracle_table.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
        L.oracle_scalar.restype = C.c_float
        L.oracle_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_angle.restype = C.c_double
        L.oracle_angle.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_abmodm.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_init_prng.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_update_seed.argtypes = [C.c_int, C.c_ulonglong, C.c_int]
        L.oracle_ranecu.restype = C.c_float
        L.oracle_ranecu.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_grid_rule_gpu.argtypes = [C.c_ulonglong, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]
        _lib = L
    return _lib


_DTYPES = {
    "woodcock": np.float32, "mfp_a": np.float32, "mfp_b": np.float32, "rayleigh_xco": np.float32, "rayleigh_pco": np.float32,
    "rayleigh_aco": np.float32, "rayleigh_bco": np.float32, "rayleigh_itlco": np.uint8, "rayleigh_ituco": np.uint8,
    "rayleigh_pmax": np.float32, "compton_fco": np.float32, "compton_uico": np.float32, "compton_fj0": np.float32,
    "compton_noscco": np.int32, "density_nominal": np.float32, "density_max": np.float32, "espc": np.float32,
    "espc_cutoff": np.float32, "espc_alias": np.int16, "voxels": np.float32, "source": np.float32, "detector": np.uint8,
}


class Oracle:
    """One loaded simulation.  cxx_host_math=False is the pinned plain-C flavour (reference CPU
    build), True the nvcc-host flavour the product implements (see mcgpu_oracle.c header)."""

    def __init__(self, in_path, cxx_host_math: bool = False, voxels=None):
        L = lib()
        err = C.create_string_buffer(512)
        if voxels is None:
            h = L.oracle_load(str(in_path).encode(), int(cxx_host_math), 0, 0, 0, 0, 0, 0, None, None, err, 512)
        else:
            mat, rho, spacing = voxels  # arrays [x,y,z]
            nx, ny, nz = mat.shape
            m = np.ascontiguousarray(mat.transpose(2, 1, 0), dtype=np.uint8)
            r = np.ascontiguousarray(rho.transpose(2, 1, 0), dtype=np.float32)
            h = L.oracle_load(str(in_path).encode(), int(cxx_host_math), nx, ny, nz, *[np.float32(v) for v in spacing],
                              m.ctypes.data, r.ctypes.data, err, 512)
        if not h:
            raise RuntimeError("oracle_load failed: " + err.value.decode())
        self.h = C.c_void_p(h)
        info = np.zeros(16, dtype=np.int64)
        L.oracle_info(self.h, info.ctypes.data)
        (self.num_projections, self.npx, self.npz, self.nx, self.ny, self.nz, self.num_values, self.num_bins, self.tpb,
         self.hpt, self.seed, self.rotation_flag, self.total_histories) = [int(v) for v in info[:13]]

    def close(self):
        if self.h:
            lib().oracle_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_histories(self, n: int):
        lib().oracle_set_histories(self.h, n)
        self.total_histories = n

    def new_image(self):
        return np.zeros((4, self.npz, self.npx), dtype=np.uint64)

    def run_cpu_rule(self, p: int = 0, threads: int = 1):
        img = self.new_image()
        n = lib().oracle_run_projection_cpu_rule(self.h, p, threads, img.ctypes.data)
        return img, int(n)

    def run_gpu_rule(self, p: int = 0, threads: int = 1):
        img = self.new_image()
        n = lib().oracle_run_projection_gpu_rule(self.h, p, threads, img.ctypes.data)
        return img, int(n)

    def run_batches(self, p, seed, hpt, b0, b1, threads=1, image=None, count_events=False):
        img = self.new_image() if image is None else image
        ev = np.zeros(8, dtype=np.uint64) if count_events else None
        lib().oracle_run_batches(self.h, p, seed, hpt, b0, b1, threads, img.ctypes.data, ev.ctypes.data if count_events else None)
        return (img, ev) if count_events else img

    def table(self, name: str) -> np.ndarray:
        ptr = C.c_void_p()
        n = lib().oracle_table(self.h, name.encode(), C.byref(ptr))
        if n < 0:
            raise KeyError(name)
        buf = (C.c_char * n).from_address(ptr.value)
        return np.frombuffer(buf, dtype=_DTYPES[name]).copy()

    def scalar(self, name: str) -> float:
        return float(lib().oracle_scalar(self.h, name.encode()))

    def angle(self, name: str) -> float:
        return float(lib().oracle_angle(self.h, name.encode()))


def run_reference_binary(binary: Path, in_path: Path, cwd: Path | None = None, timeout: float = 3600.0) -> str:
    """Run one of the reference's own executables (oracle/_ref) on an input file; returns stdout."""
    res = subprocess.run([str(binary), str(in_path)], cwd=cwd, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError(f"{binary.name} exited {res.returncode}:\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}")
    return res.stdout


def reference_host_dump(in_path: Path, out_path: Path) -> dict:
    """Run the reference's OWN host stages (oracle/ref_host_dump.cu: read_input, spectrum, CT
    trajectory, load_voxels, load_material; nvcc host semantics) and return the raw tables."""
    import struct

    subprocess.run([str(REF_HOST_DUMP), str(in_path), str(out_path)], check=True, capture_output=True, timeout=600)
    blob = Path(out_path).read_bytes()
    out, i = {}, 0
    while i < len(blob):
        tag = blob[i:i + 32].split(b"\0")[0].decode()
        (n,) = struct.unpack("<Q", blob[i + 32:i + 40])
        out[tag] = blob[i + 40:i + 40 + n]
        i += 40 + n
    return out
