"""B200-native drop-in for the MC-GPU photon-transport path of 4d-cbct-mc.

The directory name is the one the build contract fixes; it is not a valid Python
identifier, so load it with `__graft_entry__.import_package()` which registers
it as the module `cbctmc_b200`.
"""
from . import mcio, phantoms, postprocess, sharding  # noqa: F401

__all__ = ["mcio", "phantoms", "postprocess", "sharding", "engine"]


def __getattr__(name):
    # the ctypes binding loads libmcgpu_b200.so and must fail loudly when it is
    # missing, but only for code that actually asks for the engine
    if name == "engine":
        import importlib

        return importlib.import_module(".engine", __name__)
    raise AttributeError(name)
