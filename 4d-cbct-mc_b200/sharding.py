"""Multi-GPU partition of the MC-GPU path, one process per GPU (torch.distributed plumbing).

Two partitions, neither with an exchange step inside a projection:
  * projection-parallel (default, P >= world): projection p -> rank p mod world.  The seed of
    projection p is closed-form in p (MC-GPU_v1.3.cu:869 + 3456-3485), so every rank starts its
    projections independently; no collective on the data path.
  * history-split (P < world: air scan, single-projection reference runs): the reference grid's
    stream range [0, blocks*tpb) is cut into contiguous block ranges, each rank tallies its range
    and the u64 tallies are summed on rank 0 -- ncclReduce over NVLink when the backend is NCCL.
    Integer sums commute, so the result is bit-identical to a single-GPU run.  (The reference
    instead splits by a measured speed test and MPI_Reduce, MC-GPU_v1.3.cu:691-807, 1019, which
    is not reproducible.)
"""
from __future__ import annotations

import numpy as np


def projections_of_rank(rank: int, world: int, num_projections: int) -> list[int]:
    return list(range(rank, num_projections, world))


def stream_range_of_rank(rank: int, world: int, num_blocks: int, threads_per_block: int) -> tuple[int, int]:
    """Contiguous block range of the reference grid for this rank -> [stream_begin, stream_end)."""
    world = min(world, num_blocks)
    if rank >= world:
        return 0, 0
    return (num_blocks * rank // world) * threads_per_block, (num_blocks * (rank + 1) // world) * threads_per_block


def use_history_split(num_projections: int, world: int) -> bool:
    return num_projections < world


def reduce_tally(image, dst: int = 0):
    """Sum the per-rank u64 tallies onto rank `dst` with torch.distributed (NCCL: ncclReduce over
    NVLink; gloo on CPU).  `image` is a torch int64 tensor (u64 counts reinterpreted; two's
    complement addition is the same operation) and is reduced in place."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(image, dst=dst, op=dist.ReduceOp.SUM)
    return image


def as_int64_tensor(image_u64: np.ndarray, device=None):
    import torch

    t = torch.from_numpy(image_u64.view(np.int64))
    return t.to(device) if device is not None else t


class _DeviceImage:
    """The engine's device tally (uint64[4*Npix], cudaMalloc'ed by libmcgpu_b200) exposed through
    __cuda_array_interface__ as int64 (NCCL sums two's-complement words; same bits as the u64 sum)."""

    def __init__(self, ptr: int, words: int):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def device_tally_tensor(engine):
    """Zero-copy torch view (int64, on the engine's GPU) of the tally the last run call left on the device --
    what the one-process-per-GPU driver hands to torch.distributed.reduce (ncclReduce over NVLink)."""
    import torch

    info = engine.info
    ptr = engine.device_image_ptr
    if not ptr:
        raise RuntimeError("the engine has no device image (no usable GPU)")
    return torch.as_tensor(_DeviceImage(ptr, 4 * info.num_pixels_x * info.num_pixels_z), device="cuda")
