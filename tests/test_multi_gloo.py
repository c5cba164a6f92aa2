"""N>1 path on CPU: world_size-2 gloo processes shard one projection's streams (history split)
and a scan's projections (projection-parallel) exactly like bench.py / the multi-process driver
do on GPUs; the compute stand-in is the oracle (allowed in tests), the thing under test is the
partition + reduce logic of 4d-cbct-mc_b200/sharding.py."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, inp, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "oracle"))
    from __graft_entry__ import import_package

    pkg = import_package()
    import oracle_py

    ora = oracle_py.Oracle(inp, cxx_host_math=True)
    eng = pkg.engine.Engine()
    eng.load_input(inp)
    info = eng.info
    sh = pkg.sharding
    # -- history split of projection 1
    b, e = sh.stream_range_of_rank(rank, world, info.num_blocks, info.threads_per_block)
    part = ora.run_batches(1, eng.projection_seed(1), info.histories_per_thread, b, e, threads=2)
    t = sh.as_int64_tensor(part)
    sh.reduce_tally(t, dst=0)
    if rank == 0:
        np.save(Path(out_dir) / "split.npy", t.numpy().view(np.uint64))
    # -- projection-parallel: each rank simulates its own projections, no collective on the data path
    for p in sh.projections_of_rank(rank, world, info.num_projections):
        img, _ = ora.run_gpu_rule(p, threads=2)
        np.save(Path(out_dir) / f"proj{p}.npy", img)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_history_split_and_projection_parallel(pkg, oracle_py, cases, tmp_path):
    inp, cfg, _ = cases["thorax_p4"]
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(inp), str(tmp_path)), nprocs=2, join=True)
    ora = oracle_py.Oracle(inp, cxx_host_math=True)
    full, _ = ora.run_gpu_rule(1, threads=4)
    assert np.array_equal(np.load(tmp_path / "split.npy"), full)  # integer sums: bit-identical to one rank
    for p in range(4):
        ref, _ = ora.run_gpu_rule(p, threads=4)
        assert np.array_equal(np.load(tmp_path / f"proj{p}.npy"), ref)


def test_partitions_cover_everything_once(pkg):
    sh = pkg.sharding
    for world in (1, 2, 3, 4, 8):
        got = sorted(p for r in range(world) for p in sh.projections_of_rank(r, world, 894))
        assert got == list(range(894))
        for blocks in (1, 5, 521, 30999, 65000):
            ranges = [sh.stream_range_of_rank(r, world, blocks, 128) for r in range(world)]
            ranges = [x for x in ranges if x[1] > x[0]]
            assert ranges[0][0] == 0 and ranges[-1][1] == blocks * 128
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert all(x[0] % 128 == 0 and x[1] % 128 == 0 for x in ranges)
    assert sh.use_history_split(1, 8) and not sh.use_history_split(894, 8)
