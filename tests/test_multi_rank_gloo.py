"""CPU suite: the N > 1 bench plumbing with world_size-2 gloo.  Round 1 runs replicas (one image
per rank, no data-path collective - DESIGN.md section 7), so what has to be right across ranks is:
per-rank inputs differ, timings are reduced with MAX, and only rank 0 reports."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from synth import uniform_colors
    dist.init_process_group("gloo", rank=rank, world_size=world)
    colors = uniform_colors(8, 8, bench.WORKLOAD["seed"] + rank)
    ms = bench.reduce_max_ms(10.0 + 5.0 * rank, dist, device="cpu")
    total = bench.aggregate_throughput(n_pixels=64, world=world, ms_per_step=ms)
    out[rank] = (float(colors.sum()), ms, total)
    dist.destroy_process_group()


def test_replica_plumbing_world2():
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        a, b = out[0], out[1]
    assert a[0] != b[0], "ranks must quantise different images"
    assert a[1] == b[1] == 15.0, "step time is the MAX over ranks"
    assert a[2] == b[2] == pytest.approx(2 * 64 / 15e-3 / 1e6)
