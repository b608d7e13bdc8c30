"""GPU suite: chain-sharded runs.  Two ranks - two processes, here sharing cuda:0, exchanging
their moment rows over gloo - must each return exactly what a single process returns: every field of every
cluster's statistics is still the reference's sequential sum, whichever rank computed it."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, sha

pytestmark = pytest.mark.gpu

CASES = {
    "ictcp_k64": dict(side=768, K=64, cs=2, weighted=False, dither=False, km=0),
    "luv_k48_weighted_kmeans_dither": dict(side=512, K=48, cs=1, weighted=True, dither=True, km=3),
}


def _inputs(case):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import image_like_colors, saliency_like_weights, uniform_colors
    side = case["side"]
    colors = 0.5 * uniform_colors(side, side, 21) + 0.5 * image_like_colors(side, side, 22)
    w = saliency_like_weights(side, side, 23) if case["weighted"] else None
    return colors, w


def _run(case):
    import patolette_b200 as pb
    colors, w = _inputs(case)
    ok, pal, pmap, msg = pb.quantize(case["side"], case["side"], colors, case["K"], dither=case["dither"],
                                     color_space=case["cs"], tile_size=0, kmeans_niter=case["km"], weights=w)
    assert ok, msg
    return sha(pal), sha(pmap)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import patolette_b200 as pb
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb.set_sharding(rank, world, pb.torch_allgather())
    res = {}
    for name, case in CASES.items():
        res[name] = _run(case)
    pb.set_sharding(0, 1)
    out[rank] = res
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_chain_sharded_ranks_match_single_process(cuda_lib, world):
    import torch.multiprocessing as mp
    want = {name: _run(case) for name, case in CASES.items()}
    port = 29600 + os.getpid() % 2000 + world
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        got = {r: out[r] for r in range(world)}
    for r in range(world):
        assert got[r] == want, f"rank {r} of {world} differs from the single-process result"
