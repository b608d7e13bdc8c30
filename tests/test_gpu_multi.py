"""GPU suite (-m gpu), multi-GPU part: image-sharded runs (patolette_b200_sharded, one process per GPU, NCCL) must be
bit-identical to the single-GPU result for every world size.  Needs at least two GPUs (skipped otherwise)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CASES = {
    "ictcp_kmeans_dither": dict(w=1024, h=768, K=96, seed=61, color_space=2, dither=True, kmeans_niter=3),
    "luv_nodither_weighted": dict(w=900, h=700, K=64, seed=62, color_space=1, dither=False, kmeans_niter=0, weighted=True),
    "srgb_imagelike_nodither": dict(w=640, h=480, K=40, seed=63, color_space=0, dither=False, kmeans_niter=2, image_like=True),
    "tiny_some_ranks_empty": dict(w=37, h=11, K=12, seed=64, color_space=2, dither=True, kmeans_niter=0),
    "k256_2048": dict(w=2048, h=2048, K=256, seed=65, color_space=2, dither=True, kmeans_niter=10),
}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import patolette_b200 as pb
    from patolette_b200 import _lib
    from synth import make_case
    torch.cuda.set_device(rank)
    lib = _lib.load()
    assert lib.patolette_b200_set_device(rank) == 0
    dist.init_process_group("gloo", rank=rank, world_size=world)  # the unique id travels over gloo here
    assert pb.init_sharding(dist) == (rank, world)
    res = {}
    for name, spec in CASES.items():
        colors, weights, kw = make_case(spec)
        n = spec["w"] * spec["h"]
        first, count = pb.shard_range(n, rank, world)
        ok, pal, pmap, msg = pb.quantize_sharded(spec["w"], spec["h"], colors[first:first + count], spec["K"],
                                                 weights_slice=None if weights is None else weights[first:first + count], **kw)
        assert ok, msg
        res[name] = (first, count, pal.copy(), pmap.copy())
    # every split certificate refused: each cluster is evaluated twice, and all ranks must agree on the re-evaluations
    # (they are collective calls) - same result as the plain run
    assert lib.patolette_b200_set_option(b"split_certify", 2) == 0
    for name in ("luv_nodither_weighted", "ictcp_kmeans_dither"):
        spec = CASES[name]
        colors, weights, kw = make_case(spec)
        first, count = pb.shard_range(spec["w"] * spec["h"], rank, world)
        ok, pal, pmap, msg = pb.quantize_sharded(spec["w"], spec["h"], colors[first:first + count], spec["K"],
                                                 weights_slice=None if weights is None else weights[first:first + count], **kw)
        assert ok, msg
        assert np.array_equal(pal.view(np.uint64), res[name][2].view(np.uint64)) and np.array_equal(pmap, res[name][3]), \
            f"{name}: the re-evaluated run differs on rank {rank}"
    lib.patolette_b200_set_option(b"split_certify", 1)
    out[rank] = res
    lib.patolette_b200_comm_destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_image_sharded_matches_single_gpu(cuda_lib, world):
    import torch.multiprocessing as mp
    if cuda_lib.patolette_b200_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import patolette_b200 as pb
    from synth import make_case
    port = 29600 + (os.getpid() + world) % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        got = {r: out[r] for r in range(world)}
    for name, spec in CASES.items():
        colors, weights, kw = make_case(spec)
        ok, pal, pmap, msg = pb.quantize(spec["w"], spec["h"], colors, spec["K"], tile_size=0, weights=weights, **kw)
        assert ok, msg
        covered = 0
        for r in range(world):
            first, count, rpal, rmap = got[r][name]
            assert np.array_equal(rpal.view(np.uint64), pal.view(np.uint64)), f"{name}: palette of rank {r} differs"
            assert np.array_equal(rmap, pmap[first:first + count]), f"{name}: map slice of rank {r} differs"
            covered += count
        assert covered == spec["w"] * spec["h"]
