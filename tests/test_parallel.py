"""N>1 path on CPU: two gloo ranks shard a batch like split_render_data and gather the predict rows.
(The decode on each rank is played by the oracle here - this test covers the host-side sharding logic only;
the GPU tests cover the kernels.)"""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import decode, nets, weights
    from yolo_b200 import parallel
    spec = nets.spec_micro(size=(64, 96), C=10)
    B = 5                                                    # ragged split: 2 + 3
    heads = weights.synthetic_heads(B, spec, seed=42)
    lo, hi = parallel.shard_bounds(B, rank, world)
    local = decode.predict(spec, [h[lo:hi] for h in heads])
    rows = parallel.gather_rows(local, B)
    if rank == 0:
        q.put((rows, decode.predict(spec, heads), [parallel.shard_bounds(B, r, world) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, full, bounds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert bounds == [(0, 2), (2, 5)]
    np.testing.assert_array_equal(rows, full)


def test_shard_bounds_match_reference_formula():
    from yolo_b200 import parallel
    for B in (1, 5, 30, 32, 128):
        for N in (1, 2, 3, 4, 8):
            cover = []
            for i in range(N):
                lo, hi = parallel.shard_bounds(B, i, N)
                assert (lo, hi) == (int(i * B / N), int((i + 1) * B / N))      # yolo_gluon.py:117-118
                cover += list(range(lo, hi))
            assert cover == list(range(B))
    assert parallel.shard_batch(list(range(10)), 1, 4) == [2, 3, 4]
