"""CPU-only, world_size 2 over gloo: the host-side logic of the multi-GPU MSM (point-range shards, all-gather of
the per-rank partial sums, combine).  The per-shard compute is the CPU oracle here; on the GPU box the same
functions wrap vimz_msm_range_dev / vimz_point_sum (bench.py, tests/test_gpu_msm.py shard test)."""
import os
import random
import socket

import numpy as np
import pytest

from vimz_b200.sharding import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, out):
    import torch.distributed as dist
    from oracle import c as oracle_c, pyref as P
    from vimz_b200.field import affine_to_mont, ints_to_mont
    from vimz_b200.sharding import sharded_commit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = P.PALLAS
        o = oracle_c()
        g = affine_to_mont([P.generator(c)], c.p)[0]
        bases = o.gen_bases(c.curve_id, g, 11, 13, n)       # every rank can regenerate the same key
        rng = random.Random(seed)
        sc = ints_to_mont([rng.randrange(c.q) for _ in range(n)], c.q)

        def local_commit(first, count):
            return o.msm(c.curve_id, sc[first:first + count], bases[first:first + count], 1)

        def point_sum(pts):
            acc = np.zeros(12, np.uint64)
            one = ints_to_mont([1], c.q)[0]
            for p in pts:
                acc = o.point_scale_add(c.curve_id, acc, one, p)
            return acc

        res = sharded_commit(dist, n, local_commit, point_sum)
        full = o.msm(c.curve_id, sc, bases, 1)
        out[rank] = (o.to_affine(c.curve_id, res).tolist(), o.to_affine(c.curve_id, full).tolist())
    finally:
        dist.destroy_process_group()


def test_sharded_commit_world2_gloo():
    import torch.multiprocessing as mp
    world, n = 2, 301
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 5, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        got, full = out[rank]
        assert got == full          # sharded result == single-process result, on every rank
    assert out[0][0] == out[1][0]
