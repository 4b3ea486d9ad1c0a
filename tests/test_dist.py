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


def test_shard_matrix_keeps_rows_and_order():
    from vimz_b200.sharding import shard_matrix
    rows = np.array([5, 0, 3, 3, 9, 4], np.uint32)
    cols = np.arange(6, dtype=np.uint32)
    vals = np.arange(24, dtype=np.uint64).reshape(6, 4)
    r, c, v = shard_matrix((rows, cols, vals), 3, 3)
    assert r.tolist() == [2, 0, 0, 1] and c.tolist() == [0, 2, 3, 5] and np.array_equal(v, vals[[0, 2, 3, 5]])
    r, c, v = shard_matrix((rows, cols, vals), 6, 3)
    assert r.size == 0 and c.size == 0 and v.shape == (0, 4)


class _CpuShard:
    """Stand-in for vimz_b200.sharding.FoldShard with the CPU oracle doing a rank's share of the work, so the
    collective logic of ShardedFoldAccumulator runs under gloo on the CPU box."""

    def __init__(self, o, c, sh, bases, rank, world):
        from vimz_b200.field import ints_to_mont
        from vimz_b200.sharding import shard_matrix, shard_range
        self.o, self.cid, self.sh, self.q = o, c.curve_id, sh, c.q
        self.r0, self.ml = shard_range(sh.num_cons, rank, world)
        self.v0, self.vc = shard_range(sh.num_vars, rank, world)
        self.A, self.B, self.C = (shard_matrix(M, self.r0, self.ml) for M in (sh.A, sh.B, sh.C))
        self.bases = bases
        self.one = ints_to_mont([1], c.q)
        n, io = sh.num_vars, sh.num_io
        self.W1 = np.zeros((n, 4), np.uint64); self.E1 = np.zeros((self.ml, 4), np.uint64)
        self.u1 = np.zeros((1, 4), np.uint64); self.X1 = np.zeros((io, 4), np.uint64)
        self.cW = np.zeros(12, np.uint64); self.cE = np.zeros(12, np.uint64)

    def step_begin(self, W2, X2):
        o, cid, sh = self.o, self.cid, self.sh
        self.W2, self.X2 = W2, X2
        self.pW = o.msm(cid, W2[self.v0:self.v0 + self.vc], self.bases[self.v0:self.v0 + self.vc], 1)
        self.T = o.commit_T(cid, self.ml, sh.num_vars, sh.num_io, self.A, self.B, self.C, self.W1, self.u1, self.X1, W2, X2, self.one)
        self.pT = o.msm(cid, self.T, self.bases[self.r0:self.r0 + self.ml], 1)
        return self.pW, self.pT

    def step_end(self, r):
        o, cid = self.o, self.cid
        self.W1 = o.axpy(cid, self.W1, self.W2, r); self.E1 = o.axpy(cid, self.E1, self.T, r)
        tail = o.axpy(cid, np.concatenate([self.u1, self.X1]), np.concatenate([self.one, self.X2]), r)
        self.u1, self.X1 = tail[:1], tail[1:]
        self.cW = o.point_scale_add(cid, self.cW, r, self.pW); self.cE = o.point_scale_add(cid, self.cE, r, self.pT)

    def download(self):
        from vimz_b200.nova import RelaxedR1CSInstance, RelaxedR1CSWitness
        return RelaxedR1CSInstance(self.cW, self.cE, self.X1, self.u1), RelaxedR1CSWitness(self.W1, self.E1)


def _fold_worker(rank, world, port, out):
    import torch.distributed as dist
    from oracle import c as oracle_c, pyref as P
    from vimz_b200 import synthetic as S
    from vimz_b200.field import CURVES, affine_to_mont, ints_to_mont
    from vimz_b200.sharding import ShardedFoldAccumulator
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = P.PALLAS
        o = oracle_c()
        sh = S.synthetic_shape(CURVES["pallas"], "grayscale", seed=3, scale=0.004)
        g = affine_to_mont([P.generator(c)], c.p)[0]
        bases = o.gen_bases(c.curve_id, g, 11, 13, max(sh.num_cons, sh.num_vars))
        one = ints_to_mont([1], c.q)[0]

        def point_sum(pts):
            acc = np.zeros(12, np.uint64)
            for p in pts:
                acc = o.point_scale_add(c.curve_id, acc, one, p)
            return acc

        wit = []
        for k in range(2):
            Wi, Xi = S.synthetic_witness(sh, 50 + k)
            wit.append((ints_to_mont(Wi, c.q), ints_to_mont(Xi, c.q)))
        chal = [ints_to_mont([0xABCDEF0123456789ABCDEF + k], c.q) for k in range(2)]
        sharded = ShardedFoldAccumulator(dist, _CpuShard(o, c, sh, bases, rank, world), point_sum)
        whole = _CpuShard(o, c, sh, bases, 0, 1)
        for k in range(2):
            cw, ct = sharded.step_begin(*wit[k])
            ew, et = whole.step_begin(*wit[k])
            assert o.to_affine(c.curve_id, cw).tolist() == o.to_affine(c.curve_id, ew).tolist()
            assert o.to_affine(c.curve_id, ct).tolist() == o.to_affine(c.curve_id, et).tolist()
            sharded.step_end(chal[k]); whole.step_end(chal[k])
        U, W = sharded.download()
        Ue, We = whole.download()
        ok = (np.array_equal(W.W, We.W) and np.array_equal(W.E, We.E) and np.array_equal(U.u, Ue.u) and np.array_equal(U.X, Ue.X)
              and o.to_affine(c.curve_id, U.comm_W).tolist() == o.to_affine(c.curve_id, Ue.comm_W).tolist()
              and o.to_affine(c.curve_id, U.comm_E).tolist() == o.to_affine(c.curve_id, Ue.comm_E).tolist())
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_row_sharded_fold_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fold_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
