"""Point-range sharding of one MSM across ranks.

Rank g owns ck[first .. first+count) resident in its HBM and receives the matching scalar slice; each
rank runs a full local Pippenger and contributes ONE Jacobian point (96 bytes).  Elliptic-curve
addition is not an NCCL reduction, so the partial sums are all-gathered (12 x int64 per rank) and
added on every rank's GPU (vimz_point_sum).  Independent transformations need no collective at all
(replicas, one prover context per GPU).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n): the first n % world ranks get one extra element."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def gather_partial_points(dist, local_point: np.ndarray, device=None) -> np.ndarray:
    """all_gather of one Jacobian point per rank -> (world, 12) uint64 array (identical on every rank)."""
    import torch
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(local_point, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return torch.stack(parts).cpu().numpy().view(np.uint64).reshape(world, 12)


def sharded_commit(dist, n: int, local_commit: Callable[[int, int], np.ndarray], point_sum: Callable[[np.ndarray], np.ndarray],
                   device=None) -> np.ndarray:
    """commit over n points split by point range: local_commit(first, count) -> Jacobian point of this rank's
    shard; point_sum adds the gathered partial sums."""
    first, count = shard_range(n, dist.get_rank(), dist.get_world_size())
    local = local_commit(first, count)
    return point_sum(gather_partial_points(dist, local, device))


# ------------------------------------------------------------------------------------------------------------
# One fold step sharded by constraint-row range ("SpMV / cross-term / E-fold")
# ------------------------------------------------------------------------------------------------------------
class Comm:
    """`vimz_comm`: the NCCL communicator INSIDE libvimz_gpu.so (bound at run time with dlopen), so the exchange of a
    sharded step / MSM is enqueued by the library on its own stream -- no torch tensor, no Python between the kernels and
    the collective.  Rank 0 creates the 128-byte id; `from_torch_dist` ships it over an existing torch.distributed group
    (any backend), a non-Python host would use its own channel."""

    def __init__(self, device: int, rank: int, world: int, unique_id: bytes):
        import ctypes as C
        from ._lib import check, lib
        if len(unique_id) != 128:
            raise ValueError("NCCL unique id is 128 bytes")
        h = C.c_void_p()
        check(lib.vimz_comm_create(device, unique_id, rank, world, C.byref(h)))
        self._h, self.rank, self.world, self.device = h, rank, world, device

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from ._lib import check, lib
        buf = (C.c_uint8 * 128)()
        check(lib.vimz_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_dist(cls, dist, device: int) -> "Comm":
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(device, dist.get_rank(), dist.get_world_size(), box[0])

    def broadcast_dev(self, engine, d_buf: int, nbytes: int, root: int = 0) -> None:
        import ctypes as C
        from ._lib import check, lib
        check(lib.vimz_comm_broadcast_dev(engine._h, self._h, C.c_void_p(d_buf), nbytes, root))

    def commit_dev(self, ck, d_scalars: int, n: int, first: int = 0) -> np.ndarray:
        """commit over a key split by point range: this rank's slice; the partial sums are all-gathered and added on the
        GPU inside the call; every rank returns the same full commitment (vimz_msm_sharded_dev)."""
        import ctypes as C
        from ._lib import check, lib
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_msm_sharded_dev(ck.engine._h, self._h, ck._h, first, C.c_void_p(d_scalars), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            from ._lib import lib
            lib.vimz_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_matrix(M, first: int, count: int):
    """Rows [first, first+count) of a COO matrix (rows, cols, vals), re-based to row 0; entry order preserved."""
    rows, cols, vals = M
    rows = np.asarray(rows, dtype=np.uint32)
    keep = (rows >= first) & (rows < first + count)
    return (rows[keep] - np.uint32(first)).astype(np.uint32), np.asarray(cols, dtype=np.uint32)[keep], np.asarray(vals)[keep]


class _DevWords:
    """`n` int64 words of device memory at `ptr`, exposed through __cuda_array_interface__ so torch can view them."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}


class FoldShard:
    """One rank's part of a fold sharded across GPUs (`vimz_acc_init_sharded`): constraint rows
    [row_first, row_first + m_local) of A, B, C with the matching slices of E, T and of the commitment key,
    plus the variable range [var_first, var_first + var_count) of commit(W2).  W stays replicated: every rank
    needs the whole z = (W, u, X) for its rows, and folding W redundantly (n x 96 B of axpy) is cheaper than a
    collective.  step_begin returns PARTIAL commitments."""

    def __init__(self, engine, num_cons: int, num_vars: int, num_io: int, A, B, C_, make_ck: Callable, rank: int, world: int):
        import ctypes as C
        from ._lib import check, lib
        from .nova import FoldAccumulator, R1CSShape
        self.engine = engine
        self.rank, self.world = rank, world
        self.num_cons, self.num_vars, self.num_io = int(num_cons), int(num_vars), int(num_io)
        self.row_first, self.m_local = shard_range(self.num_cons, rank, world)
        self.var_first, self.var_count = shard_range(self.num_vars, rank, world)
        loc = [shard_matrix(M, self.row_first, self.m_local) for M in (A, B, C_)]
        self.shape = R1CSShape(engine, self.m_local, self.num_vars, self.num_io, *loc)
        # make_ck(first, count) -> CommitmentKey over ck[first .. first+count), resident on this rank's GPU
        self.ck_rows = make_ck(self.row_first, self.m_local)
        self.ck_vars = make_ck(self.var_first, self.var_count)
        h = C.c_void_p()
        check(lib.vimz_acc_init_sharded(engine._h, self.shape._h, self.ck_rows._h, self.ck_vars._h, self.var_first, self.var_count,
                                        C.byref(h)))
        self.acc = FoldAccumulator.__new__(FoldAccumulator)
        self.acc.shape, self.acc.ck, self.acc.engine, self.acc._h = self.shape, self.ck_rows, engine, h

    def step_begin(self, W2: np.ndarray, X2: np.ndarray):
        return self.acc.step_begin(W2, X2)

    def step_begin_dev(self, d_W2: int, X2: np.ndarray):
        return self.acc.step_begin_dev(d_W2, X2)

    def step_begin_dev_async(self, d_W2: int, X2: np.ndarray) -> int:
        """Enqueue the step; -> device address of this rank's partial (comm_W2, comm_T) pair (24 x u64)."""
        import ctypes as C
        from ._lib import check, lib
        from .field import as_fr
        X2 = as_fr(X2, self.num_io)
        p = C.c_void_p()
        check(lib.vimz_acc_step_begin_dev_async(self.acc._h, C.c_void_p(d_W2), X2.ctypes.data_as(C.c_void_p), C.byref(p)))
        return int(p.value)

    def step_combine_dev(self, d_gathered: int, world: int):
        import ctypes as C
        from ._lib import check, lib
        cw, ct = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
        check(lib.vimz_acc_step_combine_dev(self.acc._h, C.c_void_p(d_gathered), world, cw.ctypes.data_as(C.c_void_p),
                                            ct.ctypes.data_as(C.c_void_p)))
        return cw, ct

    def step_begin_sharded_dev(self, comm: "Comm", d_W2: int, X2: np.ndarray):
        """The whole sharded step inside the library: enqueue, NCCL all-gather of the partial pairs on the context stream, add
        on the GPU, one host wait -> FULL (comm_W2, comm_T) on every rank (vimz_acc_step_begin_sharded_dev)."""
        import ctypes as C
        from ._lib import check, lib
        X2, px = self.acc._fr_ptr(X2, self.num_io)
        out, pw, pt = self.acc._io()
        check(lib.vimz_acc_step_begin_sharded_dev(self.acc._h, comm._h, C.c_void_p(d_W2), px, pw, pt))
        return out[:12].copy(), out[12:].copy()

    def step_begin_sharded(self, comm: "Comm", W2, X2: np.ndarray, root: int = 0):
        """Same with the fresh witness in host memory on `root` only (None elsewhere): H2D on the root, NCCL broadcast into
        every rank's accumulator, then the step (vimz_acc_step_begin_sharded)."""
        import ctypes as C
        from ._lib import check, lib
        from .field import as_fr
        X2, px = self.acc._fr_ptr(X2, self.num_io)
        pw2 = None
        if W2 is not None:
            W2 = W2 if (type(W2) is np.ndarray and W2.dtype == np.uint64 and W2.ndim == 2 and W2.flags.c_contiguous) else as_fr(W2)
            if W2.shape != (self.num_vars, 4):
                from ._lib import VIMZ_ERR_LENGTH, InvalidWitnessLength
                raise InvalidWitnessLength(VIMZ_ERR_LENGTH, "step_begin_sharded: witness length != num_vars")
            pw2 = C.c_void_p(W2.__array_interface__["data"][0])
        out, pw, pt = self.acc._io()
        check(lib.vimz_acc_step_begin_sharded(self.acc._h, comm._h, pw2, root, px, pw, pt))
        return out[:12].copy(), out[12:].copy()

    def step_end(self, r: np.ndarray):
        self.acc.step_end(r)

    def download(self):
        """(U, W) of this shard: W.W is the whole (replicated) witness, W.E the local rows, U.comm_* partial sums."""
        return self.acc.download()

    def last_T(self) -> np.ndarray:
        return self.acc.last_T()

    def close(self):
        self.acc.close()
        self.shape.close()
        self.ck_rows.close()
        self.ck_vars.close()


class ShardedFoldAccumulator:
    """`FoldAccumulator` interface over one FoldShard per rank.  The only exchange of a step is the all-gather of
    the two partial commitments (2 x 96 B per rank) followed by point additions on every rank's GPU, so every
    rank derives the same challenge r from the same (comm_W2, comm_T).  `shard` may be any object with FoldShard's
    step/download methods and `point_sum` any callable adding (k, 12) Jacobian points (tests/test_dist.py drives
    this logic over gloo with CPU stand-ins)."""

    def __init__(self, dist, shard, point_sum: Callable[[np.ndarray], np.ndarray], device=None, comm: "Comm" = None):
        """comm != None: the exchange runs inside libvimz_gpu.so (vimz_acc_step_begin_sharded*); torch.distributed is then only
        used by download().  comm == None: the all-gather goes through torch.distributed (any backend; the CPU tests use gloo)."""
        self.dist, self.shard, self.point_sum, self.device, self.comm = dist, shard, point_sum, device, comm

    def _combine(self, cw: np.ndarray, ct: np.ndarray):
        both = np.concatenate([np.asarray(cw, np.uint64).reshape(12), np.asarray(ct, np.uint64).reshape(12)])
        import torch
        world = self.dist.get_world_size()
        t = torch.from_numpy(both.view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        parts = [torch.empty_like(t) for _ in range(world)]
        self.dist.all_gather(parts, t)
        g = torch.stack(parts).cpu().numpy().view(np.uint64).reshape(world, 2, 12)
        return self.point_sum(np.ascontiguousarray(g[:, 0])), self.point_sum(np.ascontiguousarray(g[:, 1]))

    def step_begin(self, W2: np.ndarray, X2: np.ndarray):
        return self._combine(*self.shard.step_begin(W2, X2))

    def step_begin_root(self, W2, X2: np.ndarray, root: int = 0):
        """The fresh witness exists in host memory on `root` only (library path: H2D there + NCCL broadcast + step)."""
        if self.comm is None:
            raise RuntimeError("step_begin_root needs a vimz_b200.sharding.Comm")
        return self.shard.step_begin_sharded(self.comm, W2, X2, root)

    def step_begin_dev(self, d_W2: int, X2: np.ndarray):
        if self.comm is not None:
            return self.shard.step_begin_sharded_dev(self.comm, d_W2, X2)
        if self.device is not None and hasattr(self.shard, "step_begin_dev_async"):
            return self._step_begin_on_stream(d_W2, X2)
        return self._combine(*self.shard.step_begin_dev(d_W2, X2))

    def _step_begin_on_stream(self, d_W2: int, X2: np.ndarray):
        """GPU path: the step is enqueued on the context stream, its 192-byte partial pair is all-gathered by NCCL in
        stream order on that same stream (no host copy of the partials), the shards are added on the GPU, and the host
        waits once for the two full commitments."""
        import torch
        world = self.dist.get_world_size()
        if getattr(self, "_recv", None) is None:
            self._recv = torch.empty(world * 24, dtype=torch.int64, device=self.device)
            self._stream = torch.cuda.ExternalStream(self.shard.engine.stream, device=self.device)
        d_part = self.shard.step_begin_dev_async(d_W2, X2)
        send = torch.as_tensor(_DevWords(d_part, 24), device=self.device)   # zero-copy view of the accumulator's slot pair
        with torch.cuda.stream(self._stream):
            self.dist.all_gather_into_tensor(self._recv, send)
        return self.shard.step_combine_dev(self._recv.data_ptr(), world)

    def step_end(self, r: np.ndarray):
        self.shard.step_end(r)

    def download(self):
        """Whole relaxed instance / witness on every rank: E rows concatenated in rank order, commitments summed."""
        from .nova import RelaxedR1CSInstance, RelaxedR1CSWitness
        U, W = self.shard.download()
        world = self.dist.get_world_size()
        objs = [None] * world
        self.dist.all_gather_object(objs, (np.asarray(W.E), np.asarray(U.comm_W), np.asarray(U.comm_E)))
        E = np.concatenate([o[0].reshape(-1, 4) for o in objs])
        comm_W = self.point_sum(np.stack([o[1].reshape(12) for o in objs]))
        comm_E = self.point_sum(np.stack([o[2].reshape(12) for o in objs]))
        return RelaxedR1CSInstance(comm_W, comm_E, U.X, U.u), RelaxedR1CSWitness(W.W, E)
