"""Host-side mirror of the nova-snark 0.23.0 interface for the per-step NIFS fold, backed by
libvimz_gpu.so.  Names, argument meaning and error behaviour follow the crate that zero-savvy/vimz
drives (hot loop entered at /root/reference/vimz/src/nova_snark_backend/folding.rs:35; curve cycle at
/root/reference/vimz/src/nova_snark_backend/mod.rs:19-20); the crate itself is un-vendored, so the
per-class citations are crate-relative ([EXT nova-snark] ...).

Every compute call goes to the GPU through the C ABI; nothing here does field or curve arithmetic.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import InvalidIndex, InvalidWitnessLength, VimzError, check, lib
from .field import CURVES, CurveInfo, as_fr, fr_array, ints_to_mont, mont_to_ints

NUM_CHALLENGE_BITS = 128  # [EXT nova-snark] src/constants.rs


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pinned_fr(src: np.ndarray) -> np.ndarray:
    """A copy of the (n, 4) uint64 field-element array `src` in page-locked host memory from `vimz_host_alloc` (what a host keeps
    its witness buffers in: a pageable buffer is staged by the driver at ~10 GB/s).  Freed when the array is collected."""
    import weakref
    src = np.ascontiguousarray(src, dtype=np.uint64)
    nbytes = max(src.nbytes, 32)
    p = lib.vimz_host_alloc(nbytes)
    if not p:
        raise MemoryError("vimz_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=np.uint64, count=src.size).reshape(src.shape)
    arr[...] = src
    weakref.finalize(buf, lib.vimz_host_free, C.c_void_p(p))
    return arr


class Engine:
    """One curve on one GPU (a `vimz_ctx`): plays the role of the `Group` type selected by
    `type G1` / `type G2` (mod.rs:19-20) together with its CommitmentEngine."""

    def __init__(self, curve: str = "pallas", device: int = 0):
        if curve not in CURVES:
            raise ValueError(f"unknown curve {curve!r}; expected one of {sorted(CURVES)}")
        self.curve: CurveInfo = CURVES[curve]
        self.device = device
        h = C.c_void_p()
        check(lib.vimz_ctx_create(self.curve.curve_id, device, C.byref(h)))
        self._h = h

    # -- lifetime --------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            lib.vimz_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- knobs / introspection -------------------------------------------------------------------
    def set_option(self, key: str, value: int):
        check(lib.vimz_ctx_set_option(self._h, key.encode(), int(value)))

    def sync(self):
        check(lib.vimz_ctx_sync(self._h))

    def profile(self, reset: bool = False) -> dict:
        """Device-side phase timers (enable with set_option("profile", 1)): name -> (ms, calls)."""
        out = {}
        names = ["msm_sort", "msm_accumulate", "msm_accumulate_kernel", "msm_accumulate_kernel_T", "msm_reduce", "cross_term", "axpy", "spmv",
                 "msm_entries_T", "msm_entries"]
        for i, name in enumerate(names):
            ms, calls = C.c_double(0), C.c_uint64(0)
            check(lib.vimz_ctx_profile(self._h, name.encode(), C.byref(ms), C.byref(calls), 1 if (reset and i == len(names) - 1) else 0))
            out[name] = (ms.value, calls.value)
        return out

    def lane_stats(self) -> dict:
        """Statistics of the last MSM run on each lane: bucket insertions and how many cut buckets went to the warp
        (mid) and block (giant, in chunks) roles of the combine kernel."""
        out = {}
        for lane in (0, 1):
            for what in ("entries", "nmid", "ngiant", "nchunk"):
                calls = C.c_uint64(0)
                rc = lib.vimz_ctx_profile(self._h, f"lane{lane}_{what}".encode(), None, C.byref(calls), 0)
                out[f"lane{lane}_{what}"] = int(calls.value) if rc == 0 else None
        return out

    @property
    def stream(self) -> int:
        return int(lib.vimz_ctx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib.vimz_ctx_launch_count(self._h))

    # -- group helpers ---------------------------------------------------------------------------
    def point_sum(self, pts: np.ndarray) -> np.ndarray:
        pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 12)
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_point_sum(self._h, _ptr(pts), pts.shape[0], _ptr(out)))
        return out

    def to_affine(self, pt: np.ndarray) -> np.ndarray:
        """Jacobian -> canonical affine (x, y) Montgomery limbs; identity -> zeros
        (Commitment::to_coordinates)."""
        pt = np.ascontiguousarray(pt, dtype=np.uint64).reshape(12)
        out = np.zeros(8, dtype=np.uint64)
        check(lib.vimz_point_to_affine(self._h, _ptr(pt), _ptr(out)))
        return out

    def to_affine_ints(self, pt: np.ndarray):
        a = self.to_affine(pt)
        x, y = mont_to_ints(a.reshape(2, 4), self.curve.base_modulus)
        return None if (x == 0 and y == 0) else (x, y)

    def point_scale_add(self, a: np.ndarray, r: np.ndarray, b: np.ndarray) -> np.ndarray:
        """a + r*b (RelaxedR1CSInstance::fold's commitment updates)."""
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(12)
        r = as_fr(r, 1)
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_point_scale_add(self._h, _ptr(a), _ptr(r), _ptr(b), _ptr(out)))
        return out

    def field_op(self, which: str, op: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a, b = as_fr(a), as_fr(b, None)
        out = np.zeros_like(a)
        check(lib.vimz_field_op(self._h, {"base": 0, "scalar": 1}[which], {"mul": 0, "add": 1, "sub": 2, "sqr": 3}[op],
                                _ptr(a), _ptr(b), a.shape[0], _ptr(out)))
        return out

    def scalars(self, vals: Sequence[int]) -> np.ndarray:
        """Canonical integers -> Montgomery scalar rows of this curve's scalar field."""
        return ints_to_mont(vals, self.curve.scalar_modulus)

    def scalar_ints(self, arr: np.ndarray) -> list:
        return mont_to_ints(arr, self.curve.scalar_modulus)


class CommitmentKey:
    """[EXT nova-snark] src/provider/pedersen.rs `CommitmentKey<G>{ck: Vec<G::PreprocessedGroupElement>}`.
    Uploaded once; the GPU expands it into the resident window table."""

    def __init__(self, engine: Engine, handle, n: int):
        self.engine = engine
        self._h = handle
        self.n = n

    @classmethod
    def from_bases(cls, engine: Engine, bases: np.ndarray) -> "CommitmentKey":
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        h = C.c_void_p()
        check(lib.vimz_ck_upload(engine._h, _ptr(bases), bases.shape[0], C.byref(h)))
        return cls(engine, h, bases.shape[0])

    @classmethod
    def from_device(cls, engine: Engine, d_bases: int, n: int) -> "CommitmentKey":
        h = C.c_void_p()
        check(lib.vimz_ck_upload_dev(engine._h, C.c_void_p(d_bases), n, C.byref(h)))
        return cls(engine, h, n)

    def __len__(self):
        return self.n

    @property
    def window_bits(self) -> int:
        return lib.vimz_ck_window_bits(self._h)

    @property
    def num_windows(self) -> int:
        return lib.vimz_ck_num_windows(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib.vimz_ck_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CommitmentEngine:
    """[EXT nova-snark] src/provider/pedersen.rs `CommitmentEngine<G>`."""

    @staticmethod
    def commit(ck: CommitmentKey, v: np.ndarray) -> np.ndarray:
        """commit(ck, v) = sum v_i * ck_i over ck[..len(v)]; `assert!(ck.ck.len() >= v.len())` in the
        reference surfaces here as InvalidWitnessLength.  Returns a Jacobian point (12 x u64)."""
        v = as_fr(v)
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_msm(ck.engine._h, ck._h, _ptr(v), v.shape[0], _ptr(out)))
        return out

    @staticmethod
    def commit_dev(ck: CommitmentKey, d_scalars: int, n: int, first: int = 0) -> np.ndarray:
        """Same with the scalars already resident in HBM (device pointer); `first` selects the point
        range ck[first .. first+n) for the multi-GPU shards."""
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_msm_range_dev(ck.engine._h, ck._h, first, C.c_void_p(d_scalars), n, _ptr(out)))
        return out

    @staticmethod
    def commit_async_dev(ck: CommitmentKey, d_scalars: int, n: int, d_out: int, first: int = 0) -> None:
        check(lib.vimz_msm_async_dev(ck.engine._h, ck._h, first, C.c_void_p(d_scalars), n, C.c_void_p(d_out)))


@dataclass
class R1CSWitness:
    """[EXT nova-snark] src/r1cs.rs `R1CSWitness{W}`."""
    W: np.ndarray

    def commit(self, ck: CommitmentKey) -> np.ndarray:
        return CommitmentEngine.commit(ck, self.W)


@dataclass
class R1CSInstance:
    """[EXT nova-snark] src/r1cs.rs `R1CSInstance{comm_W, X}`."""
    comm_W: np.ndarray
    X: np.ndarray


@dataclass
class RelaxedR1CSWitness:
    """[EXT nova-snark] src/r1cs.rs `RelaxedR1CSWitness{W, E}`."""
    W: np.ndarray
    E: np.ndarray

    @classmethod
    def default(cls, shape: "R1CSShape") -> "RelaxedR1CSWitness":
        return cls(fr_array(shape.num_vars), fr_array(shape.num_cons))

    def fold(self, engine: Engine, W2: R1CSWitness, T: np.ndarray, r: np.ndarray) -> "RelaxedR1CSWitness":
        """W = W1 + r*W2 ; E = E1 + r*T.  Length mismatch -> InvalidWitnessLength as in the reference."""
        W1, E1, Wv2, T = as_fr(self.W), as_fr(self.E), as_fr(W2.W), as_fr(T)
        if W1.shape[0] != Wv2.shape[0] or E1.shape[0] != T.shape[0]:
            raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "RelaxedR1CSWitness::fold: length mismatch")
        r = as_fr(r, 1)
        W, E = np.zeros_like(W1), np.zeros_like(E1)
        check(lib.vimz_fold_witness(engine._h, _ptr(r), _ptr(W1), _ptr(Wv2), W1.shape[0], _ptr(E1), _ptr(T), E1.shape[0],
                                    _ptr(W), _ptr(E)))
        return RelaxedR1CSWitness(W, E)


@dataclass
class RelaxedR1CSInstance:
    """[EXT nova-snark] src/r1cs.rs `RelaxedR1CSInstance{comm_W, comm_E, X, u}`."""
    comm_W: np.ndarray
    comm_E: np.ndarray
    X: np.ndarray
    u: np.ndarray

    @classmethod
    def default(cls, shape: "R1CSShape") -> "RelaxedR1CSInstance":
        return cls(np.zeros(12, np.uint64), np.zeros(12, np.uint64), fr_array(shape.num_io), fr_array(1))

    def fold(self, engine: Engine, U2: R1CSInstance, comm_T: np.ndarray, r: np.ndarray) -> "RelaxedR1CSInstance":
        """X = X1 + r*X2; comm_W = comm_W1 + r*comm_W2; comm_E = comm_E1 + r*comm_T; u = u1 + r."""
        r = as_fr(r, 1)
        io = as_fr(self.X).shape[0]
        # (u, X) + r * (1, X2) through the same axpy kernel as the witness fold
        one = engine.scalars([1])
        t1 = np.concatenate([as_fr(self.u, 1), as_fr(self.X)])
        t2 = np.concatenate([one, as_fr(U2.X, io)])
        out, _ = np.zeros_like(t1), None
        empty = fr_array(0)
        check(lib.vimz_fold_witness(engine._h, _ptr(r), _ptr(t1), _ptr(t2), t1.shape[0], _ptr(empty), _ptr(empty), 0,
                                    _ptr(out), _ptr(empty)))
        comm_W = engine.point_scale_add(self.comm_W, r, U2.comm_W)
        comm_E = engine.point_scale_add(self.comm_E, r, comm_T)
        return RelaxedR1CSInstance(comm_W, comm_E, out[1:].copy(), out[:1].copy())


class R1CSShape:
    """[EXT nova-snark] src/r1cs.rs `R1CSShape{num_cons, num_vars, num_io, A, B, C}` with A/B/C as COO
    triples (row, col, val) in constraint order; column space z = (W || u || X)."""

    def __init__(self, engine: Engine, num_cons: int, num_vars: int, num_io: int, A, B, C_):
        """A, B, C_: tuples (rows uint32[nnz], cols uint32[nnz], vals uint64[nnz,4] Montgomery).
        Out-of-range entries raise InvalidIndex like R1CSShape::new."""
        self.engine = engine
        self.num_cons, self.num_vars, self.num_io = int(num_cons), int(num_vars), int(num_io)
        mats = []
        for rows, cols, vals in (A, B, C_):
            rows = np.ascontiguousarray(rows, dtype=np.uint32)
            cols = np.ascontiguousarray(cols, dtype=np.uint32)
            vals = as_fr(vals, rows.shape[0])
            if cols.shape[0] != rows.shape[0]:
                raise ValueError("rows/cols length mismatch")
            mats.append((rows, cols, vals))
        h = C.c_void_p()
        args = []
        for rows, cols, vals in mats:
            args += [_ptr(rows), _ptr(cols), _ptr(vals), rows.shape[0]]
        check(lib.vimz_shape_upload(engine._h, self.num_cons, self.num_vars, self.num_io, *args, C.byref(h)))
        self._h = h
        self.nnz = tuple(m[0].shape[0] for m in mats)

    def close(self):
        if getattr(self, "_h", None):
            lib.vimz_shape_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def multiply_vec(self, z: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(Az, Bz, Cz); `z.len() != num_io + num_vars + 1` -> InvalidWitnessLength."""
        z = as_fr(z)
        Az, Bz, Cz = fr_array(self.num_cons), fr_array(self.num_cons), fr_array(self.num_cons)
        check(lib.vimz_multiply_vec(self.engine._h, self._h, _ptr(z), z.shape[0], _ptr(Az), _ptr(Bz), _ptr(Cz)))
        return Az, Bz, Cz

    def commit_T(self, ck: CommitmentKey, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness, U2: R1CSInstance,
                 W2: R1CSWitness) -> Tuple[np.ndarray, np.ndarray]:
        """(T, comm_T) with T = Az1 o Bz2 + Az2 o Bz1 - u1*Cz2 - u2*Cz1, u2 = 1."""
        Wv1, Wv2 = as_fr(W1.W), as_fr(W2.W)
        if Wv1.shape[0] != self.num_vars or Wv2.shape[0] != self.num_vars:
            raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "commit_T: witness length != num_vars")
        X1, X2 = as_fr(U1.X, self.num_io), as_fr(U2.X, self.num_io)
        u1 = as_fr(U1.u, 1)
        T = fr_array(self.num_cons)
        comm_T = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_commit_T(self.engine._h, self._h, ck._h, _ptr(Wv1), _ptr(u1), _ptr(X1), _ptr(Wv2), _ptr(X2),
                                _ptr(T), _ptr(comm_T)))
        return T, comm_T


class UnSat(VimzError):
    """nova-snark's NovaError::UnSat."""

    def __init__(self, msg: str):
        super().__init__(-100, msg)


def _check_sat(shape: "R1CSShape", ck: CommitmentKey, W: np.ndarray, E: Optional[np.ndarray], u: np.ndarray, X: np.ndarray,
               comm_W: np.ndarray, comm_E: Optional[np.ndarray]) -> None:
    eng = shape.engine
    W = as_fr(W)
    if W.shape[0] != shape.num_vars or as_fr(X).shape[0] != shape.num_io:
        raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "is_sat: witness / instance length mismatch")
    Az, Bz, Cz = shape.multiply_vec(np.concatenate([W, as_fr(u, 1), as_fr(X)]))
    lhs = eng.field_op("scalar", "mul", Az, Bz)
    rhs = eng.field_op("scalar", "mul", np.repeat(as_fr(u, 1), shape.num_cons, axis=0), Cz)
    if E is not None:
        rhs = eng.field_op("scalar", "add", rhs, as_fr(E, shape.num_cons))
    if not np.array_equal(lhs, rhs):
        raise UnSat("Az o Bz != u*Cz + E")
    if eng.to_affine_ints(CommitmentEngine.commit(ck, W)) != eng.to_affine_ints(comm_W):
        raise UnSat("comm_W does not open to W")
    if E is not None and eng.to_affine_ints(CommitmentEngine.commit(ck, as_fr(E))) != eng.to_affine_ints(comm_E):
        raise UnSat("comm_E does not open to E")


def is_sat_relaxed(shape: "R1CSShape", ck: CommitmentKey, U: "RelaxedR1CSInstance", W: "RelaxedR1CSWitness") -> None:
    """[EXT nova-snark] src/r1cs.rs R1CSShape::is_sat_relaxed (what RecursiveSNARK::verify runs,
    /root/reference/vimz/src/nova_snark_backend/folding.rs:53-55): Az o Bz = u*Cz + E and both commitments open.
    Raises UnSat; every product and both MSMs run on the GPU."""
    _check_sat(shape, ck, W.W, W.E, U.u, U.X, U.comm_W, U.comm_E)


def is_sat(shape: "R1CSShape", ck: CommitmentKey, U: "R1CSInstance", W: "R1CSWitness") -> None:
    """[EXT nova-snark] src/r1cs.rs R1CSShape::is_sat: Az o Bz = Cz (u = 1) and comm_W opens."""
    one = shape.engine.scalars([1])
    _check_sat(shape, ck, W.W, None, one, U.X, U.comm_W, None)


class TranscriptRO:
    """Stand-in for the random oracle of NIFS::prove.  The reference's RO is Poseidon over the base
    field ([EXT nova-snark] src/provider/poseidon.rs) and stays untouched host code (SURVEY.md row
    a15); the fold only needs *a* 128-bit challenge r, so the harness derives it from SHAKE-256 over
    the absorbed canonical values.  Swap in the real RO on the Rust side."""

    def __init__(self, label: bytes = b"vimz-b200"):
        self._h = hashlib.shake_256(label)

    def absorb_ints(self, *vals: int):
        for v in vals:
            self._h.update(int(v).to_bytes(32, "little"))

    def squeeze(self, num_bits: int = NUM_CHALLENGE_BITS) -> int:
        return int.from_bytes(self._h.digest(32), "little") & ((1 << num_bits) - 1)


class NIFS:
    """[EXT nova-snark] src/nifs.rs `NIFS{comm_T}`."""

    def __init__(self, comm_T: np.ndarray):
        self.comm_T = comm_T

    @staticmethod
    def prove(ck: CommitmentKey, ro: TranscriptRO, shape: R1CSShape, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness,
              U2: R1CSInstance, W2: R1CSWitness):
        """-> (NIFS, (U, W)): absorb U1, U2; (T, comm_T) = commit_T; absorb comm_T; r = squeeze(128);
        U = U1.fold(U2, comm_T, r); W = W1.fold(W2, T, r)."""
        eng = shape.engine
        for pt in (U1.comm_W, U1.comm_E, U2.comm_W):
            a = eng.to_affine_ints(pt)
            ro.absorb_ints(*(a if a else (0, 0)))
        ro.absorb_ints(*eng.scalar_ints(U1.u), *eng.scalar_ints(U1.X), *eng.scalar_ints(U2.X))
        T, comm_T = shape.commit_T(ck, U1, W1, U2, W2)
        a = eng.to_affine_ints(comm_T)
        ro.absorb_ints(*(a if a else (0, 0)))
        r = eng.scalars([ro.squeeze(NUM_CHALLENGE_BITS)])
        U = U1.fold(eng, U2, comm_T, r)
        W = W1.fold(eng, W2, T, r)
        return NIFS(comm_T), (U, W)


class FoldAccumulator:
    """Device-resident running instance (`vimz_acc`): the state RecursiveSNARK keeps in r_U/r_W for one
    curve ([EXT nova-snark] src/lib.rs), held in HBM so a fold step moves only W2 in and two
    commitments out."""

    def __init__(self, shape: R1CSShape, ck: CommitmentKey):
        self.shape, self.ck, self.engine = shape, ck, shape.engine
        h = C.c_void_p()
        check(lib.vimz_acc_init(self.engine._h, shape._h, ck._h, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.vimz_acc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        """Back to RelaxedR1CSInstance::default / RelaxedR1CSWitness::default (a new proof on the same shape and key)."""
        check(lib.vimz_acc_reset(self._h))

    def load(self, U: RelaxedR1CSInstance, W: RelaxedR1CSWitness):
        s = self.shape
        Wv, E = as_fr(W.W, s.num_vars), as_fr(W.E, s.num_cons)
        u, X = as_fr(U.u, 1), as_fr(U.X, s.num_io)
        cw = np.ascontiguousarray(U.comm_W, dtype=np.uint64).reshape(12)
        ce = np.ascontiguousarray(U.comm_E, dtype=np.uint64).reshape(12)
        check(lib.vimz_acc_load(self._h, _ptr(Wv), _ptr(E), _ptr(u), _ptr(X), _ptr(cw), _ptr(ce)))

    # The three per-step calls are the host side of a ~1 ms loop: output buffers and their ctypes pointers are made once
    # (numpy's .ctypes accessor costs microseconds per use), inputs that already are (n, 4) uint64 arrays skip the checks.
    def _io(self):
        io = getattr(self, "_io_cache", None)
        if io is None:
            out = np.zeros(24, dtype=np.uint64)
            base = out.ctypes.data
            io = self._io_cache = (out, C.c_void_p(base), C.c_void_p(base + 96))
        return io

    @staticmethod
    def _fr_ptr(a, n):
        if type(a) is bytes:  # 32*n raw bytes (little-endian Montgomery limbs): ctypes passes the buffer as is
            if len(a) != 32 * n:
                raise ValueError(f"expected {n} field elements ({32 * n} bytes), got {len(a)} bytes")
            return a, a
        if not (type(a) is np.ndarray and a.dtype == np.uint64 and a.ndim == 2 and a.shape == (n, 4) and a.flags.c_contiguous):
            a = as_fr(a, n)
        return a, C.c_void_p(a.__array_interface__["data"][0])

    def step_begin(self, W2: np.ndarray, X2: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """-> (comm_W2, comm_T).  W2 is a host array (copied H2D inside the call)."""
        s = self.shape
        W2 = as_fr(W2) if not (type(W2) is np.ndarray and W2.dtype == np.uint64 and W2.ndim == 2 and W2.flags.c_contiguous) else W2
        if W2.shape[0] != s.num_vars or W2.shape[1] != 4:
            raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "step_begin: witness length != num_vars")
        X2, px = self._fr_ptr(X2, s.num_io)
        out, pw, pt = self._io()
        check(lib.vimz_acc_step_begin(self._h, C.c_void_p(W2.__array_interface__["data"][0]), px, pw, pt))
        return out[:12].copy(), out[12:].copy()

    @staticmethod
    def _row_ptr(W2, first: int) -> int:
        """Address of row `first` of a full (n, 4) witness: a host array, or an int = device address of its first row."""
        return (W2 if isinstance(W2, int) else W2.__array_interface__["data"][0]) + 32 * first

    def stage_fresh(self, W2, first: int, count: int) -> None:
        """Enqueue the copy of W2[first : first + count] (rows of the full (n, 4) host array `W2`, or of the device buffer at
        address `W2`) behind the previous step_end and return at once; `W2` must stay alive and unchanged until the next
        step_begin_staged returns.  With the engine option "stage_commit" (default) a prefix / suffix range is also committed
        at once, beside whatever the GPU is doing for the other curve."""
        check(lib.vimz_acc_stage_fresh(self._h, C.c_void_p(self._row_ptr(W2, first)), first, count))

    def step_begin_staged(self, W2, first: int, count: int, X2: np.ndarray, wait: bool = True):
        """Upload the remaining rows W2[first : first + count] and run the step on the staged witness -> (comm_W2, comm_T);
        wait = False: enqueue only (step_wait collects the commitments)."""
        X2, px = self._fr_ptr(X2, self.shape.num_io)
        if not wait:
            self._pending = (W2, X2)
            check(lib.vimz_acc_step_begin_staged(self._h, C.c_void_p(self._row_ptr(W2, first)), first, count, px, None, None))
            return None
        out, pw, pt = self._io()
        check(lib.vimz_acc_step_begin_staged(self._h, C.c_void_p(self._row_ptr(W2, first)), first, count, px, pw, pt))
        return out[:12].copy(), out[12:].copy()

    def step_begin_async(self, W2, X2) -> None:
        """Enqueue a step (W2 / X2 host buffers must stay alive until step_wait returns).  W2: host array, or an int = device
        address of a resident witness of num_vars rows."""
        s = self.shape
        if not isinstance(W2, int):
            if not (type(W2) is np.ndarray and W2.dtype == np.uint64 and W2.ndim == 2 and W2.flags.c_contiguous):
                W2 = as_fr(W2)
            if W2.shape[0] != s.num_vars or W2.shape[1] != 4:
                raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "step_begin_async: witness length != num_vars")
        X2, px = self._fr_ptr(X2, s.num_io)
        self._pending = (W2, X2)
        check(lib.vimz_acc_step_begin_async(self._h, C.c_void_p(self._row_ptr(W2, 0)), px))

    def step_wait(self) -> Tuple[np.ndarray, np.ndarray]:
        out, pw, pt = self._io()
        check(lib.vimz_acc_step_wait(self._h, pw, pt))
        self._pending = None
        return out[:12].copy(), out[12:].copy()

    def step_begin_dev(self, d_W2: int, X2: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        X2, px = self._fr_ptr(X2, self.shape.num_io)
        out, pw, pt = self._io()
        check(lib.vimz_acc_step_begin_dev(self._h, C.c_void_p(d_W2), px, pw, pt))
        return out[:12].copy(), out[12:].copy()

    def commit_fresh(self, W2: np.ndarray, X2: np.ndarray) -> np.ndarray:
        """First half of step_begin: stage (W2, X2) in the accumulator, -> comm_W2 (r1cs_instance_and_witness)."""
        s = self.shape
        W2 = as_fr(W2)
        if W2.shape[0] != s.num_vars:
            raise InvalidWitnessLength(_lib.VIMZ_ERR_LENGTH, "commit_fresh: witness length != num_vars")
        X2, px = self._fr_ptr(X2, s.num_io)
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_acc_commit_fresh(self._h, _ptr(W2), px, _ptr(out)))
        return out

    def cross_begin(self) -> np.ndarray:
        """Second half of step_begin on the staged witness: -> comm_T (R1CSShape::commit_T inside NIFS::prove)."""
        out = np.zeros(12, dtype=np.uint64)
        check(lib.vimz_acc_cross_begin(self._h, _ptr(out)))
        return out

    def fresh_witness(self) -> Tuple[np.ndarray, np.ndarray]:
        """(W2, X2) staged by the last step_begin / commit_fresh (nova-snark's l_w_secondary / l_u_secondary.X)."""
        W2, X2 = fr_array(self.shape.num_vars), fr_array(self.shape.num_io)
        check(lib.vimz_acc_fresh_witness(self._h, _ptr(W2), _ptr(X2)))
        return W2, X2

    def step_end(self, r: np.ndarray):
        r, pr = self._fr_ptr(r, 1)
        check(lib.vimz_acc_step_end(self._h, pr))

    def last_T(self) -> np.ndarray:
        T = fr_array(self.shape.num_cons)
        check(lib.vimz_acc_last_T(self._h, _ptr(T)))
        return T

    def instance(self) -> RelaxedR1CSInstance:
        """The running RelaxedR1CSInstance only (comm_W, comm_E, X, u -- 256 bytes): what the RO absorbs as U1."""
        s = self.shape
        u, X = fr_array(1), fr_array(s.num_io)
        cw, ce = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
        check(lib.vimz_acc_download(self._h, None, None, _ptr(u), _ptr(X), _ptr(cw), _ptr(ce)))
        return RelaxedR1CSInstance(cw, ce, X, u)

    def download(self) -> Tuple[RelaxedR1CSInstance, RelaxedR1CSWitness]:
        s = self.shape
        W, E, u, X = fr_array(s.num_vars), fr_array(s.num_cons), fr_array(1), fr_array(s.num_io)
        cw, ce = np.zeros(12, np.uint64), np.zeros(12, np.uint64)
        check(lib.vimz_acc_download(self._h, _ptr(W), _ptr(E), _ptr(u), _ptr(X), _ptr(cw), _ptr(ce)))
        return RelaxedR1CSInstance(cw, ce, X, u), RelaxedR1CSWitness(W, E)
