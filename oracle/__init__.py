"""CPU oracle package (test infrastructure only -- see oracle/pyref.py and oracle/nova_cpu.c).

`c` loads oracle/liboracle.so (built by `make -C oracle`) through ctypes.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nova_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class COracle:
    """ctypes view of nova_cpu.c; all arrays are numpy uint64 with Montgomery limbs."""

    def __init__(self):
        self.lib = C.CDLL(build())
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L = self.lib
        L.oracle_msm.argtypes = [i32, vp, vp, sz, i32, vp]
        L.oracle_to_affine.argtypes = [i32, vp, vp]
        L.oracle_point_scale_add.argtypes = [i32, vp, vp, vp, vp]
        L.oracle_multiply_vec.argtypes = [i32, sz, sz, sz] + [vp, vp, vp, sz] * 3 + [vp, sz, vp, vp, vp, i32]
        L.oracle_cross_term.argtypes = [i32, sz] + [vp] * 8 + [i32]
        L.oracle_axpy.argtypes = [i32, sz, vp, vp, vp, vp, i32]
        L.oracle_field_op.argtypes = [i32, i32, i32, vp, vp, sz, vp]
        L.oracle_gen_bases.argtypes = [i32, vp, C.c_uint64, C.c_uint64, sz, vp]
        L.oracle_poseidon.argtypes = [i32, i32, i32, i32, vp, vp, vp, sz]
        self._poseidon_tables = {}

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def msm(self, curve_id, scalars, bases, nthreads=1):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        assert bases.shape[0] >= scalars.shape[0]
        out = np.zeros(12, dtype=np.uint64)
        rc = self.lib.oracle_msm(curve_id, self._p(scalars), self._p(bases), scalars.shape[0], nthreads, self._p(out))
        assert rc == 0
        return out

    def to_affine(self, curve_id, jac):
        jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(12)
        out = np.zeros(8, dtype=np.uint64)
        assert self.lib.oracle_to_affine(curve_id, self._p(jac), self._p(out)) == 0
        return out

    def point_scale_add(self, curve_id, a, r, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(12)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        out = np.zeros(12, dtype=np.uint64)
        assert self.lib.oracle_point_scale_add(curve_id, self._p(a), self._p(r), self._p(b), self._p(out)) == 0
        return out

    def multiply_vec(self, curve_id, m, num_vars, num_io, A, B, Cm, z, parallel=True):
        args = []
        keep = []
        for rows, cols, vals in (A, B, Cm):
            rows = np.ascontiguousarray(rows, dtype=np.uint32)
            cols = np.ascontiguousarray(cols, dtype=np.uint32)
            vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
            keep += [rows, cols, vals]
            args += [self._p(rows), self._p(cols), self._p(vals), rows.shape[0]]
        z = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 4)
        Az, Bz, Cz = (np.zeros((m, 4), dtype=np.uint64) for _ in range(3))
        rc = self.lib.oracle_multiply_vec(curve_id, m, num_vars, num_io, *args, self._p(z), z.shape[0],
                                          self._p(Az), self._p(Bz), self._p(Cz), 1 if parallel else 0)
        if rc == -3:
            raise ValueError("InvalidWitnessLength")
        assert rc == 0
        return Az, Bz, Cz

    def cross_term(self, curve_id, Az1, Bz1, Cz1, Az2, Bz2, Cz2, u1, nthreads=1):
        arrs = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in (Az1, Bz1, Cz1, Az2, Bz2, Cz2)]
        m = arrs[0].shape[0]
        u1 = np.ascontiguousarray(u1, dtype=np.uint64).reshape(4)
        T = np.zeros((m, 4), dtype=np.uint64)
        assert self.lib.oracle_cross_term(curve_id, m, *[self._p(a) for a in arrs], self._p(u1), self._p(T), nthreads) == 0
        return T

    def axpy(self, curve_id, a, b, r, nthreads=1):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        out = np.zeros_like(a)
        assert self.lib.oracle_axpy(curve_id, a.shape[0], self._p(a), self._p(b), self._p(r), self._p(out), nthreads) == 0
        return out

    def field_op(self, curve_id, which, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(a)
        assert self.lib.oracle_field_op(curve_id, which, op, self._p(a), self._p(b), a.shape[0], self._p(out)) == 0
        return out

    def gen_bases(self, curve_id, g_aff_mont, k0, dk, n):
        """(k0 + i*dk) * G for i < n as (n, 8) affine Montgomery rows."""
        g = np.ascontiguousarray(g_aff_mont, dtype=np.uint64).reshape(8)
        out = np.zeros((n, 8), dtype=np.uint64)
        assert self.lib.oracle_gen_bases(curve_id, self._p(g), k0, dk, n, self._p(out)) == 0
        return out

    def poseidon(self, state, curve_id=2):
        """circomlib Poseidon permutation of one width-t state of canonical integers over the curve's scalar field
        (default BN254 Fr); constants from oracle/poseidon.py, arithmetic in nova_cpu.c."""
        from . import poseidon as P
        from . import pyref
        q = {c.curve_id: c.q for c in pyref.CURVES.values()}[curve_id]
        t = len(state)
        R = 1 << 256

        def mont(vals):
            out = np.zeros((len(vals), 4), dtype=np.uint64)
            for i, v in enumerate(vals):
                m = (v % q) * R % q
                for k in range(4):
                    out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
            return out

        key = (curve_id, t)
        if key not in self._poseidon_tables:
            consts, mds = P.poseidon_params(t, q)
            self._poseidon_tables[key] = (mont(consts), mont([x for row in mds for x in row]), pow(R, -1, q))
        cm, mm, rinv = self._poseidon_tables[key]
        st = mont(list(state))
        rc = self.lib.oracle_poseidon(curve_id, t, P.N_ROUNDS_F, P.N_ROUNDS_P[t - 2], self._p(cm), self._p(mm), self._p(st), 1)
        assert rc == 0
        return [sum(int(st[i, k]) << (64 * k) for k in range(4)) * rinv % q for i in range(t)]

    def commit_T(self, curve_id, m, num_vars, num_io, A, B, Cm, W1, u1, X1, W2, X2, one_mont, nthreads=1):
        """T of R1CSShape::commit_T (the MSM is msm())."""
        z1 = np.concatenate([np.asarray(W1).reshape(-1, 4), np.asarray(u1).reshape(1, 4), np.asarray(X1).reshape(-1, 4)])
        z2 = np.concatenate([np.asarray(W2).reshape(-1, 4), np.asarray(one_mont).reshape(1, 4), np.asarray(X2).reshape(-1, 4)])
        a1, b1, c1 = self.multiply_vec(curve_id, m, num_vars, num_io, A, B, Cm, z1)
        a2, b2, c2 = self.multiply_vec(curve_id, m, num_vars, num_io, A, B, Cm, z2)
        return self.cross_term(curve_id, a1, b1, c1, a2, b2, c2, u1, nthreads)


_c = None


def c() -> COracle:
    global _c
    if _c is None:
        _c = COracle()
    return _c
