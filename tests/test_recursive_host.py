"""CPU-only: the host-side sequencing of vimz_b200/recursive.py (RecursiveSNARK mirror: base case, counter-only first call,
(1) -> (3)+(4) -> (6) order, the l_u2 / l_w2 hand-off, strict vs deferred secondary commit, verify) and vimz_b200/sonobe.py,
with the CPU oracle standing in for the resident accumulators -- no GPU, no product compute.  On the GPU box the same
classes drive libvimz_gpu.so (tests/test_gpu_recursive.py)."""
import numpy as np
import pytest

import vimz_b200.recursive as R
import vimz_b200.sonobe as SB
from vimz_b200.nova import R1CSInstance, RelaxedR1CSInstance, RelaxedR1CSWitness, UnSat
from vimz_host import synthetic as S
from vimz_host.field import CURVES, affine_to_mont, ints_to_mont, mont_to_affine, mont_to_ints
from oracle import pyref as P


class CpuEngine:
    def __init__(self, o, c):
        self.o, self.c = o, c

    def scalars(self, vals):
        return ints_to_mont(vals, self.c.q)

    def scalar_ints(self, arr):
        return mont_to_ints(arr, self.c.q)

    def to_affine_ints(self, pt):
        return mont_to_affine(self.o.to_affine(self.c.curve_id, pt), self.c.p)[0]


class CpuShape:
    def __init__(self, eng, sh):
        self.engine, self.sh = eng, sh
        self.num_cons, self.num_vars, self.num_io = sh.num_cons, sh.num_vars, sh.num_io


class CpuKey:
    def __init__(self, bases):
        self.bases = bases


CALLS = []


class CpuAcc:
    """FoldAccumulator's interface on the CPU oracle (the calls are logged so the test can check their order)."""

    def __init__(self, shape, ck):
        self.shape, self.ck, self.e = shape, ck, shape.engine
        self.o, self.cid, self.q = self.e.o, self.e.c.curve_id, self.e.c.q
        m, n, io = shape.num_cons, shape.num_vars, shape.num_io
        self.W = np.zeros((n, 4), np.uint64); self.E = np.zeros((m, 4), np.uint64)
        self.u = np.zeros((1, 4), np.uint64); self.X = np.zeros((io, 4), np.uint64)
        self.cW = np.zeros(12, np.uint64); self.cE = np.zeros(12, np.uint64)
        self.one = ints_to_mont([1], self.q)
        self.tag = "P" if m > 400 else "S"

    def _msm(self, v):
        return self.o.msm(self.cid, v, self.ck.bases, 2)

    def load(self, U, W):
        self.W, self.E, self.u, self.X, self.cW, self.cE = W.W.copy(), W.E.copy(), U.u.copy(), U.X.copy(), U.comm_W.copy(), U.comm_E.copy()

    def commit_fresh(self, W2, X2):
        CALLS.append(self.tag + ":commit_fresh")
        self.W2, self.X2 = np.array(W2), np.array(X2)
        self.cW2 = self._msm(self.W2)
        return self.cW2

    def cross_begin(self):
        CALLS.append(self.tag + ":cross_begin")
        sh = self.shape.sh
        self.T = self.o.commit_T(self.cid, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, self.W, self.u, self.X, self.W2, self.X2, self.one)
        self.cT = self._msm(self.T)
        return self.cT

    def step_begin(self, W2, X2):
        CALLS.append(self.tag + ":step_begin")
        CALLS_mute = len(CALLS)
        cw = self.commit_fresh(W2, X2)
        ct = self.cross_begin()
        del CALLS[CALLS_mute:]
        return cw, ct

    def step_end(self, r):
        CALLS.append(self.tag + ":step_end")
        o, cid = self.o, self.cid
        self.W = o.axpy(cid, self.W, self.W2, r); self.E = o.axpy(cid, self.E, self.T, r)
        tail = o.axpy(cid, np.concatenate([self.u, self.X]), np.concatenate([self.one, self.X2]), r)
        self.u, self.X = tail[:1], tail[1:]
        self.cW = o.point_scale_add(cid, self.cW, r, self.cW2); self.cE = o.point_scale_add(cid, self.cE, r, self.cT)

    def instance(self):
        return RelaxedR1CSInstance(self.cW, self.cE, self.X, self.u)

    def download(self):
        return self.instance(), RelaxedR1CSWitness(self.W, self.E)

    def fresh_witness(self):
        return self.W2, self.X2

    def close(self):
        pass


def _sat(shape, ck, W, E, u, X, comm_W, comm_E):
    e, sh = shape.engine, shape.sh
    o, cid, q = e.o, e.c.curve_id, e.c.q
    Az, Bz, Cz = o.multiply_vec(cid, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, np.concatenate([W, u, X]))
    a, b, c_, ee = (mont_to_ints(v, q) for v in (Az, Bz, Cz, E))
    uu = mont_to_ints(u, q)[0]
    if not all((x * y - uu * w - t) % q == 0 for x, y, w, t in zip(a, b, c_, ee)):
        raise UnSat("relation")
    if e.to_affine_ints(o.msm(cid, W, ck.bases, 2)) != e.to_affine_ints(comm_W):
        raise UnSat("comm_W")
    if comm_E is not None and e.to_affine_ints(o.msm(cid, E, ck.bases, 2)) != e.to_affine_ints(comm_E):
        raise UnSat("comm_E")


def _is_sat_relaxed(shape, ck, U, W):
    _sat(shape, ck, W.W, W.E, U.u, U.X, U.comm_W, U.comm_E)


def _is_sat(shape, ck, U, W):
    _sat(shape, ck, W.W, np.zeros((shape.num_cons, 4), np.uint64), shape.engine.scalars([1]), U.X, U.comm_W, None)


class _CE:
    @staticmethod
    def commit(ck, v):
        return ck._owner._msm(v) if hasattr(ck, "_owner") else ck.commit(v)


@pytest.fixture
def cpu_backend(monkeypatch, coracle):
    for mod in (R, SB):
        monkeypatch.setattr(mod, "FoldAccumulator", CpuAcc)
        monkeypatch.setattr(mod, "is_sat", _is_sat)
        monkeypatch.setattr(mod, "is_sat_relaxed", _is_sat_relaxed)

    class CE:
        @staticmethod
        def commit(ck, v):
            return coracle.msm(ck.cid, v, ck.bases, 2)

    monkeypatch.setattr(R, "CommitmentEngine", CE)

    def make(curve, circuit, scale, seed):
        c = P.CURVES[curve]
        sh = S.synthetic_shape(CURVES[curve], circuit, seed=seed, scale=scale)
        eng = CpuEngine(coracle, c)
        g = affine_to_mont([P.generator(c)], c.p)[0]
        ck = CpuKey(coracle.gen_bases(c.curve_id, g, 3, 5, max(sh.num_cons, sh.num_vars)))
        ck.cid = c.curve_id
        wits = []
        for k in range(4):
            Wi, Xi = S.synthetic_witness(sh, seed + 10 + k)
            wits.append((ints_to_mont(Wi, c.q), ints_to_mont(Xi, c.q)))
        return CpuShape(eng, sh), ck, wits

    return make


def _state(snark):
    out = []
    for acc in (snark.acc_primary, snark.acc_secondary):
        U, W = acc.download()
        out.append((W.W.tobytes(), W.E.tobytes(), U.u.tobytes(), U.X.tobytes(), acc.e.to_affine_ints(U.comm_W), acc.e.to_affine_ints(U.comm_E)))
    return out


def test_recursive_snark_sequencing_on_cpu_backend(cpu_backend):
    s1, ck1, w1 = cpu_backend("pallas", "grayscale", 0.006, 1)
    s2, ck2, w2 = cpu_backend("vesta", "secondary", 0.03, 2)
    pp = R.PublicParams(s1, ck1, s2, ck2, digest=9)
    seq = [(w1[k % 4], w2[(k + 1) % 4]) for k in range(5)]
    del CALLS[:]
    strict = R.fold_input(pp, seq, overlap_secondary=False)
    assert strict.i == 5
    # base case: the first secondary witness is committed; the first prove_step only counts; then (1) (3)+(4) (6) per step
    per_step = ["S:cross_begin", "S:step_end", "P:step_begin", "P:step_end", "S:commit_fresh"]
    assert CALLS == ["S:commit_fresh"] + per_step * 4
    # the base-case running primary instance is the first witness with u = 1 and E = 0 (from_r1cs_instance / from_r1cs_witness)
    fresh = R.RecursiveSNARK(pp, seq[0][0], seq[0][1])
    U, W = fresh.acc_primary.download()
    assert np.array_equal(W.W, seq[0][0][0]) and not W.E.any() and mont_to_ints(U.u, P.PALLAS.q) == [1]
    fresh.prove_step(*seq[0])
    assert fresh.i == 1 and not fresh.acc_secondary.W.any()               # counter only: nothing folded yet
    del CALLS[:]
    lazy = R.fold_input(pp, seq, overlap_secondary=True)
    assert CALLS == ["S:step_begin", "S:step_end", "P:step_begin", "P:step_end"] * 4     # (6) rides inside the next (1)
    assert _state(lazy) == _state(strict)                                  # both orders fold to identical pairs
    R.verify_folded_proof(strict, 5)
    R.verify_folded_proof(lazy, 5)                                         # forces the pending secondary commit first
    assert CALLS[-1] == "S:commit_fresh"
    with pytest.raises(UnSat):
        R.verify_folded_proof(strict, 6)
    bad = seq[1][1][0].copy()
    bad[0] = ints_to_mont([7], P.VESTA.q)[0]
    strict._set_fresh_secondary((bad, seq[1][1][1]))
    with pytest.raises(UnSat):
        strict.verify()


def test_sonobe_sequencing_on_cpu_backend(cpu_backend):
    s1, ck1, w1 = cpu_backend("bn254", "grayscale", 0.006, 3)
    s2, ck2, w2 = cpu_backend("grumpkin", "secondary", 0.03, 4)
    del CALLS[:]
    nova = SB.SonobeNova(s1, ck1, s2, ck2, w1[0], digest=5)
    for k in range(3):
        nova.prove_step(w1[(k + 1) % 4], w2[k % 4], w2[(k + 2) % 4])
    step = ["P:cross_begin", "P:step_end", "S:step_begin", "S:step_end", "S:step_begin", "S:step_end", "P:commit_fresh"]
    assert CALLS == ["P:commit_fresh"] + step * 3 and nova.i == 3
    nova.verify()
