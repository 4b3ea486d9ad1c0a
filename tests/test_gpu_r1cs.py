"""GPU parity: R1CSShape::multiply_vec / commit_T, RelaxedR1CS*::fold, NIFS::prove and the
device-resident FoldAccumulator vs the oracle, through the C ABI.  Bit-exact: every output vector is
compared limb for limb with the 4x64 CPU restatement, commitments by canonical affine coordinates."""
import random

import numpy as np
import pytest

import vimz_b200
from vimz_b200 import (CommitmentEngine, CommitmentKey, FoldAccumulator, NIFS, R1CSInstance, R1CSShape, R1CSWitness,
                       RelaxedR1CSInstance, RelaxedR1CSWitness, TranscriptRO)
from vimz_b200 import synthetic as S
from vimz_b200.field import CURVES, affine_to_mont, ints_to_mont, mont_to_affine, mont_to_ints
from oracle import pyref as P
from conftest import make_bases

pytestmark = pytest.mark.gpu


def rand_coo(rng, m, ncols, nnz, q):
    rows = np.array([rng.randrange(m) for _ in range(nnz)], np.uint32)
    cols = np.array([rng.randrange(ncols) for _ in range(nnz)], np.uint32)
    vals = ints_to_mont([rng.choice([1, q - 1, 2, rng.randrange(q)]) for _ in range(nnz)], q)
    return rows, cols, vals


@pytest.mark.parametrize("name", ["pallas", "vesta", "bn254", "grumpkin"])
def test_multiply_vec_random_coo(name, engines, coracle):
    """Unsorted COO with duplicate (row, col) entries and empty rows, like a raw constraint list."""
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    rng = random.Random(c.curve_id + 100)
    m, n, io = 777, 500, 2
    A, B, Cm = (rand_coo(rng, m, n + 1 + io, k, q) for k in (3000, 1500, 40))
    shape = R1CSShape(eng, m, n, io, A, B, Cm)
    z = ints_to_mont([rng.randrange(q) for _ in range(n + 1 + io)], q)
    got = shape.multiply_vec(z)
    exp = coracle.multiply_vec(c.curve_id, m, n, io, A, B, Cm, z)
    for g, e in zip(got, exp):
        assert np.array_equal(g, e)
    with pytest.raises(vimz_b200.InvalidWitnessLength):
        shape.multiply_vec(z[:-1])
    with pytest.raises(vimz_b200.InvalidIndex):
        bad = (np.array([m], np.uint32), np.array([0], np.uint32), ints_to_mont([1], q))
        R1CSShape(eng, m, n, io, bad, B, Cm)
    with pytest.raises(vimz_b200.InvalidIndex):
        bad = (np.array([0], np.uint32), np.array([n + 1 + io], np.uint32), ints_to_mont([1], q))
        R1CSShape(eng, m, n, io, A, bad, Cm)
    shape.close()


def test_empty_matrices_and_zero_io(engines, coracle):
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    e = (np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 4), np.uint64))
    A = (np.array([0, 2], np.uint32), np.array([1, 3], np.uint32), ints_to_mont([5, 7], q))
    shape = R1CSShape(eng, 4, 3, 0, A, e, e)
    z = ints_to_mont([2, 3, 4, 1], q)
    Az, Bz, Cz = shape.multiply_vec(z)
    assert mont_to_ints(Az, q) == [15, 0, 7, 0] and not Bz.any() and not Cz.any()
    shape.close()


def _setup(eng, c, scale, seed, name="grayscale"):
    sh = S.synthetic_shape(CURVES[c.name], name, seed=seed, scale=scale)
    shape = R1CSShape(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C)
    nck = max(sh.num_cons, sh.num_vars)
    bases, _ = make_bases(c, nck, seed=seed)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    return sh, shape, ck, Bm


def _affine(coracle, c, jac):
    return mont_to_affine(coracle.to_affine(c.curve_id, jac), c.p)[0]


@pytest.mark.parametrize("name", ["pallas", "bn254"])
def test_commit_T_and_folds_vs_oracle(name, engines, coracle):
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    sh, shape, ck, Bm = _setup(eng, c, 0.02, seed=21)
    one = ints_to_mont([1], q)
    rng = random.Random(9)
    # a genuinely relaxed running instance: random W1, E1, u1 (commit_T is defined for any vectors)
    W1 = ints_to_mont([rng.randrange(q) for _ in range(sh.num_vars)], q)
    X1 = ints_to_mont([rng.randrange(q) for _ in range(sh.num_io)], q)
    u1 = ints_to_mont([rng.randrange(q)], q)
    W2i, X2i = S.synthetic_witness(sh, 5)
    W2, X2 = ints_to_mont(W2i, q), ints_to_mont(X2i, q)
    U1 = RelaxedR1CSInstance(np.zeros(12, np.uint64), np.zeros(12, np.uint64), X1, u1)
    U2 = R1CSInstance(np.zeros(12, np.uint64), X2)
    T, comm_T = shape.commit_T(ck, U1, RelaxedR1CSWitness(W1, np.zeros((sh.num_cons, 4), np.uint64)), U2, R1CSWitness(W2))
    T_exp = coracle.commit_T(c.curve_id, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, W1, u1, X1, W2, X2, one, nthreads=2)
    assert np.array_equal(T, T_exp)
    assert eng.to_affine_ints(comm_T) == _affine(coracle, c, coracle.msm(c.curve_id, T_exp, Bm, 4))
    # witness fold
    r = ints_to_mont([rng.randrange(1 << 128)], q)
    E1 = ints_to_mont([rng.randrange(q) for _ in range(sh.num_cons)], q)
    Wf = RelaxedR1CSWitness(W1, E1).fold(eng, R1CSWitness(W2), T, r)
    assert np.array_equal(Wf.W, coracle.axpy(c.curve_id, W1, W2, r))
    assert np.array_equal(Wf.E, coracle.axpy(c.curve_id, E1, T, r))
    with pytest.raises(vimz_b200.InvalidWitnessLength):
        RelaxedR1CSWitness(W1, E1).fold(eng, R1CSWitness(W2[:-1]), T, r)
    with pytest.raises(vimz_b200.InvalidWitnessLength):
        shape.commit_T(ck, U1, RelaxedR1CSWitness(W1[:-1], E1), U2, R1CSWitness(W2))
    shape.close(); ck.close()


@pytest.mark.parametrize("name", ["pallas", "grumpkin"])
def test_cross_term_row_classes_stream_vs_rowclass_vs_oracle(name, engines, coracle):
    """Rows of every class of the streamed cross term -- empty, 1-3 terms, 40 terms (warp-summed in shared memory),
    300 and 500 terms, 600 and 1100 terms (own chunk, warp over global memory), duplicates, unsorted COO -- next
    to runs of short rows that fill chunks to the 256-row and the 1024-non-zero limit.  T must equal the oracle's
    with the streamed kernel and with the row-class kernel (option cross_stream = 0)."""
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    rng = random.Random(77 + c.curve_id)
    n, io = 900, 2
    ncols = n + 1 + io
    lengths = [0, 1, 2, 3, 40, 0, 300, 5, 600, 2, 1100, 33, 32, 500, 7] + [rng.choice([0, 1, 2, 3, 4, 6]) for _ in range(700)] + [12] * 90 + [64, 65, 513, 512]
    m = len(lengths)
    consts = [1, q - 1, 2, 5] + [rng.randrange(q) for _ in range(6)]

    def matrix(scale_len):
        rows, cols, vals = [], [], []
        for i, L in enumerate(lengths):
            k = L if scale_len else min(L, 2)
            for _ in range(k):
                rows.append(i); cols.append(rng.randrange(ncols)); vals.append(rng.choice(consts))
        perm = list(range(len(rows)))
        rng.shuffle(perm)
        return (np.array([rows[j] for j in perm], np.uint32), np.array([cols[j] for j in perm], np.uint32),
                ints_to_mont([vals[j] for j in perm], q))

    A, B, Cm = matrix(True), matrix(False), matrix(False)
    shape = R1CSShape(eng, m, n, io, A, B, Cm)
    bases, _ = make_bases(c, max(m, n), seed=5)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    W1 = ints_to_mont([rng.randrange(q) for _ in range(n)], q)
    X1 = ints_to_mont([rng.randrange(q) for _ in range(io)], q)
    u1 = ints_to_mont([rng.randrange(q)], q)
    W2 = ints_to_mont([rng.choice([0, 1, rng.randrange(q)]) for _ in range(n)], q)
    X2 = ints_to_mont([rng.randrange(q) for _ in range(io)], q)
    one = ints_to_mont([1], q)
    U1 = RelaxedR1CSInstance(np.zeros(12, np.uint64), np.zeros(12, np.uint64), X1, u1)
    U2 = R1CSInstance(np.zeros(12, np.uint64), X2)
    T_exp = coracle.commit_T(c.curve_id, m, n, io, A, B, Cm, W1, u1, X1, W2, X2, one, nthreads=2)
    comm_exp = _affine(coracle, c, coracle.msm(c.curve_id, T_exp, Bm, 4))
    try:
        for stream in (1, 0):
            eng.set_option("cross_stream", stream)
            T, comm_T = shape.commit_T(ck, U1, RelaxedR1CSWitness(W1, np.zeros((m, 4), np.uint64)), U2, R1CSWitness(W2))
            assert np.array_equal(T, T_exp), f"cross_stream={stream}"
            assert eng.to_affine_ints(comm_T) == comm_exp
    finally:
        eng.set_option("cross_stream", 1)
    shape.close(); ck.close()


def _oracle_fold_chain(coracle, c, sh, Bm, witnesses, challenges):
    """Reference fold of a chain of fresh instances into the default relaxed instance, all on the CPU."""
    q = c.q
    cid = c.curve_id
    one = ints_to_mont([1], q)
    m, n, io = sh.num_cons, sh.num_vars, sh.num_io
    W1 = np.zeros((n, 4), np.uint64); E1 = np.zeros((m, 4), np.uint64)
    u1 = np.zeros((1, 4), np.uint64); X1 = np.zeros((io, 4), np.uint64)
    cW = np.zeros(12, np.uint64); cE = np.zeros(12, np.uint64)
    out = []
    for (W2, X2), r in zip(witnesses, challenges):
        comm_W2 = coracle.msm(cid, W2, Bm, 4)
        T = coracle.commit_T(cid, m, n, io, sh.A, sh.B, sh.C, W1, u1, X1, W2, X2, one, nthreads=2)
        comm_T = coracle.msm(cid, T, Bm, 4)
        W1 = coracle.axpy(cid, W1, W2, r); E1 = coracle.axpy(cid, E1, T, r)
        tail = coracle.axpy(cid, np.concatenate([u1, X1]), np.concatenate([one, X2]), r)
        u1, X1 = tail[:1], tail[1:]
        cW = coracle.point_scale_add(cid, cW, r, comm_W2); cE = coracle.point_scale_add(cid, cE, r, comm_T)
        out.append(dict(comm_W2=comm_W2, comm_T=comm_T, T=T, W=W1, E=E1, u=u1, X=X1, cW=cW, cE=cE))
    return out


@pytest.mark.parametrize("name", ["pallas", "vesta", "grumpkin"])
def test_fold_accumulator_chain_vs_oracle(name, engines, coracle):
    """Three consecutive NIFS folds on the device-resident accumulator == the CPU chain, bit for bit,
    and the final relaxed instance is satisfied (is_sat_relaxed, the reference's own verify check)."""
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    sh, shape, ck, Bm = _setup(eng, c, 0.015, seed=33)
    rng = random.Random(2)
    wit = []
    for k in range(3):
        Wi, Xi = S.synthetic_witness(sh, 100 + k)
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(3)]
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)
    acc = FoldAccumulator(shape, ck)
    for k in range(3):
        comm_W2, comm_T = acc.step_begin(*wit[k])
        assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, ref[k]["comm_W2"])
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[k]["comm_T"])
        assert np.array_equal(acc.last_T(), ref[k]["T"])
        acc.step_end(chal[k])
    U, W = acc.download()
    last = ref[-1]
    assert np.array_equal(W.W, last["W"]) and np.array_equal(W.E, last["E"])
    assert np.array_equal(U.u, last["u"]) and np.array_equal(U.X, last["X"])
    assert eng.to_affine_ints(U.comm_W) == _affine(coracle, c, last["cW"])
    assert eng.to_affine_ints(U.comm_E) == _affine(coracle, c, last["cE"])
    # RecursiveSNARK::verify's check on the folded accumulator: Az o Bz = u*Cz + E and the commitments open
    z = np.concatenate([W.W, U.u, U.X])
    Az, Bz, Cz = shape.multiply_vec(z)
    a, b, cz, e = (mont_to_ints(v, q) for v in (Az, Bz, Cz, W.E))
    u = mont_to_ints(U.u, q)[0]
    assert all((x * y - u * w - t) % q == 0 for x, y, w, t in zip(a, b, cz, e))
    assert eng.to_affine_ints(CommitmentEngine.commit(ck, W.W)) == eng.to_affine_ints(U.comm_W)
    assert eng.to_affine_ints(CommitmentEngine.commit(ck, W.E)) == eng.to_affine_ints(U.comm_E)
    acc.close(); shape.close(); ck.close()


def test_nifs_prove_host_api_matches_accumulator(engines):
    """The stateless host-buffer API (NIFS::prove over commit_T / fold) and the resident accumulator give
    identical folded instances."""
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    sh, shape, ck, Bm = _setup(eng, c, 0.01, seed=8)
    Wi, Xi = S.synthetic_witness(sh, 1)
    W2 = R1CSWitness(ints_to_mont(Wi, q))
    U2 = R1CSInstance(W2.commit(ck), ints_to_mont(Xi, q))
    U1, W1 = RelaxedR1CSInstance.default(shape), RelaxedR1CSWitness.default(shape)
    nifs, (U, W) = NIFS.prove(ck, TranscriptRO(), shape, U1, W1, U2, W2)
    # replay the same challenge on the accumulator
    ro = TranscriptRO()
    for pt in (U1.comm_W, U1.comm_E, U2.comm_W):
        a = eng.to_affine_ints(pt)
        ro.absorb_ints(*(a if a else (0, 0)))
    ro.absorb_ints(*eng.scalar_ints(U1.u), *eng.scalar_ints(U1.X), *eng.scalar_ints(U2.X))
    acc = FoldAccumulator(shape, ck)
    cw, ct = acc.step_begin(W2.W, U2.X)
    assert eng.to_affine_ints(ct) == eng.to_affine_ints(nifs.comm_T)
    a = eng.to_affine_ints(ct)
    ro.absorb_ints(*(a if a else (0, 0)))
    r = eng.scalars([ro.squeeze()])
    acc.step_end(r)
    Ua, Wa = acc.download()
    assert np.array_equal(Wa.W, W.W) and np.array_equal(Wa.E, W.E) and np.array_equal(Ua.u, U.u) and np.array_equal(Ua.X, U.X)
    assert eng.to_affine_ints(Ua.comm_W) == eng.to_affine_ints(U.comm_W)
    assert eng.to_affine_ints(Ua.comm_E) == eng.to_affine_ints(U.comm_E)
    acc.close(); shape.close(); ck.close()


@pytest.mark.parametrize("circuit", ["grayscale", "brightness", "resize", "sharpness", "hash"])
def test_full_size_step_properties(circuit, engines):
    """BASELINE sizes (grayscale HD: m = 130 864, n = 128 307; brightness/contrast 315 185 x 299 829; resize
    251 968 x 244 291; sharpness 335 734 x 320 377; the running-hash step 16 672 x 16 787 --
    /root/reference/circuits/nova_snark/circuit_parameters.csv:2-9 plus the augmented circuit): two folds on the
    device, checked through size-independent properties -- the folded instance satisfies the relaxed R1CS
    relation and both commitments open -- since the CPU oracle is too slow to redo this inside a unit test budget."""
    import torch
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    sh = S.synthetic_shape(CURVES["pallas"], circuit)
    shape = R1CSShape(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C)
    nck = 1 << (max(sh.num_cons, sh.num_vars) - 1).bit_length()
    d_bases = torch.empty(nck * 8, dtype=torch.int64, device="cuda")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, 77, 1234577, nck, d_bases.data_ptr()))
    ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), nck)
    del d_bases
    acc = FoldAccumulator(shape, ck)
    rng = random.Random(1)
    for k in range(2):
        Wi, Xi = S.synthetic_witness(sh, 500 + k)
        acc.step_begin(ints_to_mont(Wi, q), ints_to_mont(Xi, q))
        acc.step_end(ints_to_mont([rng.randrange(1 << 128)], q))
    U, W = acc.download()
    Az, Bz, Cz = shape.multiply_vec(np.concatenate([W.W, U.u, U.X]))
    # check the relation on the GPU's own field kernels + spot rows with python ints
    lhs = eng.field_op("scalar", "mul", Az, Bz)
    urep = np.repeat(U.u, sh.num_cons, axis=0)
    rhs = eng.field_op("scalar", "add", eng.field_op("scalar", "mul", urep, Cz), W.E)
    assert np.array_equal(lhs, rhs)
    idx = [0, 1, sh.nbits, sh.nbits + 5, sh.num_cons - 1] + [rng.randrange(sh.num_cons) for _ in range(50)]
    a, b, cz, e = (mont_to_ints(v[idx], q) for v in (Az, Bz, Cz, W.E))
    u = mont_to_ints(U.u, q)[0]
    assert all((x * y - u * w - t) % q == 0 for x, y, w, t in zip(a, b, cz, e))
    assert W.E.any()  # a real cross term was folded in
    assert eng.to_affine_ints(CommitmentEngine.commit(ck, W.W)) == eng.to_affine_ints(U.comm_W)
    assert eng.to_affine_ints(CommitmentEngine.commit(ck, W.E)) == eng.to_affine_ints(U.comm_E)
    acc.close(); shape.close(); ck.close()


def _mid_proof_instance(coracle, c, sh, Bm, seeds=(900, 901), rs=(130, 131)):
    """A relaxed instance with the magnitudes of the MIDDLE of an HD proof, built on the CPU without folding hundreds
    of steps: z1 = R*z_a + S*z_b for two satisfying fresh instances and R, S = sums of ~130 128-bit challenges
    (~135-bit numbers).  Folding is linear, so this is exactly what ~260 folds of those two rows give:
    u1 = R + S, X1 = R*X_a + S*X_b, E1 = R*S*T_ab with T_ab the cross term of a and b.  The instance satisfies the
    relaxed relation, W1's entries are ~136-bit values (the 0/1 wires) and E1 is full-width -- the regime bench.py times."""
    q, cid = c.q, c.curve_id
    rng = random.Random(5150)
    R = sum(rng.randrange(1 << 128) for _ in range(rs[0]))
    Sv = sum(rng.randrange(1 << 128) for _ in range(rs[1]))
    one = ints_to_mont([1], q)
    m, n, io = sh.num_cons, sh.num_vars, sh.num_io
    (Wa, Xa), (Wb, Xb) = [tuple(ints_to_mont(v, q) for v in S.synthetic_witness(sh, sd)) for sd in seeds]
    Rm, Sm, RSm = ints_to_mont([R], q), ints_to_mont([Sv], q), ints_to_mont([R * Sv % q], q)
    zero_n, zero_m = np.zeros((n, 4), np.uint64), np.zeros((m, 4), np.uint64)
    W1 = coracle.axpy(cid, coracle.axpy(cid, zero_n, Wa, Rm, 4), Wb, Sm, 4)
    X1 = coracle.axpy(cid, coracle.axpy(cid, np.zeros((io, 4), np.uint64), Xa, Rm), Xb, Sm)
    u1 = ints_to_mont([(R + Sv) % q], q)
    T_ab = coracle.commit_T(cid, m, n, io, sh.A, sh.B, sh.C, Wa, one, Xa, Wb, Xb, one, nthreads=4)
    E1 = coracle.axpy(cid, zero_m, T_ab, RSm, 4)
    cW = coracle.msm(cid, W1, Bm, 8)
    cE = coracle.msm(cid, E1, Bm, 8)
    return dict(W=W1, E=E1, u=u1, X=X1, cW=cW, cE=cE)


def test_full_size_grayscale_fold_bit_exact_vs_c_oracle(coracle):
    """THE configuration bench.py times, checked against the CPU restatement limb for limb: grayscale_step_HD
    (m = 130 864, n = 128 307, commitment key 2^17 points, window c = 15), starting from a loaded mid-proof relaxed
    instance (~136-bit running scalars, full-width E), two consecutive folds.  T then spills into the tenth 15-bit window
    and its top-digit entries form a giant bucket (asserted through the lane statistics), the second step uses the
    FOLDED cached products.  Compared with oracle/nova_cpu.c: T, comm_T, comm_W2 of both steps and the folded
    W / E / u / X / comm_W / comm_E -- the reference's own acceptance test is RecursiveSNARK::verify on these values
    (/root/reference/vimz/src/nova_snark_backend/folding.rs:46-56)."""
    import torch
    c = P.PALLAS
    q, cid = c.q, c.curve_id
    eng = vimz_b200.Engine("pallas", 0)
    sh = S.synthetic_shape(CURVES["pallas"], "grayscale")
    assert (sh.num_cons, sh.num_vars) == (130_864, 128_307)
    shape = R1CSShape(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C)
    nck = 1 << 17
    k0, dk = 77, 1234577
    g = affine_to_mont([P.generator(c)], c.p)[0]
    Bm = coracle.gen_bases(cid, g, k0, dk, nck)                               # CPU copy of the key ...
    d_bases = torch.empty(nck * 8, dtype=torch.int64, device="cuda")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, k0, dk, nck, d_bases.data_ptr()))
    assert np.array_equal(d_bases.cpu().numpy().view(np.uint64).reshape(-1, 8), Bm)   # ... identical to the GPU's
    ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), nck)
    del d_bases
    assert ck.window_bits == 15 and ck.num_windows == 17
    st = _mid_proof_instance(coracle, c, sh, Bm)
    acc = FoldAccumulator(shape, ck)
    acc.load(RelaxedR1CSInstance(st["cW"], st["cE"], st["X"], st["u"]), RelaxedR1CSWitness(st["W"], st["E"]))
    one = ints_to_mont([1], q)
    rng = random.Random(77)
    W1, E1, u1, X1, cW, cE = st["W"], st["E"], st["u"], st["X"], st["cW"], st["cE"]
    for k in range(2):
        Wi, Xi = S.synthetic_witness(sh, 910 + k)
        W2, X2 = ints_to_mont(Wi, q), ints_to_mont(Xi, q)
        comm_W2, comm_T = acc.step_begin(W2, X2)
        T = coracle.commit_T(cid, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, W1, u1, X1, W2, X2, one, nthreads=4)
        spill = sum(1 for t in mont_to_ints(T, q) if 135 <= min(t, q - t).bit_length() <= 150)
        assert spill > 10_000, "T must reach into the tenth 15-bit window like the mid-proof steps bench.py times"
        stats = eng.lane_stats()
        # (with the booleanity-row fold the rows whose fresh bit is 0 insert nothing; the others carry ten or more digits)
        assert stats["lane0_entries"] > 4 * sh.num_cons and stats["lane0_ngiant"] >= 1, stats
        assert np.array_equal(acc.last_T(), T), f"T differs at step {k}"
        exp_T, exp_W2 = coracle.msm(cid, T, Bm, 8), coracle.msm(cid, W2, Bm, 8)
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, exp_T), f"comm_T differs at step {k}"
        assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, exp_W2), f"comm_W2 differs at step {k}"
        r = ints_to_mont([rng.randrange(1 << 128)], q)
        acc.step_end(r)
        W1 = coracle.axpy(cid, W1, W2, r, 4); E1 = coracle.axpy(cid, E1, T, r, 4)
        tail = coracle.axpy(cid, np.concatenate([u1, X1]), np.concatenate([one, X2]), r)
        u1, X1 = tail[:1], tail[1:]
        cW = coracle.point_scale_add(cid, cW, r, exp_W2); cE = coracle.point_scale_add(cid, cE, r, exp_T)
    U, W = acc.download()
    assert np.array_equal(W.W, W1) and np.array_equal(W.E, E1) and np.array_equal(U.u, u1) and np.array_equal(U.X, X1)
    assert eng.to_affine_ints(U.comm_W) == _affine(coracle, c, cW) and eng.to_affine_ints(U.comm_E) == _affine(coracle, c, cE)
    vimz_b200.is_sat_relaxed(shape, ck, U, W)       # and the folded instance verifies (folding.rs:53-55)
    acc.close(); shape.close(); ck.close(); eng.close()


@pytest.mark.parametrize("fold_on", [1, 0])
def test_booleanity_row_fold_is_exact(fold_on, coracle):
    """The accumulator's K_S trick (option bitrow_fold, default on: commit T + [booleanity row] Az1 and subtract the running
    K_S = sum over those rows of (A z1)_i ck_i) must give the SAME comm_T, T and folded pairs as the plain path and as the CPU
    chain -- from a loaded mid-proof instance (K_S initialised by a general MSM), over several steps (K_S folded on the side
    stream), and for a witness that VIOLATES a booleanity constraint (a fresh wire that is neither 0 nor 1: the exact slow path
    of k_masked_base_sum), since commit_T is defined for any vectors."""
    c = P.PALLAS
    q = c.q
    eng = vimz_b200.Engine("pallas", 0)
    eng.set_option("bitrow_fold", fold_on)
    eng.set_option("msm_direct_max", 0)            # the fold lives on the bucket pipeline
    sh, shape, ck, Bm = _setup(eng, c, 0.02, seed=81)
    assert sh.nbits * 8 >= sh.num_cons              # most rows of this shape are booleanity rows
    rng = random.Random(13)
    wit = []
    for k in range(6):
        Wi, Xi = S.synthetic_witness(sh, 600 + k)
        if k == 3:                                   # break three booleanity constraints: wires that are not bits
            Wi[0], Wi[5], Wi[sh.nbits - 1] = 2, rng.randrange(q), q - 1
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(6)]
    chal[3] = ints_to_mont([rng.randrange(q)], q)    # a full-width challenge: K_S += r * P_S leaves the four 32-bit pieces for the plain chain
    chal[4] = ints_to_mont([(1 << 96) + 5], q)       # pieces that are zero
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)
    acc = FoldAccumulator(shape, ck)
    st = ref[1]                                      # start from the CPU's state after two folds
    acc.load(RelaxedR1CSInstance(st["cW"], st["cE"], st["X"], st["u"]), RelaxedR1CSWitness(st["W"], st["E"]))
    for k in range(2, 6):
        comm_W2, comm_T = acc.step_begin(*wit[k])
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[k]["comm_T"]), (fold_on, k)
        assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, ref[k]["comm_W2"])
        assert np.array_equal(acc.last_T(), ref[k]["T"])                 # the stored T is the true one, not the shifted one
        acc.step_end(chal[k])
    U, W = acc.download()
    assert np.array_equal(W.W, ref[5]["W"]) and np.array_equal(W.E, ref[5]["E"]) and np.array_equal(U.u, ref[5]["u"])
    assert eng.to_affine_ints(U.comm_E) == _affine(coracle, c, ref[5]["cE"]) and eng.to_affine_ints(U.comm_W) == _affine(coracle, c, ref[5]["cW"])
    # the fold must actually have been in use: with it, a step's commit(T) inserts far fewer digits
    entries = eng.lane_stats()["lane0_entries"]
    acc.reset()
    for k in range(2):
        comm_W2, comm_T = acc.step_begin(*wit[k])
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[k]["comm_T"])
        acc.step_end(chal[k])
    acc.close(); shape.close(); ck.close(); eng.close()
    test_booleanity_row_fold_is_exact.entries[fold_on] = entries
    if len(test_booleanity_row_fold_is_exact.entries) == 2:
        on, off = test_booleanity_row_fold_is_exact.entries[1], test_booleanity_row_fold_is_exact.entries[0]
        assert on < 0.8 * off, (on, off)


test_booleanity_row_fold_is_exact.entries = {}


@pytest.mark.parametrize("opts", [{}, {"stage_commit": 1}, {"acc_order": 1}, {"stage_commit": 1, "bitrow_fold": 0}])
def test_staged_async_and_resident_entry_points_match_cpu_chain(opts, coracle):
    """Every way a step's witness can reach the accumulator gives the CPU chain's commitments and folded pair: the plain host call,
    stage_fresh (prefix / suffix, host or device memory) + step_begin_staged -- with the early commitment of the staged range
    (option stage_commit) and without --, step_begin_async + step_wait, and step_begin_dev (whose copy is a patched graph node);
    graph replays included (each variant runs more than twice), with the accumulations ordered (acc_order) or not.  A staged range
    that is re-uploaded by step_begin_staged must fall back to a full commitment."""
    import torch
    c = P.PALLAS
    q = c.q
    eng = vimz_b200.Engine("pallas", 0)
    eng.set_option("msm_direct_max", 0)            # bucket pipeline: the paths these options shape
    for k, v in opts.items():
        eng.set_option(k, v)
    sh, shape, ck, Bm = _setup(eng, c, 0.02, seed=91)
    n = sh.num_vars
    rng = random.Random(17)
    nsteps = 14
    wit = []
    for k in range(nsteps):
        Wi, Xi = S.synthetic_witness(sh, 900 + k)
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(nsteps)]
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)
    dev = [torch.from_numpy(w.view(np.int64)).to("cuda:0") for w, _ in wit]
    acc = FoldAccumulator(shape, ck)
    split = n - 1100                                 # the staged prefix is long enough for an early commitment (>= 1024 rows)
    for k in range(nsteps):
        W2, X2 = wit[k]
        mode = k % 7
        if mode == 0:
            cw, ct = acc.step_begin(W2, X2)
        elif mode == 1:                              # prefix staged from the host, rest in the step
            acc.stage_fresh(W2, 0, split)
            cw, ct = acc.step_begin_staged(W2, split, n - split, X2)
        elif mode == 2:                              # the same from device memory
            acc.stage_fresh(dev[k].data_ptr(), 0, split)
            cw, ct = acc.step_begin_staged(dev[k].data_ptr(), split, n - split, X2)
        elif mode == 3:                              # suffix staged, prefix in the step
            acc.stage_fresh(W2, n - split, split)
            cw, ct = acc.step_begin_staged(W2, 0, n - split, X2)
        elif mode == 4:
            acc.step_begin_async(dev[k].data_ptr(), X2)
            cw, ct = acc.step_wait()
        elif mode == 5:
            cw, ct = acc.step_begin_dev(dev[k].data_ptr(), X2)
        else:                                        # a WRONG prefix staged first, then everything re-uploaded: the early commitment is void
            acc.stage_fresh(wit[(k + 1) % nsteps][0], 0, split)
            cw, ct = acc.step_begin_staged(W2, 0, n, X2)
        assert eng.to_affine_ints(cw) == _affine(coracle, c, ref[k]["comm_W2"]), (opts, k, mode)
        assert eng.to_affine_ints(ct) == _affine(coracle, c, ref[k]["comm_T"]), (opts, k, mode)
        acc.step_end(chal[k])
    U, W = acc.download()
    last = ref[-1]
    assert np.array_equal(W.W, last["W"]) and np.array_equal(W.E, last["E"]) and np.array_equal(U.u, last["u"]) and np.array_equal(U.X, last["X"])
    assert eng.to_affine_ints(U.comm_W) == _affine(coracle, c, last["cW"]) and eng.to_affine_ints(U.comm_E) == _affine(coracle, c, last["cE"])
    acc.close(); shape.close(); ck.close(); eng.close()


def test_fold_from_r1cs_and_wtns_files_bn254(tmp_path, engines, coracle):
    """Real-artifact ingestion: an iden3 .r1cs + two .wtns files (circom's bn128 prime) are read,
    uploaded and folded on the BN254 engine; result == CPU chain and the folded instance is satisfied."""
    from vimz_b200 import circom_io as io
    eng, c = engines["bn254"], P.BN254
    q = c.q
    cons = [([(2, 1)], [(2, 1)], [(3, 1)]),
            ([(3, 1)], [(2, 1)], [(4, 1)]),
            ([(4, 1), (2, 1), (0, 5)], [(0, 1)], [(1, 1)])]
    rp = str(tmp_path / "cubic.r1cs")
    io.write_r1cs(rp, q, 5, 1, 0, 1, cons)
    m, n, nio, A, B, Cm = io.r1cs_to_shape_coo(io.load_r1cs(rp))
    shape = R1CSShape(eng, m, n, nio, A, B, Cm)
    bases, _ = make_bases(c, 4, seed=1)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    acc = FoldAccumulator(shape, ck)
    wit = []
    for k, x in enumerate((3, 11)):
        wp = str(tmp_path / f"w{k}.wtns")
        io.write_wtns(wp, q, [1, x ** 3 + x + 5, x, x * x, x ** 3])
        _, vals = io.load_wtns(wp)
        wit.append(io.wtns_to_witness(vals, nio, q))
    rng = random.Random(4)
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in wit]

    class Sh:  # the oracle chain helper wants .num_cons/.A...
        num_cons, num_vars, num_io = m, n, nio
    Sh.A, Sh.B, Sh.C = A, B, Cm
    ref = _oracle_fold_chain(coracle, c, Sh, Bm, wit, chal)
    for (W2, X2), r in zip(wit, chal):
        acc.step_begin(W2, X2)
        acc.step_end(r)
    U, W = acc.download()
    assert np.array_equal(W.W, ref[-1]["W"]) and np.array_equal(W.E, ref[-1]["E"]) and np.array_equal(U.X, ref[-1]["X"])
    assert eng.to_affine_ints(U.comm_E) == _affine(coracle, c, ref[-1]["cE"])
    assert eng.to_affine_ints(U.comm_W) == _affine(coracle, c, ref[-1]["cW"])
    Az, Bz, Cz = shape.multiply_vec(np.concatenate([W.W, U.u, U.X]))
    a, b, cz, e = (mont_to_ints(v, q) for v in (Az, Bz, Cz, W.E))
    u = mont_to_ints(U.u, q)[0]
    assert all((x * y - u * w - t) % q == 0 for x, y, w, t in zip(a, b, cz, e))
    acc.close(); shape.close(); ck.close()


def test_is_sat_and_is_sat_relaxed(engines):
    """RecursiveSNARK::verify's checks (R1CSShape::is_sat / is_sat_relaxed) on fresh and folded instances, and
    their failure modes (tampered witness, error vector, commitment)."""
    from vimz_b200 import UnSat, is_sat, is_sat_relaxed
    eng, c = engines["vesta"], P.VESTA
    q = c.q
    sh, shape, ck, Bm = _setup(eng, c, 0.01, seed=77)
    Wi, Xi = S.synthetic_witness(sh, 3)
    W2 = R1CSWitness(ints_to_mont(Wi, q))
    U2 = R1CSInstance(W2.commit(ck), ints_to_mont(Xi, q))
    is_sat(shape, ck, U2, W2)                                     # a fresh satisfying instance
    bad = R1CSWitness(W2.W.copy()); bad.W[5] = ints_to_mont([Wi[5] + 1], q)[0]
    with pytest.raises(UnSat):
        is_sat(shape, ck, U2, bad)
    acc = FoldAccumulator(shape, ck)
    rng = random.Random(6)
    for k in range(2):
        Wk, Xk = S.synthetic_witness(sh, 10 + k)
        acc.step_begin(ints_to_mont(Wk, q), ints_to_mont(Xk, q))
        acc.step_end(ints_to_mont([rng.randrange(1 << 128)], q))
    U, W = acc.download()
    is_sat_relaxed(shape, ck, U, W)                               # the folded accumulator verifies
    E_bad = W.E.copy(); E_bad[0] = ints_to_mont([1], q)[0]
    with pytest.raises(UnSat):
        is_sat_relaxed(shape, ck, U, RelaxedR1CSWitness(W.W, E_bad))
    U_bad = RelaxedR1CSInstance(U.comm_E, U.comm_E, U.X, U.u)     # wrong comm_W
    with pytest.raises(UnSat):
        is_sat_relaxed(shape, ck, U_bad, W)
    with pytest.raises(vimz_b200.InvalidWitnessLength):
        is_sat(shape, ck, U2, R1CSWitness(W2.W[:-1]))
    acc.close(); shape.close(); ck.close()


@pytest.mark.parametrize("name,world", [("pallas", 3), ("bn254", 2)])
def test_row_sharded_fold_matches_unsharded_chain(name, world, engines, coracle):
    """the fold sharded by constraint-row range.  `world` FoldShards (here all on cuda:0, one per would-be
    rank) step through the same chain; the sums of their partial commitments, the concatenation of their E / T rows
    and the replicated W equal the CPU chain bit for bit at every step."""
    from vimz_b200.sharding import FoldShard
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    sh = S.synthetic_shape(CURVES[c.name], "grayscale", seed=44, scale=0.012)
    nck = max(sh.num_cons, sh.num_vars)
    bases, _ = make_bases(c, nck, seed=44)
    Bm = affine_to_mont(bases, c.p)
    rng = random.Random(8)
    wit = []
    for k in range(4):
        Wi, Xi = S.synthetic_witness(sh, 300 + k)
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(4)]
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)

    def make_ck(first, count):
        return CommitmentKey.from_bases(eng, Bm[first:first + count])

    shards = [FoldShard(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, make_ck, r, world) for r in range(world)]
    assert sum(s.m_local for s in shards) == sh.num_cons and sum(s.var_count for s in shards) == sh.num_vars
    for k in range(3):
        parts = [s.step_begin(*wit[k]) for s in shards]
        comm_W2 = eng.point_sum(np.stack([p[0] for p in parts]))
        comm_T = eng.point_sum(np.stack([p[1] for p in parts]))
        assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, ref[k]["comm_W2"])
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[k]["comm_T"])
        assert np.array_equal(np.concatenate([s.last_T() for s in shards]), ref[k]["T"])
        for s in shards:
            s.step_end(chal[k])
    # fourth step through the no-host-round-trip entry points: enqueue on every shard, gather the partial pairs on the
    # device (what the NCCL all-gather does across ranks), add them with vimz_acc_step_combine_dev
    import torch
    from vimz_b200.sharding import _DevWords
    d_W2 = torch.from_numpy(wit[3][0].view(np.int64)).cuda()
    gathered = torch.empty(world * 24, dtype=torch.int64, device="cuda")
    st = torch.cuda.ExternalStream(eng.stream)
    for r, s in enumerate(shards):
        part = torch.as_tensor(_DevWords(s.step_begin_dev_async(d_W2.data_ptr(), wit[3][1]), 24), device="cuda")
        with torch.cuda.stream(st):
            gathered[24 * r:24 * (r + 1)].copy_(part, non_blocking=True)
    comm_W2, comm_T = shards[0].step_combine_dev(gathered.data_ptr(), world)
    assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, ref[3]["comm_W2"])
    assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[3]["comm_T"])
    for s in shards:
        s.step_end(chal[3])
    outs = [s.download() for s in shards]
    last = ref[-1]
    for U, W in outs:
        assert np.array_equal(W.W, last["W"]) and np.array_equal(U.u, last["u"]) and np.array_equal(U.X, last["X"])
    assert np.array_equal(np.concatenate([W.E for _, W in outs]), last["E"])
    assert eng.to_affine_ints(eng.point_sum(np.stack([U.comm_W for U, _ in outs]))) == _affine(coracle, c, last["cW"])
    assert eng.to_affine_ints(eng.point_sum(np.stack([U.comm_E for U, _ in outs]))) == _affine(coracle, c, last["cE"])
    for s in shards:
        s.close()
    # range errors of vimz_acc_init_sharded
    import ctypes as C
    shape = R1CSShape(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C)
    small = CommitmentKey.from_bases(eng, Bm[:10])
    h = C.c_void_p()
    rc = vimz_b200.lib.vimz_acc_init_sharded(eng._h, shape._h, small._h, small._h, 0, 10, C.byref(h))
    assert rc == vimz_b200._lib.VIMZ_ERR_LENGTH          # ck_rows shorter than the rows
    rc = vimz_b200.lib.vimz_acc_init_sharded(eng._h, shape._h, small._h, small._h, sh.num_vars - 5, 10, C.byref(h))
    assert rc == vimz_b200._lib.VIMZ_ERR_LENGTH          # variable range leaves the witness
    shape.close(); small.close()


def test_library_level_sharded_step_and_msm_world1(engines, coracle):
    """The multi-GPU entry points with the exchange inside the library (vimz_comm_* / vimz_acc_step_begin_sharded* /
    vimz_msm_sharded_dev; NCCL bound with dlopen) on a one-rank communicator: same values as the plain entry points and as
    the CPU chain.  (N > 1 runs the same code under torchrun: bench.py's sharded_step / msm legs check it there.)"""
    import torch
    from vimz_b200.sharding import Comm, FoldShard, ShardedFoldAccumulator
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    assert vimz_b200.lib.vimz_comm_nccl_version() >= 20000
    comm = Comm(0, 0, 1, Comm.unique_id())
    assert (comm.rank, comm.world) == (0, 1)
    sh = S.synthetic_shape(CURVES["pallas"], "grayscale", seed=45, scale=0.012)
    bases, logs = make_bases(c, max(sh.num_cons, sh.num_vars), seed=45)
    Bm = affine_to_mont(bases, c.p)
    rng = random.Random(18)
    wit = []
    for k in range(3):
        Wi, Xi = S.synthetic_witness(sh, 400 + k)
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(3)]
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)
    shard = FoldShard(eng, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, lambda f, n: CommitmentKey.from_bases(eng, Bm[f:f + n]), 0, 1)
    acc = ShardedFoldAccumulator(None, shard, eng.point_sum, device="cuda", comm=comm)
    d_W = [torch.from_numpy(w.view(np.int64)).cuda() for w, _ in wit]
    for k in range(3):
        if k == 1:   # witness in host memory on the root: H2D + (one-rank) broadcast inside the call
            cw, ct = acc.step_begin_root(wit[k][0], wit[k][1], root=0)
        else:
            cw, ct = acc.step_begin_dev(d_W[k].data_ptr(), wit[k][1])
        assert eng.to_affine_ints(cw) == _affine(coracle, c, ref[k]["comm_W2"]) and eng.to_affine_ints(ct) == _affine(coracle, c, ref[k]["comm_T"])
        acc.step_end(chal[k])
    U, W = shard.download()
    assert np.array_equal(W.W, ref[-1]["W"]) and np.array_equal(W.E, ref[-1]["E"])
    assert eng.to_affine_ints(U.comm_E) == _affine(coracle, c, ref[-1]["cE"])
    with pytest.raises(vimz_b200.VimzError):
        shard.step_begin_sharded(comm, None, wit[0][1], root=0)        # the root must pass the witness
    # sharded commit through the same communicator
    n = len(bases)
    sc = [rng.randrange(q) for _ in range(n)]
    d_S = torch.from_numpy(ints_to_mont(sc, q).view(np.int64)).cuda()
    got = comm.commit_dev(shard.ck_rows, d_S.data_ptr(), sh.num_cons, 0)
    assert eng.to_affine_ints(got) == P.scalar_mul(c, sum(s_ * k_ for s_, k_ in zip(sc[:sh.num_cons], logs)) % q, P.generator(c))
    shard.close(); comm.close()


def test_commit_T_on_shape_without_rows(engines):
    """num_cons = 0 (what a row shard beyond the last constraint looks like): T is empty and comm_T the identity, whatever
    an earlier commit left in the workspace."""
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    bases, _ = make_bases(c, 40, seed=9)
    ck = CommitmentKey.from_bases(eng, affine_to_mont(bases, c.p))
    CommitmentEngine.commit(ck, ints_to_mont(list(range(1, 41)), q))           # leaves a histogram / digit array behind
    e = (np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 4), np.uint64))
    shape = R1CSShape(eng, 0, 3, 1, e, e, e)
    W = RelaxedR1CSWitness(ints_to_mont([5, 6, 7], q), fr_zero(0))
    U = RelaxedR1CSInstance(np.zeros(12, np.uint64), np.zeros(12, np.uint64), ints_to_mont([3], q), ints_to_mont([2], q))
    T, comm_T = shape.commit_T(ck, U, W, R1CSInstance(np.zeros(12, np.uint64), ints_to_mont([4], q)), R1CSWitness(ints_to_mont([1, 0, 1], q)))
    assert T.shape == (0, 4) and eng.to_affine_ints(comm_T) is None
    shape.close(); ck.close()


def fr_zero(n):
    return np.zeros((n, 4), np.uint64)


def test_row_sharded_fold_with_empty_shards(engines, coracle):
    """More ranks than constraint rows / variables: shards with zero rows, zero variables or both still step, and the
    sums over all shards equal the CPU fold."""
    from vimz_b200.sharding import FoldShard
    eng, c = engines["pallas"], P.PALLAS
    q = c.q
    m, n, io, world = 5, 3, 2, 7
    rng = random.Random(77)
    A, B, Cm = (rand_coo(rng, m, n + 1 + io, k, q) for k in (9, 7, 4))
    bases, _ = make_bases(c, max(m, n), seed=5)
    Bm = affine_to_mont(bases, c.p)

    class Sh:  # what _oracle_fold_chain reads
        num_cons, num_vars, num_io = m, n, io
    Sh.A, Sh.B, Sh.C = A, B, Cm
    wit = [(ints_to_mont([rng.randrange(q) for _ in range(n)], q), ints_to_mont([rng.randrange(q) for _ in range(io)], q)) for _ in range(2)]
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(2)]
    ref = _oracle_fold_chain(coracle, c, Sh, Bm, wit, chal)

    def make_ck(first, count):
        return CommitmentKey.from_bases(eng, Bm[first:first + count])

    shards = [FoldShard(eng, m, n, io, A, B, Cm, make_ck, r, world) for r in range(world)]
    assert sum(1 for s in shards if s.m_local == 0) >= 2 and sum(1 for s in shards if s.var_count == 0) >= 4
    for k in range(2):
        parts = [s.step_begin(*wit[k]) for s in shards]
        assert eng.to_affine_ints(eng.point_sum(np.stack([p[0] for p in parts]))) == _affine(coracle, c, ref[k]["comm_W2"])
        assert eng.to_affine_ints(eng.point_sum(np.stack([p[1] for p in parts]))) == _affine(coracle, c, ref[k]["comm_T"])
        for s in shards:
            s.step_end(chal[k])
    outs = [s.download() for s in shards]
    assert np.array_equal(np.concatenate([W.E for _, W in outs]), ref[-1]["E"])
    assert all(np.array_equal(W.W, ref[-1]["W"]) for _, W in outs)
    assert eng.to_affine_ints(eng.point_sum(np.stack([U.comm_E for U, _ in outs]))) == _affine(coracle, c, ref[-1]["cE"])
    assert eng.to_affine_ints(eng.point_sum(np.stack([U.comm_W for U, _ in outs]))) == _affine(coracle, c, ref[-1]["cW"])
    for s in shards:
        s.close()


@pytest.mark.parametrize("cache", [0, 1])
def test_accumulator_cached_products_vs_recomputed(cache, coracle):
    """The resident accumulator keeps (Az1, Bz1, Cz1) and folds them (A(z1 + r z2) = Az1 + r Az2) instead of recomputing
    them every step (option cross_cache, default on).  Both variants, started from a LOADED mid-proof instance (whose
    products are computed once by multiply_vec), must reproduce the CPU chain bit for bit; so must the row-class kernel
    (cross_stream = 0), which fills the cache with an explicit mat-vec."""
    c = P.PALLAS
    q = c.q
    eng = vimz_b200.Engine("pallas", 0)
    eng.set_option("cross_cache", cache)
    sh, shape, ck, Bm = _setup(eng, c, 0.012, seed=71)
    rng = random.Random(3)
    wit = []
    for k in range(5):
        Wi, Xi = S.synthetic_witness(sh, 500 + k)
        wit.append((ints_to_mont(Wi, q), ints_to_mont(Xi, q)))
    chal = [ints_to_mont([rng.randrange(1 << 128)], q) for _ in range(5)]
    ref = _oracle_fold_chain(coracle, c, sh, Bm, wit, chal)
    # start from the CPU's state after two folds
    acc = FoldAccumulator(shape, ck)
    st = ref[1]
    acc.load(RelaxedR1CSInstance(st["cW"], st["cE"], st["X"], st["u"]), RelaxedR1CSWitness(st["W"], st["E"]))
    for k in range(2, 5):
        if k == 4:
            eng.set_option("cross_stream", 0)      # row-class kernel for the last step
        comm_W2, comm_T = acc.step_begin(*wit[k])
        assert eng.to_affine_ints(comm_T) == _affine(coracle, c, ref[k]["comm_T"]), (cache, k)
        assert eng.to_affine_ints(comm_W2) == _affine(coracle, c, ref[k]["comm_W2"])
        assert np.array_equal(acc.last_T(), ref[k]["T"])
        acc.step_end(chal[k])
    U, W = acc.download()
    assert np.array_equal(W.W, ref[4]["W"]) and np.array_equal(W.E, ref[4]["E"]) and np.array_equal(U.u, ref[4]["u"])
    assert eng.to_affine_ints(U.comm_E) == _affine(coracle, c, ref[4]["cE"])
    vimz_b200.is_sat_relaxed(shape, ck, U, W)
    acc.close(); shape.close(); ck.close(); eng.close()


@pytest.mark.parametrize("name", list(P.CURVES))
def test_fold_chain_golden_fixture(name, engines):
    """The committed fold-chain vectors (tests/golden/fold_chain.json, python big-int model): the resident accumulator
    reproduces T, both fresh commitments and the folded instance / witness of every step, on all four curves."""
    from test_oracle import golden_coo, golden_pt, load_fold_golden
    g = load_fold_golden()[name]
    eng, c = engines[name], P.CURVES[name]
    q = c.q
    m, n, io = g["num_cons"], g["num_vars"], g["num_io"]
    A, B, Cm = (golden_coo(g[k], q) for k in "ABC")
    ck = CommitmentKey.from_bases(eng, affine_to_mont([golden_pt(b) for b in g["ck"]], c.p))
    shape = R1CSShape(eng, m, n, io, A, B, Cm)
    acc = FoldAccumulator(shape, ck)
    ints = lambda hs: [int(h, 16) for h in hs]
    for st in g["steps"]:
        comm_W2, comm_T = acc.step_begin(ints_to_mont(ints(st["W2"]), q), ints_to_mont(ints(st["X2"]), q))
        assert eng.to_affine_ints(comm_W2) == golden_pt(st["comm_W2"]) and eng.to_affine_ints(comm_T) == golden_pt(st["comm_T"])
        assert mont_to_ints(acc.last_T(), q) == ints(st["T"])
        acc.step_end(ints_to_mont([int(st["r"], 16)], q))
        U, W = acc.download()
        assert mont_to_ints(W.W, q) == ints(st["W"]) and mont_to_ints(W.E, q) == ints(st["E"])
        assert mont_to_ints(U.u, q) == [int(st["u"], 16)] and mont_to_ints(U.X, q) == ints(st["X"])
        assert eng.to_affine_ints(U.comm_W) == golden_pt(st["comm_W"]) and eng.to_affine_ints(U.comm_E) == golden_pt(st["comm_E"])
    acc.close(); shape.close(); ck.close()
