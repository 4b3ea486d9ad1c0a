"""CPU-only: pins the oracle itself (oracle/pyref.py and oracle/nova_cpu.c).

The reference holds no golden vector for the fold path (parity unpinned at the
nova-snark boundary), so the oracle is pinned by (a) number-theoretic facts about the four curves,
(b) agreement of two independent restatements (Python big-int vs C 4x64 Montgomery), (c) the
definition-level identities the domain offers, and (d) the committed fixtures in tests/golden/."""
import json
import os
import random

import numpy as np
import pytest

from oracle import pyref as P
from vimz_b200.field import affine_to_mont, ints_to_mont, mont_to_affine, mont_to_ints
from conftest import make_bases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def is_probable_prime(n, rounds=16):
    if n < 4:
        return n in (2, 3)
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rng = random.Random(7)
    for _ in range(rounds):
        a = rng.randrange(2, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@pytest.mark.parametrize("name", list(P.CURVES))
def test_curve_constants(name):
    c = P.CURVES[name]
    assert is_probable_prime(c.p) and is_probable_prime(c.q)
    G = P.generator(c)
    assert P.on_curve(c, G)
    assert P.scalar_mul(c, c.q - 1, G) == P.aff_neg(c, G)  # [q]G = O  <=>  [q-1]G = -G
    assert P.aff_add(c, P.scalar_mul(c, c.q - 1, G), G) is None


def test_cycle_structure():
    assert P.PALLAS.q == P.VESTA.p and P.VESTA.q == P.PALLAS.p
    assert P.BN254.q == P.GRUMPKIN.p and P.GRUMPKIN.q == P.BN254.p


@pytest.mark.parametrize("name", list(P.CURVES))
def test_c_field_ops_match_bigint(name, coracle):
    c = P.CURVES[name]
    rng = random.Random(11)
    for which, mod in ((0, c.p), (1, c.q)):
        a = [rng.randrange(mod) for _ in range(300)] + [0, 1, mod - 1, mod - 1, 0, 1]
        b = [rng.randrange(mod) for _ in range(300)] + [0, mod - 1, mod - 1, 1, 5, 1]
        A, B = ints_to_mont(a, mod), ints_to_mont(b, mod)
        assert mont_to_ints(coracle.field_op(c.curve_id, which, 0, A, B), mod) == [x * y % mod for x, y in zip(a, b)]
        assert mont_to_ints(coracle.field_op(c.curve_id, which, 1, A, B), mod) == [(x + y) % mod for x, y in zip(a, b)]
        assert mont_to_ints(coracle.field_op(c.curve_id, which, 2, A, B), mod) == [(x - y) % mod for x, y in zip(a, b)]


@pytest.mark.parametrize("name", list(P.CURVES))
@pytest.mark.parametrize("n", [1, 3, 31, 33, 257])
def test_msm_three_ways(name, n, coracle):
    """definition (naive) == restated cpu_best_multiexp (python) == C oracle == closed form via discrete logs."""
    c = P.CURVES[name]
    rng = random.Random(n * 31 + c.curve_id)
    bases, logs = make_bases(c, n, seed=n + c.curve_id)
    sc = [rng.randrange(c.q) for _ in range(n)]
    if n >= 3:
        sc[0], sc[1], sc[2] = 0, 1, c.q - 1
    expect = P.scalar_mul(c, sum(s * k for s, k in zip(sc, logs)) % c.q, P.generator(c))
    if n <= 33:
        assert P.msm_naive(c, sc, bases) == expect
    assert P.cpu_best_multiexp(c, sc, bases, num_threads=4) == expect
    Bm, Sm = affine_to_mont(bases, c.p), ints_to_mont(sc, c.q)
    for nt in (1, 4):
        got = mont_to_affine(coracle.to_affine(c.curve_id, coracle.msm(c.curve_id, Sm, Bm, nt)), c.p)[0]
        assert got == expect


def test_msm_identity_and_cancellation(coracle):
    c = P.PALLAS
    G = P.generator(c)
    bases = [G, P.aff_neg(c, G), None, G]
    Bm = affine_to_mont(bases, c.p)
    # s*G + s*(-G) + 7*O + 0*G = O
    Sm = ints_to_mont([5, 5, 7, 0], c.q)
    j = coracle.msm(c.curve_id, Sm, Bm, 1)
    assert mont_to_affine(coracle.to_affine(c.curve_id, j), c.p)[0] is None
    # doubling inside a bucket: G + G
    Sm = ints_to_mont([1, 0, 0, 1], c.q)
    j = coracle.msm(c.curve_id, Sm, Bm, 1)
    assert mont_to_affine(coracle.to_affine(c.curve_id, j), c.p)[0] == P.scalar_mul(c, 2, G)


@pytest.mark.parametrize("cbits,nwin", [(8, 32), (13, 20), (16, 16), (17, 15)])
def test_signed_digits(cbits, nwin):
    rng = random.Random(cbits)
    q = P.PALLAS.q
    for s in [0, 1, q - 1, (1 << 254), (1 << 254) - 1, (1 << (cbits - 1)), (1 << (cbits - 1)) + 1] + [rng.randrange(q) for _ in range(200)]:
        d = P.signed_digits(s, cbits, nwin)
        assert all(abs(x) <= (1 << (cbits - 1)) for x in d)


def _tiny_shape():
    """x*x = y ; y*x = z ; (z + x + 5)*1 = out   over z = (W=[x,y,z], u, X=[out])  -- the classic cubic."""
    # columns: 0:x 1:y 2:z 3:u(one) 4:out
    A = [(0, 0, 1), (1, 1, 1), (2, 2, 1), (2, 0, 1), (2, 3, 5)]
    B = [(0, 0, 1), (1, 0, 1), (2, 3, 1)]
    C = [(0, 1, 1), (1, 2, 1), (2, 4, 1)]
    return P.R1CSShape(3, 3, 1, A, B, C)


def test_r1cs_definitions_by_hand():
    q = P.PALLAS.q
    S = _tiny_shape()
    x = 3
    W, X = [x, x * x, x ** 3], [x ** 3 + x + 5]
    Az, Bz, Cz = S.multiply_vec(q, W + [1] + X)
    assert Az == [3, 9, 35] and Bz == [3, 3, 1] and Cz == [9, 27, 35]
    assert S.is_sat_relaxed(q, W, [0, 0, 0], 1, X)
    with pytest.raises(ValueError):
        S.multiply_vec(q, W + [1])
    # fold two satisfying instances: the folded relaxed instance must satisfy Az o Bz = u Cz + E
    x2 = 4
    W2, X2 = [x2, x2 * x2, x2 ** 3], [x2 ** 3 + x2 + 5]
    T = S.cross_term(q, W, 1, X, W2, X2)
    # by hand for row 0: Az1*Bz2 + Az2*Bz1 - u1*Cz2 - Cz1 = 3*4 + 4*3 - 16 - 9 = -1
    assert T[0] == (12 + 12 - 16 - 9) % q
    r = 0x1234567890ABCDEF
    Wf, Ef = P.fold_witness(q, W, [0, 0, 0], W2, T, r)
    Xf = [(a + r * b) % q for a, b in zip(X, X2)]
    assert S.is_sat_relaxed(q, Wf, Ef, (1 + r) % q, Xf)


def test_c_r1cs_matches_python(coracle):
    c = P.VESTA
    q = c.q
    rng = random.Random(5)
    m, n, io = 40, 30, 2
    def rand_mat(nnz):
        return [(rng.randrange(m), rng.randrange(n + 1 + io), rng.choice([1, q - 1, 2, rng.randrange(q)])) for _ in range(nnz)]
    A, B, Cm = rand_mat(90), rand_mat(70), rand_mat(50)
    S = P.R1CSShape(m, n, io, A, B, Cm)
    W1 = [rng.randrange(q) for _ in range(n)]
    W2 = [rng.randrange(2) for _ in range(n)]
    X1 = [rng.randrange(q) for _ in range(io)]
    X2 = [rng.randrange(q) for _ in range(io)]
    u1 = rng.randrange(q)
    def pack(M):
        return (np.array([e[0] for e in M], np.uint32), np.array([e[1] for e in M], np.uint32), ints_to_mont([e[2] for e in M], q))
    z1 = W1 + [u1] + X1
    got = coracle.multiply_vec(c.curve_id, m, n, io, pack(A), pack(B), pack(Cm), ints_to_mont(z1, q))
    exp = S.multiply_vec(q, z1)
    for g, e in zip(got, exp):
        assert mont_to_ints(g, q) == e
    T = coracle.commit_T(c.curve_id, m, n, io, pack(A), pack(B), pack(Cm), ints_to_mont(W1, q), ints_to_mont([u1], q),
                         ints_to_mont(X1, q), ints_to_mont(W2, q), ints_to_mont(X2, q), ints_to_mont([1], q), nthreads=3)
    assert mont_to_ints(T, q) == S.cross_term(q, W1, u1, X1, W2, X2)
    r = rng.randrange(1 << 128)
    got = coracle.axpy(c.curve_id, ints_to_mont(W1, q), ints_to_mont(W2, q), ints_to_mont([r], q), nthreads=2)
    assert mont_to_ints(got, q) == [(a + r * b) % q for a, b in zip(W1, W2)]
    with pytest.raises(ValueError):
        coracle.multiply_vec(c.curve_id, m, n, io, pack(A), pack(B), pack(Cm), ints_to_mont(z1[:-1], q))


def test_golden_fixtures(coracle):
    """tests/golden/msm_vectors.json was produced by tests/golden/make_golden.py from the python big-int
    model; the C oracle must reproduce every vector."""
    with open(os.path.join(GOLDEN, "msm_vectors.json")) as f:
        data = json.load(f)
    assert data["vectors"]
    for v in data["vectors"]:
        c = P.CURVES[v["curve"]]
        bases = [None if b is None else (int(b[0], 16), int(b[1], 16)) for b in v["bases"]]
        sc = [int(s, 16) for s in v["scalars"]]
        expect = None if v["result"] is None else (int(v["result"][0], 16), int(v["result"][1], 16))
        j = coracle.msm(c.curve_id, ints_to_mont(sc, c.q), affine_to_mont(bases, c.p), 2)
        assert mont_to_affine(coracle.to_affine(c.curve_id, j), c.p)[0] == expect


# Published known answers for alt_bn128 / BN254 G1 (EIP-196 test vectors, also py_ecc / go-ethereum bn256 tests):
BN254_2G = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
            9918110051302171585080402603319702774565515993150576347155970296011118125764)
BN254_3G = (3353031288059533942658390886683067124040920775575537747144343083137631628272,
            19321533766552368860946552437480515441416830039777911637913418824951667761761)


def test_bn254_published_known_answers(coracle):
    """External pin of the group law (not derived from this repository's code): 2G and 3G on BN254 G1."""
    c = P.BN254
    G = P.generator(c)
    assert P.scalar_mul(c, 2, G) == BN254_2G and P.scalar_mul(c, 3, G) == BN254_3G
    assert P.aff_add(c, G, G) == BN254_2G and P.aff_add(c, BN254_2G, G) == BN254_3G
    Bm = affine_to_mont([G, G], c.p)
    for sc, exp in (([2, 0], BN254_2G), ([1, 1], BN254_2G), ([2, 1], BN254_3G)):
        j = coracle.msm(c.curve_id, ints_to_mont(sc, c.q), Bm, 1)
        assert mont_to_affine(coracle.to_affine(c.curve_id, j), c.p)[0] == exp


def load_fold_golden():
    with open(os.path.join(GOLDEN, "fold_chain.json")) as f:
        return json.load(f)["curves"]


def golden_coo(entries, q):
    rows = np.array([e[0] for e in entries], np.uint32)
    cols = np.array([e[1] for e in entries], np.uint32)
    vals = ints_to_mont([int(e[2], 16) for e in entries], q) if entries else np.zeros((0, 4), np.uint64)
    return rows, cols, vals


def golden_pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


@pytest.mark.parametrize("name", list(P.CURVES))
def test_fold_chain_golden_c_oracle(name, coracle):
    """tests/golden/fold_chain.json (python big-int model, tests/golden/make_fold_golden.py): the C restatement reproduces
    every intermediate of three consecutive NIFS folds -- T, both fresh commitments, the folded witness and instance."""
    g = load_fold_golden()[name]
    c = P.CURVES[name]
    q, cid = c.q, c.curve_id
    m, n, io = g["num_cons"], g["num_vars"], g["num_io"]
    A, B, Cm = (golden_coo(g[k], q) for k in "ABC")
    Bm = affine_to_mont([golden_pt(b) for b in g["ck"]], c.p)
    one = ints_to_mont([1], q)
    aff = lambda j: mont_to_affine(coracle.to_affine(cid, j), c.p)[0]
    ints = lambda hs: [int(h, 16) for h in hs]
    W1 = np.zeros((n, 4), np.uint64); E1 = np.zeros((m, 4), np.uint64)
    u1 = np.zeros((1, 4), np.uint64); X1 = np.zeros((io, 4), np.uint64)
    cW = np.zeros(12, np.uint64); cE = np.zeros(12, np.uint64)
    for st in g["steps"]:
        W2, X2, r = ints_to_mont(ints(st["W2"]), q), ints_to_mont(ints(st["X2"]), q), ints_to_mont([int(st["r"], 16)], q)
        comm_W2 = coracle.msm(cid, W2, Bm[:n], 1)
        T = coracle.commit_T(cid, m, n, io, A, B, Cm, W1, u1, X1, W2, X2, one)
        comm_T = coracle.msm(cid, T, Bm[:m], 1)
        assert mont_to_ints(T, q) == ints(st["T"])
        assert aff(comm_W2) == golden_pt(st["comm_W2"]) and aff(comm_T) == golden_pt(st["comm_T"])
        W1 = coracle.axpy(cid, W1, W2, r); E1 = coracle.axpy(cid, E1, T, r)
        tail = coracle.axpy(cid, np.concatenate([u1, X1]), np.concatenate([one, X2]), r)
        u1, X1 = tail[:1], tail[1:]
        cW = coracle.point_scale_add(cid, cW, r, comm_W2); cE = coracle.point_scale_add(cid, cE, r, comm_T)
        assert mont_to_ints(W1, q) == ints(st["W"]) and mont_to_ints(E1, q) == ints(st["E"])
        assert mont_to_ints(u1, q) == [int(st["u"], 16)] and mont_to_ints(X1, q) == ints(st["X"])
        assert aff(cW) == golden_pt(st["comm_W"]) and aff(cE) == golden_pt(st["comm_E"])


@pytest.mark.parametrize("circuit", ["grayscale", "blur4k", "sharpness4k", "hash"])
def test_synthetic_step_shapes_are_satisfied(circuit, coracle):
    """The synthetic step circuits bench.py folds (published sizes; the 4K ones are the x3-width estimates): a scaled-down
    instance of each is satisfied by its generated witness (A z o B z = C z with u = 1), mostly 0/1 for the pixel circuits."""
    from vimz_b200 import synthetic as S
    from vimz_b200.field import CURVES
    cv = CURVES["pallas"]
    q = cv.scalar_modulus
    sh = S.synthetic_shape(cv, circuit, seed=5, scale=0.01 if circuit != "hash" else 0.2)
    Wi, Xi = S.synthetic_witness(sh, 9)
    assert len(Wi) == sh.num_vars and len(Xi) == sh.num_io
    z = ints_to_mont(list(Wi) + [1] + list(Xi), q)
    Az, Bz, Cz = coracle.multiply_vec(cv.curve_id, sh.num_cons, sh.num_vars, sh.num_io, sh.A, sh.B, sh.C, z)
    a, b, c_ = (mont_to_ints(v, q) for v in (Az, Bz, Cz))
    assert all((x * y - w) % q == 0 for x, y, w in zip(a, b, c_))
    if circuit != "hash":
        assert sum(1 for w in Wi if w < 2) / len(Wi) > 0.6
    full = S.STEP_CIRCUITS[circuit]
    assert sh.num_cons == int((full[0] + S.NOVA_AUGMENTED) * (0.01 if circuit != "hash" else 0.2))


# ---- implementation-independent pin of the Pasta (and BN254 / Grumpkin) group law: the GLV endomorphism ------------------
# All four curves have j-invariant 0 (y^2 = x^3 + b), so phi(x, y) = (zeta * x, y) with zeta a primitive cube root of unity of
# the BASE field is an endomorphism and acts as multiplication by a primitive cube root of unity lambda of the SCALAR field:
# phi(P) = [lambda] P.  Hence for ANY correct commit(): commit(ck, lambda * v) has the same y as commit(ck, v) and x multiplied
# by a non-trivial cube root of unity -- a statement about field elements that needs no second implementation of the group law
# (a wrong scalar multiplication lands on such a point with probability ~2^-254).  zeta and lambda come from modular
# exponentiation alone.  The Pasta curves have no published vectors that could be restated here; this identity
# pins the metric's curves independently of oracle/pyref.py's own addition formulas.
def cube_roots_of_unity(modulus):
    g = 2
    while pow(g, (modulus - 1) // 3, modulus) == 1:
        g += 1
    z = pow(g, (modulus - 1) // 3, modulus)
    assert z != 1 and pow(z, 3, modulus) == 1
    return z, z * z % modulus


def check_endomorphism(c, commit_affine, scalars):
    """commit_affine(list of canonical scalars) -> affine (x, y) or None.  Returns the zeta the implementation pairs with
    lambda_1 (so callers can check that it is the same for every input)."""
    lam1, lam2 = cube_roots_of_unity(c.q)
    zetas = cube_roots_of_unity(c.p)
    base = commit_affine(scalars)
    assert base is not None
    seen = []
    for lam in (lam1, lam2):
        pt = commit_affine([s * lam % c.q for s in scalars])
        assert pt is not None and pt[1] == base[1], "phi keeps y"
        ratio = pt[0] * pow(base[0], -1, c.p) % c.p
        assert ratio in zetas, "x must be multiplied by a primitive cube root of unity"
        seen.append(ratio)
    assert seen[0] != seen[1] and seen[0] * seen[1] % c.p == 1      # lambda^2 <-> zeta^2
    return seen[0]


@pytest.mark.parametrize("name", list(P.CURVES))
def test_endomorphism_identity_pins_the_oracles(name, coracle):
    c = P.CURVES[name]
    rng = random.Random(31 + c.curve_id)
    G = P.generator(c)
    # python oracle: single point
    z0 = check_endomorphism(c, lambda sc: P.scalar_mul(c, sc[0], G), [rng.randrange(1, c.q)])
    # C oracle: a 200-point MSM, same pairing of lambda and zeta
    n = 200
    bases = [P.scalar_mul(c, rng.randrange(1, c.q), G) for _ in range(8)]
    bases = [bases[i % 8] if i < 8 else P.aff_add(c, bases[i % 8], bases[(i * 5 + 1) % 8]) for i in range(n)]
    Bm = affine_to_mont(bases, c.p)

    def c_commit(sc):
        return mont_to_affine(coracle.to_affine(c.curve_id, coracle.msm(c.curve_id, ints_to_mont(sc, c.q), Bm, 2)), c.p)[0]

    z1 = check_endomorphism(c, c_commit, [rng.randrange(c.q) for _ in range(n)])
    assert z0 == z1
