"""GPU parity: CommitmentEngine::commit (Pippenger over the resident window table) vs the oracle.

Bit-exact bar: the affine coordinates of the commitment must equal the oracle's, for every curve,
for the sizes the oracle finishes in seconds, for the committed golden vectors, and -- at large n --
through closed forms (bases with known discrete logs) and linearity."""
import json
import os
import random

import numpy as np
import pytest

import vimz_b200
from vimz_b200 import CommitmentEngine, CommitmentKey
from vimz_b200.field import affine_to_mont, ints_to_mont, mont_to_affine
from oracle import pyref as P
from conftest import make_bases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def gpu_commit_affine(eng, ck, scalars_mont):
    return eng.to_affine_ints(CommitmentEngine.commit(ck, scalars_mont))


def oracle_commit_affine(coracle, c, scalars_mont, bases_mont, nt=4):
    return mont_to_affine(coracle.to_affine(c.curve_id, coracle.msm(c.curve_id, scalars_mont, bases_mont, nt)), c.p)[0]


@pytest.mark.parametrize("name", list(P.CURVES))
def test_golden_vectors(name, engines):
    eng = engines[name]
    c = P.CURVES[name]
    with open(os.path.join(GOLDEN, "msm_vectors.json")) as f:
        vecs = [v for v in json.load(f)["vectors"] if v["curve"] == name]
    assert vecs
    for v in vecs:
        bases = [None if b is None else (int(b[0], 16), int(b[1], 16)) for b in v["bases"]]
        sc = [int(s, 16) for s in v["scalars"]]
        expect = None if v["result"] is None else (int(v["result"][0], 16), int(v["result"][1], 16))
        ck = CommitmentKey.from_bases(eng, affine_to_mont(bases, c.p))
        assert gpu_commit_affine(eng, ck, ints_to_mont(sc, c.q)) == expect
        ck.close()


@pytest.mark.parametrize("name", list(P.CURVES))
@pytest.mark.parametrize("n", [1, 2, 127, 128, 1000])
def test_msm_vs_oracle_small(name, n, engines, coracle):
    eng, c = engines[name], P.CURVES[name]
    rng = random.Random(n + 17 * c.curve_id)
    bases, logs = make_bases(c, n, seed=n)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    for dist in ("uniform", "bits", "edge"):
        if dist == "uniform":
            sc = [rng.randrange(c.q) for _ in range(n)]
        elif dist == "bits":  # witness-like: ~93% 0/1, some bytes, some full width
            sc = [rng.randrange(2) if rng.random() < 0.93 else (rng.randrange(256) if rng.random() < 0.3 else rng.randrange(c.q)) for _ in range(n)]
        else:
            pool = [0, 1, c.q - 1, c.q - 2, 1 << 254, (1 << 254) - 1, 2, 1 << 127, (1 << 128) - 1]
            sc = [pool[i % len(pool)] % c.q for i in range(n)]
        Sm = ints_to_mont(sc, c.q)
        got = gpu_commit_affine(eng, ck, Sm)
        assert got == oracle_commit_affine(coracle, c, Sm, Bm), (name, n, dist)
        # closed form through the discrete logs (independent of both MSM implementations)
        assert got == P.scalar_mul(c, sum(s * k for s, k in zip(sc, logs)) % c.q, P.generator(c))
    ck.close()


@pytest.mark.parametrize("window", [8, 11, 16, 17, 20, 22])
def test_msm_windows_agree(window, engines, coracle):
    """Every window size must give the same group element (prefix commit too).  Windows above 16 bits have more than
    32 768 buckets: the multi-block scan (k_scan_blocksum / top / apply) and the deeper reduction geometry that the
    2^20 .. 2^24-point MSMs of bench.py run."""
    c = P.PALLAS
    eng = vimz_b200.Engine("pallas", 0)
    eng.set_option("msm_window", window)
    n = 600
    rng = random.Random(window)
    bases, _ = make_bases(c, n, seed=99)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    assert ck.window_bits == window
    for m in (n, 333, 1, 0):
        Sm = ints_to_mont([rng.randrange(c.q) for _ in range(m)], c.q)
        got = gpu_commit_affine(eng, ck, Sm)
        exp = oracle_commit_affine(coracle, c, Sm, Bm) if m else None
        assert got == exp
    ck.close()
    eng.close()


def test_msm_empty_identity_and_repeats(engines):
    c = P.PALLAS
    eng = engines["pallas"]
    G = P.generator(c)
    bases = [G, P.aff_neg(c, G), None, G, G, G]
    ck = CommitmentKey.from_bases(eng, affine_to_mont(bases, c.p))
    assert gpu_commit_affine(eng, ck, ints_to_mont([], c.q)) is None                     # empty vector
    assert gpu_commit_affine(eng, ck, ints_to_mont([0] * 6, c.q)) is None                # all-zero scalars
    assert gpu_commit_affine(eng, ck, ints_to_mont([5, 5, 7, 0, 0, 0], c.q)) is None     # P + (-P), identity base
    assert gpu_commit_affine(eng, ck, ints_to_mont([1, 0, 0, 1, 0, 0], c.q)) == P.scalar_mul(c, 2, G)   # P + P in one bucket
    assert gpu_commit_affine(eng, ck, ints_to_mont([1, 0, 0, 1, 1, 1], c.q)) == P.scalar_mul(c, 4, G)
    assert gpu_commit_affine(eng, ck, ints_to_mont([c.q - 1] * 1, c.q)) == P.aff_neg(c, G)
    with pytest.raises(vimz_b200.InvalidWitnessLength):                                  # ck shorter than v
        CommitmentEngine.commit(ck, ints_to_mont([1] * 7, c.q))
    ck.close()


def test_msm_big_bucket_path(engines, coracle):
    """All scalars equal -> every digit lands in one bucket per window: exercises the split-bucket tasks."""
    c = P.VESTA
    eng = engines["vesta"]
    n = 20000
    bases, logs = make_bases(c, n, seed=4)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    for s in (1, 3, c.q - 1, 0x1234567890ABCDEF1234567890ABCDEF):
        got = gpu_commit_affine(eng, ck, ints_to_mont([s] * n, c.q))
        assert got == P.scalar_mul(c, s * sum(logs) % c.q, P.generator(c))
    ck.close()


@pytest.mark.parametrize("name", ["pallas", "bn254"])
def test_msm_large_closed_form_and_linearity(name, engines):
    """n = 2^17 (the grayscale-HD commitment-key size) with device-generated bases k_i*G:
    result == (sum s_i k_i) G, and msm(a) + msm(b) == msm(a + b)."""
    import torch
    c = P.CURVES[name]
    eng = engines[name]
    n = 1 << 17
    k0, dk = 0x1234567, 0x89ABCDE
    d_bases = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, k0, dk, n, d_bases.data_ptr()))
    ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), n)
    # spot-check the generator against the oracle on a prefix
    host = d_bases[: 8 * 64].cpu().numpy().view(np.uint64).reshape(-1, 8)
    pts = mont_to_affine(host, c.p)
    G = P.generator(c)
    assert pts[0] == P.scalar_mul(c, k0, G) and pts[63] == P.scalar_mul(c, k0 + 63 * dk, G)
    rng = np.random.default_rng(5)
    a = rng.integers(0, 1 << 62, size=(n,), dtype=np.uint64)
    b = rng.integers(0, 1 << 62, size=(n,), dtype=np.uint64)
    def full(v, hi):  # scalars v + hi * 2^192 (spread over all windows), canonical ints
        return [int(x) + (int(h) << 192) for x, h in zip(v, hi)]
    sa, sb = full(a, b >> 3), full(b, a >> 3)
    logs = [(k0 + i * dk) for i in range(n)]
    ca = CommitmentEngine.commit(ck, ints_to_mont(sa, c.q))
    cb = CommitmentEngine.commit(ck, ints_to_mont(sb, c.q))
    cab = CommitmentEngine.commit(ck, ints_to_mont([(x + y) % c.q for x, y in zip(sa, sb)], c.q))
    assert eng.to_affine_ints(ca) == P.scalar_mul(c, sum(s * k for s, k in zip(sa, logs)) % c.q, G)
    assert eng.to_affine_ints(eng.point_sum(np.stack([ca, cb]))) == eng.to_affine_ints(cab)
    # point-range shards (the multi-GPU split) add up to the whole
    Sm = torch.from_numpy(ints_to_mont(sa, c.q).view(np.int64)).cuda()
    parts = []
    for g in range(4):
        first, cnt = g * (n // 4), n // 4
        parts.append(CommitmentEngine.commit_dev(ck, Sm.data_ptr() + first * 32, cnt, first=first))
    assert eng.to_affine_ints(eng.point_sum(np.stack(parts))) == eng.to_affine_ints(ca)
    ck.close()


@pytest.mark.parametrize("log2n,dist", [(20, "uniform"), (20, "witness"), (20, "edge"), (22, "uniform"), (22, "witness")])
def test_msm_bench_sizes_closed_form(log2n, dist):
    """The MSM half of the metric at the sizes bench.py times: Pallas, 2^20 points (window c = 17) and 2^22 (c = 20),
    device-generated bases (k0 + i*dk)*G, the three scalar distributions of SURVEY 8d.  The commitment must equal
    (sum s_i k_i)*G -- one big-integer scalar multiplication by the python oracle, independent of any MSM code."""
    import torch
    from vimz_host import synthetic as S
    c = P.PALLAS
    eng = vimz_b200.Engine("pallas", 0)
    n = 1 << log2n
    k0, dk = 77, 1234577
    d_bases = torch.empty(n * 8, dtype=torch.int64, device="cuda")
    vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, k0, dk, n, d_bases.data_ptr()))
    ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), n)
    G = P.generator(c)
    host = d_bases[-8:].cpu().numpy().view(np.uint64).reshape(-1, 8)            # the LAST base against the oracle
    assert mont_to_affine(host, c.p)[0] == P.scalar_mul(c, k0 + (n - 1) * dk, G)
    del d_bases
    assert ck.window_bits == {20: 17, 22: 20}[log2n], "bench.py's MSM numbers are quoted on these windows"
    if dist == "uniform":
        Sm = S.uniform_scalars_mont(n, c.q, 11 + log2n)
    elif dist == "witness":
        Sm = S.witness_like_scalars_mont(n, c.q, 12 + log2n)
    else:
        pat = ints_to_mont([0, 1, c.q - 1], c.q)
        Sm = np.ascontiguousarray(pat[np.arange(n) % 3])
    d_S = torch.from_numpy(Sm.view(np.int64)).cuda()
    got = eng.to_affine_ints(CommitmentEngine.commit_dev(ck, d_S.data_ptr(), n))
    assert got == P.scalar_mul(c, S.closed_form_log(Sm, k0, dk, c.q), G), (log2n, dist)
    # a ragged prefix and a point-range shard pair through the same table
    n1 = n // 3 + 5
    a = CommitmentEngine.commit_dev(ck, d_S.data_ptr(), n1)
    b = CommitmentEngine.commit_dev(ck, d_S.data_ptr() + n1 * 32, n - n1, first=n1)
    assert eng.to_affine_ints(a) == P.scalar_mul(c, S.closed_form_log(Sm[:n1], k0, dk, c.q), G)
    assert eng.to_affine_ints(eng.point_sum(np.stack([a, b]))) == got
    ck.close()
    eng.close()


def test_point_helpers(engines, coracle):
    c = P.GRUMPKIN
    eng = engines["grumpkin"]
    G = P.generator(c)
    A = P.scalar_mul(c, 1234567, G)
    B = P.scalar_mul(c, 7654321, G)
    one = ints_to_mont([1], c.p)[0]
    def jac(pt):
        return np.concatenate([affine_to_mont([pt], c.p)[0], one]) if pt else np.zeros(12, np.uint64)
    r = 0xDEADBEEFCAFEBABE0123456789ABCDEF
    got = eng.point_scale_add(jac(A), ints_to_mont([r], c.q), jac(B))
    assert eng.to_affine_ints(got) == P.aff_add(c, A, P.scalar_mul(c, r, B))
    assert eng.to_affine_ints(eng.point_scale_add(jac(None), ints_to_mont([r], c.q), jac(B))) == P.scalar_mul(c, r, B)
    assert eng.to_affine_ints(eng.point_scale_add(jac(A), ints_to_mont([0], c.q), jac(B))) == A
    exp = coracle.to_affine(c.curve_id, coracle.point_scale_add(c.curve_id, jac(A), ints_to_mont([r], c.q)[0], jac(B)))
    assert np.array_equal(eng.to_affine(got), exp)
    assert eng.to_affine_ints(eng.point_sum(np.stack([jac(A), jac(B), jac(None), jac(A)]))) == P.aff_add(c, P.aff_add(c, A, B), A)


@pytest.mark.parametrize("window", [8, 12])
def test_msm_skewed_buckets_all_combine_classes(window, engines, coracle):
    """Scalar multisets built so that, with few buckets (small window) and many entries, buckets are cut into
    a handful (thread combine), dozens (warp combine) and thousands (block-chunk combine) of segments."""
    c = P.PALLAS
    eng = vimz_b200.Engine("pallas", 0)
    eng.set_option("msm_window", window)
    n = 60000
    bases, logs = make_bases(c, n, seed=window)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    rng = random.Random(window)
    G = P.generator(c)
    for mix in ("giant+uniform", "few-values", "ones"):
        if mix == "giant+uniform":   # 70 % ones (one giant bucket), the rest uniform
            sc = [1 if rng.random() < 0.7 else rng.randrange(c.q) for _ in range(n)]
        elif mix == "few-values":    # 40 distinct scalars: every touched bucket is heavy
            pool = [rng.randrange(c.q) for _ in range(40)]
            sc = [pool[rng.randrange(40)] for _ in range(n)]
        else:
            sc = [1] * n
        got = gpu_commit_affine(eng, ck, ints_to_mont(sc, c.q))
        assert got == P.scalar_mul(c, sum(s * k for s, k in zip(sc, logs)) % c.q, G), (window, mix)
    ck.close()
    eng.close()


def test_bn254_published_known_answers_gpu(engines):
    """EIP-196 vectors for 2G / 3G on BN254 G1 through the GPU commit path (external pin, see tests/test_oracle.py)."""
    from test_oracle import BN254_2G, BN254_3G
    c = P.BN254
    eng = engines["bn254"]
    G = P.generator(c)
    ck = CommitmentKey.from_bases(eng, affine_to_mont([G, G], c.p))
    assert gpu_commit_affine(eng, ck, ints_to_mont([2, 0], c.q)) == BN254_2G
    assert gpu_commit_affine(eng, ck, ints_to_mont([1, 1], c.q)) == BN254_2G      # P + P inside one bucket
    assert gpu_commit_affine(eng, ck, ints_to_mont([2, 1], c.q)) == BN254_3G
    assert gpu_commit_affine(eng, ck, ints_to_mont([c.q - 2, 0], c.q)) == P.aff_neg(c, BN254_2G)
    ck.close()


@pytest.mark.parametrize("direct_c", [0, 6, 8, 9])
def test_msm_direct_table_ranges_and_digit_widths(direct_c, coracle):
    """Short keys commit from the direct multiples table (k_msm_direct): every digit width gives the oracle's group
    element, prefixes and point-range shards (`first` > 0) index the table correctly, and the key reports its geometry."""
    import torch
    c = P.VESTA
    eng = vimz_b200.Engine("vesta", 0)
    if direct_c:
        eng.set_option("msm_direct_c", direct_c)
    n = 3001
    rng = random.Random(direct_c)
    bases, logs = make_bases(c, n, seed=70 + direct_c)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    if direct_c == 6:    # 43 windows exceed the table's window limit: the key silently takes the bucket pipeline instead
        assert ck.window_bits != 6
    else:                # a direct key: fixed digit width, not the cost-model window
        assert ck.window_bits == (direct_c or 10)
    G = P.generator(c)
    for dist in ("uniform", "bits", "edge"):
        if dist == "uniform":
            sc = [rng.randrange(c.q) for _ in range(n)]
        elif dist == "bits":
            sc = [rng.randrange(2) if rng.random() < 0.93 else rng.randrange(c.q) for _ in range(n)]
        else:
            pool = [0, 1, c.q - 1, (c.q - 1) // 2, (c.q + 1) // 2, 1 << 254, (1 << 254) - 1, 127, 128, 129, 511, 512, 513]
            sc = [pool[i % len(pool)] % c.q for i in range(n)]
        Sm = ints_to_mont(sc, c.q)
        whole = gpu_commit_affine(eng, ck, Sm)
        assert whole == P.scalar_mul(c, sum(s * k for s, k in zip(sc, logs)) % c.q, G), (direct_c, dist)
        assert gpu_commit_affine(eng, ck, Sm[:777]) == oracle_commit_affine(coracle, c, Sm[:777], Bm[:777])   # prefix
        d_S = torch.from_numpy(Sm.view(np.int64)).cuda()
        parts, first = [], 0
        for cnt in (1000, 1, 0, 2000):                   # ragged point-range shards, one of them empty
            parts.append(CommitmentEngine.commit_dev(ck, d_S.data_ptr() + first * 32, cnt, first=first))
            first += cnt
        assert first == n and eng.to_affine_ints(eng.point_sum(np.stack(parts))) == whole
    ck.close()
    eng.close()


def test_concurrent_commits_on_one_context(engines, coracle):
    """CompressedSNARK::prove / RecursiveSNARK::verify commit from several rayon workers at once.  Calls from
    different threads on ONE context are serialised inside the library (ctypes releases the GIL, so they really overlap):
    every thread must get exactly the commitment a sequential call gives."""
    import threading
    c = P.PALLAS
    eng = engines["pallas"]
    n = 3000
    bases, _ = make_bases(c, n, seed=123)
    Bm = affine_to_mont(bases, c.p)
    ck = CommitmentKey.from_bases(eng, Bm)
    rng = random.Random(9)
    vecs = [ints_to_mont([rng.randrange(c.q) for _ in range(n - 17 * t)], c.q) for t in range(6)]
    expect = [gpu_commit_affine(eng, ck, v) for v in vecs]
    assert expect[0] == oracle_commit_affine(coracle, c, vecs[0], Bm)
    got = [[None] * 4 for _ in vecs]
    errs = []

    def worker(t):
        try:
            for rep in range(4):
                got[t][rep] = gpu_commit_affine(eng, ck, vecs[t])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(vecs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for t in range(len(vecs)):
        assert got[t] == [expect[t]] * 4
    ck.close()


@pytest.mark.parametrize("name", list(P.CURVES))
def test_endomorphism_identity_on_gpu_commitments(name, engines):
    """Implementation-independent pin (see tests/test_oracle.py::check_endomorphism): commit(ck, lambda * v) must be
    (zeta * x, y) of commit(ck, v) for primitive cube roots of unity lambda (scalar field) and zeta (base field), with the
    SAME pairing of lambda and zeta for every key and vector -- on device-generated bases, for a short key (direct table)
    and a 2^15-point key (bucket pipeline), without any oracle group arithmetic."""
    import torch
    from test_oracle import check_endomorphism
    c, eng = P.CURVES[name], engines[name]
    rng = random.Random(5 + c.curve_id)
    seen = set()
    for n in (700, 1 << 15):
        d_bases = torch.empty(n * 8, dtype=torch.int64, device="cuda")
        vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, 1000 + n, 987654321, n, d_bases.data_ptr()))
        ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), n)
        sc = [rng.randrange(c.q) if rng.random() < 0.5 else rng.randrange(1 << 130) for _ in range(n)]
        seen.add(check_endomorphism(c, lambda v: gpu_commit_affine(eng, ck, ints_to_mont(v, c.q)), sc))
        ck.close()
    assert len(seen) == 1
