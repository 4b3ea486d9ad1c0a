"""GPU parity: fp.cuh Montgomery arithmetic vs python big integers, through the C ABI (vimz_field_op)."""
import random

import pytest

from oracle import pyref as P
from vimz_b200.field import ints_to_mont, mont_to_ints

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(P.CURVES))
@pytest.mark.parametrize("which", ["base", "scalar"])
def test_field_ops_bit_exact(name, which, engines):
    c = P.CURVES[name]
    mod = c.p if which == "base" else c.q
    rng = random.Random(hash((name, which)) & 0xFFFF)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 255) % mod, (1 << 256) % mod, (1 << 128) - 1, (1 << 64), 0xFFFFFFFF, mod >> 1]
    a = [x for x in edge for _ in edge] + [rng.randrange(mod) for _ in range(5000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(mod) for _ in range(5000)]
    A, B = ints_to_mont(a, mod), ints_to_mont(b, mod)
    eng = engines[name]
    assert mont_to_ints(eng.field_op(which, "mul", A, B), mod) == [x * y % mod for x, y in zip(a, b)]
    assert mont_to_ints(eng.field_op(which, "add", A, B), mod) == [(x + y) % mod for x, y in zip(a, b)]
    assert mont_to_ints(eng.field_op(which, "sub", A, B), mod) == [(x - y) % mod for x, y in zip(a, b)]


@pytest.mark.parametrize("name", list(P.CURVES))
@pytest.mark.parametrize("which", ["base", "scalar"])
def test_field_square_bit_exact(name, which, engines):
    """fp_sqr (the dedicated squaring of fp_sqr.cuh on the Pasta fields: the two squarings of every mixed addition) equals the
    big-integer square and, limb for limb in Montgomery form, the product a * a."""
    import numpy as np
    c = P.CURVES[name]
    mod = c.p if which == "base" else c.q
    rng = random.Random(hash((name, which, "sqr")) & 0xFFFF)
    top = (1 << 254) - 1
    edge = [0, 1, 2, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1, (1 << 255) % mod, (1 << 256) % mod, (1 << 128) - 1, (1 << 64), 0xFFFFFFFF,
            top % mod, (top - 0xFFFFFFFF) % mod, (1 << 253), sum(0xFFFFFFFF << (64 * i) for i in range(4)) % mod,
            sum(0xFFFFFFFF << (64 * i + 32) for i in range(4)) % mod]
    # values whose Montgomery FORM (x * 2^256 mod p) hits the carry-heavy patterns, not only the values themselves
    rinv = pow(1 << 256, -1, mod)
    edge += [e * rinv % mod for e in edge]
    a = edge + [rng.randrange(mod) for _ in range(20000)]
    A = ints_to_mont(a, mod)
    eng = engines[name]
    sq = eng.field_op(which, "sqr", A, A)
    assert mont_to_ints(sq, mod) == [x * x % mod for x in a]
    assert np.array_equal(sq, eng.field_op(which, "mul", A, A))


def test_field_op_matches_c_oracle_limbs(engines, coracle):
    """Limb-for-limb (Montgomery representation) equality with the 4x64 CPU restatement."""
    import numpy as np
    c = P.BN254
    rng = random.Random(3)
    a = ints_to_mont([rng.randrange(c.p) for _ in range(4096)], c.p)
    b = ints_to_mont([rng.randrange(c.p) for _ in range(4096)], c.p)
    for op_i, op in enumerate(["mul", "add", "sub"]):
        assert np.array_equal(engines["bn254"].field_op("base", op, a, b), coracle.field_op(c.curve_id, 0, op_i, a, b))
