"""CPU checks of the dedicated Pasta squaring (vimz_b200/csrc/fp_sqr.cuh): the generator's limb-level model -- the very carry
chains the header executes as inline PTX -- against x*x*R^-1 mod p, and the committed header against the generator's output.
The GPU side of the same routine is tests/test_gpu_field.py::test_field_square_bit_exact."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gen():
    spec = importlib.util.spec_from_file_location("gen_fp_sqr", os.path.join(ROOT, "tools", "gen_fp_sqr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_limb_model_matches_big_integer_square():
    g = _gen()
    assert g.selftest(n=1500)


def test_limb_model_on_carry_heavy_montgomery_forms():
    """Inputs whose LIMBS (not values) are all-ones / alternating patterns: every deferred carry of the triangle and every ripple of
    the reduction rows is exercised; the model asserts internally that no carry is dropped."""
    g = _gen()
    rng = random.Random(11)
    for p in (g.P_PALLAS, g.P_VESTA):
        rinv = pow(1 << 256, -1, p)
        pats = [(1 << 254) - 1, ((1 << 254) - 1) ^ 0xFFFFFFFF, int("3fffffff" + "00000000ffffffff" * 3 + "00000000", 16) >> 32,
                int("3" + "f" * 63, 16), int("2" + "a" * 63, 16), int("1" + "5" * 63, 16), p - 1, p - 0xFFFFFFFF, (p >> 1) + 1]
        for _ in range(300):
            x = 0
            for i in range(8):
                x |= rng.choice([0, 0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 1, rng.getrandbits(32)]) << (32 * i)
            pats.append(x & ((1 << 254) - 1))
        for x in pats:
            x %= p
            limbs = [(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)]
            assert g.model(limbs, p) == x * x * rinv % p


def test_committed_header_is_the_generators_output():
    g = _gen()
    with open(os.path.join(ROOT, "vimz_b200", "csrc", "fp_sqr.cuh")) as f:
        assert f.read() == g.emit(), "vimz_b200/csrc/fp_sqr.cuh is stale: python tools/gen_fp_sqr.py --emit > vimz_b200/csrc/fp_sqr.cuh"
