"""TEST INFRASTRUCTURE (oracle) -- circomlib Poseidon over BN254 Fr and the vimz row-wise image running hash.

SURVEY.md section 8(f)-4: the fold path itself has no golden vector in the reference, but the IVC state it
carries does -- `marketplace/image-data/*.hash` hold the final running hash z_720 of eight 720p images.  This
module restates the hash so those fixtures can be reproduced from the PNGs:

  * Poseidon(nInputs) of circomlib 2.0.5 (`circuits/poseidon.circom`, an un-vendored npm dependency: the reference
    includes it at /root/reference/circuits/src/utils/hashers.circom:4): x^5 S-box, R_F = 8 full rounds,
    R_P = N_ROUNDS_P[t - 2] partial rounds, state = (0, inputs...), output = state[0].  The round constants and the
    MDS matrix are the published Grain-LFSR parameters of the Poseidon paper (generate_parameters_grain.sage with
    field = 1, sbox = 0, n = 254), regenerated here bit for bit instead of being copied.  The optimised
    circuit in circomlib (constants C, S, M, P) computes the same permutation.
  * hashers (/root/reference/circuits/src/utils/hashers.circom): PairHasher :7-16, _WindowFoldHasher :40-73
    (INCLUDING its round count `(LENGTH + WINDOW - 1) \\ WINDOW`, which for LENGTH = 128, WINDOW = 8 stops after
    8 + 15*7 = 113 elements -- the last 15 field elements of a row never enter the hash), HeadTailHasher :115-120.
  * running hash (/root/reference/circuits/image_running_hash.circom:8-18 and
    /root/reference/pyvimz/pyvimz/image_hasher.py:89-111): acc_0 = 0, acc_{i+1} = HeadTailHasher(128)(acc_i, row_i)
    with rows packed 10 pixels per field element (/root/reference/pyvimz/pyvimz/img/ops.py:4-33).

Pinned by: the published circomlib vector poseidon([1, 2]) and the reference's own eight `.hash` fixtures
(tests/test_poseidon_golden.py).  The fast path runs the permutation in oracle/nova_cpu.c (`oracle_poseidon`).
"""
from __future__ import annotations

from functools import lru_cache
from typing import List, Sequence

BN254_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
N_ROUNDS_F = 8
# circomlib poseidon.circom: N_ROUNDS_P[t - 2]
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]


class _Grain:
    """80-bit Grain LFSR of the Poseidon reference parameter script (self-shrinking output)."""

    def __init__(self, field: int, sbox: int, n: int, t: int, r_f: int, r_p: int):
        bits: List[int] = []
        for value, width in ((field, 2), (sbox, 4), (n, 12), (t, 12), (r_f, 10), (r_p, 10)):
            bits += [(value >> (width - 1 - i)) & 1 for i in range(width)]
        bits += [1] * 30
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._step()

    def _step(self) -> int:
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self) -> int:
        while True:
            if self._step() == 1:
                return self._step()
            self._step()  # discard the pair

    def bits(self, n: int) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


@lru_cache(maxsize=None)
def poseidon_params(t: int, p: int = BN254_FR):
    """(round constants [(R_F + R_P) * t], MDS matrix t x t) for width t."""
    n = p.bit_length()
    r_p = N_ROUNDS_P[t - 2]
    g = _Grain(1, 0, n, t, N_ROUNDS_F, r_p)
    consts = []
    while len(consts) < (N_ROUNDS_F + r_p) * t:
        v = g.bits(n)
        if v < p:
            consts.append(v)
    while True:  # Cauchy matrix 1 / (x_i + y_j) from 2t distinct field elements
        xs_ys = [g.bits(n) % p for _ in range(2 * t)]
        if len(set(xs_ys)) != 2 * t:
            continue
        xs, ys = xs_ys[:t], xs_ys[t:]
        if any((x + y) % p == 0 for x in xs for y in ys):
            continue
        mds = [[pow((x + y) % p, p - 2, p) for y in ys] for x in xs]
        return consts, mds


def permute(state: Sequence[int], p: int = BN254_FR) -> List[int]:
    """The Poseidon permutation (plain form: add constants, S-box, MDS) on a width-t state."""
    t = len(state)
    consts, mds = poseidon_params(t, p)
    r_p = N_ROUNDS_P[t - 2]
    s = list(state)
    half = N_ROUNDS_F // 2
    for r in range(N_ROUNDS_F + r_p):
        s = [(x + consts[r * t + i]) % p for i, x in enumerate(s)]
        if r < half or r >= half + r_p:
            s = [pow(x, 5, p) for x in s]
        else:
            s[0] = pow(s[0], 5, p)
        s = [sum(mds[i][j] * s[j] for j in range(t)) % p for i in range(t)]
    return s


def poseidon(inputs: Sequence[int], p: int = BN254_FR, fast: bool = True) -> int:
    """circomlib Poseidon(len(inputs)): state (0, inputs...), output state[0]."""
    if fast:
        from . import c as oracle_c
        return oracle_c().poseidon([0] + [x % p for x in inputs])[0]
    return permute([0] + [x % p for x in inputs], p)[0]


def window_fold_hash(array: Sequence[int], window: int = 8, fast: bool = True) -> int:
    """_WindowFoldHasher(LENGTH, WINDOW_SIZE) -- hashers.circom:40-73, round count and all."""
    length = len(array)
    num_rounds = (length + window - 1) // window
    first = min(length, window)
    h = poseidon(array[:first], fast=fast)
    processed = first
    for _ in range(num_rounds - 1):
        remaining = length - processed
        cur = remaining if remaining < window - 1 else window - 1
        h = poseidon([h] + list(array[processed:processed + cur]), fast=fast)
        processed += cur
    return h


def head_tail_hash(head: int, tail: Sequence[int], fast: bool = True) -> int:
    """HeadTailHasher(TAIL_LENGTH) -- hashers.circom:115-120."""
    return poseidon([head, window_fold_hash(tail, 8, fast)], fast=fast)


def image_running_hash(rows: Sequence[Sequence[int]], fast: bool = True) -> int:
    """acc_0 = 0; acc <- HeadTailHasher(width)(acc, row) for every packed row (image_hasher.py:89-111)."""
    acc = 0
    for row in rows:
        acc = head_tail_hash(acc, row, fast)
    return acc
