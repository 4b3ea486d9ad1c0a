"""Deterministic synthetic step circuits with the dimensions and witness statistics of the vimz
Circom steps, for tests and bench.py.

The real `.r1cs` / witness files cannot be produced in this image (no circom / node; they are
git-ignored in the reference: /root/reference/.gitignore:4-10), so the fold is exercised on shapes
built to the published sizes (/root/reference/circuits/nova_snark/circuit_parameters.csv:2-9, plus
~10k constraints/variables of Nova's augmented circuit):

  * "bit" rows      b*(b-1) = 0          A={(i,b,1)} B={(i,b,1),(i,one,-1)} C={}       (Num2Bits / LessEqThan bits)
  * "pack" rows     (sum 2^k b_k)*1 = x  one per 240-bit pixel word or 19-bit comparator (pixels.circom:6-29)
  * "dense" rows    (sum a z)*(sum b z) = w   8-20 terms, coefficients from a small set of full-width
                    constants (Poseidon MDS / round constants)
  * "copy" rows     w*1 = w              pad num_cons up to the published count
Witnesses satisfy the shape by construction; >85 % of W is 0/1 like the real circuits (row a10).
Host-side numpy / python integers only -- nothing here is on the measured path.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

from .field import CurveInfo, ints_to_mont

NOVA_AUGMENTED = 10_000  # constraints/variables added by NovaAugmentedCircuit on the primary curve (approx.)

# name -> (step-circuit constraints, wires, words of 240 bits, comparators, comparator bits, steps for HD)
STEP_CIRCUITS: Dict[str, Tuple[int, int, int, int, int, int]] = {
    "grayscale": (120_864, 118_307, 256, 2_560, 19, 720),
    "brightness": (305_185, 289_829, 256, 15_360, 14, 720),
    "contrast": (305_185, 289_829, 256, 15_360, 14, 720),
    "resize": (241_968, 234_291, 512, 7_680, 12, 240),
    "crop": (672_273, 671_633, 128, 0, 0, 720),
    "blur": (248_934, 241_257, 512, 7_680, 12, 720),
    "sharpness": (325_734, 310_377, 512, 15_360, 12, 720),
    "hash": (6_672, 6_787, 0, 0, 0, 720),
    # 4K variants (BASELINE config "blur/sharpness convolution steps on 4K.png"): no such circuit exists in the reference (the
    # row width is hard-wired to 128 words, circuits/src/blur_step.circom:17), so these are the x3-width ESTIMATES of
    # three times the HD constraints / wires / packed words / comparators, 2160 rows per image
    "blur4k": (746_802, 723_771, 1_536, 23_040, 12, 2_160),
    "sharpness4k": (977_202, 931_131, 1_536, 46_080, 12, 2_160),
    # the secondary-curve shape: TrivialTestCircuit inside the augmented circuit, ~10.5k rows
    "secondary": (500, 500, 0, 0, 0, 0),
}


@dataclass
class SyntheticShape:
    name: str
    num_cons: int
    num_vars: int
    num_io: int
    A: Tuple[np.ndarray, np.ndarray, np.ndarray]
    B: Tuple[np.ndarray, np.ndarray, np.ndarray]
    C: Tuple[np.ndarray, np.ndarray, np.ndarray]
    # witness recipe
    nbits: int = 0
    packs: List[Tuple[int, int, int]] = field(default_factory=list)          # (first bit var, nbits, target var)
    nfree: int = 0
    free_base: int = 0
    dense: List[Tuple[List[Tuple[int, int]], List[Tuple[int, int]], int]] = field(default_factory=list)
    modulus: int = 0

    @property
    def nnz(self) -> int:
        return self.A[0].shape[0] + self.B[0].shape[0] + self.C[0].shape[0]

    def coo_ints(self):
        """COO triples with canonical integer values (for the python oracle)."""
        from .field import mont_to_ints
        out = []
        for rows, cols, vals in (self.A, self.B, self.C):
            iv = mont_to_ints(vals, self.modulus)
            out.append(list(zip(rows.tolist(), cols.tolist(), iv)))
        return out


def _dims(name: str, scale: float):
    cons, wires, words, ncmp, cmpbits, _ = STEP_CIRCUITS[name]
    m = int((cons + NOVA_AUGMENTED) * scale)
    n = int((wires + NOVA_AUGMENTED) * scale)
    words = max(1, int(words * scale)) if words else 0
    ncmp = int(ncmp * scale)
    return m, n, words, ncmp, cmpbits


def synthetic_shape(curve: CurveInfo, name: str = "grayscale", seed: int = 0xB200, scale: float = 1.0,
                    word_bits: int = 240) -> SyntheticShape:
    """Build the shape for step circuit `name` (scaled by `scale` for the small parity cases)."""
    q = curve.scalar_modulus
    rng = random.Random(seed)
    m, n, words, ncmp, cmpbits = _dims(name, scale)
    num_io = 2
    # variable budget
    nbits = words * word_bits + ncmp * cmpbits
    npack = words + ncmp
    while nbits + npack > int(0.95 * min(n, m)):   # tiny scales: shrink the bit section
        if words > 1:
            words -= 1
        elif ncmp > 0:
            ncmp -= 1
        elif word_bits > 8:
            word_bits //= 2
        else:
            break
        nbits = words * word_bits + ncmp * cmpbits
        npack = words + ncmp
    rest = n - nbits - npack
    assert rest >= 4, "shape too small"
    ndense = min(max(rest * 5 // 6, 1), m - nbits - npack)
    nfree = rest - ndense
    ncopy = m - nbits - npack - ndense
    assert nfree >= 1 and ncopy >= 0
    one_col = n  # column of u
    # variable layout: [bits][pack targets][free][dense targets]
    pack_base, free_base, dense_base = nbits, nbits + npack, nbits + npack + nfree

    consts = [1, q - 1, 2, 3, 5, 7] + [rng.randrange(q) for _ in range(58)]
    rowsA, colsA, valsA = [], [], []
    rowsB, colsB, valsB = [], [], []
    rowsC, colsC, valsC = [], [], []

    # bit rows (vectorised)
    bit_rows = np.arange(nbits, dtype=np.uint32)
    # pack rows
    packs = []
    row = nbits
    b0 = 0
    pow2 = [1 << k for k in range(word_bits)]
    for j in range(npack):
        width = word_bits if j < words else cmpbits
        tgt = pack_base + j
        packs.append((b0, width, tgt))
        rowsA += [row] * width
        colsA += list(range(b0, b0 + width))
        valsA += pow2[:width]
        rowsB.append(row); colsB.append(one_col); valsB.append(1)
        rowsC.append(row); colsC.append(tgt); valsC.append(1)
        b0 += width
        row += 1
    # dense rows: target variable w_j defined from earlier variables (and X, one)
    dense = []
    ncols = n + 1 + num_io
    for j in range(ndense):
        tgt = dense_base + j
        hi = dense_base + j  # may reference any earlier variable
        def pick():
            r = rng.random()
            if r < 0.05:
                return n + rng.randrange(0, 1 + num_io)  # one / X
            if r < 0.55 and j > 0:
                return dense_base + rng.randrange(max(0, j - 64), j)  # recent dense vars (Poseidon state)
            return rng.randrange(pack_base, hi) if hi > pack_base else rng.randrange(0, max(1, hi))
        ta = [(pick(), rng.choice(consts)) for _ in range(rng.randint(8, 20))]
        tb = [(pick(), rng.choice(consts)) for _ in range(rng.choice([1, 1, 1, 8, 12]))]
        dense.append((ta, tb, tgt))
        for c_, v_ in ta:
            rowsA.append(row); colsA.append(c_); valsA.append(v_)
        for c_, v_ in tb:
            rowsB.append(row); colsB.append(c_); valsB.append(v_)
        rowsC.append(row); colsC.append(tgt); valsC.append(1)
        row += 1
    # copy rows
    for j in range(ncopy):
        v = rng.randrange(pack_base, n)
        rowsA.append(row); colsA.append(v); valsA.append(1)
        rowsB.append(row); colsB.append(one_col); valsB.append(1)
        rowsC.append(row); colsC.append(v); valsC.append(1)
        row += 1
    assert row == m and max(colsA + colsB + colsC + [0]) < ncols

    def mont_vals(vals):
        # few distinct values: convert each once
        uniq = {}
        for v in vals:
            if v not in uniq:
                uniq[v] = len(uniq)
        table = ints_to_mont(list(uniq.keys()), q)
        idx = np.fromiter((uniq[v] for v in vals), dtype=np.int64, count=len(vals))
        return table[idx] if len(vals) else np.zeros((0, 4), np.uint64)

    one_m = ints_to_mont([1, q - 1], q)
    A = (np.concatenate([bit_rows, np.asarray(rowsA, np.uint32)]),
         np.concatenate([bit_rows, np.asarray(colsA, np.uint32)]),
         np.concatenate([np.repeat(one_m[:1], nbits, axis=0), mont_vals(valsA)]))
    B = (np.concatenate([np.repeat(bit_rows, 2), np.asarray(rowsB, np.uint32)]),
         np.concatenate([np.stack([bit_rows, np.full(nbits, one_col, np.uint32)], 1).reshape(-1), np.asarray(colsB, np.uint32)]),
         np.concatenate([np.tile(one_m, (nbits, 1)), mont_vals(valsB)]))
    Cm = (np.asarray(rowsC, np.uint32), np.asarray(colsC, np.uint32), mont_vals(valsC))
    return SyntheticShape(name, m, n, num_io, A, B, Cm, nbits=nbits, packs=packs, nfree=nfree, free_base=free_base,
                          dense=dense, modulus=q)


def synthetic_witness(shape: SyntheticShape, seed: int, bits: np.ndarray | None = None):
    """A satisfying (W, X) as python integers.  `bits` (0/1 array of length shape.nbits) lets the caller
    feed pixel-derived bits; default is seeded uniform bits."""
    q = shape.modulus
    rng = random.Random(seed)
    n, io = shape.num_vars, shape.num_io
    W = [0] * n
    if bits is None:
        bits = np.random.default_rng(seed).integers(0, 2, size=shape.nbits, dtype=np.uint8)
    bl = bits.tolist()
    W[: shape.nbits] = bl
    for b0, width, tgt in shape.packs:
        acc = 0
        for k in range(width):
            if bl[b0 + k]:
                acc |= 1 << k
        W[tgt] = acc
    for j in range(shape.nfree):
        W[shape.free_base + j] = rng.randrange(256) if rng.random() < 0.8 else rng.randrange(1 << 19)
    X = [rng.randrange(q) for _ in range(io)]
    z_tail = [1] + X
    for ta, tb, tgt in shape.dense:
        sa = 0
        for c_, v_ in ta:
            sa += v_ * (W[c_] if c_ < n else z_tail[c_ - n])
        sb = 0
        for c_, v_ in tb:
            sb += v_ * (W[c_] if c_ < n else z_tail[c_ - n])
        W[tgt] = (sa % q) * (sb % q) % q
    return W, X


def uniform_scalars_mont(n: int, modulus: int, seed: int) -> np.ndarray:
    """n uniform Montgomery-form scalars below 2^(bits-1) <= modulus (dense `T`/`E`-like vectors) without
    python big-int work: any value < modulus is the Montgomery form of some scalar."""
    g = np.random.default_rng(seed)
    a = g.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + g.integers(0, 2, size=(n, 4), dtype=np.uint64)
    top_bits = modulus.bit_length() - 1 - 192
    a[:, 3] &= np.uint64((1 << top_bits) - 1)
    return a


def witness_like_scalars_mont(n: int, modulus: int, seed: int, p_bit: float = 0.93, p_byte: float = 0.02) -> np.ndarray:
    """Distribution B of SURVEY.md section 8d: 93 % {0,1}, 2 % bytes, 5 % uniform -- Montgomery form."""
    g = np.random.default_rng(seed)
    out = uniform_scalars_mont(n, modulus, seed + 1)
    sel = g.random(n)
    small = ints_to_mont(list(range(256)), modulus)
    bits = g.integers(0, 2, size=n)
    byts = g.integers(0, 256, size=n)
    is_bit = sel < p_bit
    is_byte = (sel >= p_bit) & (sel < p_bit + p_byte)
    out[is_bit] = small[bits[is_bit]]
    out[is_byte] = small[byts[is_byte]]
    return out


def closed_form_log(scalars_mont: np.ndarray, k0: int, dk: int, modulus: int, first: int = 0) -> int:
    """sum_i s_i * (k0 + (first + i) * dk) mod q for (n, 4) Montgomery-form scalars s_i: the discrete log (base G) of
    commit(ck, s) when ck_i = (k0 + i*dk) * G (vimz_gen_bases_dev / oracle_gen_bases), so the commitment of millions of
    points can be checked against ONE scalar multiplication, independent of any MSM implementation.
    Exact integer arithmetic in numpy: 32-bit half-limbs summed in blocks of 1024 so nothing exceeds 2^63."""
    a = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
    n = a.shape[0]
    if n == 0:
        return 0
    pad = (-n) % 1024
    if pad:
        a = np.concatenate([a, np.zeros((pad, 4), np.uint64)])
    halves = np.stack([a & np.uint64(0xFFFFFFFF), a >> np.uint64(32)], axis=2).reshape(-1, 8)   # (n, 8): 32-bit digits, LE
    blocks = halves.reshape(-1, 1024, 8)
    j = np.arange(1024, dtype=np.uint64).reshape(1, 1024, 1)
    s0 = blocks.sum(axis=1)                 # (nb, 8)  < 2^42
    s1 = (blocks * j).sum(axis=1)           # (nb, 8)  < 2^52
    tot0 = [int(x) for x in s0.sum(axis=0)]                                    # sum of digit d over all i
    base = np.arange(blocks.shape[0], dtype=np.uint64) * np.uint64(1024)       # block offsets < 2^32
    tot1 = []
    for d in range(8):                      # sum_i i * digit_d(i) = sum_blocks (base * s0 + s1), in python integers
        tot1.append(sum(int(b) * int(x) + int(y) for b, x, y in zip(base.tolist(), s0[:, d].tolist(), s1[:, d].tolist())))
    S = sum(tot0[d] << (32 * d) for d in range(8))          # sum_i mont_i
    S1 = sum(tot1[d] << (32 * d) for d in range(8))         # sum_i i * mont_i
    rinv = pow(1 << 256, -1, modulus)
    return ((k0 + first * dk) * S + dk * S1) % modulus * rinv % modulus
