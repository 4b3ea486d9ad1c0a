"""CPU ORACLE (test infrastructure, NOT product code) -- big-int restatement of the
nova-snark 0.23.0 per-step NIFS fold path that zero-savvy/vimz drives.

PARITY UNPINNED at the nova-snark boundary: the arithmetic of this path lives in
un-vendored crates (nova-snark 0.23.0, halo2curves 0.1.0, pasta_curves 0.5.1,
pasta-msm 0.1.4 -- pins in /root/reference/vimz/Cargo.lock:3577,2711,3958,3945) and the
reference holds no golden vector for it.  What pins this oracle
instead: every output is a canonical value (field element mod p, affine coordinates of a
uniquely defined group element), so an independent big-int implementation is bit-identical
to any correct implementation; curve constants are checked numerically in
tests/test_oracle.py (moduli prime, generators on curve and of the stated order).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import
this module.

Call sites in the reference this follows:
  * create_recursive_circuit   /root/reference/vimz/src/nova_snark_backend/folding.rs:35
  * RecursiveSNARK::verify     /root/reference/vimz/src/nova_snark_backend/folding.rs:53-55
  * type G1 / type G2          /root/reference/vimz/src/nova_snark_backend/mod.rs:19-20
Algorithm restated from SURVEY.md Appendix A ([EXT nova-snark@0.23.0] src/r1cs.rs,
src/nifs.rs, src/provider/pedersen.rs, src/provider/mod.rs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

R_BITS = 256
R = 1 << R_BITS

# ----------------------------------------------------------------------------------------
# Curve cycle constants.  Each curve: y^2 = x^3 + b over F_base,
# group order = scalar-field modulus.
# ----------------------------------------------------------------------------------------
PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
VESTA_P = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
BN254_P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
BN254_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


@dataclass(frozen=True)
class Curve:
    name: str
    curve_id: int
    p: int  # base field modulus (coordinates)
    q: int  # scalar field modulus (group order)
    b: int  # y^2 = x^3 + b
    gx: int
    gy: int


PALLAS = Curve("pallas", 0, PALLAS_P, VESTA_P, 5, PALLAS_P - 1, 2)
VESTA = Curve("vesta", 1, VESTA_P, PALLAS_P, 5, VESTA_P - 1, 2)
BN254 = Curve("bn254", 2, BN254_P, BN254_R, 3, 1, 2)
GRUMPKIN = Curve(
    "grumpkin", 3, BN254_R, BN254_P, BN254_R - 17, 1,
    17631683881184975370165255887551781615748388533673675138860,
)
CURVES = {c.name: c for c in (PALLAS, VESTA, BN254, GRUMPKIN)}
CURVES_BY_ID = {c.curve_id: c for c in CURVES.values()}


# ----------------------------------------------------------------------------------------
# Montgomery form helpers: the host layout of halo2curves / pasta_curves field types is
# [u64;4] little-endian limbs holding a*R mod p, R = 2^256.
# ----------------------------------------------------------------------------------------
def to_mont(a: int, p: int) -> int:
    return (a % p) * R % p


def from_mont(a: int, p: int) -> int:
    return a * pow(R, -1, p) % p


def int_to_le32(a: int) -> bytes:
    return int(a).to_bytes(32, "little")


def le32_to_int(b: bytes) -> int:
    return int.from_bytes(b, "little")


def mont_mul(a: int, b: int, p: int) -> int:
    """Montgomery product a*b*R^-1 mod p (what one GPU/CPU modmul computes)."""
    return a * b * pow(R, -1, p) % p


# ----------------------------------------------------------------------------------------
# Group law, canonical (non-Montgomery) integers.  Affine point = (x, y) or None = identity.
# ----------------------------------------------------------------------------------------
Aff = Optional[Tuple[int, int]]


def on_curve(c: Curve, P: Aff) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - c.b) % c.p == 0


def aff_neg(c: Curve, P: Aff) -> Aff:
    if P is None:
        return None
    return (P[0], (-P[1]) % c.p)


def aff_add(c: Curve, P: Aff, Q: Aff) -> Aff:
    if P is None:
        return Q
    if Q is None:
        return P
    p = c.p
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    y3 = (lam * (x1 - x3) - y1) % p
    return (x3, y3)


# Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3; identity Z = 0.
Jac = Tuple[int, int, int]
JAC_ID: Jac = (0, 1, 0)


def jac_from_aff(P: Aff) -> Jac:
    return JAC_ID if P is None else (P[0], P[1], 1)


def jac_to_aff(c: Curve, P: Jac) -> Aff:
    X, Y, Z = P
    if Z % c.p == 0:
        return None
    zi = pow(Z, -1, c.p)
    zi2 = zi * zi % c.p
    return (X * zi2 % c.p, Y * zi2 * zi % c.p)


def jac_double(c: Curve, P: Jac) -> Jac:
    p = c.p
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return JAC_ID
    A = X * X % p
    B = Y * Y % p
    C = B * B % p
    D = 2 * ((X + B) * (X + B) - A - C) % p
    E = 3 * A % p
    F = E * E % p
    X3 = (F - 2 * D) % p
    Y3 = (E * (D - X3) - 8 * C) % p
    Z3 = 2 * Y * Z % p
    return (X3, Y3, Z3)


def jac_add(c: Curve, P: Jac, Q: Jac) -> Jac:
    p = c.p
    if P[2] == 0:
        return Q
    if Q[2] == 0:
        return P
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    Z1Z1 = Z1 * Z1 % p
    Z2Z2 = Z2 * Z2 % p
    U1 = X1 * Z2Z2 % p
    U2 = X2 * Z1Z1 % p
    S1 = Y1 * Z2 * Z2Z2 % p
    S2 = Y2 * Z1 * Z1Z1 % p
    if U1 == U2:
        if S1 == S2:
            return jac_double(c, P)
        return JAC_ID
    H = (U2 - U1) % p
    Rr = (S2 - S1) % p
    HH = H * H % p
    HHH = H * HH % p
    V = U1 * HH % p
    X3 = (Rr * Rr - HHH - 2 * V) % p
    Y3 = (Rr * (V - X3) - S1 * HHH) % p
    Z3 = Z1 * Z2 * H % p
    return (X3, Y3, Z3)


def scalar_mul(c: Curve, k: int, P: Aff) -> Aff:
    """Plain double-and-add; k is reduced mod the group order."""
    k %= c.q
    acc = JAC_ID
    base = jac_from_aff(P)
    while k:
        if k & 1:
            acc = jac_add(c, acc, base)
        base = jac_double(c, base)
        k >>= 1
    return jac_to_aff(c, acc)


def generator(c: Curve) -> Aff:
    return (c.gx, c.gy)


# ----------------------------------------------------------------------------------------
# MSM.  `msm_naive` is the definition (sum of independent scalar multiplications) and is
# what parity is anchored on.  `cpu_best_multiexp` restates nova-snark's bucket method
# (unsigned windows, c = ceil(ln n), segments = 256/c + 1, running-sum) --
# same group element, used to cross-check the C oracle's window logic.
# ----------------------------------------------------------------------------------------
def msm_naive(c: Curve, scalars: Sequence[int], bases: Sequence[Aff]) -> Aff:
    assert len(scalars) <= len(bases)
    acc = JAC_ID
    for s, P in zip(scalars, bases):
        if P is None or s % c.q == 0:
            continue
        acc = jac_add(c, acc, jac_from_aff(scalar_mul(c, s, P)))
    return jac_to_aff(c, acc)


def window_size(n: int) -> int:
    """[EXT nova-snark provider::cpu_multiexp_serial] c = 1 if n<4; 3 if n<32; else ceil(ln n)."""
    if n < 4:
        return 1
    if n < 32:
        return 3
    return int(math.ceil(math.log(n)))


def cpu_multiexp_serial(c: Curve, scalars: Sequence[int], bases: Sequence[Aff]) -> Jac:
    n = len(scalars)
    w = window_size(n)
    segments = 256 // w + 1
    acc = JAC_ID
    for seg in range(segments - 1, -1, -1):
        for _ in range(w):
            acc = jac_double(c, acc)
        buckets: List[Jac] = [JAC_ID] * ((1 << w) - 1)
        for s, P in zip(scalars, bases):
            d = ((s % c.q) >> (seg * w)) & ((1 << w) - 1)
            if d and P is not None:
                buckets[d - 1] = jac_add(c, buckets[d - 1], jac_from_aff(P))
        running = JAC_ID
        for bk in reversed(buckets):
            running = jac_add(c, running, bk)
            acc = jac_add(c, acc, running)
    return acc


def cpu_best_multiexp(c: Curve, scalars: Sequence[int], bases: Sequence[Aff], num_threads: int = 8) -> Aff:
    n = len(scalars)
    if n > num_threads:
        chunk = n // num_threads
        acc = JAC_ID
        for s0 in range(0, n, chunk):
            acc = jac_add(c, acc, cpu_multiexp_serial(c, scalars[s0:s0 + chunk], bases[s0:s0 + chunk]))
        return jac_to_aff(c, acc)
    return jac_to_aff(c, cpu_multiexp_serial(c, scalars, bases))


def commit(c: Curve, ck: Sequence[Aff], v: Sequence[int]) -> Aff:
    """[EXT nova-snark src/provider/pedersen.rs] CommitmentEngine::commit(ck, v) = sum v_i * ck_i over
    the prefix ck[..len(v)] (asserts ck.len() >= v.len())."""
    assert len(ck) >= len(v)
    if len(v) >= 64:
        return cpu_best_multiexp(c, v, ck[: len(v)])
    return msm_naive(c, v, ck[: len(v)])


# ----------------------------------------------------------------------------------------
# R1CS (scalars live in the curve's scalar field F_q).  Matrices are COO triples in
# constraint order as in nova-snark 0.23.0's R1CSShape.
# ----------------------------------------------------------------------------------------
Coo = List[Tuple[int, int, int]]


@dataclass
class R1CSShape:
    num_cons: int
    num_vars: int
    num_io: int
    A: Coo
    B: Coo
    C: Coo

    def multiply_vec(self, q: int, z: Sequence[int]):
        """[EXT src/r1cs.rs R1CSShape::multiply_vec] -> (Az, Bz, Cz); z = W || u || X."""
        if len(z) != self.num_io + self.num_vars + 1:
            raise ValueError("InvalidWitnessLength")
        out = []
        for M in (self.A, self.B, self.C):
            Mz = [0] * self.num_cons
            for row, col, val in M:
                Mz[row] = (Mz[row] + val * z[col]) % q
            out.append(Mz)
        return tuple(out)

    def cross_term(self, q, W1, u1, X1, W2, X2):
        """[EXT src/r1cs.rs R1CSShape::commit_T], the vector part:
        T = Az1 o Bz2 + Az2 o Bz1 - u1*Cz2 - u2*Cz1 with u2 = 1."""
        Az1, Bz1, Cz1 = self.multiply_vec(q, list(W1) + [u1] + list(X1))
        Az2, Bz2, Cz2 = self.multiply_vec(q, list(W2) + [1] + list(X2))
        return [
            (Az1[i] * Bz2[i] + Az2[i] * Bz1[i] - u1 * Cz2[i] - Cz1[i]) % q
            for i in range(self.num_cons)
        ]

    def is_sat_relaxed(self, q, W, E, u, X) -> bool:
        """Az o Bz = u*Cz + E (commitment checks are done by the caller)."""
        Az, Bz, Cz = self.multiply_vec(q, list(W) + [u] + list(X))
        return all((Az[i] * Bz[i] - u * Cz[i] - E[i]) % q == 0 for i in range(self.num_cons))


def fold_witness(q, W1, E1, W2, T, r):
    """[EXT src/r1cs.rs RelaxedR1CSWitness::fold] W = W1 + r*W2 ; E = E1 + r*T."""
    W = [(a + r * b) % q for a, b in zip(W1, W2)]
    E = [(a + r * b) % q for a, b in zip(E1, T)]
    return W, E


def fold_instance(c: Curve, comm_W1: Aff, comm_E1: Aff, u1: int, X1, comm_W2: Aff, X2, comm_T: Aff, r: int):
    """[EXT src/r1cs.rs RelaxedR1CSInstance::fold] X = X1 + r*X2; comm_W = comm_W1 + r*comm_W2;
    comm_E = comm_E1 + r*comm_T; u = u1 + r."""
    q = c.q
    X = [(a + r * b) % q for a, b in zip(X1, X2)]
    comm_W = aff_add(c, comm_W1, scalar_mul(c, r, comm_W2))
    comm_E = aff_add(c, comm_E1, scalar_mul(c, r, comm_T))
    return comm_W, comm_E, (u1 + r) % q, X


def nifs_prove(c: Curve, ck, shape: R1CSShape, U1, W1, U2, W2, squeeze):
    """[EXT src/nifs.rs NIFS::prove] with the RO abstracted as `squeeze(comm_T) -> r`
    (Poseidon RO is untouched host code).
    U1 = (comm_W, comm_E, u, X); W1 = (W, E); U2 = (comm_W, X); W2 = W."""
    q = c.q
    comm_W1, comm_E1, u1, X1 = U1
    Wv1, E1 = W1
    comm_W2, X2 = U2
    T = shape.cross_term(q, Wv1, u1, X1, W2, X2)
    comm_T = commit(c, ck, T)
    r = squeeze(comm_T)
    U = fold_instance(c, comm_W1, comm_E1, u1, X1, comm_W2, X2, comm_T, r)
    W = fold_witness(q, Wv1, E1, W2, T, r)
    return comm_T, T, U, W


# ----------------------------------------------------------------------------------------
# Signed-digit recoding used by the GPU MSM (checked here so the host tests can validate the
# device decomposition independently of the curve arithmetic).
# ----------------------------------------------------------------------------------------
def signed_digits(s: int, c: int, nwin: int) -> List[int]:
    """s = sum d_j 2^(c j), d_j in [-2^(c-1), 2^(c-1)]; the top window absorbs the final carry."""
    out = []
    carry = 0
    for j in range(nwin):
        d = ((s >> (c * j)) & ((1 << c) - 1)) + carry
        carry = 0
        if d > (1 << (c - 1)) and j != nwin - 1:
            d -= 1 << c
            carry = 1
        out.append(d)
    assert sum(d << (c * j) for j, d in enumerate(out)) == s
    return out
