"""Host-side field/curve constants and buffer helpers (no arithmetic on the hot path).

Buffers crossing the C ABI are numpy uint64 arrays whose rows are the 4 x u64 little-endian
Montgomery limbs of halo2curves / pasta_curves field elements: scalars
`(n, 4)`, affine points `(n, 8)`, Jacobian points `(12,)`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

R = 1 << 256

PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
VESTA_P = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
BN254_P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
BN254_R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


@dataclass(frozen=True)
class CurveInfo:
    name: str
    curve_id: int
    base_modulus: int    # coordinates live here
    scalar_modulus: int  # group order; R1CS scalars live here
    b: int


CURVES = {
    "pallas": CurveInfo("pallas", 0, PALLAS_P, VESTA_P, 5),
    "vesta": CurveInfo("vesta", 1, VESTA_P, PALLAS_P, 5),
    "bn254": CurveInfo("bn254", 2, BN254_P, BN254_R, 3),
    "grumpkin": CurveInfo("grumpkin", 3, BN254_R, BN254_P, BN254_R - 17),
}


def ints_to_mont(vals: Iterable[int], modulus: int) -> np.ndarray:
    """Canonical integers -> (n, 4) uint64 Montgomery rows."""
    buf = b"".join(((int(v) % modulus) * R % modulus).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def mont_to_ints(arr: np.ndarray, modulus: int) -> list:
    """(n, 4) uint64 Montgomery rows -> canonical integers."""
    rinv = pow(R, -1, modulus)
    raw = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") * rinv % modulus for i in range(0, len(raw), 32)]


def raw_to_ints(arr: np.ndarray) -> list:
    raw = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def affine_to_mont(points: Sequence, modulus: int) -> np.ndarray:
    """[(x, y) | None] -> (n, 8) uint64; identity encoded as (0, 0)."""
    flat = []
    for p in points:
        if p is None:
            flat += [0, 0]
        else:
            flat += [p[0], p[1]]
    return ints_to_mont(flat, modulus).reshape(-1, 8)


def mont_to_affine(arr: np.ndarray, modulus: int) -> list:
    vals = mont_to_ints(np.ascontiguousarray(arr).reshape(-1, 4), modulus)
    out = []
    for i in range(0, len(vals), 2):
        out.append(None if vals[i] == 0 and vals[i + 1] == 0 else (vals[i], vals[i + 1]))
    return out


def fr_array(n: int) -> np.ndarray:
    return np.zeros((n, 4), dtype=np.uint64)


def as_fr(a, n: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("expected an (n, 4) uint64 array of Montgomery field elements")
    if n is not None and a.shape[0] != n:
        raise ValueError(f"expected {n} field elements, got {a.shape[0]}")
    return a
