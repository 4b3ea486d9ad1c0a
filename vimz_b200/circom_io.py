"""Data formats on either side of the fold path: the callers'
inputs, not the hot path.  Host-side numpy/python only.

  * vimz input JSON rows: 10 pixels x 24 bit packed into one field element, pixel 0 in the low bits, R in the
    low byte -- restates `compress_by_rows` (/root/reference/pyvimz/pyvimz/img/ops.py:4-33) and the Circom
    decompressor (/root/reference/circuits/src/utils/pixels.circom:6-29); pinned by tests/golden/pyvimz_rows.json,
    generated from the reference's own python code (tests/golden/make_pyvimz_golden.py).
  * per-step input maps: restates `prepare_step_input` (/root/reference/vimz/src/nova_snark_backend/input.rs:57-112).
  * iden3 `.r1cs` / `.wtns` readers (Appendix C.1 / C.2; read by nova-scotia's `load_r1cs`, call site
    /root/reference/vimz/src/nova_snark_backend/folding.rs:22) producing the COO triples `R1CSShape` takes, with
    nova-scotia's wire -> (W || u || X) column mapping.  The format is iden3's published one; nova-scotia is
    un-vendored, so this reader is validated on files written by `write_r1cs` / `write_wtns` below.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .field import ints_to_mont

PACKING_FACTOR = 10  # /root/reference/vimz/src/lib.rs:10


# ---------------------------------------------------------------------------------------------------
# pixel packing
# ---------------------------------------------------------------------------------------------------
def pack_pixels(row: np.ndarray) -> List[int]:
    """One image row (W,) grayscale or (W, 3) RGB -> integers, 10 pixels per 240-bit value."""
    row = np.asarray(row)
    if row.ndim == 1:
        px = row.astype(np.uint64)                      # grayscale: the 24-bit pixel is the value itself
    else:
        r, g, b = (row[:, k].astype(np.uint64) for k in range(3))
        px = r | (g << np.uint64(8)) | (b << np.uint64(16))
    out = []
    for s in range(0, len(px), PACKING_FACTOR):
        v = 0
        for k, p in enumerate(px[s:s + PACKING_FACTOR].tolist()):
            v |= int(p) << (24 * k)
        out.append(v)
    return out


def compress_by_rows(image) -> List[List[str]]:
    """Same strings as the reference's compress_by_rows: fixed 6 hex digits per pixel, most significant pixel first."""
    image = np.asarray(image)
    out = []
    for row in image:
        vals = pack_pixels(row)
        n = row.shape[0]
        strs = []
        for i, v in enumerate(vals):
            npx = min(PACKING_FACTOR, n - i * PACKING_FACTOR)
            strs.append("0x" + format(v, "0%dx" % (6 * npx)))
        out.append(strs)
    return out


def unpack_pixels(value: int, count: int = PACKING_FACTOR) -> List[Tuple[int, int, int]]:
    """Inverse of the packing (what pixels.circom's decompressor constrains): [(r, g, b)] * count."""
    return [((value >> (24 * k)) & 0xFF, (value >> (24 * k + 8)) & 0xFF, (value >> (24 * k + 16)) & 0xFF) for k in range(count)]


# ---------------------------------------------------------------------------------------------------
# per-step inputs (vimz/src/nova_snark_backend/input.rs)
# ---------------------------------------------------------------------------------------------------
ROWS_PER_STEP = {"hd": (3, 2), "fhd": (3, 2), "4k": (2, 1), "8k": (2, 1)}  # Resolution::ratio_to_lower (transformation.rs:115-123)


def prepare_step_inputs(transformation: str, original: Sequence, transformed: Optional[Sequence], resolution: str = "hd",
                        redact: Optional[Sequence] = None) -> List[Dict[str, object]]:
    """-> one {signal name: value} map per fold step, keys `row_orig` / `row_tran` / `block` / `redact`."""
    t = transformation.lower()
    steps = []
    if t in ("brightness", "contrast", "grayscale"):
        for i in range(len(original)):
            steps.append({"row_orig": original[i], "row_tran": transformed[i]})
    elif t in ("blur", "sharpness"):            # original is zero-padded by one row on each side
        for i in range(len(original) - 2):
            steps.append({"row_orig": list(original[i:i + 3]), "row_tran": transformed[i]})
    elif t in ("crop", "hash"):
        for i in range(len(original)):
            steps.append({"row_orig": original[i]})
    elif t == "redact":
        for i in range(len(original)):
            steps.append({"block": original[i], "redact": redact[i]})
    elif t == "resize":
        o, r = ROWS_PER_STEP[resolution.lower()]
        for i in range(len(original) // o):
            steps.append({"row_orig": list(original[i * o:(i + 1) * o]), "row_tran": list(transformed[i * r:(i + 1) * r])})
    else:
        raise ValueError(f"unknown transformation {transformation!r}")
    return steps


# ---------------------------------------------------------------------------------------------------
# iden3 .r1cs / .wtns
# ---------------------------------------------------------------------------------------------------
def _sections(buf: bytes, magic: bytes):
    if buf[:4] != magic:
        raise ValueError(f"bad magic {buf[:4]!r}, expected {magic!r}")
    version, nsec = struct.unpack_from("<II", buf, 4)
    pos, out = 12, {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", buf, pos)
        pos += 12
        out.setdefault(typ, (pos, size))
        pos += size
    return version, out


def load_r1cs(path: str):
    """-> dict(prime, n_wires, n_pub_out, n_pub_in, n_prv_in, constraints=[(A, B, C)]) with A/B/C lists of
    (wire, coeff int)."""
    buf = open(path, "rb").read()
    version, sec = _sections(buf, b"r1cs")
    if version != 1 or 1 not in sec or 2 not in sec:
        raise ValueError("unsupported .r1cs")
    pos, _ = sec[1]
    (fs,) = struct.unpack_from("<I", buf, pos)
    prime = int.from_bytes(buf[pos + 4:pos + 4 + fs], "little")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_cons = struct.unpack_from("<IIIIQI", buf, pos + 4 + fs)
    pos, _ = sec[2]
    cons = []
    for _ in range(n_cons):
        lcs = []
        for _k in range(3):
            (nt,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            terms = []
            for _t in range(nt):
                (w,) = struct.unpack_from("<I", buf, pos)
                terms.append((w, int.from_bytes(buf[pos + 4:pos + 4 + fs], "little")))
                pos += 4 + fs
            lcs.append(terms)
        cons.append(tuple(lcs))
    return {"prime": prime, "n_wires": n_wires, "n_pub_out": n_pub_out, "n_pub_in": n_pub_in, "n_prv_in": n_prv_in,
            "n_labels": n_labels, "constraints": cons}


def write_r1cs(path: str, prime: int, n_wires: int, n_pub_out: int, n_pub_in: int, n_prv_in: int, constraints) -> None:
    fs = 32
    hdr = struct.pack("<I", fs) + prime.to_bytes(fs, "little") + struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in, n_wires,
                                                                            len(constraints))
    body = b""
    for lcs in constraints:
        for terms in lcs:
            body += struct.pack("<I", len(terms))
            for w, cf in terms:
                body += struct.pack("<I", w) + (cf % prime).to_bytes(fs, "little")
    labels = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        for typ, payload in ((1, hdr), (2, body), (3, labels)):
            f.write(struct.pack("<IQ", typ, len(payload)) + payload)


def load_wtns(path: str) -> Tuple[int, List[int]]:
    buf = open(path, "rb").read()
    version, sec = _sections(buf, b"wtns")
    if version != 2 or 1 not in sec or 2 not in sec:
        raise ValueError("unsupported .wtns")
    pos, _ = sec[1]
    (fs,) = struct.unpack_from("<I", buf, pos)
    prime = int.from_bytes(buf[pos + 4:pos + 4 + fs], "little")
    (n,) = struct.unpack_from("<I", buf, pos + 4 + fs)
    pos, _ = sec[2]
    return prime, [int.from_bytes(buf[pos + i * fs:pos + (i + 1) * fs], "little") for i in range(n)]


def write_wtns(path: str, prime: int, values: Sequence[int]) -> None:
    fs = 32
    hdr = struct.pack("<I", fs) + prime.to_bytes(fs, "little") + struct.pack("<I", len(values))
    body = b"".join((int(v) % prime).to_bytes(fs, "little") for v in values)
    with open(path, "wb") as f:
        f.write(b"wtns" + struct.pack("<II", 2, 2))
        f.write(struct.pack("<IQ", 1, len(hdr)) + hdr)
        f.write(struct.pack("<IQ", 2, len(body)) + body)


def r1cs_to_shape_coo(r1cs: dict, modulus: Optional[int] = None):
    """Circom wires -> nova's column space z = (W || u || X) the way nova-scotia's CircomCircuit allocates them:
    wire 0 is the constant one (column num_vars), wires 1 .. n_pub_out + n_pub_in are the public IO X, the rest
    is the witness W.  -> (num_cons, num_vars, num_io, A, B, C) with COO triples (rows, cols, Montgomery vals)."""
    q = modulus or r1cs["prime"]
    num_io = r1cs["n_pub_out"] + r1cs["n_pub_in"]
    num_vars = r1cs["n_wires"] - 1 - num_io

    def col(w):
        if w == 0:
            return num_vars
        if w <= num_io:
            return num_vars + w
        return w - 1 - num_io

    mats = []
    for k in range(3):
        rows, cols, vals = [], [], []
        for i, lcs in enumerate(r1cs["constraints"]):
            for w, cf in lcs[k]:
                rows.append(i); cols.append(col(w)); vals.append(cf)
        mats.append((np.asarray(rows, np.uint32), np.asarray(cols, np.uint32),
                     ints_to_mont(vals, q) if vals else np.zeros((0, 4), np.uint64)))
    return len(r1cs["constraints"]), num_vars, num_io, mats[0], mats[1], mats[2]


def wtns_to_witness(values: Sequence[int], num_io: int, modulus: int):
    """.wtns values (wire order) -> (W, X) Montgomery arrays in nova's split."""
    X = ints_to_mont(values[1:1 + num_io], modulus)
    W = ints_to_mont(values[1 + num_io:], modulus)
    return W, X
