"""CPU-only: data formats around the fold.  Pixel packing is pinned against a fixture
produced by the reference's own python code; the .r1cs/.wtns readers round-trip files in iden3's layout and the
resulting shape is satisfied by its witness under the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import pyref as P
from vimz_b200 import circom_io as io
from vimz_b200.field import CURVES, mont_to_ints

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def fix():
    return json.load(open(os.path.join(GOLDEN, "pyvimz_rows.json")))


def test_compress_by_rows_matches_reference(fix):
    rgb, gray = np.array(fix["rgb"], np.uint8), np.array(fix["gray"], np.uint8)
    assert io.compress_by_rows(rgb) == fix["original"]          # RGB rows, 128 elements each for HD
    assert io.compress_by_rows(gray) == fix["transformed"]      # grayscale rows
    assert io.compress_by_rows(rgb[:2, :23]) == fix["ragged"]   # ragged tail (23 = 2*10 + 3 pixels)
    assert len(fix["original"][0]) == 128 and fix["width"] == 1280


def test_known_answer_from_circuit_fixture():
    """/root/reference/circuits/src/utils/decompress_input.json: 0x060504030201 -> pixels [1,2,3], [4,5,6], 0..."""
    px = io.unpack_pixels(0x060504030201)
    assert px[0] == (1, 2, 3) and px[1] == (4, 5, 6) and all(p == (0, 0, 0) for p in px[2:])
    assert io.pack_pixels(np.array([[1, 2, 3], [4, 5, 6]], np.uint8)) == [0x060504030201]


def test_pack_unpack_roundtrip(fix):
    rgb = np.array(fix["rgb"], np.uint8)
    vals = io.pack_pixels(rgb[0])
    assert all(v < (1 << 240) for v in vals)
    back = [p for v in vals for p in io.unpack_pixels(v)]
    assert back == [tuple(int(c) for c in p) for p in rgb[0]]
    assert [int(s, 16) for s in fix["original"][0]] == vals


def test_prepare_step_inputs_shapes(fix):
    o, t = fix["original"], fix["transformed"]
    g = io.prepare_step_inputs("grayscale", o, t)
    assert len(g) == 6 and g[2] == {"row_orig": o[2], "row_tran": t[2]}
    z = [["0x00"] * 128]
    b = io.prepare_step_inputs("blur", z + o + z, t)
    assert len(b) == 6 and b[0]["row_orig"] == [z[0], o[0], o[1]] and b[5]["row_orig"][2] == z[0]
    r = io.prepare_step_inputs("resize", o, t[:4], "hd")
    assert len(r) == 2 and len(r[1]["row_orig"]) == 3 and r[1]["row_tran"] == t[2:4]
    assert io.prepare_step_inputs("hash", o, None)[5] == {"row_orig": o[5]}
    with pytest.raises(ValueError):
        io.prepare_step_inputs("rotate", o, t)


def test_r1cs_wtns_roundtrip_and_satisfaction(tmp_path):
    q = CURVES["bn254"].scalar_modulus      # circom's default prime (build_circuits.sh:47)
    # out = x^3 + x + 5 with wires: 0 one, 1 out (public output), 2 x (private input), 3 x^2, 4 x^3
    cons = [([(2, 1)], [(2, 1)], [(3, 1)]),
            ([(3, 1)], [(2, 1)], [(4, 1)]),
            ([(4, 1), (2, 1), (0, 5)], [(0, 1)], [(1, 1)])]
    rp, wp = str(tmp_path / "c.r1cs"), str(tmp_path / "c.wtns")
    io.write_r1cs(rp, q, 5, 1, 0, 1, cons)
    x = 3
    io.write_wtns(wp, q, [1, x ** 3 + x + 5, x, x * x, x ** 3])
    r = io.load_r1cs(rp)
    assert r["prime"] == q and r["n_wires"] == 5 and len(r["constraints"]) == 3 and r["constraints"] == [tuple(c) for c in cons]
    prime, w = io.load_wtns(wp)
    assert prime == q and w[1] == 35
    m, n, nio, A, B, C = io.r1cs_to_shape_coo(r)
    assert (m, n, nio) == (3, 3, 1)
    W, X = io.wtns_to_witness(w, nio, q)

    def triples(M):
        return list(zip(M[0].tolist(), M[1].tolist(), mont_to_ints(M[2], q)))
    S = P.R1CSShape(m, n, nio, triples(A), triples(B), triples(C))
    assert S.is_sat_relaxed(q, mont_to_ints(W, q), [0] * m, 1, mont_to_ints(X, q))
    with pytest.raises(ValueError):
        io.load_wtns(rp)
