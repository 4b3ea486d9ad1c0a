#!/usr/bin/env python3
"""Generates tests/golden/fold_chain.json from the python big-int oracle (oracle/pyref.py): a hand-sized relaxed-R1CS
(6 constraints, 5 variables, 2 public IO) folded three times on every curve with fixed 128-bit challenges, starting from
the default (all-zero) relaxed instance like RecursiveSNARK does.  Every intermediate of NIFS::prove is recorded -- T, comm_T,
comm_W2, and the folded (W, E, u, X, comm_W, comm_E) -- as canonical integers / affine coordinates, so the C restatement and
the CUDA path are checked against a committed file and not only against each other.  Run from the repo root:
    python tests/golden/make_fold_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as P  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
M, N, IO, STEPS = 6, 5, 2, 3


def hx(v):
    return "0x%064x" % v


def pt(a):
    return None if a is None else [hx(a[0]), hx(a[1])]


def shape_for(q, rng):
    """z = (w0..w4, u, x0, x1).  Rows: booleanity of w0, a product, a linear row with a big coefficient, a row using the
    public IO and u, an empty-C row with duplicate (row, col) entries in A, and a row that is empty in all three matrices."""
    one, neg = 1, q - 1
    u, x0, x1 = N, N + 1, N + 2
    big = rng.randrange(q)
    A = [(0, 0, one), (1, 1, one), (2, 2, big), (2, 3, one), (3, x0, one), (3, 4, 2), (4, 1, one), (4, 1, one)]
    B = [(0, 0, one), (0, u, neg), (1, 2, one), (2, u, one), (3, x1, one), (4, 3, neg)]
    C = [(1, 3, one), (2, 4, one), (3, 0, one), (3, u, 5)]
    return P.R1CSShape(M, N, IO, A, B, C)


def main():
    rng = random.Random(0xF01D)
    out = {"about": "NIFS fold chain from the default relaxed instance; python big-int model (oracle/pyref.py)", "curves": {}}
    for name, c in P.CURVES.items():
        q = c.q
        G = P.generator(c)
        ck = [P.scalar_mul(c, rng.randrange(1, q), G) for _ in range(max(M, N))]
        shape = shape_for(q, rng)
        W1, E1, u1, X1 = [0] * N, [0] * M, 0, [0] * IO
        cW, cE = None, None
        steps = []
        for k in range(STEPS):
            W2 = [rng.randrange(2), rng.randrange(q), rng.randrange(256), rng.randrange(q), rng.randrange(q)]
            X2 = [rng.randrange(q), rng.randrange(q)]
            r = rng.randrange(1 << 128)
            comm_W2 = P.commit(c, ck, W2)
            comm_T, T, U, W = P.nifs_prove(c, ck, shape, (cW, cE, u1, X1), (W1, E1), (comm_W2, X2), W2, lambda _ct: r)
            cW, cE, u1, X1 = U
            W1, E1 = W
            steps.append({"W2": [hx(v) for v in W2], "X2": [hx(v) for v in X2], "r": hx(r), "comm_W2": pt(comm_W2), "T": [hx(v) for v in T],
                          "comm_T": pt(comm_T), "W": [hx(v) for v in W1], "E": [hx(v) for v in E1], "u": hx(u1), "X": [hx(v) for v in X1],
                          "comm_W": pt(cW), "comm_E": pt(cE)})
        # (the fresh witnesses are random, not satisfying: this fixture pins the ARITHMETIC of the fold, not satisfiability)
        out["curves"][name] = {"num_cons": M, "num_vars": N, "num_io": IO,
                               "A": [[r_, c_, hx(v)] for r_, c_, v in shape.A], "B": [[r_, c_, hx(v)] for r_, c_, v in shape.B],
                               "C": [[r_, c_, hx(v)] for r_, c_, v in shape.C], "ck": [pt(b) for b in ck], "steps": steps}
    with open(os.path.join(OUT, "fold_chain.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote fold_chain.json")


if __name__ == "__main__":
    main()
