#!/usr/bin/env python3
"""Generates tests/golden/*.json from the python big-int oracle (oracle/pyref.py).

The reference's own implementation of the fold path is Rust (nova-snark 0.23.0) and cannot run in
this image, and the reference holds no vectors for it, so these fixtures come from the definition
(naive double-and-add MSM with python integers).  They freeze the oracle's answers so that any later
edit of the oracle or the kernels is checked against a committed file.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as P  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    return "0x%064x" % v


def main():
    rng = random.Random(0xB200)
    vectors = []
    for name, c in P.CURVES.items():
        G = P.generator(c)
        for n, kind in ((1, "uniform"), (5, "edge"), (24, "uniform"), (40, "bits")):
            bases = [P.scalar_mul(c, rng.randrange(1, c.q), G) for _ in range(n)]
            if kind == "edge":
                sc = [0, 1, c.q - 1, 2, (1 << 254) % c.q]
                bases[3] = None            # identity base
                bases[4] = bases[1]        # repeated base
            elif kind == "bits":
                sc = [rng.randrange(2) if rng.random() < 0.9 else rng.randrange(c.q) for _ in range(n)]
            else:
                sc = [rng.randrange(c.q) for _ in range(n)]
            res = P.msm_naive(c, sc, bases)
            vectors.append({
                "curve": name, "kind": kind,
                "bases": [None if b is None else [hx(b[0]), hx(b[1])] for b in bases],
                "scalars": [hx(s) for s in sc],
                "result": None if res is None else [hx(res[0]), hx(res[1])],
            })
    with open(os.path.join(OUT, "msm_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/pyref.py msm_naive)", "vectors": vectors}, f, indent=1)
    print("wrote", len(vectors), "MSM vectors")


if __name__ == "__main__":
    main()
