#!/usr/bin/env python3
"""Generates tests/golden/running_hash.json: the reference's own known answers for the row-wise image running hash
(/root/reference/marketplace/image-data/*.hash, produced by the reference's circom witness generator) next to what
oracle/poseidon.py computes from the matching PNGs, plus the accumulator after each of the first six rows of
source_image/HD.png (the rows committed in pyvimz_rows.json) so the check also runs where /root/reference is absent.
Run from the repo root:  python tests/golden/make_running_hash_golden.py"""
import glob
import json
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import poseidon as P  # noqa: E402
from vimz_b200.circom_io import compress_by_rows  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/marketplace/image-data"
final = {}
for path in sorted(glob.glob(os.path.join(DATA, "*.hash"))):
    name = os.path.basename(path)[:-5]
    expected = open(path).read().strip()
    rows = [[int(h, 16) for h in r] for r in compress_by_rows(np.array(Image.open(os.path.join(DATA, name + ".png"))))]
    got = str(P.image_running_hash(rows))
    print(name, "OK" if got == expected else "MISMATCH", got)
    assert got == expected, name
    final[name] = expected
rows6 = json.load(open(os.path.join(OUT, "pyvimz_rows.json")))["original"]
acc, accs = 0, []
for r in rows6:
    acc = P.head_tail_hash(acc, [int(h, 16) for h in r])
    accs.append(str(acc))
json.dump({"generator": "tests/golden/make_running_hash_golden.py",
           "final_hashes_source": "/root/reference/marketplace/image-data/*.hash (reference fixtures, reproduced by oracle/poseidon.py at generation time)",
           "final_hashes": final,
           "hd_first_rows_accumulators": accs,
           "poseidon_1_2": hex(P.poseidon([1, 2]))}, open(os.path.join(OUT, "running_hash.json"), "w"), indent=1)
print("written", os.path.join(OUT, "running_hash.json"))
