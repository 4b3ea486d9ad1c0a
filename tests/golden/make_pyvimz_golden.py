#!/usr/bin/env python3
"""Generates tests/golden/pyvimz_rows.json by IMPORTING the reference's own python code
(/root/reference/pyvimz/pyvimz/img/{ops,transformations}.py) on the reference's own image
(/root/reference/source_image/HD.png, first 6 rows) -- the reference is only present in the build container,
so the fixture is committed.  Run from the repo root:  python tests/golden/make_pyvimz_golden.py"""
import json
import os
import sys

import numpy as np
from PIL import Image

sys.path.insert(0, "/root/reference/pyvimz")
from pyvimz.img.ops import compress_by_rows  # noqa: E402
from pyvimz.img.transformations import convert_to_grayscale  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
img = Image.open("/root/reference/source_image/HD.png").convert("RGB")
rows = 6
crop = img.crop((0, 0, img.size[0], rows))
gray = convert_to_grayscale(crop)
fix = {
    "generator": "tests/golden/make_pyvimz_golden.py (reference pyvimz.img.ops.compress_by_rows / convert_to_grayscale)",
    "source": "source_image/HD.png rows 0..5", "width": img.size[0],
    "rgb": np.asarray(crop).tolist(), "gray": np.asarray(gray).tolist(),
    "original": compress_by_rows(np.asarray(crop)), "transformed": compress_by_rows(gray),
    "ragged": compress_by_rows(np.asarray(crop)[:2, :23]),
}
json.dump(fix, open(os.path.join(OUT, "pyvimz_rows.json"), "w"))
print("rows", rows, "elements per row", len(fix["original"][0]), "bytes", os.path.getsize(os.path.join(OUT, "pyvimz_rows.json")))
