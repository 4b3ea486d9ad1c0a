#!/usr/bin/env python3
"""Run a few Pallas MSMs (uniform scalars, 2^LOG2 resident points) -- the target of the ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 2 -c 1 -o gpurun_out/acc python tools/profile_msm.py 20"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vimz_b200  # noqa: E402
from vimz_b200 import CommitmentEngine, CommitmentKey  # noqa: E402
from vimz_b200 import synthetic as S  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 1 << log2n
eng = vimz_b200.Engine("pallas", 0)
d_bases = torch.empty(n * 8, dtype=torch.int64, device="cuda")
vimz_b200._lib.check(vimz_b200.lib.vimz_gen_bases_dev(eng._h, 77, 1234577, n, d_bases.data_ptr()))
ck = CommitmentKey.from_device(eng, d_bases.data_ptr(), n)
sc = torch.from_numpy(S.uniform_scalars_mont(n, eng.curve.scalar_modulus, 1).view(np.int64)).cuda()
for _ in range(iters):
    out = CommitmentEngine.commit_dev(ck, sc.data_ptr(), n)
print("msm ok", log2n, ck.window_bits, ck.num_windows, out[:2])
