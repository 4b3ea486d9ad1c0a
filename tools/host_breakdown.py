#!/usr/bin/env python3
"""Wall-clock breakdown of one fold step on the host side at the fold index bench.py times (PREFOLD 260): latency of
step_begin per curve (graph launch -> both lanes -> result on the host), the stand-in RO, step_end.
usage: python tools/host_breakdown.py [prefold]   (VIMZ_WINDOW_PALLAS etc. are honoured through bench.GpuFold)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

prefold = int(sys.argv[1]) if len(sys.argv) > 1 else 260
prim = bench.GpuFold("pallas", os.environ.get("VIMZ_TL_CIRCUIT", "grayscale"), bench.SEED, 0, torch)
sec = bench.GpuFold("vesta", "secondary", bench.SEED + 1, 0, torch)
for k in range(prefold):
    sec.step(k, True); prim.step(k, True)
acc = {"begin_p": 0.0, "ro_p": 0.0, "end_p": 0.0, "begin_s": 0.0, "ro_s": 0.0, "end_s": 0.0}
N = 100
t_all = time.perf_counter()
for k in range(prefold, prefold + N):
    for tag, f in (("s", sec), ("p", prim)):
        i = k % len(f.wits)
        t0 = time.perf_counter()
        cw, ct = f.acc.step_begin_dev(f.dev_ptr[i], f.X2_bytes[i])
        t1 = time.perf_counter()
        r = ((bench.challenge_from(ct.tobytes(), k) << 256) % f.q).to_bytes(32, "little")
        t2 = time.perf_counter()
        f.acc.step_end(r)
        t3 = time.perf_counter()
        acc["begin_" + tag] += t1 - t0; acc["ro_" + tag] += t2 - t1; acc["end_" + tag] += t3 - t2
prim.eng.sync(); sec.eng.sync()
tot = time.perf_counter() - t_all
print({k: round(v / N * 1e6, 1) for k, v in acc.items()}, "us per step; total", round(tot / N * 1e6, 1), "us/step")
print("primary lanes (0 = commit(T), 1 = commit(W2)):", prim.eng.lane_stats(), "window", prim.ck.window_bits)
print("secondary lanes:", sec.eng.lane_stats(), "window", sec.ck.window_bits)
