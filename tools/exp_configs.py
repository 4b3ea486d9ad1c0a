import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
sec = bench.GpuFold("vesta", "secondary", bench.SEED + 1, 0, torch)
for k in range(20):
    sec.step(k, True)
for nw in (3, 6, 12, 24):
    t0 = time.time()
    f = bench.GpuFold("pallas", "brightness", bench.SEED, 0, torch, num_witnesses=nw)
    t1 = time.time()
    engines = [f.eng, sec.eng]
    def sr(k):
        sec.step(k, True); f.step(k, True)
    for k in range(nw + 8):
        sr(k)
    ms, _ = bench.timed_region(torch, engines, lambda k: sr(100 + k), 10, None)
    print("witnesses", nw, "build s", round(t1 - t0, 1), "steps/s", round(10 / (ms * 1e-3), 1), f.eng.lane_stats(), flush=True)
    f.close()
