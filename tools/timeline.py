#!/usr/bin/env python3
"""Per-kernel timeline of one fold step as it really runs (both stream lanes concurrent, CUDA-graph replay), from CUPTI
through torch.profiler -- the image has no nsys.  Prints, for the last profiled step, every kernel / memcpy with its
stream, start offset and duration, and the host-side latency of the two step_begin calls.
usage: python tools/timeline.py [prefold] [resident|e2e|staged]"""
import json, os, sys, time, tempfile
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

prefold = int(sys.argv[1]) if len(sys.argv) > 1 else 260
mode = sys.argv[2] if len(sys.argv) > 2 else "resident"
resident = mode == "resident"


def one(k):
    if mode == "staged":
        i = k % len(sec.wits)
        sec.acc.step_begin_async(sec.pin_np[i], sec.X2_bytes[i])
        prim.stage(k)
        cw, ct = sec.acc.step_wait()
        sec.acc.step_end(((bench.challenge_from(ct.tobytes(), k) << 256) % sec.q).to_bytes(32, "little"))
        prim.step_staged(k)
    else:
        sec.step(k, resident); prim.step(k, resident)
prim = bench.GpuFold("pallas", os.environ.get("VIMZ_TL_CIRCUIT", "grayscale"), bench.SEED, 0, torch)
sec = bench.GpuFold("vesta", "secondary", bench.SEED + 1, 0, torch)
for k in range(prefold):
    sec.step(k, True); prim.step(k, True)
for k in range(prefold, prefold + 4):
    one(k)
marks = []
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for k in range(prefold + 4, prefold + 7):
        t0 = time.perf_counter_ns()
        if mode == "staged":
            i = k % len(sec.wits)
            sec.acc.step_begin_async(sec.pin_np[i], sec.X2_bytes[i])
            prim.stage(k)
            cw, ct = sec.acc.step_wait()
            sec.acc.step_end(((bench.challenge_from(ct.tobytes(), k) << 256) % sec.q).to_bytes(32, "little"))
        else:
            sec.step(k, resident)
        t1 = time.perf_counter_ns()
        if mode == "staged":
            prim.step_staged(k)
        else:
            prim.step(k, resident)
        t2 = time.perf_counter_ns()
        marks.append((t0, t1, t2))
    prim.eng.sync(); sec.eng.sync()
path = os.path.join(tempfile.gettempdir(), "vimz_trace.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
gpu.sort(key=lambda e: e["ts"])
# last step = after the second-to-last secondary cross term ... simply take the last third by count of cross-term kernels
cross = [i for i, e in enumerate(gpu) if "k_cross_term_stream" in e["name"]]
start = cross[-2] if len(cross) >= 2 else 0
# include the few nodes (memcpy / memset / digits) that precede the secondary cross term of the last step
while start > 0 and gpu[start]["ts"] - gpu[start - 1]["ts"] - gpu[start - 1].get("dur", 0) < 15 and "k_point_scale" not in gpu[start - 1]["name"] and "k_axpy" not in gpu[start - 1]["name"]:
    start -= 1
t0 = gpu[start]["ts"]
print(f"{'start us':>9s} {'dur us':>8s} {'stream':>6s}  name")
import re
for e in gpu[start:]:
    name = re.sub(r"^void ", "", e["name"]); name = re.sub(r"\(.*", "", name).replace("vimz::", "")
    print(f"{e['ts'] - t0:9.1f} {e.get('dur', 0):8.1f} {str(e.get('args', {}).get('stream', '?')):>6s}  {name[:70]}")
for (a, b, c) in marks:
    print(f"host: secondary step {(b - a) / 1e3:.1f} us, primary step {(c - b) / 1e3:.1f} us")
