#!/bin/bash
python tools/pipe_exp.py 260 100 2>&1 | tail -8
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/timeline.py 260 > gpurun_out/ks4_timeline.txt 2>/dev/null; tail -42 gpurun_out/ks4_timeline.txt
