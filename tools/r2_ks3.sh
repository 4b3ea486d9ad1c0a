#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/ks3_pytest.log 2>&1; tail -3 gpurun_out/ks3_pytest.log
python tools/pipe_exp.py 260 100 2>&1 | tail -8
VIMZ_OPTS=stage_commit=1 python tools/pipe_exp.py 260 100 2>&1 | tail -8
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/timeline.py 260 > gpurun_out/ks3_timeline.txt 2>/dev/null; tail -42 gpurun_out/ks3_timeline.txt
