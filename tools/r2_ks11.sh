#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q > gpurun_out/ks11_pytest.log 2>&1; tail -2 gpurun_out/ks11_pytest.log
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/host_breakdown.py 2>&1 | tail -3 | head -1
python tools/timeline.py 260 > gpurun_out/ks11_timeline.txt 2>/dev/null; tail -24 gpurun_out/ks11_timeline.txt | grep "reduce\|combine"
