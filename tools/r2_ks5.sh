#!/bin/bash
for o in 1 0 1 0; do echo "acc_order=$o"; VIMZ_OPTS=acc_order=$o python tools/host_breakdown.py 2>&1 | tail -3 | head -1; done
python -m pytest tests/test_gpu_r1cs.py -m gpu -x -q > gpurun_out/ks5_pytest.log 2>&1; tail -2 gpurun_out/ks5_pytest.log
python tools/timeline.py 260 > gpurun_out/ks5_timeline.txt 2>/dev/null; tail -26 gpurun_out/ks5_timeline.txt
