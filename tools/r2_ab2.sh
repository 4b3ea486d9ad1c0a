#!/bin/bash
# A/B of build/variants/*.so against the default library: host-side step breakdown (begin_p = primary critical lane), two runs each
for lib in vimz_b200/libvimz_gpu.so build/variants/*.so; do
  for r in 1 2; do
    echo -n "$lib: "; VIMZ_GPU_LIB=$PWD/$lib python tools/host_breakdown.py 2>&1 | tail -3 | head -1
  done
done
for v in "$@"; do
  VIMZ_GPU_LIB=$PWD/build/variants/$v.so python tools/timeline.py 260 > gpurun_out/ab2_timeline_$v.txt 2>/dev/null
  VIMZ_GPU_LIB=$PWD/build/variants/$v.so python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q 2>&1 | tail -1
done
