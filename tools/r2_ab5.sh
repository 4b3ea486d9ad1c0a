#!/bin/bash
# A/B: chunk sums per quad in k_reduce_tail (2 = default, 4, 8: fewer blocks per bit-plane sum)
for lib in vimz_b200/libvimz_gpu.so build/variants/tail4.so build/variants/tail8.so; do
  for r in 1 2; do echo -n "$lib: "; VIMZ_GPU_LIB=$PWD/$lib python tools/host_breakdown.py 2>&1 | tail -3 | head -1; done
done
VIMZ_GPU_LIB=$PWD/build/variants/tail4.so python tools/timeline.py 260 2>/dev/null | tail -14 | grep "reduce\|combine"
