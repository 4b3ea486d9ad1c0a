"""vimz_host -- host-side layouts and synthetic inputs shared by the product binding (vimz_b200), the CPU oracle
harness and bench.py.  Pure numpy / python integers: importing this package does NOT load libvimz_gpu.so, so the
reference arm of bench.py (`--impl reference`) and the CPU tests can build their inputs without touching the product.

  field      4 x u64 Montgomery buffer helpers and the curve constants of the two cycles
  synthetic  step-circuit shapes / witnesses with the published sizes, scalar distributions, closed-form MSM sums
"""
