#!/bin/bash
# Round evidence: launch list of the default bench command (per-kernel device times, cold-cache & serialised under ncu:
# compare SHARES), then ncu --set full of the accumulation kernel inside a fold step, of a 2^20 MSM, and of the direct kernel.
TAG=${1:-r1s3}
# ncu replays every kernel: keep the pre-folds short (the captures below then sit at folds 38..41 of the run, where T still
# fits nine 15-bit windows; bench.py itself measures around fold 360, see PREFOLD there)
export VIMZ_BENCH_PREFOLD=32
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 20 > gpurun_out/${TAG}_launches_bench.log 2>&1
W=$(python tools/step_window.py gpurun_out/${TAG}_launches.csv 38 41)
echo "window $W"
python tools/launch_table.py gpurun_out/${TAG}_launches.csv $W > gpurun_out/${TAG}_launches_fold_step.txt
head -30 gpurun_out/${TAG}_launches_fold_step.txt
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 80 -c 4 -o gpurun_out/${TAG}_acc_step \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 2 -c 1 -o gpurun_out/${TAG}_acc_2p20 \
    python tools/profile_msm.py 20 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_msm_direct|k_cross_term_stream|k_reduce_tail|k_msm_combine_all|k_msm_scatter" -s 400 -c 20 -o gpurun_out/${TAG}_step_kernels \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --msm-log2 > /dev/null 2>&1
for f in acc_step acc_2p20 step_kernels; do
  ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_$f.csv 2>/dev/null
done
# source-level hot spots of the accumulation kernel (needs -lineinfo + --import-source), then drop the big reports:
# gpurun copies back at most 64 MiB
ncu -i gpurun_out/${TAG}_acc_2p20.ncu-rep --page source --csv > gpurun_out/${TAG}_acc_2p20_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_step_kernels.ncu-rep gpurun_out/${TAG}_acc_step.ncu-rep
ls -la gpurun_out | head -20
