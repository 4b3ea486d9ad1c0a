#!/bin/bash
# Round-2 evidence on the final build, at the fold index bench.py times (PREFOLD 260: T reaches into its tenth window, a giant
# bucket exists).  Outputs under gpurun_out/ (summaries are copied to profiles/ by hand).
TAG=${1:-r2}
BENCH="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-configs --msm-log2"
# 1. launch list (per-kernel device time; ncu serialises the lanes and runs cold: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > /dev/null 2>&1
W=$(python tools/step_window.py gpurun_out/${TAG}_launches.csv 266 269)
python tools/launch_table.py gpurun_out/${TAG}_launches.csv $W > gpurun_out/${TAG}_launches_fold_step.txt
cat gpurun_out/${TAG}_launches_fold_step.txt
rm -f gpurun_out/${TAG}_launches.csv
# 2. the real timeline (both lanes concurrent, graph replay) from CUPTI
python tools/timeline.py 260 > gpurun_out/${TAG}_timeline_fold_step.txt 2>/dev/null
python tools/timeline.py 260 e2e > gpurun_out/${TAG}_timeline_fold_step_e2e.txt 2>/dev/null
# 3. ncu --set full: accumulation (2 launches per step on the primary curve -> skip 2 * 264), then the other step kernels
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 528 -c 4 -o gpurun_out/${TAG}_acc_step $BENCH > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_acc_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_acc_step.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/${TAG}_acc_step.csv > gpurun_out/${TAG}_ncu_msm_accumulate_fold_step.txt
ncu -i gpurun_out/${TAG}_acc_step.ncu-rep --page source --csv > gpurun_out/${TAG}_acc_step_source.csv 2>/dev/null
python tools/ncu_hot.py k_msm_accumulate < gpurun_out/${TAG}_acc_step_source.csv > gpurun_out/${TAG}_ncu_msm_accumulate_hotspots.txt 2>/dev/null
# per step: matvec x2, finish x2, direct x2, tail x2, combine x2, scatter x2, chunks x2, masked sum, pow2 parts, scale-add parts = 17 matching launches -> skip 17 * 264
ncu --set full --clock-control none -k "regex:k_matvec_stream|k_cross_finish|k_msm_direct|k_reduce_tail|k_msm_combine_all|k_msm_scatter|k_reduce_chunks|k_masked_base_sum|k_point_pow2_parts|k_point_scale_add_parts" -s 4488 -c 17 \
    -o gpurun_out/${TAG}_step_kernels $BENCH > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_step_kernels.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_kernels.csv 2>/dev/null
python tools/ncu_summary.py < gpurun_out/${TAG}_step_kernels.csv > gpurun_out/${TAG}_ncu_fold_step_kernels.txt
# 2^20 MSM accumulate
ncu --set full --clock-control none -k regex:k_msm_accumulate -s 2 -c 1 -o gpurun_out/${TAG}_acc_2p20 python tools/profile_msm.py 20 4 > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_acc_2p20.ncu-rep --page raw --csv | python tools/ncu_summary.py > gpurun_out/${TAG}_ncu_msm_accumulate_2p20.txt 2>/dev/null
rm -f gpurun_out/${TAG}_step_kernels.ncu-rep gpurun_out/${TAG}_acc_step.ncu-rep gpurun_out/${TAG}_acc_2p20.ncu-rep gpurun_out/${TAG}_acc_step_source.csv
# 4. compute-sanitizer over the kernels new or changed this round (TMA/cp.async mat-vec, cross finish, deferred giants, aggregated scatter,
#    cached products, library-level sharded step)
SEL='(cached_products or row_classes or skewed or big_bucket or library_level or (chain_vs_oracle and pallas) or booleanity or (staged_async and opts1)) and not direct'
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool  pytest -k \"$SEL\""
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_msm.py tests/test_gpu_r1cs.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|rror" | head -8
done > gpurun_out/${TAG}_sanitizer.txt 2>&1
cat gpurun_out/${TAG}_sanitizer.txt
./tools/mul_latency > gpurun_out/${TAG}_mul_latency.json
ls -la gpurun_out | grep ${TAG}_
