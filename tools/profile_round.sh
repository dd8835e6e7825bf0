#!/bin/bash
# ncu evidence of one measurement round (run on the GPU box through gpurun): launch lists + --set full captures.
#   tools/profile_round.sh <tag>      -> gpurun_out/launches_<tag>.csv, launches_<tag>_ecckd.csv, prof_<kernel>_<tag>.ncu-rep
TAG=${1:-r1g}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
# launch lists (cold cache, serialised: compare shares only)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_${TAG}_ecckd_tc64.csv \
    python bench.py --workload tripleclouds_ecckd64 --ncol 10000 --steps 2 --warmup 1 --no-cpu-baseline >> $OUT/bench_under_ncu_${TAG}.log 2>&1
# full captures: one warm launch (skip the first 3) of each top kernel
for k in gas_lw_kernel lw_up_kernel sw_adding_kernel gas_sw_kernel sw_flux_kernel cloud_gen_warp_kernel lw_down_kernel lw_flux_kernel; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline >> $OUT/bench_under_ncu_${TAG}.log 2>&1
  # keep the raw page only (gpurun returns at most 64 MiB): profiles/<tag>_<kernel>_raw.csv
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null && rm -f $OUT/prof_${k}_${TAG}.ncu-rep
done
for k in ckd_lw_kernel ckd_sw_kernel tc_sw_kernel tc_lw_kernel; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python bench.py --workload tripleclouds_ecckd64 --ncol 10000 --steps 1 --warmup 1 --no-cpu-baseline >> $OUT/bench_under_ncu_${TAG}.log 2>&1
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null && rm -f $OUT/prof_${k}_${TAG}.ncu-rep
done
ls -la $OUT | tail -20
