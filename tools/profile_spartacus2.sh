#!/bin/bash
# SPARTACUS layer kernels after the cooperative matrix exponential: launch list + ncu --set full of the two heavy launches.
#   tools/profile_spartacus2.sh <tag>
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file $OUT/${TAG}_launches_spartacus.csv \
    python bench.py --workload spartacus_rrtmg --ncol 4000 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
for k in sp_sw_layer_kernel sp_lw_layer_kernel sp_sw_sweep_kernel sp_lw_sweep_kernel; do
  # launches per step: mode 0, mode 1; skip to a mode-1 launch of a later step
  SKIP=3; case $k in *sweep*) SKIP=1;; esac
  $NCU --set full --import-source on -k regex:$k --launch-skip $SKIP -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python bench.py --workload spartacus_rrtmg --ncol 4000 --steps 1 --warmup 1 --no-cpu-baseline >> $OUT/bench_under_ncu_${TAG}.log 2>&1
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page source --print-source cuda,sass --csv > $OUT/${TAG}_${k}_source.csv 2>/dev/null
  python tools/ncu_srclines.py $OUT/${TAG}_${k}_source.csv 60 > $OUT/${TAG}_${k}_lines.txt
  rm -f $OUT/prof_${k}_${TAG}.ncu-rep
done
ls -la $OUT | tail -6
