#!/bin/bash
# ncu --set full captures of named kernels of the default bench workload (run on the GPU box through gpurun):
#   tools/profile_kernels.sh <tag> <kernel-regex>...   -> gpurun_out/<tag>_<kernel>_raw.csv, gpurun_out/<tag>_<kernel>_lines.txt
# BENCH_ARGS adds arguments to bench.py (e.g. --workload ...).
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
for k in "$@"; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline $BENCH_ARGS >> $OUT/bench_under_ncu_${TAG}.log 2>&1
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page source --print-source cuda,sass --csv > $OUT/src_${k}.csv 2>/dev/null
  python tools/ncu_srclines.py $OUT/src_${k}.csv 45 > $OUT/${TAG}_${k}_lines.txt 2>&1
  rm -f $OUT/prof_${k}_${TAG}.ncu-rep $OUT/src_${k}.csv
done
python tools/ncu_summary.py table $OUT/${TAG}_*_raw.csv > $OUT/${TAG}_table.md 2>&1
cat $OUT/${TAG}_table.md
