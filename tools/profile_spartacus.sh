#!/bin/bash
# SPARTACUS measurement round (run on the GPU box through gpurun): bench line, ncu launch list, --set full captures of the
# four SPARTACUS kernels.   tools/profile_spartacus.sh <tag> [ncol]
TAG=${1:-r1i}
NCOL=${2:-50000}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python bench.py --workload spartacus_rrtmg --ncol $NCOL --steps 3 --warmup 3 > $OUT/bench_${TAG}_spartacus.json 2> $OUT/bench_${TAG}_spartacus.err
tail -c 3000 $OUT/bench_${TAG}_spartacus.json
python bench.py --workload spartacus_rrtmg --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_spartacus_reference.json 2>> $OUT/bench_${TAG}_spartacus.err
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $OUT/${TAG}_launches_spartacus.csv \
    python bench.py --workload spartacus_rrtmg --ncol 4000 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
for k in sp_sw_layer_kernel sp_lw_layer_kernel sp_sw_sweep_kernel sp_lw_sweep_kernel; do
  $NCU --set full --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o $OUT/prof_${k}_${TAG} \
      python bench.py --workload spartacus_rrtmg --ncol 4000 --steps 1 --warmup 1 --no-cpu-baseline >> $OUT/bench_under_ncu_${TAG}.log 2>&1
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page raw --csv > $OUT/${TAG}_${k}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_${k}_${TAG}.ncu-rep --page source --csv > $OUT/${TAG}_${k}_source.csv 2>/dev/null
  rm -f $OUT/prof_${k}_${TAG}.ncu-rep
done
ls -la $OUT | tail -12
