set -x
OUT=gpurun_out; mkdir -p $OUT
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/r2x_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_r2x.log 2>&1
python tools/ncu_summary.py launches $OUT/r2x_launches.csv > $OUT/r2x_launches.md; head -16 $OUT/r2x_launches.md
STEPS=5 bash tools/bench_workloads.sh cloudless_ecckd32 mcica_ecckd32 tripleclouds_ecckd64 tripleclouds_rrtmg tripleclouds_mixed spartacus_rrtmg > $OUT/bench_r2x_workloads.txt 2>&1; cat $OUT/bench_r2x_workloads.txt
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > $OUT/bench_r2x_reference.json; cut -c1-200 $OUT/bench_r2x_reference.json
