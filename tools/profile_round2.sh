#!/bin/bash
# Round-2 ncu evidence of the default workload (run on the GPU box through gpurun):
#   launch list -> gpurun_out/<tag>_launches.csv ; --set full raw pages of every kernel of the step -> gpurun_out/<tag>_<kernel>_raw.csv
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
ncu --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
bash tools/profile_kernels.sh $TAG gas_prep_kernel gas_col_kernel gas_lw_band_kernel gas_sw_kernel cloud_prep_kernel cloud_optics_kernel \
    cloud_gen_warp_kernel lw_down_kernel lw_up_kernel lw_flux_kernel sw_adding_kernel sw_flux_kernel
python tools/ncu_summary.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.md
cat $OUT/${TAG}_launches.md
