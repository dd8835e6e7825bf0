#!/bin/bash
# register-budget sweep of the SPARTACUS kernels (resident CTAs per SM asked of the compiler), 20 000 columns
for cfg in "2 2" "4 2" "2 4" "4 4"; do   # compiled-in budgets: 2 and 4 CTAs/SM (3, 6 and 8 were measured and dropped)
  set -- $cfg
  echo "== MINB_LAYER=$1 MINB_SWEEP=$2"
  ECRAD_B200_SP_MINB_LAYER=$1 ECRAD_B200_SP_MINB_SWEEP=$2 bash tools/bench_workloads.sh spartacus_rrtmg 2>&1 | tail -2
done
