#!/bin/bash
# device-resident throughput vs column tile size (ECRAD_B200_TILE), default workload
for t in 2048 4096 5000 10000; do
  echo "tile $t"
  ECRAD_B200_TILE=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('  value %.0f col/s  ms/step %.2f  e2e %.0f col/s' % (d['value'], d['ms_per_step'], d['e2e']['value']))
"
done
