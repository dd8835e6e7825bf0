#!/bin/bash
# one bench line per extra workload (stage times from the serialised pass)
for w in "$@"; do python bench.py --workload $w --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline ${EXTRA} 2>&1 | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print(d['config']['workload']); print('  value %.0f col/s  ms/step %.2f  e2e %.0f col/s launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    print('  ', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})
"; done
