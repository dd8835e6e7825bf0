#!/bin/bash
# quick GPU check: parity tests + bench stage times
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('value %.0f col/s  ms/step %.2f  e2e %.0f col/s' % (d['value'], d['ms_per_step'], d['e2e']['value']))
    print({k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})
"
