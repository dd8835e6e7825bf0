#!/bin/bash
# compute-sanitizer over one small SPARTACUS run (32 columns, 137 levels) and one McICA run: memcheck, then racecheck of the
# shared-memory protocols (warp-cooperative matrix exponential, TMA rings).   tools/sanitize_spartacus.sh [out-dir]
OUT=${1:-gpurun_out}
mkdir -p $OUT
cat > /tmp/san_case.py <<'P'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from ecrad_b200.radiation_interface import setup_radiation
raw = {k: np.array(v, dtype=np.float64) for k, v in np.load('tests/golden/ecrad_meridian_inputs.npz').items()}
for kw in (dict(sw_solver_name='SPARTACUS', lw_solver_name='SPARTACUS', do_3d_effects=True), dict(use_aerosols=True)):
    cfg = RadiationConfig(**kw).consolidate()
    h = setup_radiation(cfg)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), 32, 137)
    print(kw.get('sw_solver_name', 'McICA'), float(np.nansum(out['sw_up'])), float(np.nansum(out['lw_up'])))
    h.finalize()
P
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > $OUT/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|SPARTACUS|McICA" $OUT/sanitizer_$tool.log | head -12
done
